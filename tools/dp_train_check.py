#!/usr/bin/env python
"""Data-parallel invariance on real GPUs: one PNOTrainer step on the global batch on one GPU vs the
same step with the batch sharded over the ranks of a torchrun job (NCCL gradient all-reduce).  Prints the relative
difference of the loss and of every parameter after the step.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_train_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import pno_physics_bench_b200 as pno  # noqa: E402
from pno_physics_bench_b200.training import ELBOLoss, PNOTrainer  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def make():
    torch.manual_seed(0)
    return pno.ProbabilisticNeuralOperator(3, 128, 2, 8, engine="tf32x3")


g = torch.Generator().manual_seed(1)
B = 4 * world
x, y = torch.randn(B, 3, 32, 32, generator=g), torch.randn(B, 1, 32, 32, generator=g)

# sharded step: every rank takes its slice, same weight-noise key, activation noise indexed by the global batch offset
m_dp = make()
tr_dp = PNOTrainer(m_dp, ELBOLoss(kl_weight=1e-4), device=str(dev), seed=7)
per = B // world
loss_dp = tr_dp.train_step(x[rank * per:(rank + 1) * per], y[rank * per:(rank + 1) * per], batch_offset=rank * per)

# reference: the whole batch on every GPU with the gradient synchronisation switched off
m_1 = make()
tr_1 = PNOTrainer(m_1, ELBOLoss(kl_weight=1e-4), device=str(dev), seed=7)
if tr_1.grad_sync is not None:
    tr_1.grad_sync.enabled = False
loss_1 = tr_1.train_step(x, y, batch_offset=0)
worst = 0.0
for (k, a), (_, b_) in zip(m_dp.state_dict().items(), m_1.state_dict().items()):
    a, b_ = (torch.view_as_real(t) if t.is_complex() else t for t in (a, b_))
    d = float((a.double() - b_.double()).norm() / b_.double().norm().clamp_min(1e-30))
    worst = max(worst, d)
if rank == 0:
    print(f"world {world}: max relative parameter difference after one step (sharded vs whole batch) = {worst:.3e}")
    print("DP_CHECK_OK" if worst < 1e-5 else "DP_CHECK_FAILED")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
