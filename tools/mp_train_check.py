#!/usr/bin/env python
"""Mode-parallel invariance on real GPUs: one PNOTrainer step with the spectral weights sharded by modes and
the batch sharded over the ranks of a torchrun job vs the same step of the unsharded model on the whole batch on one GPU.  Prints
the relative difference of the loss and of every parameter after the step.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/mp_train_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import pno_physics_bench_b200 as pno  # noqa: E402
from pno_physics_bench_b200.mode_parallel import ModeParallelPNO  # noqa: E402
from pno_physics_bench_b200.training import ELBOLoss, PNOTrainer  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def make():
    torch.manual_seed(0)
    m = pno.ProbabilisticNeuralOperator(3, 128, 2, 8, engine="tf32x3")
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():           # trained-like weights: the mean path matters, not only the noise
        for layer in m.fourier_layers:
            for w, lv in ((layer.weights1, layer.weights1_log_var), (layer.weights2, layer.weights2_log_var)):
                w.copy_(torch.complex(torch.randn(w.shape, generator=g), torch.randn(w.shape, generator=g)) / 16)
                lv.copy_(torch.rand(lv.shape, generator=g) * 4 - 10)
    return m


# PNO_MP_BATCH: batch per rank (default 4); PNO_MP_MAX_BATCH: batch rows per channel-mix launch (default 128) -- small values
# exercise the batch chunking of the global batch (8 ranks x 32 fields) and the K chunks of the weight-gradient kernel
import pno_physics_bench_b200.mode_parallel as mp_mod  # noqa: E402
if os.environ.get("PNO_MP_MAX_BATCH"):
    mp_mod.MIX_MAX_BATCH = int(os.environ["PNO_MP_MAX_BATCH"])
g = torch.Generator().manual_seed(1)
B = int(os.environ.get("PNO_MP_BATCH", "4")) * world
x, y = torch.randn(B, 3, 32, 32, generator=g), torch.randn(B, 1, 32, 32, generator=g)

# sharded step: modes and batch split over the ranks
mp = ModeParallelPNO(make(), engine="tf32x3").to(dev)
tr_mp = PNOTrainer(mp, ELBOLoss(kl_weight=1e-4), device=str(dev), seed=7)
per = B // world
loss_mp = tr_mp.train_step(x[rank * per:(rank + 1) * per], y[rank * per:(rank + 1) * per], batch_offset=rank * per)
sd_mp = mp.gather_state_dict()

# reference: the unsharded model on the whole batch, on every GPU, without gradient synchronisation
m_1 = make()
tr_1 = PNOTrainer(m_1, ELBOLoss(kl_weight=1e-4), device=str(dev), seed=7)
if tr_1.grad_sync is not None:
    tr_1.grad_sync.enabled = False
loss_1 = tr_1.train_step(x, y, batch_offset=0)
worst, worst_key = 0.0, ""
for k, b_ in m_1.state_dict().items():
    a = sd_mp[k]
    a, b_ = (torch.view_as_real(t.contiguous()) if t.is_complex() else t for t in (a, b_))
    d = float((a.double() - b_.double()).norm() / b_.double().norm().clamp_min(1e-30))
    if d > worst:
        worst, worst_key = d, k
kl_d = abs(float(loss_mp["kl"]) - float(loss_1["kl"])) / abs(float(loss_1["kl"]))
if rank == 0:
    print(f"world {world}: max relative parameter difference after one step (mode-parallel vs unsharded) = {worst:.3e} ({worst_key}); "
          f"KL rel diff {kl_d:.2e}")
    print("MP_CHECK_OK" if worst < 1e-5 and kl_d < 1e-6 else "MP_CHECK_FAILED")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
