#!/usr/bin/env python
"""Sample-sharding invariance on real GPUs: predict_with_uncertainty with the S samples sharded over the
ranks of a torchrun job (NCCL all-reduce of the moments) vs all S samples on one GPU.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/mc_shard_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import pno_physics_bench_b200 as pno  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = pno.ProbabilisticNeuralOperator(3, 128, 2, 8).to(dev)
x = torch.randn(4, 3, 32, 32, generator=torch.Generator().manual_seed(1)).to(dev)
S = 13                                          # uneven split
mean, std = model.predict_with_uncertainty(x, num_samples=S, seed=99)                             # sharded over the ranks
mean1, std1 = model.predict_with_uncertainty(x, num_samples=S, seed=99, shard_samples=False)      # all samples on this rank
em = float((mean - mean1).norm() / mean1.norm())
es = float((std - std1).norm() / std1.norm())
if rank == 0:
    print(f"world {world}: rel-L2 sharded vs unsharded: mean {em:.2e}, std {es:.2e}")
    print("MC_SHARD_OK" if em < 1e-6 and es < 1e-6 else "MC_SHARD_FAILED")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
