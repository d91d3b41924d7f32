"""Host-side logic of the multi-GPU path on CPU: world size 2, gloo backend.

Covers what does not need a GPU: the MC-sample shard ranges, the shared Philox seed broadcast, the moment all-reduce
(sum y, sum y^2 merged across ranks gives the unsharded mean / unbiased std) and the bucketed gradient averaging, including
the strided kernel-native spectral parameters and complex gradients.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pno_physics_bench_b200 import distributed as D
from pno_physics_bench_b200 import functional as F_


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = {}
        # --- shared seed: every rank ends up with rank 0's draw
        torch.manual_seed(100 + rank)
        res["seed"] = D.shared_seed(None, torch.device("cpu"))
        res["seed_fixed"] = D.shared_seed(77, torch.device("cpu"))
        # --- MC moments: S = 7 "samples" (uneven split), merged moments equal the unsharded statistics
        S = 7
        g = torch.Generator().manual_seed(5)
        ys = torch.randn(S, 3, 1, 4, 4, generator=g, dtype=torch.float64) * 3 + 1
        first, count = D.mc_shard(S)
        res["shard"] = (first, count)
        s1 = torch.zeros(3, 1, 4, 4, dtype=torch.float64)
        s2 = torch.zeros_like(s1)
        for s in range(first, first + count):
            s1 += ys[s]
            s2 += ys[s] ** 2
        D.allreduce_moments(s1, s2)
        mean = s1 / S
        std = torch.sqrt(torch.clamp(s2 - S * mean * mean, min=0) / (S - 1))     # pno_mc_finalize's formula
        res["mean_err"] = float((mean - ys.mean(0)).abs().max())
        res["std_err"] = float((std - ys.std(0)).abs().max())
        # --- gradient averaging: a strided spectral parameter (complex), its log-var, and a plain conv weight
        w = torch.nn.Parameter(F_.alloc_spectral(2, 3, 4, 5, torch.complex64))
        lv = torch.nn.Parameter(F_.alloc_spectral(2, 3, 4, 5, torch.float32))
        cw = torch.nn.Parameter(torch.zeros(3, 2, 1, 1))
        nograd = torch.nn.Parameter(torch.zeros(4))
        gg = torch.Generator().manual_seed(9)
        full = [torch.randn(2, 2, 3, 4, 5, 2, generator=gg), torch.randn(2, 2, 3, 4, 5, generator=gg),
                torch.randn(2, 3, 2, 1, 1, generator=gg)]
        w.grad = torch.view_as_complex(full[0][rank].contiguous()).clone()
        w.grad = F_.alloc_spectral(2, 3, 4, 5, torch.complex64).copy_(w.grad)      # strided like the kernels' outputs
        lv.grad = F_.alloc_spectral(2, 3, 4, 5, torch.float32).copy_(full[1][rank])
        cw.grad = full[2][rank].clone()
        D.allreduce_gradients([w, lv, cw, nograd], large_bytes=256)     # w, lv: reduced in their own storage; cw: the small bucket
        res["gw"] = float((torch.view_as_real(w.grad) - full[0].mean(0)).abs().max())
        res["glv"] = float((lv.grad - full[1].mean(0)).abs().max())
        res["gcw"] = float((cw.grad - full[2].mean(0)).abs().max())
        res["gnone"] = nograd.grad is None            # a parameter without a gradient stays without one (like a single-GPU run)
        # --- the same through autograd hooks (the trainer's path): collectives are launched while backward is running
        a = torch.nn.Parameter(torch.zeros(300))      # "large" (>= 256 bytes): async all-reduce in place from its hook
        b = torch.nn.Parameter(torch.zeros(5))        # small: flat bucket at finish()
        unused = torch.nn.Parameter(torch.zeros(3))
        sync = D.GradientSynchronizer([a, b, unused], large_bytes=256)
        xa = torch.arange(300, dtype=torch.float32) * (rank + 1)
        ((a * xa).sum() + (b * (rank + 2.0)).sum()).backward()
        sync.finish()
        res["hook_a"] = float((a.grad - torch.arange(300, dtype=torch.float32) * 1.5).abs().max())
        res["hook_b"] = float((b.grad - 2.5).abs().max())
        res["hook_unused"] = unused.grad is None
        sync.remove()
        # --- mode-parallel exchanges: [all modes, local batch] <-> [my modes, global batch] and back
        from pno_physics_bench_b200 import mode_parallel as MP
        M, Bl, C = 12, 3, 5
        gspec = torch.randn(2, M, world * Bl, C, generator=torch.Generator().manual_seed(11))      # the global spectrum (both planes)
        mine = gspec[:, :, rank * Bl:(rank + 1) * Bl].contiguous()
        sh = MP.modes_from_batch(mine)
        n = M // world
        res["a2a_fwd"] = float((sh - gspec[:, rank * n:(rank + 1) * n]).abs().max())
        back = MP.batch_from_modes(sh)
        res["a2a_back"] = float((back - mine).abs().max())
        res["windows"] = MP.mode_windows(3, 2, rank, world)
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world_size_2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert r0["seed"] == r1["seed"] and r0["seed_fixed"] == r1["seed_fixed"] == 77
    assert r0["shard"] == (0, 4) and r1["shard"] == (4, 3)
    for r in (r0, r1):
        assert r["mean_err"] < 1e-12 and r["std_err"] < 1e-12
        assert r["gw"] < 1e-6 and r["glv"] < 1e-6 and r["gcw"] < 1e-6 and r["gnone"] is True
        assert r["hook_a"] < 1e-6 and r["hook_b"] < 1e-6 and r["hook_unused"] is True
        assert r["a2a_fwd"] == 0.0 and r["a2a_back"] == 0.0
    assert r0["windows"] == [(0, 6)] and r1["windows"] == [(6, 6)]
