#!/usr/bin/env python
"""CUDA-event time of every kernel of one PNO block at the cfg-3 shape (64x64, hid 256, modes 20, batch 64), L2 flushed
between launches, next to each kernel's algorithmic HBM bytes.

    python tools/kernel_times.py [engine ...]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pno_physics_bench_b200 import _lib  # noqa: E402


def block_kernel_times(engine, B=64, C=256, H=64, W=64, m=20, iters=5, device="cuda"):
    lib = _lib.load()
    M = 2 * m * m
    plan = _lib.plan(H, W, m, m, torch.device(device, torch.cuda.current_device()))
    g = torch.Generator(device=device).manual_seed(0)
    x = torch.randn(B, C, H, W, device=device, generator=g)
    loc, y = torch.empty_like(x), torch.empty_like(x)
    xr, xi, yr, yi = (torch.randn(M, B, C, device=device, generator=g) for _ in range(4))
    mu1, mu2 = (torch.randn(m, m, C, C, 2, device=device, generator=g) / C for _ in range(2))
    lv1, lv2 = (torch.full((m, m, C, C), -4.6, device=device) for _ in range(2))
    wm, wl = torch.randn(C, C, device=device, generator=g) / 16, torch.randn(C, C, device=device, generator=g) / 64
    bm, bl = torch.zeros(C, device=device), torch.full((C,), -3.0, device=device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)      # > 126 MB L2
    st, eng = _lib.stream_ptr(), _lib.ENGINES[engine]
    calls = {
        "local": lambda it: lib.pno_local_fwd(eng, x.data_ptr(), B, C, C, H * W, wm.data_ptr(), bm.data_ptr(), wl.data_ptr(),
                                              bl.data_ptr(), 1234, it, 2, 0, loc.data_ptr(), 0, st),
        "dft_fwd": lambda it: lib.pno_dft_fwd(plan, eng, x.data_ptr(), B, C, 0, xr.data_ptr(), xi.data_ptr(), st),
        "mix": lambda it: lib.pno_mix_fwd(plan, eng, xr.data_ptr(), xi.data_ptr(), B, C, C, mu1.data_ptr(), mu2.data_ptr(),
                                          lv1.data_ptr(), lv2.data_ptr(), 1234, it, 1, yr.data_ptr(), yi.data_ptr(), st),
        "dft_inv": lambda it: lib.pno_dft_inv(plan, eng, yr.data_ptr(), yi.data_ptr(), B, C, 0, loc.data_ptr(), _lib.ACT_GELU,
                                              y.data_ptr(), 0, st),
    }
    act = 4.0 * B * C * H * W            # one fp32 activation tensor
    spec = 4.0 * M * B * C * 2           # one truncated spectrum (re + im)
    nbytes = {"local": 2 * act + 8.0 * C * C, "dft_fwd": act + spec, "mix": 2 * spec + 12.0 * C * C * M,
              "dft_inv": spec + 2 * act}
    out = {}
    for name, fn in calls.items():
        ts = []
        for it in range(iters + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(fn(it), name)
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(e0.elapsed_time(e1) * 1e3)
        us = sum(ts) / len(ts)
        out[name] = {"us": round(us, 1), "min_us": round(min(ts), 1), "bytes": nbytes[name], "gbs": round(nbytes[name] / us / 1e3, 1)}
    return out


if __name__ == "__main__":
    for engine in (sys.argv[1:] or ["tf32", "tf32x3"]):
        r = block_kernel_times(engine)
        tot = sum(v["us"] for v in r.values())
        print(engine, json.dumps(r), f"block total {tot:.0f} us")
