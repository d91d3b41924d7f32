#!/usr/bin/env python
"""Channel-mix kernel at the cfg-3 layer shape: sampled (Philox in the generators) vs deterministic weights (mu only), to
separate the cost of the noise generation from the cost of the operand pipeline.   python tools/mix_ablation.py [engine]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pno_physics_bench_b200 as pno  # noqa: E402
from pno_physics_bench_b200 import _lib  # noqa: E402

engine = sys.argv[1] if len(sys.argv) > 1 else "tf32"
lib = _lib.load()
B, C, m = 64, 256, 20
layer = pno.SpectralConv2d_Probabilistic(C, C, m, m).cuda()
M = 2 * m * m
xr, xi = torch.randn(M, B, C, device="cuda"), torch.randn(M, B, C, device="cuda")
yr, yi = torch.empty(M, B, C, device="cuda"), torch.empty(M, B, C, device="cuda")
plan = _lib.plan(64, 64, m, m, xr.device)
n = [t.detach().permute(2, 3, 0, 1) for t in (layer.weights1, layer.weights2, layer.weights1_log_var, layer.weights2_log_var)]
for sampled in (True, False):
    def run():
        _lib.check(lib.pno_mix_fwd(plan, _lib.ENGINES[engine], xr.data_ptr(), xi.data_ptr(), B, C, C, n[0].data_ptr(), n[1].data_ptr(),
                                   n[2].data_ptr() if sampled else 0, n[3].data_ptr() if sampled else 0, 1, 2, 3, yr.data_ptr(),
                                   yi.data_ptr(), _lib.stream_ptr()), "mix")
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    byts = 4 * M * B * C * 4 + (12 if sampled else 8) * M * C * C            # 2 spectra in + out, mu (+ log_var)
    print(f"{engine} sampled={sampled}: {us:.1f} us, {byts / us / 1e3:.0f} GB/s")
