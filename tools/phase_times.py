#!/usr/bin/env python
"""Barrier-wait cycle counters of the tensor-core inverse-transform kernel at the cfg-3 layer shape (diagnostics)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pno_physics_bench_b200 import _lib  # noqa: E402

lib = _lib.load()
engine = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
B, C, H, W, m = 64, 256, 64, 64, 20
M = 2 * m * m
plan = _lib.plan(H, W, m, m, torch.device("cuda", 0))
yr, yi = torch.randn(M, B, C, device="cuda"), torch.randn(M, B, C, device="cuda")
add = torch.randn(B, C, H, W, device="cuda")
y = torch.empty_like(add)
dbg = torch.zeros(148 * 21, dtype=torch.int64, device="cuda")
for use_add, act in ((False, 0), (True, 1)):
    dbg.zero_()
    lib.pno_debug_set_buffer(dbg.data_ptr())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    _lib.check(lib.pno_dft_inv(plan, _lib.ENGINES[engine], yr.data_ptr(), yi.data_ptr(), B, C, 0, add.data_ptr() if use_add else 0,
                               act, y.data_ptr(), 0, _lib.stream_ptr()))
    ev1.record()
    torch.cuda.synchronize()
    lib.pno_debug_set_buffer(0)
    d = dbg[: 148 * 21].view(148, 3, 7).double().mean(0) / 1e3
    print(f"addend={use_add} act={act}: {ev0.elapsed_time(ev1) * 1e3:.0f} us; kcycles per CTA (mean)")
    print("  mma issuer : wait b1_full %.0f  d1_empty %.0f  zt_full %.0f  d2_empty %.0f  total %.0f | S1 exec %.0f  S2 exec %.0f" % tuple(d[0]))
    print("  transposer : wait zt_empty %.0f  d1_full %.0f  total %.0f" % (d[1][0], d[1][1], d[1][4]))
    print("  epilogue   : wait d2_full %.0f  total %.0f | tmem+addend %.0f  act %.0f  store %.0f" % (d[2][2], d[2][4], d[2][0], d[2][1], d[2][3]))
