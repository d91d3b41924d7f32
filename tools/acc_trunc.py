#!/usr/bin/env python
"""How does the error of a tcgen05 kind::tf32 product grow with the number of MMAs chained into one fp32 accumulator?
Operands are exactly representable in tf32 (no operand rounding), so what is measured is the accumulation alone:
D[128][N] = A[128][K] B[K][N] through pno_selftest_umma (K/8 MMAs into one tensor-memory accumulator) against fp64."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pno_physics_bench_b200 import _lib  # noqa: E402

lib = _lib.load()


def tf32_trunc(t):
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


for k in (8, 16, 32, 64, 128, 256):
    gen = torch.Generator().manual_seed(k)
    n = 64
    a = tf32_trunc(torch.randn(128, k, generator=gen)).cuda()
    b = tf32_trunc(torch.randn(k, n, generator=gen)).cuda()
    d = torch.empty(128, n, device="cuda")
    _lib.check(lib.pno_selftest_umma(0, n, k, a.data_ptr(), b.data_ptr(), a.data_ptr(), b.t().contiguous().data_ptr(), d.data_ptr(),
                                     _lib.stream_ptr()), "selftest")
    torch.cuda.synchronize()
    want = a.double() @ b.double()
    err = d.double() - want
    rel = float(err.norm() / want.norm())
    bias = float((err * want.sign()).mean() / want.abs().mean())     # < 0: magnitudes shrink (round toward zero)
    f32 = (a @ b).double() if False else None
    print(f"K {k:4d} ({k // 8:3d} MMAs): rel-L2 {rel:.3e}, mean signed relative error {bias:+.3e}")
