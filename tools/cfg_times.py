#!/usr/bin/env python
"""BASELINE configs 1 and 5 (shapes outside the tensor-core tilings): time per step on the CUDA-core engine.

cfg1: PNO 64x64, in 3, hid 32, L4, modes 12, batch 8, fwd+bwd with sample=True (MSE + 1e-4 KL)
cfg5: PNO 256x256, hid 256, L4, modes 48, batch 2 per GPU: sampled forward (and fwd+bwd)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pno_physics_bench_b200 as pno  # noqa: E402


def timed(fn, n):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


which = sys.argv[1] if len(sys.argv) > 1 else "1"
engine = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
torch.manual_seed(0)
if which == "1":
    model = pno.ProbabilisticNeuralOperator(3, 32, 4, 12, engine=engine).cuda().train()
    x, y = torch.randn(8, 3, 64, 64, device="cuda"), torch.randn(8, 1, 64, 64, device="cuda")

    def step():
        model.zero_grad(set_to_none=True)
        loss = torch.mean((model(x, sample=True) - y) ** 2) + 1e-4 * model.kl_divergence()
        loss.backward()
    print(f"cfg1 fwd+bwd ({engine}): {timed(step, 10):.2f} ms / step (reference CPU: 131 ms)")
else:
    model = pno.ProbabilisticNeuralOperator(3, 256, 4, 48, engine=engine).cuda().train()
    x, y = torch.randn(2, 3, 256, 256, device="cuda"), torch.randn(2, 1, 256, 256, device="cuda")
    with torch.no_grad():
        print(f"cfg5 sampled forward, batch 2 ({engine}): {timed(lambda: model(x, sample=True), 3):.1f} ms")

    def step():
        model.zero_grad(set_to_none=True)
        loss = torch.mean((model(x, sample=True) - y) ** 2) + 1e-4 * model.kl_divergence()
        loss.backward()
    print(f"cfg5 fwd+bwd, batch 2 ({engine}): {timed(step, 2):.1f} ms, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
