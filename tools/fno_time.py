#!/usr/bin/env python
"""BASELINE config 4: deterministic FNO baseline path (sample=False) 128x128, hidden 128, modes 32, batch 64 inference.

    python tools/fno_time.py [engine]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pno_physics_bench_b200 as pno  # noqa: E402

engine = sys.argv[1] if len(sys.argv) > 1 else "tf32"
torch.manual_seed(0)
model = pno.FourierNeuralOperator(3, 128, 4, 32, engine=engine).cuda().eval()
x = torch.randn(64, 3, 128, 128, device="cuda")
with torch.no_grad():
    for _ in range(3):
        y = model(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        y = model(x)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"engine {engine}: {ms:.2f} ms per batch of 64 -> {64 / ms * 1e3:.0f} fields/s (reference CPU: 8.6 fields/s)")
