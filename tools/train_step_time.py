#!/usr/bin/env python
"""BASELINE config 2: one PNOTrainer step (sampled forward, ELBO loss, backward, clip 1.0, AdamW) at 64x64, hidden 256,
4 layers, modes 20, batch 32 -- CUDA-event time per step and the per-kernel launch mix.

    python tools/train_step_time.py [engine] [steps]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pno_physics_bench_b200 as pno  # noqa: E402
from pno_physics_bench_b200.training import ELBOLoss, PNOTrainer  # noqa: E402

engine = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
model = pno.ProbabilisticNeuralOperator(3, 256, 4, 20, engine=engine)
trainer = PNOTrainer(model, ELBOLoss(kl_weight=1e-4), device="cuda", seed=1)
x = torch.randn(32, 3, 64, 64, device="cuda")
y = torch.randn(32, 1, 64, 64, device="cuda")
for _ in range(2):
    losses = trainer.train_step(x, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(steps):
    losses = trainer.train_step(x, y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"engine {engine}: {ms:.1f} ms / step ({32 / ms * 1e3:.0f} fields/s), host {1e3 * (time.perf_counter() - t0) / steps:.1f} ms, "
      f"loss {float(losses['total']):.4e}, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
