#!/usr/bin/env python
"""Small driver for ncu: a few MC samples of the cfg-3 model (64x64, hid 256, modes 20, batch 64) through the public API.

    python tools/profile_step.py [engine] [num_samples]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pno_physics_bench_b200 as pno  # noqa: E402

engine = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
samples = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
model = pno.ProbabilisticNeuralOperator(3, 256, 4, 20, engine=engine).cuda()
x = torch.randn(64, 3, 64, 64, device="cuda")
mean, std = model.predict_with_uncertainty(x, num_samples=samples, seed=1)
torch.cuda.synchronize()
print("ok", float(mean.abs().mean()), float(std.mean()))
