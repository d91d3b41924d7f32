import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pno_physics_bench_b200 as pno
torch.manual_seed(0)
model = pno.ProbabilisticNeuralOperator(3, 256, 4, 48, engine="tf32").cuda().train()
x = torch.randn(2, 3, 256, 256, device="cuda")
with torch.no_grad():
    for _ in range(2):
        y = model(x, sample=True)
torch.cuda.synchronize()
print("ok")
