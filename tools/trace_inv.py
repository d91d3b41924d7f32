#!/usr/bin/env python
"""Event timeline of CTA 0 of the tensor-core inverse-transform kernel (diagnostics; cfg-3 layer shape)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pno_physics_bench_b200 import _lib  # noqa: E402

lib = _lib.load()
engine = sys.argv[1] if len(sys.argv) > 1 else "tf32"
B, C, H, W, m = 64, 256, 64, 64, 20
M = 2 * m * m
plan = _lib.plan(H, W, m, m, torch.device("cuda", 0))
yr, yi = torch.randn(M, B, C, device="cuda"), torch.randn(M, B, C, device="cuda")
add = torch.randn(B, C, H, W, device="cuda")
y = torch.empty_like(add)
dbg = torch.zeros(4096 + 6 * 512, dtype=torch.int64, device="cuda")
for rep in range(2):
    dbg.zero_()
    lib.pno_debug_set_buffer(dbg.data_ptr())
    _lib.check(lib.pno_dft_inv(plan, _lib.ENGINES[engine], yr.data_ptr(), yi.data_ptr(), B, C, 0, add.data_ptr(), 1, y.data_ptr(), 0,
                               _lib.stream_ptr()))
    torch.cuda.synchronize()
    lib.pno_debug_set_buffer(0)
names = {0: {1: "I b1_full", 2: "I d1_empty", 3: "I S1 issued", 4: "I zt_full", 5: "I d2_empty", 6: "I S2 issued"},
         1: {1: "T d1_full", 2: "T zt_empty", 3: "T tile done", 4: "T arrived", 5: "T tmem loaded"},
         2: {1: "E d2_full", 2: "E done", 3: "E add_full"}, 5: {1: "P tile complete", 2: "P store read out", 3: "P load issued"}, 3: {1: "L b1_empty", 2: "L loaded", 3: "L arrived"},
         4: {1: "E' d2_full", 2: "E' done"}}
ev = []
d = dbg.cpu()
for role in range(6):
    for n in range(512):
        v = int(d[4096 + role * 512 + n])
        if v:
            ev.append((v >> 8, names[role][v & 255]))
ev.sort()
limit = int(sys.argv[2]) if len(sys.argv) > 2 else 150
t_prev = 0
for t, name in ev[:limit]:
    print(f"{t:9d} (+{t - t_prev:6d})  {name}")
    t_prev = t
