#!/usr/bin/env python
"""Event timeline of CTA 0 of the tensor-core forward-transform kernel (diagnostics; needs a PNO_TRACE build:
`make -C pno_physics_bench_b200/csrc clean && make -C pno_physics_bench_b200/csrc EXTRA=-DPNO_TRACE`)."""
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
x = torch.randn(B, C, H, W, device="cuda")
xr, xi = torch.empty(M, B, C, device="cuda"), torch.empty(M, B, C, device="cuda")
dbg = torch.zeros(4096 + 5 * 512, dtype=torch.int64, device="cuda")
for rep in range(2):
    dbg.zero_()
    lib.pno_debug_set_buffer(dbg.data_ptr())
    _lib.check(lib.pno_dft_fwd(plan, _lib.ENGINES[engine], x.data_ptr(), B, C, 0, xr.data_ptr(), xi.data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    lib.pno_debug_set_buffer(0)
names = {0: {1: "I S1 issued", 2: "I b2_full", 3: "I S2 issued", 4: "I d2_empty", 5: "I x landed"},
         1: {1: "T d1_full", 2: "T b2_empty", 3: "T tmem loaded", 4: "T arrived"},
         2: {1: "E d2_full", 2: "E done"}, 3: {1: "P stage free"}}
ev = []
d = dbg.cpu()
for role in range(5):
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
