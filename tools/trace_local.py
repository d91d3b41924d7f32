#!/usr/bin/env python
"""Event timeline of CTA 0 of the tensor-core local-branch kernel (diagnostics; needs a PNO_TRACE build)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pno_physics_bench_b200 import _lib  # noqa: E402

lib = _lib.load()
engine = sys.argv[1] if len(sys.argv) > 1 else "tf32"
B, C, HW = 64, 256, 4096
x = torch.randn(B, C, HW, device="cuda")
out = torch.empty_like(x)
wm, wl = torch.randn(C, C, device="cuda") / 16, torch.randn(C, C, device="cuda") / 64
bm, bl = torch.zeros(C, device="cuda"), torch.full((C,), -3.0, device="cuda")
dbg = torch.zeros(4096 + 5 * 512, dtype=torch.int64, device="cuda")
for rep in range(2):
    dbg.zero_()
    lib.pno_debug_set_buffer(dbg.data_ptr())
    _lib.check(lib.pno_local_fwd(_lib.ENGINES[engine], x.data_ptr(), B, C, C, HW, wm.data_ptr(), bm.data_ptr(), wl.data_ptr(),
                                 bl.data_ptr(), 1234, rep, 2, 0, out.data_ptr(), 0, _lib.stream_ptr()))
    torch.cuda.synchronize()
    lib.pno_debug_set_buffer(0)
names = {0: {1: "I tempty", 2: "I stage landed", 3: "I tile issued"}, 2: {1: "E tfull", 2: "E done"}, 3: {1: "P stage free"}}
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
