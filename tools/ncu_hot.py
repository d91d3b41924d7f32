#!/usr/bin/env python
"""Top stall-sample SASS instructions of one kernel in an .ncu-rep (needs ncu on PATH).
    python tools/ncu_hot.py gpurun_out/prof.ncu-rep regex:k_tc_mix [topN]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(lines[start:]))
hdr = rows[0]
ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = []
for k, r in enumerate(rows[1:]):
    if len(r) <= isamp:
        continue
    try:
        data.append((int(r[isamp]), k, r[isrc].strip(), int(r[iexec])))
    except ValueError:
        pass
tot = sum(d[0] for d in data)
print(f"total samples {tot}, instructions {len(data)}")
for s, k, src, ex in sorted(data, reverse=True)[:top]:
    print(f"{s:8d} {100.0 * s / tot:5.1f}%  #{k:5d} exec={ex:9d}  {src[:110]}")
