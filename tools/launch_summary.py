#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, average, share."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot = collections.OrderedDict()
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] == "ns" else v * 1e3 if row["Metric Unit"] == "ms" else v
    k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")[:48]
    tot.setdefault(k, [0, 0.0])
    tot[k][0] += 1
    tot[k][1] += v
s = sum(v for _, v in tot.values())
for k, (c, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:48s} n={c:3d} total={v:9.1f} us avg={v / c:8.1f} us share={v / s * 100:5.1f}%")
print(f"total {s:.1f} us over {sum(c for c, _ in tot.values())} launches")
