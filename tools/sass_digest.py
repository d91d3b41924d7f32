#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-native SASS instructions in libpno_b200.so (the proof that the tensor-core / TMA paths are
what is compiled in): tcgen05.mma = UTC*MMA, tcgen05.ld = LDTM, TMA = UTMALDG / UTMASTG / UTMAPF, tcgen05.commit = UTCBAR.

    python tools/sass_digest.py [profiles/r02_sass_digest.txt]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "pno_physics_bench_b200", "libpno_b200.so")
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_digest.txt")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = collections.OrderedDict([("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
                                ("UTMAPF", r"\bUTMAPF"), ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("HMMA", r"\bHMMA")])
cnt, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
    elif cur:
        for k, p in pats.items():
            if re.search(p, line):
                cnt[cur][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
rows, tot = [], collections.Counter()
for (f, c), d in zip(cnt.items(), names):
    tot.update(c)
    if c.get("UTC*MMA") or c.get("UTMALDG") or c.get("LDTM"):
        d = d.replace("(anonymous namespace)::", "").replace("void ", "")
        rows.append((re.sub(r"\(.*", "", d), c))
with open(out, "w") as o:
    o.write("SASS digest of pno_physics_bench_b200/libpno_b200.so (`cuobjdump -sass`, sm_100a; tools/sass_digest.py), one line per kernel\n"
            "instantiation: tcgen05.mma (UTC*MMA), tcgen05.ld (LDTM), TMA loads / stores / L2 prefetches (UTMALDG / UTMASTG / UTMAPF),\n"
            "tcgen05.commit (UTCBAR), mbarrier operations (SYNCS); the legacy mma.sync path (HMMA) must be 0.\n")
    o.write("whole library: " + ", ".join(f"{k} {tot.get(k, 0)}" for k in pats) +
            f"; {len(rows)} of {len(cnt)} kernels carry tensor-core / TMA instructions\n\n")
    o.write(f"{'kernel':84s} " + " ".join(f"{k:>7s}" for k in pats) + "\n")
    for d, c in sorted(rows):
        o.write(f"{d[:84]:84s} " + " ".join(f"{c.get(k, 0):7d}" for k in pats) + "\n")
print(open(out).read())
