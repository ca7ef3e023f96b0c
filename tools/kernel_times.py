#!/usr/bin/env python
"""Sums an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(l for l in open(sys.argv[1], errors="ignore") if l.startswith('"')))
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
t = defaultdict(float); c = defaultdict(int)
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    name = r[ki].split("(")[0]
    t[name] += v; c[name] += 1
tot = sum(t.values())
print("| kernel | launches | ms | share |\n|---|---|---|---|")
for k in sorted(t, key=lambda k: -t[k]):
    print("| %s | %d | %.2f | %.1f%% |" % (k, c[k], t[k], 100 * t[k] / tot))
print("| total | %d | %.2f | |" % (sum(c.values()), tot))
