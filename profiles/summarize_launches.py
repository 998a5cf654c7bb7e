#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > vi:
        agg[r[ki].split("(")[0][-60:]].append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':62s} {'n':>5s} {'mean_us':>10s} {'total_us':>11s} {'share':>6s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:62s} {len(v):5d} {sum(v) / len(v) / 1e3:10.2f} {sum(v) / 1e3:11.1f} {sum(v) / tot:6.1%}")
