#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
   python tools/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.txt"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
n = 0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(row["Metric Value"].replace(",", ""))
    n += 1
tot = sum(v[1] for v in agg.values())
print(f"{n} launches, total {tot / 1e3:.1f} us (ncu per-launch times are cold-cache and serialised: compare SHARES)")
print(f"{'kernel':70s} {'count':>6s} {'us total':>12s} {'share':>7s} {'us/launch':>10s}")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {c:6d} {t / 1e3:12.1f} {100 * t / tot:6.1f}% {t / 1e3 / c:10.1f}")
