#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
Usage: python tools/launch_summary.py launches.csv"""
import csv, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = OrderedDict()
for r in rows:
    name = r[4].split("(")[0][:70]; t = float(r[-1].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tot/1e6:.3f} ms total (gpu__time_duration.sum, serialised and cold-cache under ncu: shares, not absolutes)")
for k, (n, t) in agg.items():
    print(f"{k:72s} n={n:3d} total={t/1e6:9.3f} ms  avg={t/n/1e6:8.4f} ms  {100*t/tot:5.1f}%")
