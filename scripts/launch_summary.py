#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count / total / first launches."""
import csv
import sys
from collections import OrderedDict

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = [(r["Kernel Name"].split("(")[0], float(r["Metric Value"].replace(",", "")), r["Grid Size"]) for r in csv.DictReader(lines)]
tot = OrderedDict()
for k, v, g in rows:
    c = tot.setdefault(k, [0, 0.0])
    c[0] += 1
    c[1] += v
for k, (n, v) in tot.items():
    print(f"{k[:60]:60s} n={n:4d} total={v / 1e6:9.3f} ms")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
for k, v, g in rows[:n]:
    print(f"  {k[:50]:50s} {v / 1e3:10.1f} us  grid {g}")
