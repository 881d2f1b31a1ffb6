#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.  usage: tools/launch_summary.py file.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
ci = {h: i for i, h in enumerate(rows[hdr])}
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) < len(ci) or not r[ci["Metric Value"]].replace(",", "").replace(".", "").isdigit():
        continue
    a = agg.setdefault(r[ci["Kernel Name"]][:90], [0, 0.0])
    a[0] += 1
    a[1] += float(r[ci["Metric Value"]].replace(",", ""))
unit = next((r[ci["Metric Unit"]] for r in rows[hdr + 1:] if len(r) >= len(ci)), "?")
for k, (n, t) in agg.items():
    print(f"{n:4d} launches {t:14.1f} {unit} total {t / n:12.1f} avg  {k}")
