#!/usr/bin/env python
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` log into a per-kernel launch list.

usage: python profiles/launch_list.py gpurun_out/rNN_launches.csv "<command that was profiled>" > profiles/rNN_launches.txt
"""
import csv, sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = OrderedDict()
for r in rows:
    k = r[4]
    ns = float(r[14].replace(",", ""))
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ns
tot = sum(v[1] for v in agg.values()) or 1.0
print("ncu --metrics gpu__time_duration.sum --clock-control none: %s" % (sys.argv[2] if len(sys.argv) > 2 else ""))
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%6d launches %12.1f us total %10.1f us avg %5.1f%%  %s" % (n, ns / 1e3, ns / 1e3 / n, 100 * ns / tot, k[:120]))
