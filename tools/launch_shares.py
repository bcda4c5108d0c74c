#!/usr/bin/env python
"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: launch_shares.py launches.csv ['header comment']"""
import csv, re, sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = f.read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
for r in csv.DictReader(lines[start:]):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append(r)
agg = OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"])          # drop the argument list
    name = re.sub(r"\((bool|int)\)", "", name)
    key = (name, r["Block Size"], r["Grid Size"])
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += float(r["Metric Value"].replace(",", "")) / 1e3
total = sum(a[1] for a in agg.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
for (name, blk, grid), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-62s %-14s %-14s n=%3d tot=%9.1f us avg=%8.1f us %5.1f%%" % (name[:62], blk, grid, n, us, us / n, 100 * us / total))
