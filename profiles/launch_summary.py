#!/usr/bin/env python3
"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv):  python profiles/launch_summary.py x.csv"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    agg.setdefault(row["Kernel Name"], []).append((float(row["Metric Value"].replace(",", "")), row.get("Grid Size"), row.get("Block Size")))
tot = sum(x[0] for v in agg.values() for x in v)
print("%d kernels, %d launches, %.1f us serialised" % (len(agg), sum(len(v) for v in agg.values()), tot / 1e3))
for k, v in sorted(agg.items(), key=lambda kv: -sum(x[0] for x in kv[1])):
    s = sum(x[0] for x in v)
    print("%-44s n=%2d  %8.1f us  %5.1f %%  grid %s block %s" % (k.split("(")[0][:44], len(v), s / 1e3, 100 * s / tot, v[0][1], v[0][2]))
