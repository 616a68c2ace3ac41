#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python profiles/summarize_launches.py gpurun_out/launches_X.csv"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = row["Kernel Name"].split("(")[0]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"# {sys.argv[1]}: {sum(a[0] for a in agg.values())} launches, {tot:.1f} us total (cold-cache, serialised)")
print(f"{'kernel':45s} {'n':>5s} {'total_us':>10s} {'avg_us':>8s} {'share':>6s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:45s} {n:5d} {t:10.1f} {t / n:8.1f} {100 * t / tot:5.1f}%")
