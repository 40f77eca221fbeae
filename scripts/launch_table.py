#!/usr/bin/env python
"""Aggregate an ncu launch list (gpu__time_duration.sum per launch) per kernel for the LAST step in the file."""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lines = [l for l in open(path) if not l.startswith("==")]
rows = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((row["Kernel Name"], float(row["Metric Value"].replace(",", "")), row["Metric Unit"]))
n = len(rows) // nsteps
last = rows[-n:]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v, u in last:
    k = re.sub(r"\(.*", "", k).replace("xlx::<unnamed>::", "").replace("void ", "")
    agg[k][0] += 1
    agg[k][1] += v / (1000.0 if u == "ns" else 1.0)
tot = sum(v for _, v in agg.values())
print(f"{n} launches per step, {tot / 1000:.2f} ms of kernel time (serialised, cold-cache: compare shares)")
for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{v / tot * 100:6.2f}%  {c:5d}  {v:10.1f} us  {k[:100]}")
