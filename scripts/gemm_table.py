#!/usr/bin/env python
"""Aggregate a XLX_GEMM_LOG csv (bench.py roofline pass: 2 steps) into a per-shape table for the last step."""
import collections
import csv
import sys

rows = list(csv.DictReader(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/gemm_shapes.csv")))
rows = rows[len(rows) // 2:]
agg = collections.OrderedDict()
for r in rows:
    k = (int(r['M']), int(r['N']), int(r['K']), int(r['a_mn']), int(r['b_mn']), int(r['epi']), int(r['grid']))
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r['us'])
tot = sum(v[1] for v in agg.values())
print(f"GEMM time per step: {tot / 1000:.2f} ms in {sum(v[0] for v in agg.values())} launches")
print(f"{'M':>6} {'N':>5} {'K':>6} amn bmn   epi grid  cnt   avg_us   tot_ms  TF/s(alg)  share")
for k, (c, us) in sorted(agg.items(), key=lambda x: -x[1][1]):
    M, N, K, a, b, e, g = k
    tf = 2 * M * N * K * c / us / 1e6
    print(f"{M:6d} {N:5d} {K:6d}  {a}   {b}  {e:5d} {g:4d} {c:4d} {us / c:8.1f} {us / 1000:8.2f} {tf:9.1f}  {us / tot * 100:5.1f}%")
