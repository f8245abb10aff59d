#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (+ grid size):
   python tools/launch_breakdown.py gpurun_out/launches_newton.csv [top]"""
import csv, collections, re, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 10 and r[0].isdigit()]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in rows:
    name = re.sub(r"b200mf::", "", r[4]); name = re.sub(r"\(.*", "", name).replace("void ", "")
    ns = float(r[-1].replace(",", ""))
    k = (name[:90], r[8])
    agg[k][0] += 1; agg[k][1] += ns; tot += ns
print(f"{len(rows)} launches, {tot/1e6:.3f} ms of kernel time")
for (name, grid), (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{ns/1e6:9.3f} ms {100*ns/tot:5.1f}%  x{n:<6d} {ns/n/1e3:9.1f} us  grid {grid:>14s}  {name}")
