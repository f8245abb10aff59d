#!/bin/bash
# round 2: launch list of one Newton step with the final code (graph + smoother hook), aggregated by kernel
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_newton_launches.csv python tools/newton_step_profile.py > gpurun_out/r2_newton_profile.log 2>&1
tail -1 gpurun_out/r2_newton_profile.log | cut -c1-400
python - <<'PY'
import csv, collections, json, re
rows = [r for r in csv.reader(open("gpurun_out/r2_newton_launches.csv", errors="ignore")) if len(r) > 10 and r[0].isdigit()]
try:
    n = json.loads([l for l in open("gpurun_out/r2_newton_profile.log") if l.startswith("{")][-1])["gpu_launches"]
except Exception:
    n = len(rows)
rows = rows[-n:]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for r in rows:
    name = re.sub(r"b200mf::", "", r[4]); name = re.sub(r"\(.*", "", name).replace("void ", "")
    ns = float(r[-1].replace(",", "")); k = (name[:80], r[8]); agg[k][0] += 1; agg[k][1] += ns; tot += ns
out = [f"{len(rows)} kernels of the measured Newton step (last {n} of the process), {tot/1e6:.3f} ms of kernel time (ncu per-launch durations: cold caches, serialised)"]
for (name, grid), (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    out.append(f"{ns/1e6:9.3f} ms {100*ns/tot:5.1f}%  x{c:<6d} {ns/c/1e3:9.1f} us  grid {grid:>14s}  {name}")
open("gpurun_out/r2_newton_launch_breakdown.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out))
PY
rm -f gpurun_out/r2_newton_launches.csv
