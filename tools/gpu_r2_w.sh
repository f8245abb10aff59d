#!/bin/bash
# round 2, GPU call W: where the hooked smoother loses its time - one smoother application, cell kernels vs the rest
mkdir -p gpurun_out
for h in 1 0; do B200MF_SMOOTHER_HOOK=$h timeout 300 python tools/smoother_hook_ab.py 64 2>&1 | tail -1; done | tee gpurun_out/r2w_hook_ab.jsonl
B200MF_SMOOTHER_HOOK=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2w_launches_hook.csv python tools/smoother_hook_ab.py 48 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(l for l in open("gpurun_out/r2w_launches_hook.csv") if l.startswith('"'))]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    agg[r[ki][:70]].append(float(r[vi].replace(",", "")))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{sum(v)/1e3:10.1f} us total  x{len(v):4d}  {sum(v)/len(v)/1e3:9.1f} us each  {k}")
PY
timeout 600 python -m pytest tests/test_gpu_heat.py tests/test_gpu_solver.py -x -q 2>&1 | tail -3
