#!/bin/bash
# round 2, final 1-GPU call: full GPU suite, smoke, default bench line + reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2f_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | grep -v "^Exception\|^Traceback\|^  File\|^TypeError" | tail -2
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r2f_bench_reference.json 2>> gpurun_out/r2f_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2f_bench.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["roofline"]["frac"], d["e2e"]["value"], d["cpu_baseline"]["value"])
print({q: (round(v["frac_kernel"], 3), round(v["frac_step"], 3)) for q, v in d["sweep"].items()})
for k in ("newton_step", "heat_step", "tube3d_step"):
    b = d[k]; print(k, {x: b.get(x) for x in ("newton_step_s", "time_step_s", "cg_iterations", "gpu_launches")})
r = json.loads([l for l in open("gpurun_out/r2f_bench_reference.json") if l.startswith("{")][-1]); print("reference arm", r["value"], r["cpu_baseline"]["cores"])
PY
tail -3 gpurun_out/r2f_bench.err
