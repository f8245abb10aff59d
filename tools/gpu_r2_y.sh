#!/bin/bash
# round 2, GPU call Y: full GPU suite + solver blocks with the smoother hook on / off
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | cut -c1-300
for tag in hook nohook; do
  if [ $tag = nohook ]; then export B200MF_SMOOTHER_HOOK=0; fi
  timeout 600 python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 > gpurun_out/r2y_bench_$tag.json 2> gpurun_out/r2y_bench_$tag.err
done
unset B200MF_SMOOTHER_HOOK
python - <<'PY'
import json
for tag in ("hook", "nohook"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2y_bench_{tag}.json") if l.startswith("{")][-1])
    except Exception as e:
        print(tag, "no json", e); continue
    for k in ("newton_step", "heat_step", "tube3d_step"):
        b = d.get(k) or {}
        print(tag, k, {x: b.get(x) for x in ("newton_step_s", "update_matrix_free_s", "cg_solve_s", "time_step_s", "cg_iterations", "newton_iterations", "gpu_launches", "seconds_per_newton_iteration")}, (b.get("mg_graph") or {}).get("launches_per_replay"))
PY
