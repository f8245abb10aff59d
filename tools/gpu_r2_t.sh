#!/bin/bash
# round 2, GPU call T: multigrid lower levels as a CUDA graph - solver/Newton/heat tests + the solver blocks of bench.py with and without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_newton.py tests/test_gpu_heat.py -x -q 2>&1 | tail -15
timeout 600 python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 > gpurun_out/r2t_bench_graph.json 2> gpurun_out/r2t_bench_graph.err
B200MF_MG_GRAPH=0 timeout 600 python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 > gpurun_out/r2t_bench_nograph.json 2> gpurun_out/r2t_bench_nograph.err
python - <<'PY'
import json
for tag in ("graph", "nograph"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2t_bench_{tag}.json") if l.startswith("{")][-1])
    except Exception as e:
        print(tag, "no json", e); continue
    for k in ("newton_step", "heat_step", "tube3d_step"):
        b = d.get(k) or {}
        print(tag, k, {x: b.get(x) for x in ("newton_step_s", "update_matrix_free_s", "cg_solve_s", "time_step_s", "cg_iterations", "newton_iterations", "gpu_launches", "seconds_per_newton_iteration", "mg_graph")})
PY
tail -5 gpurun_out/r2t_bench_graph.err
