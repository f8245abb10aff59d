#!/bin/bash
# round 2, GPU call K: stacked scalar kernel: parity (Laplace + heat-step tests), A/B against the generic kernel, bench line with heat_step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_heat.py tests/test_gpu_solver.py -x -q -k "laplace or heat or gmg or chebyshev" > gpurun_out/r2k_pytest.log 2>&1
tail -6 gpurun_out/r2k_pytest.log
for stack in 1 0; do
  for deg in 4 3 2; do
    B200MF_SCALAR_STACK=$stack python tools/bench_laplace.py --degree $deg --cells $((256/deg)) --tag stack$stack 2>&1 | tail -1
  done
done > gpurun_out/r2k_laplace_ab.jsonl
cat gpurun_out/r2k_laplace_ab.jsonl
( time python bench.py --no-cpu > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err ) 2> gpurun_out/r2k_bench.time
tail -3 gpurun_out/r2k_bench.err; grep real gpurun_out/r2k_bench.time
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2k_bench.json'))
print(json.dumps(d.get('heat_step'), indent=1))
PY
