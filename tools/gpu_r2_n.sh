#!/bin/bash
# round 2, GPU call N: full GPU suite (hygiene, heat, coarse CG, tube golden) + default bench + racecheck on small vmults
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1
tail -8 gpurun_out/r2n_pytest.log
( time python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err ) 2> gpurun_out/r2n_bench.time
tail -3 gpurun_out/r2n_bench.err; grep real gpurun_out/r2n_bench.time
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2n_bench.json'))
print('value %.3e frac %.3f e2e %.3e cpu %.3e' % (d['value'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value']))
print(d['cpu_baseline']['sample'])
for k,v in d.get('sweep',{}).items(): print(k, '%.3e' % v['dofs_per_s'], 'kernel %.3f step %.3f' % (v['frac_kernel'], v['frac_step']))
print({k:v for k,v in d['newton_step'].items() if k.endswith('_s') or k in ('cg_iterations','gpu_launches')})
print({k:v for k,v in d['heat_step'].items() if k.endswith('_s') or k in ('cg_iterations',)}, d['heat_step']['system_vmult']['roofline']['frac'])
print(d.get('tube3d_step'))
PY
for p in 2 3 4 5; do
  timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest "tests/test_gpu_parity.py::test_nh_vmult_fp64[$p]" -q -x 2>&1 | grep -E "RACECHECK SUMMARY|passed|failed|Race reported|ERROR SUMMARY" | head -8
done > gpurun_out/r2n_racecheck.txt 2>&1
cat gpurun_out/r2n_racecheck.txt
