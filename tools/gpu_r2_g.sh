#!/bin/bash
# round 2, GPU call C: full GPU test suite (new: faces, Newton steps) + N = 5 variants with the asynchronous gather + sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1
tail -15 gpurun_out/r2g_pytest.log
: > gpurun_out/r2g_variants.jsonl
: > gpurun_out/r2g_variants.err
run() { # tag degree
  export B200MF_LIBRARY=$PWD/gpurun_variants/$1/libb200mf.so
  echo "== $1" >> gpurun_out/r2g_variants.err
  timeout 300 python -m pytest "tests/test_gpu_parity.py::test_nh_vmult_fp64[$2]" -q -x 2>&1 | tail -1 >> gpurun_out/r2g_variants.err
  timeout 300 python bench.py --degree $2 --steps 10 --warmup 3 --no-cpu --no-e2e --no-extras 2>> gpurun_out/r2g_variants.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']
    print(json.dumps({'tag': '$1', 'degree': $2, 'dofs_per_s': d['value'], 'ms': d['ms_per_step'], 'kernel_ms': r['kernel_ms'], 'frac': r['frac'], 'frac_step': r['frac_whole_step']}))
" >> gpurun_out/r2g_variants.jsonl
  unset B200MF_LIBRARY
}
for t in a5 c5 g5; do run $t 4; done
cat gpurun_out/r2g_variants.jsonl
cat gpurun_out/r2g_variants.err | grep -v "^==" | head -20
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --sweep > gpurun_out/r2g_sweep.json 2>> gpurun_out/r2g_variants.err
python -c "
import json
d=json.load(open('gpurun_out/r2g_sweep.json'))
for k,v in d['sweep'].items(): print(k, '%.3e' % v['dofs_per_s'], 'kernel %.3f step %.3f' % (v['frac_kernel'], v['frac_step']))
print(d.get('newton_step'))
"
