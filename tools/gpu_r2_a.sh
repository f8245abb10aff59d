#!/bin/bash
# round 2, GPU call A: parity of the sliced-table kernel (all tests), then the N = 5 / 4 / 3 tuning variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
tail -5 gpurun_out/r2a_pytest.log
: > gpurun_out/r2a_variants.jsonl
run() { # tag degree
  export B200MF_LIBRARY=$PWD/gpurun_variants/$1/libb200mf.so
  echo "== $1" >> gpurun_out/r2a_variants.err
  timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "test_nh_vmult_fp64 and $2" >> gpurun_out/r2a_variants.err 2>&1
  timeout 300 python bench.py --degree $2 --steps 10 --warmup 3 --no-cpu --no-e2e 2>> gpurun_out/r2a_variants.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']
    print(json.dumps({'tag': '$1', 'degree': $2, 'dofs_per_s': d['value'], 'ms': d['ms_per_step'], 'kernel_ms': r['kernel_ms'], 'frac': r['frac'], 'frac_step': r['frac_whole_step']}))
" >> gpurun_out/r2a_variants.jsonl
  unset B200MF_LIBRARY
}
for t in a5 b5 c5 d5 f5 g5; do run $t 4; done
run h4 3
run i3 2
unset B200MF_LIBRARY
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --sweep > gpurun_out/r2a_sweep.json 2>> gpurun_out/r2a_variants.err
cat gpurun_out/r2a_variants.jsonl
python -c "
import json
d=json.load(open('gpurun_out/r2a_sweep.json'))
for k,v in d['sweep'].items(): print(k, '%.3e' % v['dofs_per_s'], '%.3f' % v['frac'])
"
