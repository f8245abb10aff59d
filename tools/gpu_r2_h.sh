#!/bin/bash
# round 2, GPU call F: variants with per-slot "empty" mbarriers instead of a block barrier per plane + ncu of the default
mkdir -p gpurun_out
: > gpurun_out/r2h_variants.jsonl
: > gpurun_out/r2h_variants.err
run() { # tag degree
  export B200MF_LIBRARY=$PWD/gpurun_variants/$1/libb200mf.so
  echo "== $1" >> gpurun_out/r2h_variants.err
  timeout 300 python -m pytest "tests/test_gpu_parity.py::test_nh_vmult_fp64[$2]" -q -x 2>&1 | tail -1 >> gpurun_out/r2h_variants.err
  timeout 300 python bench.py --degree $2 --steps 10 --warmup 3 --no-cpu --no-e2e --no-extras 2>> gpurun_out/r2h_variants.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']
    print(json.dumps({'tag': '$1', 'degree': $2, 'dofs_per_s': d['value'], 'ms': d['ms_per_step'], 'kernel_ms': r['kernel_ms'], 'frac': r['frac'], 'frac_step': r['frac_whole_step']}))
" >> gpurun_out/r2h_variants.jsonl
}
for t in v1 v2 v3 v4 v5 v6; do run $t 4; done
cat gpurun_out/r2h_variants.jsonl
grep -v "^==" gpurun_out/r2h_variants.err | head
