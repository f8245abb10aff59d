#!/bin/bash
# round 2, GPU call Q: plane-spread variants at Q4 + one complete racecheck record of the sliced kernel
mkdir -p gpurun_out
: > gpurun_out/r2q_variants.jsonl
run() {
  export B200MF_LIBRARY=$PWD/gpurun_variants/$1/libb200mf.so
  timeout 300 python -m pytest "tests/test_gpu_parity.py::test_nh_vmult_fp64[$2]" -q -x 2>&1 | tail -1
  timeout 300 python bench.py --degree $2 --steps 10 --warmup 3 --no-cpu --no-e2e --no-extras 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']
    print(json.dumps({'tag': '$1', 'degree': $2, 'dofs_per_s': d['value'], 'kernel_ms': r['kernel_ms'], 'frac': r['frac'], 'frac_step': r['frac_whole_step']}))
" >> gpurun_out/r2q_variants.jsonl
  unset B200MF_LIBRARY
}
for t in w0 w1 w2; do run $t 4; done
cat gpurun_out/r2q_variants.jsonl
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest "tests/test_gpu_parity.py::test_nh_vmult_fp64[4]" -q -x > gpurun_out/r2q_racecheck_q4_full.txt 2>&1
grep -c "Race reported" gpurun_out/r2q_racecheck_q4_full.txt; grep -A12 "Race reported" gpurun_out/r2q_racecheck_q4_full.txt | head -60 | cut -c1-250
gzip -f gpurun_out/r2q_racecheck_q4_full.txt
