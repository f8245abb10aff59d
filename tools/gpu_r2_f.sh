#!/bin/bash
# round 2, GPU call F: variants with per-slot "empty" mbarriers instead of a block barrier per plane + ncu of the default
mkdir -p gpurun_out
: > gpurun_out/r2f_variants.jsonl
: > gpurun_out/r2f_variants.err
run() { # tag degree
  export B200MF_LIBRARY=$PWD/gpurun_variants/$1/libb200mf.so
  echo "== $1" >> gpurun_out/r2f_variants.err
  timeout 300 python -m pytest "tests/test_gpu_parity.py::test_nh_vmult_fp64[$2]" -q -x 2>&1 | tail -1 >> gpurun_out/r2f_variants.err
  timeout 300 python bench.py --degree $2 --steps 10 --warmup 3 --no-cpu --no-e2e --no-extras 2>> gpurun_out/r2f_variants.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']
    print(json.dumps({'tag': '$1', 'degree': $2, 'dofs_per_s': d['value'], 'ms': d['ms_per_step'], 'kernel_ms': r['kernel_ms'], 'frac': r['frac'], 'frac_step': r['frac_whole_step']}))
" >> gpurun_out/r2f_variants.jsonl
}
for t in a5 b5 c5 g5; do run $t 4; done
cat gpurun_out/r2f_variants.jsonl
grep -v "^==" gpurun_out/r2f_variants.err | head
export B200MF_LIBRARY=$PWD/gpurun_variants/a5/libb200mf.so
p=4
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:NHOp -s 3 -c 1 -o gpurun_out/r2f_q$p -f python bench.py --degree $p --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/r2f_ncu_q$p.log 2>&1
ncu -i gpurun_out/r2f_q$p.ncu-rep --page raw --csv > gpurun_out/r2f_ncu_q${p}_raw.csv 2>/dev/null
ncu -i gpurun_out/r2f_q$p.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2f_ncu_q${p}_source.csv.gz
rm -f gpurun_out/r2f_q$p.ncu-rep
