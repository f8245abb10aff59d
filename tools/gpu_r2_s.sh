#!/bin/bash
# round 2, GPU call S: A/B of the inactive-lane guard in the sliced kernels (Q4-Q6), FP32 level-operator roofline (Q2-Q4) + ncu of FP32 Q3
mkdir -p gpurun_out
: > gpurun_out/r2s_variants.jsonl
line() {  # tag degree dtype
  timeout 300 python bench.py --degree $2 --dtype $3 --steps 10 --warmup 3 --no-cpu --no-e2e --no-extras 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if not l.startswith('{'): continue
    d = json.loads(l); r = d['roofline']
    print(json.dumps({'tag': '$1', 'degree': $2, 'dtype': '$3', 'dofs_per_s': d['value'], 'ms_per_step': d['ms_per_step'], 'kernel_ms': r['kernel_ms'], 'frac': r['frac'], 'frac_step': r['frac_whole_step'], 'kernel': r.get('kernel')}))
" >> gpurun_out/r2s_variants.jsonl
}
for rep in 1 2; do
  for d in 4 5 6; do
    line guard $d f64
    B200MF_LIBRARY=$PWD/gpurun_variants/g0/libb200mf.so line noguard $d f64
  done
done
for d in 2 3 4; do line main $d f32; done
cat gpurun_out/r2s_variants.jsonl
p=3
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:NHOp -s 3 -c 1 -o gpurun_out/r2s_f32_q$p -f python bench.py --degree $p --dtype f32 --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/r2s_ncu_f32_q$p.log 2>&1
ncu -i gpurun_out/r2s_f32_q$p.ncu-rep --page raw --csv > gpurun_out/r2s_ncu_f32_q${p}_raw.csv 2>/dev/null
ncu -i gpurun_out/r2s_f32_q$p.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2s_ncu_f32_q${p}_source.csv.gz
rm -f gpurun_out/r2s_f32_q$p.ncu-rep
ls -la gpurun_out | tail -5
