#!/bin/bash
# round 2, GPU call D: new bench.py default line (sweep / numbering / newton_step / parity blocks) + ncu of Q4 (async gather)
mkdir -p gpurun_out
( time python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err ) 2> gpurun_out/r2d_bench.time
tail -3 gpurun_out/r2d_bench.err; cat gpurun_out/r2d_bench.time
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2d_bench.json'))
print('value %.3e frac %.3f e2e %.3e' % (d['value'], d['roofline']['frac'], d['e2e']['value']))
for k,v in d.get('sweep',{}).items(): print(k, '%.3e' % v['dofs_per_s'], 'kernel %.3f step %.3f' % (v['frac_kernel'], v['frac_step']))
for k,v in d.get('numbering',{}).items():
    if isinstance(v, dict): print(k, '%.3e' % v['dofs_per_s'], 'kernel %.3f' % v['frac_kernel'])
print(d.get('newton_step')); print(d.get('parity')); print(d.get('run_to_run')); print(d.get('cpu_baseline'))
PY
for p in 4; do
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:NHOp -s 3 -c 1 -o gpurun_out/r2d_q$p -f python bench.py --degree $p --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/r2d_ncu_q$p.log 2>&1
ncu -i gpurun_out/r2d_q$p.ncu-rep --page raw --csv > gpurun_out/r2d_ncu_q${p}_raw.csv 2>/dev/null
ncu -i gpurun_out/r2d_q$p.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2d_ncu_q${p}_source.csv.gz
rm -f gpurun_out/r2d_q$p.ncu-rep
done
