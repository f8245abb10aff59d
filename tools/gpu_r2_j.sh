#!/bin/bash
# round 2, GPU call J: full GPU suite + default bench line (pipelined host vmult)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1
tail -8 gpurun_out/r2j_pytest.log
( time python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err ) 2> gpurun_out/r2j_bench.time
tail -3 gpurun_out/r2j_bench.err; grep real gpurun_out/r2j_bench.time
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench.json'))
print('value %.3e frac %.3f e2e %.3e (%.2f ms)' % (d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step']))
for k,v in d.get('sweep',{}).items(): print(k, '%.3e' % v['dofs_per_s'], 'kernel %.3f step %.3f' % (v['frac_kernel'], v['frac_step']))
print(d.get('newton_step',{}).get('newton_step_s'), d.get('parity',{}).get('rel_l2'))
PY
B200MF_HOST_PIPELINE=0 python bench.py --steps 5 --no-cpu --no-extras | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('serial host path: e2e %.3e (%.2f ms)' % (d['e2e']['value'], d['e2e']['ms_per_step']))"
