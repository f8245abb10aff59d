#!/bin/bash
# device-resident vmult rate of every caching level / formulation at Q3 (+ the degree sweep of tensor2)
mkdir -p gpurun_out
: > gpurun_out/levels.jsonl
for c in tensor2 scalar_referential tensor4 tensor4_ns; do
  timeout 600 python bench.py --caching $c --no-cpu --no-e2e --steps 10 >> gpurun_out/levels.jsonl 2>> gpurun_out/levels.err
done
timeout 600 python bench.py --caching tensor2 --formulation 0 --no-cpu --no-e2e --steps 10 >> gpurun_out/levels.jsonl 2>> gpurun_out/levels.err
timeout 600 python bench.py --dtype f32 --no-cpu --no-e2e --steps 10 >> gpurun_out/levels.jsonl 2>> gpurun_out/levels.err
timeout 900 python bench.py --sweep --no-cpu --no-e2e --steps 10 >> gpurun_out/levels.jsonl 2>> gpurun_out/levels.err
python - <<'PY'
import json
for l in open('gpurun_out/levels.jsonl'):
    d=json.loads(l); print(d['config']['workload'], d['dtype'], '%.3f ms'%d['ms_per_step'], '%.3e DoF/s'%d['value'], 'frac %.3f'%d['roofline']['frac'])
    if 'sweep' in d:
        for k,v in d['sweep'].items(): print('  ',k,'%.3f ms %.3e DoF/s frac %.3f'%(v['ms'],v['dofs_per_s'],v['frac']))
PY
tail -3 gpurun_out/levels.err
