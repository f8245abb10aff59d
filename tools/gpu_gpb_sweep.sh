#!/bin/bash
# tuning: device-resident vmult rate per degree and warps-per-block (library built with -DB200MF_GPB_SWEEP)
mkdir -p gpurun_out; : > gpurun_out/gpb.jsonl
for p in 1 2 3 4; do for g in 1 2 4; do
  B200MF_GPB=$g timeout 300 python bench.py --degree $p --no-cpu --no-e2e --steps 10 2>>gpurun_out/gpb.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print(json.dumps({'p':$p,'gpb':$g,'ms':d['ms_per_step'],'kernel_ms':r['kernel_ms'],'frac':r['frac']}))" | tee -a gpurun_out/gpb.jsonl
done; done
for g in 1 2 4; do
  B200MF_GPB=$g timeout 300 python bench.py --degree 3 --dtype f32 --no-cpu --no-e2e --steps 10 2>>gpurun_out/gpb.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print(json.dumps({'p':3,'dtype':'f32','gpb':$g,'ms':d['ms_per_step'],'kernel_ms':r['kernel_ms'],'frac':r['frac']}))" | tee -a gpurun_out/gpb.jsonl
done
tail -3 gpurun_out/gpb.err
