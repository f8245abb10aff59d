#!/bin/bash
# multi-GPU call (gpurun --gpus N): parity tests incl. the distributed ones, then the weak-scaling bench at 1..N
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_multi.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_multi.log
: > gpurun_out/scale.jsonl
timeout 600 python bench.py --no-cpu --no-e2e >> gpurun_out/scale.jsonl 2>> gpurun_out/scale.err
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --no-cpu --no-e2e >> gpurun_out/scale.jsonl 2>> gpurun_out/scale.err
  fi
done
python - <<'PY'
import json
for l in open('gpurun_out/scale.jsonl'):
    if not l.startswith('{'): continue
    d=json.loads(l); print(d['n_gpus'], '%.3f ms'%d['ms_per_step'], '%.3e DoF/s'%d['value'], 'kernel share %.3f'%d['roofline']['kernel_share_of_step'], d.get('ghost_exchange'))
PY
tail -5 gpurun_out/scale.err
