#!/bin/bash
# round 2, multi-GPU call: parity tests of the distributed paths (p2p + nccl exchange, vmult_add, Newton/GMG) and the
# weak-scaling bench at N = $1 (default 2) with both exchange paths
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2m_topo.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2m_pytest_multi.log 2>&1
tail -15 gpurun_out/r2m_pytest_multi.log
for mode in 1 0; do
  B200MF_P2P=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2m_scale${N}_p2p$mode.json 2> gpurun_out/r2m_scale${N}_p2p$mode.err
  tail -3 gpurun_out/r2m_scale${N}_p2p$mode.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2m_scale${N}_p2p$mode.json'))
    print('N=$N p2p=$mode value %.3e ms %.4f loop_ms %.4f e2e %.3e' % (d['value'], d['ms_per_step'], d.get('ms_per_step_plain_loop_rank0',0), d['e2e']['value']), d.get('timed_region'), d.get('exchange_path'))
    print(d.get('ghost_exchange')); print(d.get('parity')); print(d.get('step_ms_rank0'))
except Exception as e: print('no line', e)
PY
done

