#!/bin/bash
# round 2, GPU call Z (2 GPUs): distributed tests with the recorded lower MG levels + smoother hook, in both exchange modes; N=2 bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r2z_pytest_multi_2gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2z_pytest_multi_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2z_scale2.json 2> gpurun_out/r2z_scale2.err
grep '^{' gpurun_out/r2z_scale2.json | python -c "
import sys, json
d = json.loads(sys.stdin.readlines()[-1])
print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step', 'ms_per_step_plain_loop', 'ms_per_step_graph_replay_rank0', 'parity', 'ghost_exchange')})
print('e2e', d.get('e2e'))"
tail -3 gpurun_out/r2z_scale2.err
