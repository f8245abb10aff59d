#!/bin/bash
# round 2, 8-GPU call: weak scaling bench (p2p, then nccl) and the tube config (BASELINE configs[3]) on 8 GPUs
mkdir -p gpurun_out
N=8
for mode in 1 0; do
  B200MF_P2P=$mode timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2p_scale${N}_p2p$mode.json 2> gpurun_out/r2p_scale${N}_p2p$mode.err
  tail -2 gpurun_out/r2p_scale${N}_p2p$mode.err | cut -c1-300
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/bench_tube_multi.py --degree 5 --first 3 --levels 3 > gpurun_out/r2p_tube8.json 2> gpurun_out/r2p_tube8.err
tail -3 gpurun_out/r2p_tube8.err | cut -c1-400
grep "^{" gpurun_out/r2p_tube8.json | cut -c1-1500
python - <<'PY'
import json
for mode in (1,0):
    try:
        for l in open(f'gpurun_out/r2p_scale8_p2p{mode}.json'):
            if l.startswith('{'): d=json.loads(l)
        print('N=8 p2p',mode,'value %.4e ms %.4f loop_ms %.4f e2e %.3e'%(d['value'],d['ms_per_step'],d.get('ms_per_step_plain_loop_rank0',0),d['e2e']['value']), '|', d.get('exchange_path'))
        g=d['ghost_exchange']; print('   serial %.4f overl %.4f none %.4f'%(g['ms_serial_exchange'],g['ms_overlapped_exchange'],g['ms_no_exchange']), d['parity']['rel_l2'], d['step_ms_rank0'])
    except Exception as e: print(mode, e)
PY
