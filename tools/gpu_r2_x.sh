#!/bin/bash
# round 2, GPU call X: hooked smoother after the update-kernel fix
mkdir -p gpurun_out
for h in 1 0; do B200MF_SMOOTHER_HOOK=$h timeout 300 python tools/smoother_hook_ab.py 64 2>&1 | grep '^{'; done | tee gpurun_out/r2x_hook_ab.jsonl
timeout 600 python -m pytest tests/test_gpu_heat.py tests/test_gpu_solver.py -q 2>&1 | tail -40 | cut -c1-250
