#!/bin/bash
# parity tests + device-resident sweep + heat solver bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --sweep --no-cpu --no-e2e --steps 10 > gpurun_out/sweep.json 2> gpurun_out/sweep.err; python tools/show_bench.py gpurun_out/sweep.json; tail -3 gpurun_out/sweep.err
: > gpurun_out/solver.jsonl
timeout 900 python tools/bench_solver.py heat >> gpurun_out/solver.jsonl 2> gpurun_out/solver.err || echo "heat failed"
cut -c1-900 gpurun_out/solver.jsonl; tail -5 gpurun_out/solver.err
