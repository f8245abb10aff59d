#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/solver.jsonl
timeout 600 python tools/bench_solver.py heat --levels 3 >> gpurun_out/solver.jsonl 2> gpurun_out/solver.err || echo "heat small failed"
timeout 600 python tools/bench_solver.py newton --levels 3 >> gpurun_out/solver.jsonl 2>> gpurun_out/solver.err || echo "newton small failed"
timeout 1500 python tools/bench_solver.py newton >> gpurun_out/solver.jsonl 2>> gpurun_out/solver.err || echo "newton failed"
timeout 1500 python tools/bench_solver.py heat >> gpurun_out/solver.jsonl 2>> gpurun_out/solver.err || echo "heat failed"
cat gpurun_out/solver.jsonl | cut -c1-1500
tail -5 gpurun_out/solver.err
