#!/bin/bash
# the round's evidence run on one B200: parity tests, bench line, launch list, ncu --set full of the top kernel, solver benches
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --sweep > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python tools/show_bench.py gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/launches_bench.log 2>&1
for p in 2 3 4; do
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:NHOp -s 3 -c 1 -f -o gpurun_out/prof_q$p python bench.py --degree $p --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_q$p.log 2>&1
  echo "ncu q$p rc=$?"
done
: > gpurun_out/solver.jsonl
timeout 900 python tools/bench_solver.py newton >> gpurun_out/solver.jsonl 2> gpurun_out/solver.err || echo "newton failed"
timeout 900 python tools/bench_solver.py heat >> gpurun_out/solver.jsonl 2>> gpurun_out/solver.err || echo "heat failed"
cut -c1-700 gpurun_out/solver.jsonl; tail -3 gpurun_out/solver.err
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/launches_newton.csv python tools/bench_solver.py newton --coarse 4 --levels 5 > gpurun_out/launches_newton.log 2>&1; echo "ncu newton rc=$?"
python tools/launch_breakdown.py gpurun_out/launches_newton.csv 22
