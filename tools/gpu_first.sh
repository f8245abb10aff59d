#!/bin/bash
# one gpurun call: parity tests, bench + sweep, launch list, ncu full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --sweep > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/launches_bench.log 2>&1
for p in 2 3 4; do
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:NHOp -s 3 -c 1 -f -o gpurun_out/prof_q$p python bench.py --degree $p --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_q$p.log 2>&1
  echo "ncu q$p rc=$?"
done
ls -la gpurun_out
