#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --sweep > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python tools/show_bench.py gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/launches_bench.log 2>&1
for p in 2 3 4; do
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:NHOp -s 3 -c 1 -f -o gpurun_out/prof_q$p python bench.py --degree $p --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_q$p.log 2>&1
  echo "ncu q$p rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:NHOp -s 3 -c 1 -f -o gpurun_out/prof_q3_f32 python bench.py --degree 3 --dtype f32 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_q3f32.log 2>&1
