#!/bin/bash
# the round's evidence run on one B200: parity tests, bench lines, launch list, ncu --set full of the top kernel, solver benches
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --sweep > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python tools/show_bench.py gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/launches_bench.log 2>&1
# (gpurun_out/ may hold 64 MiB: source import only for the headline kernel)
for p in 2 3 4; do
  src=""; [ $p = 3 ] && src="--import-source on"
  timeout 600 ncu --set full --clock-control none $src --kernel-name-base demangled -k regex:NHOp -s 3 -c 1 -f -o gpurun_out/prof_q$p python bench.py --degree $p --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_q$p.log 2>&1
  echo "ncu q$p rc=$?"
done
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k regex:NHOp -s 3 -c 1 -f -o gpurun_out/prof_q3_f32 python bench.py --degree 3 --dtype f32 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_q3f32.log 2>&1
: > gpurun_out/solver.jsonl
timeout 900 python tools/bench_solver.py newton >> gpurun_out/solver.jsonl 2> gpurun_out/solver.err || echo "newton failed"
timeout 900 python tools/bench_solver.py heat >> gpurun_out/solver.jsonl 2>> gpurun_out/solver.err || echo "heat failed"
cut -c1-600 gpurun_out/solver.jsonl; tail -3 gpurun_out/solver.err
(cd tools/microbench && nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_pipes fp64_pipes.cu && ./fp64_pipes) > gpurun_out/fp64_pipes.jsonl 2>&1; cat gpurun_out/fp64_pipes.jsonl
timeout 300 python tools/bench_laplace.py --tag final >> gpurun_out/laplace.jsonl 2>> gpurun_out/solver.err; tail -1 gpurun_out/laplace.jsonl | cut -c1-400

# gpurun_out/ may hold 64 MiB and one report is ~20 MB: export the pages as CSV on the box and drop the reports
for r in gpurun_out/prof_*.ncu-rep; do
  ncu -i $r --page raw --csv > ${r%.ncu-rep}_raw.csv 2>/dev/null
done
ncu -i gpurun_out/prof_q3.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/prof_q3_source.csv.gz
rm -f gpurun_out/prof_*.ncu-rep
du -sh gpurun_out
