#!/bin/bash
# ncu --set full capture of the NH tensor2 vmult kernel for the degrees given as arguments
mkdir -p gpurun_out
for p in "$@"; do
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:NHOp -s 3 -c 1 -f -o gpurun_out/prof_q$p python bench.py --degree $p --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_q$p.log 2>&1
  echo "ncu q$p rc=$?"; tail -3 gpurun_out/ncu_q$p.log | cut -c1-300
done
ls -la gpurun_out
