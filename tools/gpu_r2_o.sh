#!/bin/bash
# round 2, 2-GPU call O: distributed tests (incl. tube, default overlap with p2p) + racecheck of the sliced kernels + ncu captures for profiles/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2o_pytest_multi.log 2>&1
tail -6 gpurun_out/r2o_pytest_multi.log
for p in 3 4 5 6; do
  echo "== racecheck test_nh_vmult_fp64[$p]"
  timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest "tests/test_gpu_parity.py::test_nh_vmult_fp64[$p]" -q -x 2>&1 | grep -E "RACECHECK SUMMARY|passed|failed|ERROR SUMMARY|Error: Race" | sort | uniq -c | head -6
done > gpurun_out/r2o_racecheck.txt 2>&1
cat gpurun_out/r2o_racecheck.txt
for p in 3 4 5; do
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:NHOp -s 3 -c 1 -o gpurun_out/r2o_q$p -f python bench.py --degree $p --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/r2o_ncu_q$p.log 2>&1
ncu -i gpurun_out/r2o_q$p.ncu-rep --page raw --csv > gpurun_out/r2o_ncu_q${p}_raw.csv 2>/dev/null
ncu -i gpurun_out/r2o_q$p.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2o_ncu_q${p}_source.csv.gz
rm -f gpurun_out/r2o_q$p.ncu-rep
done
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k regex:scalar_stack_kernel -s 3 -c 1 -o gpurun_out/r2o_heat -f python tools/bench_laplace.py --degree 4 --cells 64 > gpurun_out/r2o_ncu_heat.log 2>&1
ncu -i gpurun_out/r2o_heat.ncu-rep --page raw --csv > gpurun_out/r2o_ncu_heat_q4_raw.csv 2>/dev/null
rm -f gpurun_out/r2o_heat.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2o_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > /dev/null 2>&1
ls -la gpurun_out | grep r2o
