#!/bin/bash
# round 2, GPU call R: full GPU suite, racecheck of the sliced kernels after the inactive-lane guard, smoke, default bench + reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2r_pytest.txt
for d in 4 5 6; do
  timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest "tests/test_gpu_parity.py::test_nh_vmult_fp64[$d]" -q -x > gpurun_out/r2r_racecheck_q$d.txt 2>&1
  echo "Q$d: $(grep -c 'hazard' gpurun_out/r2r_racecheck_q$d.txt) hazard lines; $(grep 'RACECHECK SUMMARY' gpurun_out/r2r_racecheck_q$d.txt | tail -1)" | tee -a gpurun_out/r2r_racecheck_summary.txt
  grep -B2 -A10 "hazard" gpurun_out/r2r_racecheck_q$d.txt | head -30 | cut -c1-220
  gzip -f gpurun_out/r2r_racecheck_q$d.txt
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; tail -c 3000 gpurun_out/r2r_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/r2r_bench_reference.json 2>> gpurun_out/r2r_bench.err; cat gpurun_out/r2r_bench_reference.json | cut -c1-600
