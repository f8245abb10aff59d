#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/variants.jsonl
for v in default c3 c2 c6m2; do
  if [ $v = default ]; then unset B200MF_LIBRARY; else export B200MF_LIBRARY=$PWD/gpurun_variants/libb200mf_$v.so; fi
  timeout 300 python tools/bench_laplace.py --tag $v >> gpurun_out/variants.jsonl 2>> gpurun_out/variants.err
done
unset B200MF_LIBRARY
timeout 300 python tools/bench_laplace.py --tag default-mapped --mapped >> gpurun_out/variants.jsonl 2>> gpurun_out/variants.err
for p in 1 2 3 5 6; do timeout 300 python tools/bench_laplace.py --tag default --degree $p --cells $((256/p)) >> gpurun_out/variants.jsonl 2>> gpurun_out/variants.err; done
cat gpurun_out/variants.jsonl; tail -3 gpurun_out/variants.err
