#!/bin/bash
# Tuning builds: tools/build_variant.sh <tag> "<extra nvcc flags>"  ->  gpurun_variants/<tag>/libb200mf.so
# (select at run time with B200MF_LIBRARY=gpurun_variants/<tag>/libb200mf.so)
set -e
tag=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$root/gpurun_variants/$tag"
make -C "$root/matrix-free-dealii-precice_b200/csrc" BUILD=build_$tag OUT="$root/gpurun_variants/$tag/libb200mf.so" EXTRA="$*" 2>&1 | grep -E "error|warning: v|Error" || true
ls -la "$root/gpurun_variants/$tag/libb200mf.so"
