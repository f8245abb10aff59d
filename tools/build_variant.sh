#!/bin/bash
# tuning: tools/build_variant.sh <name> <extra nvcc flags...>  -> gpurun_variants/libb200mf_<name>.so (separate object dir)
name=$1; shift
cd "$(dirname "$0")/../matrix-free-dealii-precice_b200/csrc"
mkdir -p ../../gpurun_variants
make -s BUILD=build_$name OUT=../../gpurun_variants/libb200mf_$name.so EXTRA="$*" 2>&1 | grep -v "^$" | head
