#!/bin/bash
# build_variant.sh NAME [nvcc -D flags...]: liblcs_b200 with tuning defines -> variants/NAME.so (tuning experiments only)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared --expt-relaxed-constexpr "$@" \
  -o variants/$name.so lagrangiancoherence_b200/csrc/{advect,prefilter,epilogue,filters,seams}.cu
echo built variants/$name.so
