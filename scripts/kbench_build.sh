#!/bin/bash
# Builds kbench variants: scripts/kbench_build.sh name "-Dflags" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p build/kbench
while [ $# -gt 0 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Iinclude \
    -Iceres-solver-cuda_b200/include -Iceres-solver-cuda_b200/examples $flags \
    -o build/kbench/kb_$name scripts/kbench.cu &
done
wait
ls build/kbench
