#!/bin/bash
# builds the micro-benchmarks next to their sources (tools/ubench/build/ travels to the GPU box, git ignores it)
set -e
D=$(cd "$(dirname "$0")" && pwd)
mkdir -p "$D/build"
for f in "$D"/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -lineinfo $EXTRA -I"$D/../../include" \
       -I"$D/../../advanced-soft-actor-critic_b200/asac_b200/csrc" "$f" -o "$D/build/$(basename "${f%.cu}")" -lcuda
done
