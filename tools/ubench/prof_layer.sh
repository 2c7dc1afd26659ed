#!/bin/bash
# ncu source-level capture of one layer-chain variant: tools/ubench/prof_layer.sh <variant>
V=$1
ncu --set full --clock-control none --import-source on -k regex:k_chain --launch-skip 3 -c 1 -f -o gpurun_out/prof_ubench_v$V \
    tools/ubench/build/layer_chain $V > gpurun_out/prof_ubench_v$V.log 2>&1
echo "rc=$?"
