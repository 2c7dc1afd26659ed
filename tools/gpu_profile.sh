#!/bin/bash
# Runs on the GPU box (under gpurun): ncu launch list of the bench's timed region and one
# `--set full` capture of the heaviest kernels.  Outputs land in gpurun_out/.
#   tools/gpu_profile.sh <tag> [kernel-regex] [config]
set -u
TAG=${1:-r02}
KRE=${2:-k_policy_backward|k_value_pass|k_q_backward}
CONFIG=${3:-c2}
BENCH="python bench.py --config ${CONFIG} --no-sub-results --steps 3 --warmup 3 --cpu-seconds 0.5"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_${TAG}.csv ${BENCH} > gpurun_out/launches_${TAG}.log 2>&1
echo "launch list rc=$?"
# same list without ncu's cache flush between launches (warm L2 / instruction cache)
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_warm_${TAG}.csv ${BENCH} > gpurun_out/launches_warm_${TAG}.log 2>&1
echo "warm launch list rc=$?"
if [ "${SKIP_FULL:-0}" = "1" ]; then exit 0; fi
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:${KRE}" -c 8 \
    -f -o gpurun_out/prof_${TAG} python bench.py --config ${CONFIG} --no-sub-results --steps 2 --warmup 3 --cpu-seconds 0.5 \
    > gpurun_out/prof_${TAG}.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/ | tail -5
