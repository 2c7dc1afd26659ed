#!/bin/bash
# compute-sanitizer over the GPU tests of the update path (memcheck: all of them; racecheck: the stage-by-stage tests)
TAG=${1:-r02r}
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_sac.py tests/test_gpu_learner.py tests/test_gpu_rep.py tests/test_gpu_per.py tests/test_gpu_discrete.py -x -q -k "not sweep" > gpurun_out/${TAG}_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/${TAG}_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_sac.py -x -q -k "every_stage_matches_oracle or fused" > gpurun_out/${TAG}_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/${TAG}_sanitizer_racecheck.log
