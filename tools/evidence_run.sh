#!/bin/bash
# One consolidated evidence run on the GPU box: tests, smoke, default bench + reference arm, ncu launch lists and
# full captures.  tools/evidence_run.sh <tag>
TAG=${1:-r02q}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/${TAG}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 200 --warmup 5 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err
python tools/phase_breakdown.py > gpurun_out/${TAG}_phase_breakdown.txt 2>&1
tools/gpu_profile.sh ${TAG}_c2 "k_policy_backward|k_value_pass|k_q_backward|k_reduce_adam|k_step_epilogue" c2 > gpurun_out/${TAG}_profile_c2.log 2>&1
SKIP_FULL=1 tools/gpu_profile.sh ${TAG}_c3 x c3 > gpurun_out/${TAG}_profile_c3.log 2>&1
SKIP_FULL=1 tools/gpu_profile.sh ${TAG}_c4 x c4 > gpurun_out/${TAG}_profile_c4.log 2>&1
tail -3 gpurun_out/${TAG}_gpu_tests.log; cat gpurun_out/${TAG}_smoke.log | tail -2
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('c2', round(d['value']), round(d['value_warm_l2']), round(d['e2e']['value']), d['clocks'])
for k,v in d.get('other_configs',{}).items(): print(k, round(v['value']), round(v.get('value_warm_l2',0)), round(v['e2e']['value']))
r=json.loads(open('gpurun_out/${TAG}_bench_reference_arm.json').read().strip().splitlines()[-1])
print('reference arm', r.get('value'), r.get('cpu_baseline'))
PY
