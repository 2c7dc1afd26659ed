#!/bin/bash
# Same-box ablation of the round-2 schedule switches on the default bench workload (config 2): each line is the full
# product path with ONE switch turned off.  tools/ablation.sh > gpurun_out/<tag>_switch_ablation.txt
run() {
  env "$@" python bench.py --steps 1000 --warmup 20 --no-sub-results --cpu-seconds 0.5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-28s %7.1f us/step L2-flushed  %7.1f warm  %6.0f steps/s  e2e %6.0f' % ('$*', d['ms_per_step']*1000, 1e6/d['value_warm_l2'], d['value'], d['e2e']['value']))"
}
run ASAC_NONE=1
run ASAC_PDL=0
run ASAC_LATE_WAIT=0
run ASAC_Q_HANDOFF=0
run ASAC_PI_HANDOFF=0
run ASAC_PUSH_COMBINE=0
run ASAC_INGEST_FUSED=0
run ASAC_SAMPLE_AHEAD=0
run ASAC_NONE=1
