#!/bin/bash
export ASAC_PDL=1
for combo in "2 1" "2 0" "1 1" "1 0" "0 1" "0 0"; do
  set -- $combo
  ASAC_ADAM_NARROW=$1 ASAC_POST_LATE=$2 python bench.py --steps 600 --warmup 20 --no-sub-results --cpu-seconds 0.5 > gpurun_out/bis.json 2>gpurun_out/bis.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bis.json').read().strip().splitlines()[-1])
print('narrow=$1 post_late=$2', 'cold us %.1f' % (d['ms_per_step']*1000), 'warm us %.1f' % (1e6/d['value_warm_l2']))
PY
done
