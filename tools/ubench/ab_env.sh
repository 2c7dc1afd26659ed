#!/bin/bash
# A/B of an environment switch on the default bench: tools/ubench/ab_env.sh VAR v1 v2 ... [-- extra bench args]
VAR=$1; shift
for v in "$@"; do
  env $VAR=$v python bench.py --steps 1000 --warmup 20 --no-sub-results --cpu-seconds 1 > gpurun_out/ab_${VAR}_$v.json 2> gpurun_out/ab_${VAR}_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/ab_${VAR}_$v.json').read().strip().splitlines()[-1])
print('$VAR=$v', 'cold us %.1f' % (d['ms_per_step']*1000), 'warm us %.1f' % (1e6/d['value_warm_l2']), 'e2e %.0f' % d['e2e']['value'])
PY
done
