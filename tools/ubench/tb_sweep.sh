for tb in 0 8 16; do
  ASAC_TILE_BATCH=$tb python bench.py --steps 1000 --warmup 20 --no-sub-results --cpu-seconds 1 > gpurun_out/r02n_tb$tb.json 2> gpurun_out/r02n_tb$tb.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02n_tb$tb.json').read().strip().splitlines()[-1])
print('TB=$tb', 'value', round(d['value']), 'warm', round(d['value_warm_l2']), 'us', round(d['ms_per_step']*1000,1))
print('   ', {k['name'] if isinstance(k,dict) and 'name' in k else str(k)[:40]: (k.get('us') if isinstance(k,dict) else None) for k in d.get('kernels',[])})
PY
done
