#!/bin/bash
# development aid: end-to-end arm of bench.py by chunk size and number of lanes
for cfg in "16 3" "16 4" "8 4" "32 3" "8 6" "4 6"; do
  set -- $cfg
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-chunk $1 --e2e-streams $2 > /tmp/e2e.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('/tmp/e2e.json'))
e=d['e2e']
print('chunk', e['chunk'], 'streams', e['streams'], 'e2e', round(e['value']), 'copy_only', round(e['copy_only_ops_per_s']), 'h2d GB/s', round(e['h2d_gbs_per_rank'],1))
PY
done
