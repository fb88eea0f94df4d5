#!/bin/bash
# usage: scripts/sweep.sh  (on the GPU box) — trace-kernel tuning sweep, reduced spp, NOT a headline number
for r in 4 8 12 16 24 32; do
  echo "REFILL=$r"; CRB_REFILL=$r python bench.py --samples 16 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('  Msamples/s %.1f  ms/step %.2f trace %.2f shade %.2f gen %.2f frac %.3f'%(d['value'],d['ms_per_step'],r['ms_trace'],r['ms_shade'],r['ms_raygen'],r['frac']))
"
done
