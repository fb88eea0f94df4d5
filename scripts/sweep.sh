#!/bin/bash
run() {
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('  Msamples/s %.1f  ms/step %.2f trace %.2f shade %.2f gen %.2f frac %.3f'%(d['value'],d['ms_per_step'],r['ms_trace'],r['ms_shade'],r['ms_raygen'],r['frac']))
"
}
for r in 4 8 12 16 24; do for ml in 8; do echo "REFILL=$r MIN_LANES=$ml"; CRB_REFILL=$r CRB_MIN_LANES=$ml run; done; done
echo "REFILL=8 MINB=6"; CRB_REFILL=8 CRB_MINB=6 run
