#!/bin/bash
run() {
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('  Msamples/s %.1f  ms/step %.2f trace %.2f shade %.2f gen %.2f frac %.3f'%(d['value'],d['ms_per_step'],r['ms_trace'],r['ms_shade'],r['ms_raygen'],r['frac']))
"
}
for r in 8 16 24 32; do for ml in 4 8 12; do echo "REFILL=$r MIN_LANES=$ml"; CRB_REFILL=$r CRB_MIN_LANES=$ml run; done; done
