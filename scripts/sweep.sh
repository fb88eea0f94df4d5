#!/bin/bash
# usage: scripts/sweep.sh  (on the GPU box) — trace-kernel tuning sweep at the full workload
run() {
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('  Msamples/s %.1f  ms/step %.2f trace %.2f shade %.2f gen %.2f frac %.3f'%(d['value'],d['ms_per_step'],r['ms_trace'],r['ms_shade'],r['ms_raygen'],r['frac']))
"
}
for mb in 4 6 8; do for r in 16 24; do for sl in 4 8; do
  echo "MINB=$mb REFILL=$r SLICE=$sl"; CRB_MINB=$mb CRB_REFILL=$r CRB_NODE_SLICE=$sl run
done; done; done
