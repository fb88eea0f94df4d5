#!/bin/bash
run() {
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('  Msamples/s %.1f  ms/step %.2f trace %.2f shade %.2f gen %.2f frac %.3f launches %d e2e %.1f'%(d['value'],d['ms_per_step'],r['ms_trace'],r['ms_shade'],r['ms_raygen'],r['frac'],d['gpu_launches'],d['e2e']['value']))
"
}
for pool in 2097152 4194304 8388608 16777216 33554432; do echo "pool=$pool (100 spp)"; run --pool $pool; done
echo f32 16M; run --precision f32 --pool 16777216
