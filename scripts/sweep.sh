#!/bin/bash
# Sweeps the trace-engine knobs (environment overrides read by render_impl / make_dev_scene).
run() { python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | python scripts/benchline.py; }
for mb in 8 10; do for ml in 4 8 12 16; do echo "MINB=$mb MIN_LANES=$ml"; CRB_MINB=$mb CRB_MIN_LANES=$ml run; done; done
for sl in 8 16 64; do echo "SLICE=$sl"; CRB_NODE_SLICE=$sl run; done
