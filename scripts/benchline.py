#!/usr/bin/env python
"""Reads one bench.py JSON line on stdin and prints the figures used when tuning kernels."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline", {})
    print(f"Msamples/s {d['value']:.1f}  ms/step {d['ms_per_step']:.2f}  trace {r.get('ms_trace', 0):.2f}  shade {r.get('ms_shade', 0):.2f}  "
          f"gen {r.get('ms_raygen', 0):.2f}  frac {r.get('frac', 0):.3f}  e2e {d['e2e']['value']:.1f}  launches {d.get('gpu_launches')}")
