#!/usr/bin/env python
"""A/B of trace-engine knobs on ONE box, one process: every config's scene is committed once, then rendered under each
environment setting (the library reads its CRB_* knobs at render time).  Prints ms_trace / ms_total of the last render.

  python scripts/ab_fast.py "CRB_FAST_MINB=5" "CRB_FAST_MINB=8 CRB_REFILL_RT=16" ...   [CONFIGS=book1,teapot SPP=32,64]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from crucible_b200 import abi, demo_builder
from crucible_b200.gpu import GpuScene

DEFAULT_SPP = {"book1": 32, "cornell": 100, "teapot": 64, "instanced": 8}
configs = os.environ.get("CONFIGS", "book1,cornell,teapot,instanced").split(",")
settings = sys.argv[1:] or [""]
for name in configs:
    sc = demo_builder.CONFIGS[name](samples=int(os.environ.get("SPP_" + name.upper(), DEFAULT_SPP.get(name, 16))))
    desc, cam = sc.describe(), sc.scene_cam.to_abi()
    gs = GpuScene(desc, 0)
    for spec in settings:
        kv = dict(x.split("=", 1) for x in spec.split())
        for k, v in kv.items():
            os.environ[k] = v
        best = None
        for _ in range(int(os.environ.get("RENDERS", 3))):
            _, _, st = gs.render(cam, seed=1, time_kernels=True, want_rgb=False, want_rgb8=False)
            if best is None or st["ms_total"] < best["ms_total"]:
                best = st
        for k in kv:
            del os.environ[k]
        print(json.dumps({"config": name, "env": spec, "ms_trace": round(best["ms_trace"], 2), "ms_shade": round(best["ms_shade"], 2),
                          "ms_total": round(best["ms_total"], 2), "engine": best["trace_engine"], "retried": best["retried_rays"],
                          "msamples_per_s": round(best["samples"] / best["ms_total"] / 1e3, 1)}), flush=True)
    gs.close()
