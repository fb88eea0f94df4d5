#!/usr/bin/env python
"""Strong-scaling proxy on ONE GPU: the per-rank share of a still at N ranks (interleaved row blocks are statistically
identical, so rank 0 stands for all) against the whole frame.  efficiency ~ T(1) / (N * T(N)); the framebuffer exchange
is not included (measured separately by bench.py --gpus N).   python scripts/scale_proxy.py [config] [spp]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from crucible_b200 import demo_builder
from crucible_b200.gpu import GpuScene

name = sys.argv[1] if len(sys.argv) > 1 else "book1"
kw = {"samples": int(sys.argv[2])} if len(sys.argv) > 2 else {}
sc = demo_builder.CONFIGS[name](**kw)
desc, cam = sc.describe(), sc.scene_cam.to_abi()
gs = GpuScene(desc, 0)
t1 = None
for world in (1, 2, 4, 8):
    best = None
    for _ in range(4):
        _, _, st = gs.render(cam, seed=1, row_world=world, row_rank=0, time_kernels=True, want_rgb=False, want_rgb8=False)
        if best is None or st["ms_total"] < best["ms_total"]:
            best = st
    if world == 1:
        t1 = best["ms_total"]
    print(json.dumps({"config": name, "world": world, "ms_total": round(best["ms_total"], 3), "ms_trace": round(best["ms_trace"], 3),
                      "ms_shade": round(best["ms_shade"], 3), "ms_raygen": round(best["ms_raygen"], 3), "iterations": best["iterations"],
                      "launches": best["launches"], "efficiency_proxy": round(t1 / (world * best["ms_total"]), 4)}), flush=True)
gs.close()
