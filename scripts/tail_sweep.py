#!/usr/bin/env python
"""One rank's share of the 8-GPU book1 frame (rows of rank 0 of 8) on one GPU, for different hand-over points to the tail
kernel (env CRB_TAIL = paths left when k_tail takes over).  Prints ms per frame."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crucible_b200 import demo_builder
from crucible_b200.gpu import GpuScene
sc = demo_builder.book1_end_scene(image_width=1920, samples=100)
gs, cam = GpuScene(sc.describe(), 0), sc.scene_cam.to_abi()
for world in (8, 4, 1):
    for tail in os.environ.get("TAILS", "16384,65536,262144,1048576").split(","):
        os.environ["CRB_TAIL"] = tail
        best = 1e9
        for _ in range(4):
            _, _, st = gs.render(cam, seed=1, row_world=world, row_rank=0, want_rgb=False)
            best = min(best, st["ms_total"])
        print(json.dumps({"world": world, "tail": int(tail), "ms_total": round(best, 3), "iterations": st["iterations"], "launches": st["launches"],
                          "ideal_ms": None}), flush=True)
