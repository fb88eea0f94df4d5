#!/usr/bin/env python
"""One rank's share of the 8-GPU book1 frame (rows of rank 0 of 8) on one GPU: k_tail with rebalancing rounds
(CRB_TAIL_ROUND = bounces per round, 0 = the single-round tail of round 2h) at several hand-over points (CRB_TAIL)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crucible_b200 import demo_builder
from crucible_b200.gpu import GpuScene
sc = demo_builder.book1_end_scene(image_width=1920, samples=100)
gs, cam = GpuScene(sc.describe(), 0), sc.scene_cam.to_abi()
ref = None
for world in (8, 1):
    for spec in os.environ.get("SPECS", "0:65536 4:65536 2:65536 3:65536 6:65536 4:131072 4:262144 3:262144 4:524288 0:65536").split():
        rnd, tail = spec.split(":")
        os.environ["CRB_TAIL_ROUND"], os.environ["CRB_TAIL"] = rnd, tail
        best = 1e9
        for _ in range(4 if world == 8 else 2):
            rgb, _, st = gs.render(cam, seed=1, row_world=world, row_rank=0, want_rgb=True)
            best = min(best, st["ms_total"])
        if world == 8:
            if ref is None:
                ref = rgb.copy()
            same = bool((rgb == ref).all())
        print(json.dumps({"world": world, "round": int(rnd), "tail": int(tail), "ms_total": round(best, 3), "iterations": st["iterations"],
                          "launches": st["launches"], "rays": st["rays"], "identical_image": same if world == 8 else None}), flush=True)
