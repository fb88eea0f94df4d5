#!/usr/bin/env python
"""Two renders of rank 0's share of a still at `world` ranks (ncu launch lists of the per-rank step: scripts/scale_proxy.py).
  python scripts/render_rank.py [world] [max_depth]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from crucible_b200 import demo_builder
from crucible_b200.gpu import GpuScene

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sc = demo_builder.CONFIGS["book1"]()
gs = GpuScene(sc.describe(), 0)
cam = sc.scene_cam.to_abi()
if len(sys.argv) > 2:
    cam.max_depth = int(sys.argv[2])
for _ in range(2):
    _, _, st = gs.render(cam, seed=1, row_world=world, row_rank=0, want_rgb=False, want_rgb8=False)
print(st)
