#!/usr/bin/env python
"""Small driver for ncu captures of the kernels a bench run launches rarely: k_trace_batch (cr_trace_batch on
1 M first-bounce rays of book1), k_tail and k_resolve (one 1080p render at 8 spp in f64 and f32)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from crucible_b200 import abi, demo_builder
from crucible_b200.gpu import GpuScene

sc = demo_builder.book1_end_scene(image_width=1920, samples=8, seed=1)
desc, cam = sc.describe(), sc.scene_cam.to_abi()
gs = GpuScene(desc, 0)
rng = np.random.default_rng(1)
n = 1 << 20
o = np.array([13.0, 2.0, 3.0]) + rng.normal(scale=0.05, size=(n, 3))
d = -o + rng.normal(scale=3.0, size=(n, 3))
rays = np.concatenate([o, d, np.zeros((n, 1))], 1)
for prec in (abi.CR_PRECISION_F64, abi.CR_PRECISION_F32):
    hits = gs.trace_batch(rays, precision=prec)
    _, _, st = gs.render(cam, seed=1, precision=prec)
    print(prec, int((hits["prim_index"] >= 0).sum()), st["rays"], st["launches"])
gs.close()
