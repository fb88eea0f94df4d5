#!/usr/bin/env python
"""Last check of the round on the GPU box: the Cornell box (quads in shared memory at the padded stride) in f32 and f64 through
the shared-memory engine against the reference-order engine, then smoke()."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from crucible_b200 import abi, demo_builder
from crucible_b200.gpu import GpuScene
sc = demo_builder.CONFIGS["cornell"](image_width=160, samples=64)
gs, cam = GpuScene(sc.describe(), 0), sc.scene_cam.to_abi()
for prec, name in ((abi.CR_PRECISION_F64, "f64"), (abi.CR_PRECISION_F32, "f32")):
    a, _, sa = gs.render(cam, seed=3, precision=prec)
    b, _, sb = gs.render(cam, seed=3, precision=prec, reference_order=True)
    d = np.abs(a - b)
    print(name, "engine", sa["trace_engine"], "vs", sb["trace_engine"], "max diff", float(d.max()), "mean diff", float(d.mean()), "mean", float(a.mean()), flush=True)
    assert sa["trace_engine"] == 3
    assert (d.max() == 0.0) if name == "f64" else (d.mean() < 2e-3)
import __graft_entry__ as g
g.smoke()
