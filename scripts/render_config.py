#!/usr/bin/env python
"""One render of one BASELINE config (for ncu captures and quick timings on the GPU box).

  python scripts/render_config.py instanced --samples 8 [--precision f32] [--renders 2] [--pool N]
Prints one JSON line per render (CrStats of a time_kernels run)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from crucible_b200 import abi, demo_builder
from crucible_b200.gpu import GpuScene

ap = argparse.ArgumentParser()
ap.add_argument("config", choices=list(demo_builder.CONFIGS))
ap.add_argument("--samples", type=int, default=0)
ap.add_argument("--width", type=int, default=0)
ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
ap.add_argument("--renders", type=int, default=1)
ap.add_argument("--pool", type=int, default=0)
args = ap.parse_args()
kw = {}
if args.samples:
    kw["samples"] = args.samples
if args.width:
    kw["image_width"] = args.width
sc = demo_builder.CONFIGS[args.config](**kw)
desc, cam = sc.describe(), sc.scene_cam.to_abi()
gs = GpuScene(desc, 0)
prec = abi.CR_PRECISION_F64 if args.precision == "f64" else abi.CR_PRECISION_F32
for k in range(args.renders):
    _, _, st = gs.render(cam, seed=1, precision=prec, time_kernels=True, want_rgb=False, want_rgb8=False, pool_paths=args.pool)
    print(json.dumps({"config": args.config, "precision": args.precision, "render": k, "bvh": gs.bvh_info(),
                      "msamples_per_s": st["samples"] / st["ms_total"] / 1e3, "mrays_per_s": st["rays"] / st["ms_total"] / 1e3, **st}), flush=True)
gs.close()
