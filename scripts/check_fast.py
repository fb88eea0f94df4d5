#!/usr/bin/env python
"""Order-free engine vs reference-order engine vs oracle on the BASELINE scenes (GPU box).
  python scripts/check_fast.py [config ...]      (instanced:N = N mesh copies)
Per config: ids / t bit-exact on three ray batches for both engines, retried rays, and a timed render with each engine
(images must be bit-identical: same paths, order-independent accumulation)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from conftest import random_rays
from scenes_util import scene_bounds

from crucible_b200 import abi, demo_builder
from crucible_b200.gpu import GpuScene
from oracle import binding as oracle

SPP = {"book1": 100, "cornell": 200, "teapot": 256, "instanced": 16}
for arg in sys.argv[1:] or ["book1", "teapot", "cornell", "instanced:200"]:
    name, _, copies = arg.partition(":")
    kw = {"samples": SPP[name]}
    if copies:
        kw["copies"] = int(copies)
    sc = demo_builder.CONFIGS[name](**kw)
    desc, cam = sc.describe(), sc.scene_cam.to_abi()
    t0 = time.time()
    gs = GpuScene(desc, 0)
    ci = gs.commit_info()
    out = {"config": arg, "commit_s": round(time.time() - t0, 3), "ms_search_tree": round(ci["ms_search_tree"], 1), "bvh": gs.bvh_info()}
    if not os.environ.get("NO_ORACLE"):
        orc = oracle.OracleScene(desc)
        lo, hi = scene_bounds(desc)
        n = 1 << 19
        batches = {"primary": orc.gen_rays(cam, 0, cam.image_width * cam.image_height)[:: max(1, cam.image_width * cam.image_height // n)],
                   "bounce": orc.gen_rays(cam, 3, n, seed=7), "random": random_rays(n, lo, hi, 42)}
        for b, rays in batches.items():
            exp = orc.trace_batch(rays)
            for eng, ref in (("fast", False), ("ref", True)):
                got = gs.trace_batch(rays, reference_order=ref)
                bad = int((got["prim_index"] != exp["prim_index"]).sum())
                hit = exp["prim_index"] >= 0
                bad_t = int((got["t"][hit] != exp["t"][hit]).sum())
                out[f"{b}_{eng}"] = {"id_mismatch": bad, "t_mismatch": bad_t, "retried": gs.last_retried() if not ref else None}
    imgs = {}
    for eng, ref in (("fast", False), ("ref", True)):
        for prec, pname in ((abi.CR_PRECISION_F64, "f64"), (abi.CR_PRECISION_F32, "f32")):
            gs.render(cam, seed=1, precision=prec, want_rgb8=False, reference_order=ref)
            rgb, _, st = gs.render(cam, seed=1, precision=prec, want_rgb8=False, reference_order=ref, time_kernels=True)
            imgs[(eng, pname)] = rgb
            out[f"render_{eng}_{pname}"] = {"ms_total": round(st["ms_total"], 2), "ms_trace": round(st["ms_trace"], 2), "ms_shade": round(st["ms_shade"], 2),
                                            "msamples_per_s": round(st["samples"] / st["ms_total"] / 1e3, 1), "rays": st["rays"], "retried": st["retried_rays"]}
    out["f64_images_identical"] = bool(np.array_equal(imgs[("fast", "f64")], imgs[("ref", "f64")]))
    out["f32_mean_diff"] = float(abs(imgs[("fast", "f32")].mean() - imgs[("ref", "f32")].mean()))
    print(json.dumps(out), flush=True)
    gs.close()
