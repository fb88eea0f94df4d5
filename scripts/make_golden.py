#!/usr/bin/env python
"""Writes tests/golden/<config>.npz from the ORACLE (oracle/oracle.cpp at the current commit): per config a strided
subset of each SURVEY 8d ray batch (rays + CrHit records) and one 64 px, 4 spp f64 render.  Run once when the oracle
is at a reviewed state; tests/test_golden.py then fails if the oracle stops reproducing the files.

  python scripts/make_golden.py [config ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import golden
from oracle import binding as oracle

names = sys.argv[1:] or list(golden.CONFIGS) + ["nested"]
if "nested" in names:  # nested elements: leaf order + 2048 seeded random rays of tests/scenes_util.nested_scene(1)
    from conftest import random_rays
    from scenes_util import nested_scene, scene_bounds

    names.remove("nested")
    d = nested_scene(1)
    orc = oracle.OracleScene(d)
    lo, hi = scene_bounds(d)
    rays = random_rays(2048, lo, hi, 77)
    np.savez_compressed(golden.path("nested"), rays=rays, hits=orc.trace_batch(rays), leaf_order=orc.bvh_leaf_order())
    print("nested", os.path.getsize(golden.path("nested")), "bytes")
for name in names:
    sc = golden.build(name)
    desc, cam = sc.describe(), sc.scene_cam.to_abi()
    orc = oracle.OracleScene(desc)
    out = {"bvh": np.array([orc.bvh_info()[k] for k in ("n_nodes", "max_depth", "n_visible")], np.int64)}
    for bname, rays in golden.batches(desc, cam, orc).items():
        idx = golden.keep_indices(len(rays))
        sub = np.ascontiguousarray(rays[idx])
        out[bname + "_idx"] = idx
        out[bname + "_rays"] = sub
        out[bname + "_hits"] = orc.trace_batch(sub)
    small = golden.small_camera(name)
    rgb, rgb8, st = orc.render(small, seed=golden.RENDER["seed"])
    out["render_rgb"], out["render_rgb8"], out["render_rays"] = rgb, rgb8, np.array([st["rays"]], np.int64)
    os.makedirs(golden.GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(golden.path(name), **out)
    print(name, {k: v.shape for k, v in out.items()}, os.path.getsize(golden.path(name)), "bytes")
