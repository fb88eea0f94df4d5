#!/usr/bin/env python
"""Every kernel of the library once, at modest sizes, for ONE ncu pass (north_star: "every kernel backed by a committed
ncu capture"; the dominant kernels also have full-size --set full captures of their own, see profiles/README.md).

  ncu --metrics <scripts/summarize_ncu.py METRICS> --clock-control none --csv ... python scripts/profile_all.py

Scenarios: book1 640 px x 8 spp in f64 (order-free engine, tree in shared memory and in global memory; reference order at
both register budgets) and f32; cr_trace_batch on 256 K rays (both engines, both precisions); the Cornell box (quads,
k_shade_emissive); a scene with object keyframes (the ANIM builds of trace / shade / tail); a scene with nested
elements; a 200 K-triangle mesh scene committed with the device BVH builder (k_bvh_*, k_flatten_*, the search-tree
kernels) and rendered; cr_measure_fma_peak."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

from crucible_b200 import abi, demo_builder
from crucible_b200.gpu import GpuScene


def render(gs, cam, **kw):
    _, _, st = gs.render(cam, seed=1, want_rgb=False, **kw)
    return st


def pinned(**env):
    for k in ("CRB_TRAVERSAL", "CRB_MINB"):
        os.environ.pop(k, None)
    os.environ.update(env)


sc = demo_builder.book1_end_scene(image_width=640, samples=8, seed=1)
desc, cam = sc.describe(), sc.scene_cam.to_abi()
gs = GpuScene(desc, 0)
for env in ({"CRB_TRAVERSAL": "s"}, {"CRB_TRAVERSAL": "f"}, {"CRB_TRAVERSAL": "r", "CRB_MINB": "8"}, {"CRB_TRAVERSAL": "r", "CRB_MINB": "10"}):
    pinned(**env)
    print("book1 f64", env, render(gs, cam)["trace_engine"])
pinned()
print("book1 f32", render(gs, cam, precision=abi.CR_PRECISION_F32)["trace_engine"])
rng = np.random.default_rng(1)
n = 1 << 18
o = np.array([13.0, 2.0, 3.0]) + rng.normal(scale=0.05, size=(n, 3))
d = -o + rng.normal(scale=3.0, size=(n, 3))
rays = np.concatenate([o, d, np.zeros((n, 1))], 1)
for prec in (abi.CR_PRECISION_F64, abi.CR_PRECISION_F32):
    for ro in (False, True):
        hits = gs.trace_batch(rays, precision=prec, reference_order=ro)
        print("trace_batch", prec, ro, int((hits["prim_index"] >= 0).sum()))
gs.close()

sc = demo_builder.cornell_box(image_width=256, samples=16)
gs = GpuScene(sc.describe(), 0)
print("cornell", render(gs, sc.scene_cam.to_abi())["rays"])
gs.close()

from test_object_animation import _moving_scene  # noqa: E402  (the keyframed scene of the parity tests)

sc = _moving_scene(image_width=320, samples=8)
gs = GpuScene(sc.describe(), 0)
print("animated", render(gs, sc.scene_cam.to_abi())["rays"])
gs.close()

from scenes_util import nested_scene  # noqa: E402

gs = GpuScene(nested_scene(1), 0)
cam2 = demo_builder.book1_end_scene(image_width=320, samples=8).scene_cam.to_abi()
print("nested", render(gs, cam2)["rays"])
gs.close()

sc = demo_builder.instanced_teapots(image_width=640, samples=4, copies=32, grid=6, spacing=5.0)  # 202 K triangles: device builders
gs = GpuScene(sc.describe(), 0)
print("mesh", gs.commit_info()["builder"], render(gs, sc.scene_cam.to_abi())["rays"])
gs.close()

f64p, f32p = abi.C.c_double(), abi.C.c_double()
abi.check(abi.load().cr_measure_fma_peak(0, abi.C.byref(f64p), abi.C.byref(f32p)))
print("fma peaks", f64p.value, f32p.value)
