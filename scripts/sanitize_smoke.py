"""A small pass over every kernel family for `compute-sanitizer --tool memcheck` (GPU box):
device BVH build + flatten, trace batch, f64 / f32 render, tail kernel, animated builds, frame loop.
  compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_smoke.py"""
import sys
import tempfile

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np

from crucible_b200 import abi, demo_builder, gpu
from crucible_b200.gpu import GpuScene
from scenes_util import random_scene

d = random_scene(200, 3000, 50, seed=3)
host = GpuScene(d, 0, bvh_builder=abi.CR_BVH_HOST)
dev = GpuScene(d, 0, bvh_builder=abi.CR_BVH_DEVICE)
assert host.bvh_nodes().tobytes() == dev.bvh_nodes().tobytes()
for which in range(4):
    assert np.array_equal(host.device_records(which), dev.device_records(which))
rng = np.random.default_rng(1)
rays = np.concatenate([rng.normal(size=(5000, 3)) * 8, rng.normal(size=(5000, 3)), np.zeros((5000, 1))], 1)
a, b = host.trace_batch(rays), dev.trace_batch(rays)
assert np.array_equal(a["prim_index"], b["prim_index"]) and np.array_equal(a["t"], b["t"])
dev.trace_batch(rays, precision=abi.CR_PRECISION_F32)
sc = demo_builder.book1_end_scene(image_width=96, samples=3, seed=1)
gs = GpuScene(sc.describe(), 0)
cam = sc.scene_cam.to_abi()
r64, _, st = gs.render(cam, seed=1)
r32, _, _ = gs.render(cam, seed=1, precision=abi.CR_PRECISION_F32, pool_paths=4096)
assert abs(r64.mean() - r32.mean()) < 0.05
sc = demo_builder.cornell_box(image_width=48, samples=2)
GpuScene(sc.describe(), 0).render(sc.scene_cam.to_abi(), seed=1)
sc = demo_builder.book1_walkthrough(image_width=64, samples=2, duration=3 / 24.0)
with tempfile.TemporaryDirectory() as tmp:
    g = GpuScene(sc.describe(), 0)
    gpu.render_frames(g, sc.scene_cam.to_abi(), tmp, 3, 0, 1, seed=1, fmt=abi.CR_PPM_P6)
print("sanitize smoke ok:", st["rays"], "rays")
# round 2: every trace engine pinned in turn (order-free in shared / global memory, reference order), nested elements
import os

from scenes_util import nested_scene

sc = demo_builder.book1_end_scene(image_width=96, samples=3, seed=1)
gs, cam = GpuScene(sc.describe(), 0), sc.scene_cam.to_abi()
for env in ({"CRB_TRAVERSAL": "s"}, {"CRB_TRAVERSAL": "f"}, {"CRB_TRAVERSAL": "r", "CRB_MINB": "8"}, {"CRB_TRAVERSAL": "r", "CRB_MINB": "10"}):
    for k in ("CRB_TRAVERSAL", "CRB_MINB"):
        os.environ.pop(k, None)
    os.environ.update(env)
    img, _, st2 = gs.render(cam, seed=1)
    assert np.array_equal(img, r64), env
for k in ("CRB_TRAVERSAL", "CRB_MINB"):
    os.environ.pop(k, None)
n = GpuScene(nested_scene(1), 0)
n.trace_batch(rays)
n.render(cam, seed=1)
print("round-2 engines ok")
