"""Frozen oracle outputs (tests/golden/*.npz, written by scripts/make_golden.py FROM THE ORACLE at a known state).

The reference has no golden vector for hit(), the BVH, scatter or rendering, and no Rust toolchain exists
here, so the restated oracle is the pin.  These fixtures freeze that pin: per config a strided subset of each seeded
SURVEY 8d ray batch (rays + the oracle's full CrHit records) and one small f64 render.  `tests/test_golden.py` checks
that the oracle still reproduces them (CPU); the GPU parity tests check the CUDA path against the same records, so an
edit to oracle.cpp can no longer move the target silently."""
import os

import numpy as np
from conftest import random_rays
from scenes_util import scene_bounds

from crucible_b200 import demo_builder

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N_KEEP = 2048
N_BATCH = 1 << 20

# name -> (builder name, kwargs): the BASELINE configs at their own resolution; config 4 with 64 meshes on an 8 x 8 grid instead of 1582 on 40 x 40 so the
# CPU suite stays light (the full-size scene is compared with the live oracle on the GPU box, test_full_size_parity.py)
CONFIGS = {
    "book1": ("book1", dict(image_width=1920, samples=4)),
    "teapot": ("teapot", dict(image_width=1920, samples=4)),
    "cornell": ("cornell", dict(image_width=1024, samples=4)),
    "instanced64": ("instanced", dict(image_width=3840, samples=4, copies=64, grid=8, spacing=5.0)),
}
RENDER = dict(image_width=64, samples=4, seed=1)


def build(name):
    builder, kw = CONFIGS[name]
    return demo_builder.CONFIGS[builder](**kw)


def small_camera(name):
    """The same scene's camera at 64 px, 4 spp (the frozen render)."""
    builder, kw = CONFIGS[name]
    kw = dict(kw, image_width=RENDER["image_width"], samples=RENDER["samples"])
    return demo_builder.CONFIGS[builder](**kw).scene_cam.to_abi()


def keep_indices(n):
    return np.unique(np.linspace(0, n - 1, N_KEEP).astype(np.int64))


def batches(desc, cam, orc):
    lo, hi = scene_bounds(desc)
    wh = cam.image_width * cam.image_height
    return {"primary": orc.gen_rays(cam, 0, wh), "first_bounce": orc.gen_rays(cam, 3, N_BATCH, seed=7),
            "random": random_rays(N_BATCH, lo, hi, 42)}


def path(name):
    return os.path.join(GOLDEN_DIR, name + ".npz")


def load(name):
    return np.load(path(name))


def check_prefix(name, bname, rays, got, uv_tol=1e-5):
    """`got` (CrHit records of the FULL seeded batch `rays`, from the CUDA path) against the frozen oracle records."""
    z = load(name)
    idx = z[bname + "_idx"]
    assert np.array_equal(rays[idx], z[bname + "_rays"]), f"{name}/{bname}: the seeded ray batch itself changed"
    exp = z[bname + "_hits"]
    g = got[idx]
    for f in ("prim_index", "obj_id", "front_face", "material"):
        assert np.array_equal(g[f], exp[f]), (name, bname, f)
    hit = exp["prim_index"] >= 0
    for f in ("t", "p", "n"):  # same IEEE operations in the same order: bit-identical
        assert np.array_equal(g[f][hit], exp[f][hit]), (name, bname, f)
    for f in ("u", "v"):  # acos / atan2: CUDA libdevice vs glibc, a few ulp
        assert np.all(np.abs(g[f][hit] - exp[f][hit]) <= uv_tol), (name, bname, f)
