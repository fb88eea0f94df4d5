"""Small seeded scenes shared by the parity tests."""
import numpy as np

from crucible_b200 import abi
from crucible_b200.scene import SceneDesc


def _mats(kinds):
    """One material per entry of `kinds` + the textures they need (solid, checker, image)."""
    texs, mats = [], []
    t = abi.CrTexture(); t.kind = abi.CR_TEX_SOLID; t.color[:] = (0.8, 0.3, 0.2); texs.append(t)
    t = abi.CrTexture(); t.kind = abi.CR_TEX_SOLID; t.color[:] = (0.1, 0.2, 0.9); texs.append(t)
    t = abi.CrTexture(); t.kind = abi.CR_TEX_CHECKER; t.even, t.odd, t.inv_scale = 0, 1, 1.0 / 0.32; texs.append(t)
    for k in kinds:
        m = abi.CrMaterial()
        m.kind = k
        m.scatter_prob = 1.0
        if k == abi.CR_MAT_LAMBERTIAN:
            m.tex = 2
        elif k == abi.CR_MAT_METAL:
            m.albedo[:] = (0.7, 0.6, 0.5)
            m.fuzz = 0.1
        elif k == abi.CR_MAT_DIELECTRIC:
            m.ior = 1.5
        else:
            m.emit[:] = (4.0, 4.0, 4.0)
        mats.append(m)
    return mats, texs


def random_scene(n_sph=0, n_tri=0, n_quad=0, seed=0, extent=10.0, kinds=(0, 1, 2), interleave=True):
    """Random soup of spheres / triangles / quads; batches interleaved so insertion order mixes kinds."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    d = SceneDesc()
    d.materials, d.textures = _mats(kinds)
    nm = len(d.materials)
    batches = []
    if n_sph:
        c = (rng.random((n_sph, 3)) * 2 - 1) * extent
        r = rng.random((n_sph, 1)) * 0.6 + 0.05
        batches.append((abi.CR_PRIM_SPHERE, np.concatenate([c, r], 1)))
    if n_tri:
        a = (rng.random((n_tri, 3)) * 2 - 1) * extent
        b = a + (rng.random((n_tri, 3)) * 2 - 1) * 1.5
        c = a + (rng.random((n_tri, 3)) * 2 - 1) * 1.5
        batches.append((abi.CR_PRIM_TRIANGLE, np.concatenate([a, b, c], 1)))
    if n_quad:
        q = (rng.random((n_quad, 3)) * 2 - 1) * extent
        u = (rng.random((n_quad, 3)) * 2 - 1) * 1.5
        v = (rng.random((n_quad, 3)) * 2 - 1) * 1.5
        batches.append((abi.CR_PRIM_QUAD, np.concatenate([q, u, v], 1)))
    base = 0
    if interleave and len(batches) > 1:
        # split every batch in chunks and round-robin them, like alternating add_element calls
        chunks = []
        for kind, data in batches:
            for part in np.array_split(data, 4):
                if len(part):
                    chunks.append((kind, part))
        order = rng.permutation(len(chunks))
        batches = [chunks[i] for i in order]
    for kind, data in batches:
        n = len(data)
        mat = rng.integers(0, nm, n).astype(np.int32)
        oid = (base + np.arange(n)).astype(np.int32)
        d.batches.append((kind, data, mat, oid))
        base += n
    return d


def nested_scene(seed, n=160, hidden=True):
    """Top-level primitives mixed with: a HitList of spheres and triangles (one member hidden), a BVHWrapper of 40
    primitives (one hidden), a HitList that holds a BVHWrapper and a HitList, an empty HitList and a single-member wrapper."""
    base = random_scene(n, n // 2, 12, seed)
    prims = []  # (kind, row, mat, oid) in insertion order
    for kind, data, mat, oid in base.batches:
        for i in range(len(data)):
            prims.append((kind, data[i], int(mat[i]), int(oid[i])))
    rng = np.random.Generator(np.random.Philox(key=seed + 1000))
    rng.shuffle(prims)
    d = SceneDesc()
    d.materials, d.textures = base.materials, base.textures
    it = iter(prims)

    def take(k):
        for _ in range(k):
            kind, row, mat, oid = next(it)
            d.batches.append((kind, row[None, :], np.array([mat], np.int32), np.array([oid], np.int32)))

    take(30)
    d.begin_group(abi.CR_GROUP_HITLIST)
    take(9)
    d.end_group()
    take(25)
    d.begin_group(abi.CR_GROUP_BVH)
    take(40)
    d.end_group()
    d.begin_group(abi.CR_GROUP_HITLIST)
    take(3)
    d.begin_group(abi.CR_GROUP_BVH)
    take(17)
    d.end_group()
    take(1)
    d.begin_group(abi.CR_GROUP_HITLIST)
    take(4)
    d.end_group()
    d.end_group()
    d.begin_group(abi.CR_GROUP_HITLIST)  # empty list: the empty box, never hit
    d.end_group()
    d.begin_group(abi.CR_GROUP_BVH)
    take(1)
    d.end_group()
    take(len(prims) - 30 - 9 - 25 - 40 - 3 - 17 - 1 - 4 - 1)
    if hidden:
        d.hidden = [3, 33, 70, 110, 131]  # top level, inside the first list, inside the wrapper, nested wrapper, inner list
    return d


def scene_bounds(desc):
    lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
    for kind, data, _, _ in desc.batches:
        if kind < 0:  # group markers
            continue
        if kind == abi.CR_PRIM_SPHERE:
            pts = [data[:, :3] - data[:, 3:4], data[:, :3] + data[:, 3:4]]
        elif kind == abi.CR_PRIM_TRIANGLE:
            pts = [data[:, 0:3], data[:, 3:6], data[:, 6:9]]
        else:
            pts = [data[:, 0:3], data[:, 0:3] + data[:, 3:6], data[:, 0:3] + data[:, 6:9], data[:, 0:3] + data[:, 3:6] + data[:, 6:9]]
        for p in pts:
            p = p[np.all(np.abs(p) < 100, axis=1)]  # ignore the r=1000 ground sphere for ray generation
            if len(p):
                lo, hi = np.minimum(lo, p.min(0)), np.maximum(hi, p.max(0))
    return lo, hi


def compare_hits(got, exp, rtol=1e-5, exact_t=True):
    """North-star gate: prim ids + front_face bit-exact; t, normal, uv within 1e-5 relative."""
    assert np.array_equal(got["prim_index"], exp["prim_index"]), \
        f"{np.count_nonzero(got['prim_index'] != exp['prim_index'])} prim_index mismatches of {len(exp)}"
    assert np.array_equal(got["obj_id"], exp["obj_id"])
    assert np.array_equal(got["front_face"], exp["front_face"])
    assert np.array_equal(got["material"], exp["material"])
    hit = exp["prim_index"] >= 0
    if exact_t:
        # same IEEE operations in the same order: t, p and the normal are BIT-identical
        assert np.array_equal(got["t"][hit], exp["t"][hit])
        assert np.array_equal(got["p"][hit], exp["p"][hit])
        assert np.array_equal(got["n"][hit], exp["n"][hit])
    t = exp["t"][hit]
    assert np.all(np.abs(got["t"][hit] - t) <= rtol * np.abs(t))
    assert np.all(np.linalg.norm(got["n"][hit] - exp["n"][hit], axis=1) <= rtol)
    # u, v go through acos/atan2 (libm vs CUDA: a few ulp)
    assert np.all(np.abs(got["u"][hit] - exp["u"][hit]) <= rtol)
    assert np.all(np.abs(got["v"][hit] - exp["v"][hit]) <= rtol)


def compare_both_engines(gs, exp, rays, tmin=0.001, tmax=float("inf"), max_retried=None):
    """The same batch through the order-free engine (default) and through the reference-order kernel
    (CR_TRACE_REFERENCE_ORDER): both must reproduce the oracle's records.  Returns (hits of the default engine, number of
    rays the order-free engine handed back to the reference-order kernel)."""
    got = gs.trace_batch(rays, tmin, tmax)
    retried = gs.last_retried()
    compare_hits(got, exp)
    compare_hits(gs.trace_batch(rays, tmin, tmax, reference_order=True), exp)
    if max_retried is not None:
        assert retried <= max_retried, f"{retried} rays went back to the reference-order kernel"
    return got, retried
