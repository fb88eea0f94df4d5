"""SURVEY 8 f3 — the device BVH builder must produce the SAME tree as the recursion of
BVHWrapper::help_generate (bvhwrapper.rs:46-94), which the host builder runs as written and which
tests/test_abi_and_host.py pins against the oracle's independent build.

CPU tests: structure of the node export, builder selection and its error behaviour.
GPU tests: host-built and device-built node arrays are equal BYTE FOR BYTE (boxes including the sign of a
zero, children, axes, skip links) over ragged sizes, ties, duplicates, hidden primitives and the demo scenes;
closest-hit ids of a device-built scene equal the oracle's.
"""
import numpy as np
import pytest
from conftest import random_rays
from scenes_util import compare_hits, random_scene, scene_bounds

from crucible_b200 import abi, demo_builder
from crucible_b200.gpu import GpuScene
from crucible_b200.scene import SceneDesc

REF_LEAF, REF_NONE = 0x80000000, 0x7FFFFFFF


def node_count(span):  # nodes help_generate creates for a span (bvhwrapper.rs:46-80)
    return 1 if span <= 2 else 1 + node_count(span // 2) + node_count(span - span // 2)


def check_structure(nodes, n_visible):
    """Preorder invariants of the exported tree."""
    assert len(nodes) == (node_count(n_visible) if n_visible else 0)
    if not len(nodes):
        return
    spans = np.zeros(len(nodes), np.int64)

    def walk(i, span):
        spans[i] = span
        n = nodes[i]
        assert n["skip"] == i + node_count(span)
        assert n["axis"] in (0, 1, 2)
        if span <= 2:
            assert n["left"] & REF_LEAF
            assert (n["right"] == REF_NONE) if span == 1 else bool(n["right"] & REF_LEAF)
            return
        assert n["left"] == i + 1 and n["right"] == i + 1 + node_count(span // 2)
        for c in (n["left"], n["right"]):
            if c != 0:  # (the root box is re-derived from its children, bvhwrapper.rs:34-44)
                assert np.all(nodes[c]["lo"] >= n["lo"]) and np.all(nodes[c]["hi"] <= n["hi"])
        walk(int(n["left"]), span // 2)
        walk(int(n["right"]), span - span // 2)

    walk(0, n_visible)
    # longest axis of the stored box (bvh.rs:82-94); the root keeps the axis of the fold, its box is re-derived
    ext = nodes["hi"] - nodes["lo"]
    sx, sy, sz = ext[:, 0], ext[:, 1], ext[:, 2]
    axis = np.where(sx > sy, np.where(sx > sz, 0, 2), np.where(sy > sz, 1, 2))
    assert np.array_equal(axis[1:], nodes["axis"][1:])


@pytest.mark.parametrize("n_sph,n_tri,n_quad,seed", [(1, 0, 0, 1), (2, 0, 0, 2), (3, 0, 0, 3), (0, 7, 0, 4), (100, 300, 40, 6), (0, 1000, 0, 5)])
def test_node_export_is_a_preorder_tree(crlib, n_sph, n_tri, n_quad, seed):
    d = random_scene(n_sph, n_tri, n_quad, seed)
    gs = GpuScene(d, device=-1, bvh_builder=abi.CR_BVH_HOST)
    nodes = gs.bvh_nodes()
    check_structure(nodes, gs.bvh_info()["n_visible"])
    assert gs.commit_info()["builder"] == abi.CR_BVH_HOST


def test_builder_selection_without_a_device(crlib):
    d = random_scene(10, 0, 0, 1)
    # AUTO on a host-only scene uses the host builder; DEVICE has nothing to run on and says so
    assert GpuScene(d, device=-1, bvh_builder=abi.CR_BVH_AUTO).commit_info()["builder"] == abi.CR_BVH_HOST
    with pytest.raises(abi.CrucibleError) as e:
        GpuScene(d, device=-1, bvh_builder=abi.CR_BVH_DEVICE)
    assert e.value.code == abi.CR_ERR_NO_DEVICE
    lib = abi.load()
    h = lib.cr_scene_create(-1)
    assert lib.cr_scene_set_bvh_builder(h, 3) == abi.CR_ERR_INVALID
    assert lib.cr_scene_set_bvh_builder(h, -1) == abi.CR_ERR_INVALID
    lib.cr_scene_destroy(h)


# ---- GPU: the two builders agree byte for byte ----------------------------------------------------------------

def both(desc, device):
    host = GpuScene(desc, device, bvh_builder=abi.CR_BVH_HOST)
    dev = GpuScene(desc, device, bvh_builder=abi.CR_BVH_DEVICE)
    assert host.commit_info()["builder"] == abi.CR_BVH_HOST and dev.commit_info()["builder"] == abi.CR_BVH_DEVICE
    return host, dev


def assert_same_tree(desc, device):
    host, dev = both(desc, device)
    assert host.bvh_info() == dev.bvh_info()
    a, b = host.bvh_nodes(), dev.bvh_nodes()
    assert a.shape == b.shape
    if a.tobytes() != b.tobytes():
        for f in ("left", "right", "axis", "skip"):
            bad = np.flatnonzero(a[f] != b[f])
            assert not len(bad), f"{len(bad)} nodes differ in {f}, first at {bad[0]}: host {a[f][bad[0]]} device {b[f][bad[0]]}"
        for f in ("lo", "hi"):
            bad = np.flatnonzero((a[f].view(np.uint64) != b[f].view(np.uint64)).any(axis=1))
            assert not len(bad), f"{len(bad)} boxes differ in {f}, first at {bad[0]}: host {a[f][bad[0]]} device {b[f][bad[0]]}"
        raise AssertionError("node arrays differ in padding only?")
    assert np.array_equal(host.bvh_leaf_order(), dev.bvh_leaf_order())
    # what the trace kernels read: nodes (f64, outward-rounded f32, BIGBOX bits) and triangle records, flattened by
    # the host loop for the host builder and by kernels for the device builder
    for which in range(4):
        assert np.array_equal(host.device_records(which), dev.device_records(which)), f"device records {which} differ"
    host.close()
    dev.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n_sph,n_tri,n_quad,seed", [(1, 0, 0, 1), (2, 0, 0, 2), (3, 0, 0, 3), (4, 0, 0, 4), (5, 0, 0, 5), (0, 7, 0, 6),
                                                     (485, 0, 0, 7), (100, 300, 40, 8), (0, 1023, 0, 9), (0, 1024, 0, 10), (0, 1025, 0, 11),
                                                     (0, 2049, 0, 12), (700, 3000, 397, 13), (0, 40001, 0, 14)])
def test_device_build_equals_host_build(gpu_device, n_sph, n_tri, n_quad, seed):
    assert_same_tree(random_scene(n_sph, n_tri, n_quad, seed), gpu_device)


@pytest.mark.gpu
def test_device_build_demo_scenes_and_hidden(gpu_device):
    for sc in (demo_builder.book1_end_scene(seed=3), demo_builder.load_teapot(sky=None), demo_builder.cornell_box()):
        assert_same_tree(sc.describe(), gpu_device)
    sc = demo_builder.book1_end_scene(seed=3)
    sc.hide_element("large_metal")
    sc.hide_element("small17")
    assert_same_tree(sc.describe(), gpu_device)


def _tie_scene(seed, n=6000):
    """Coordinates drawn from a handful of values: every level sorts long runs of EQUAL keys, so the tree depends
    on the stability of every sort (box_compare returns Equal, the span keeps its current order), and on
    duplicates (identical primitives)."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    d = random_scene(0, 0, 0, seed)
    d.materials, d.textures = random_scene(1, 0, 0, seed).materials, random_scene(1, 0, 0, seed).textures
    grid = np.array([-2.0, -1.0, 0.0, 1.0, 2.0])
    c = grid[rng.integers(0, 5, (n, 3))]
    r = np.array([0.25, 0.5])[rng.integers(0, 2, (n, 1))]
    d.batches.append((abi.CR_PRIM_SPHERE, np.concatenate([c, r], 1), np.zeros(n, np.int32), np.arange(n, dtype=np.int32)))
    return d


@pytest.mark.gpu
def test_device_build_is_stable_under_ties(gpu_device):
    for seed in (1, 2):
        assert_same_tree(_tie_scene(seed), gpu_device)


@pytest.mark.gpu
def test_device_build_keeps_the_sign_of_zero(gpu_device):
    """Interval::tight_enclose (utils.rs:631-635) folds with <= / >=, so among -0.0 and +0.0 the FIRST one in span
    order survives; f64::partial_cmp calls them Equal, so the sort must not reorder them either."""
    rng = np.random.Generator(np.random.Philox(key=5))
    n = 5000
    d = _tie_scene(3, n=1)
    d.batches.clear()
    tri = (rng.random((n, 9)) * 2 - 1) * 4.0
    z = np.where(rng.random(n) < 0.5, -0.0, 0.0)
    for col in (0, 3, 6):  # a.x = b.x = c.x = +-0: the triangle's box has lo.x = hi.x = that zero
        tri[:, col] = z
    ys = np.where(rng.random(n) < 0.5, -0.0, 0.0)
    tri[: n // 2, 1] = ys[: n // 2]  # half of the triangles touch y = +-0 with one vertex
    d.batches.append((abi.CR_PRIM_TRIANGLE, tri, np.zeros(n, np.int32), np.arange(n, dtype=np.int32)))
    host, dev = both(d, gpu_device)
    a = host.bvh_nodes()
    assert np.signbit(a["lo"][:, 0]).any() and (~np.signbit(a["lo"][:, 0])).any(), "the scene must exercise both zeros"
    host.close()
    dev.close()
    assert_same_tree(d, gpu_device)


@pytest.mark.gpu
def test_auto_builder_threshold(gpu_device):
    small = GpuScene(random_scene(0, 1000, 0, 1), gpu_device)
    assert small.commit_info()["builder"] == abi.CR_BVH_HOST
    big_desc = random_scene(0, 33000, 0, 2)
    big = GpuScene(big_desc, gpu_device)
    ci = big.commit_info()
    assert ci["builder"] == abi.CR_BVH_DEVICE and ci["levels"] == big.bvh_info()["max_depth"] and ci["ms_device"] > 0
    # non-finite coordinates never reach a builder (add_prims rejects them), so box_compare's NaN -> Equal branch
    # (bvhwrapper.rs:92) cannot make the two builders disagree
    kind, data, mat, oid = big_desc.batches[0]
    data = data.copy()
    data[17, 0] = np.nan
    big_desc.batches[0] = (kind, data, mat, oid)
    with pytest.raises(abi.CrucibleError) as e:
        GpuScene(big_desc, gpu_device, bvh_builder=abi.CR_BVH_DEVICE)
    assert e.value.code == abi.CR_ERR_INVALID


@pytest.mark.gpu
def test_instanced_mesh_device_build_and_trace(gpu_device, oracle):
    """Config 4's structure at 1/50 size: teapot copies on a grid (long runs of equal y keys), earth spheres, ground."""
    sc = demo_builder.instanced_teapots(copies=30, grid=6)
    d = sc.describe()
    assert_same_tree(d, gpu_device)
    gs = GpuScene(d, gpu_device)  # AUTO: 189 k triangles -> device build
    assert gs.commit_info()["builder"] == abi.CR_BVH_DEVICE
    orc = oracle.OracleScene(d)
    lo, hi = scene_bounds(d)
    rays = random_rays(20000, lo, hi, 42)
    compare_hits(gs.trace_batch(rays), orc.trace_batch(rays))


def test_device_records_need_a_device(crlib):
    gs = GpuScene(random_scene(5, 5, 0, 1), device=-1)
    with pytest.raises(abi.CrucibleError) as e:
        gs.device_records(0)
    assert e.value.code == abi.CR_ERR_NO_DEVICE
    ci = gs.commit_info()
    assert ci["builder"] == abi.CR_BVH_HOST and ci["levels"] == gs.bvh_info()["max_depth"] and ci["ms_upload"] == 0.0
    assert ci["ms_total"] >= ci["ms_build"] >= 0.0


@pytest.mark.gpu
def test_config4_full_size_builders_agree(gpu_device):
    """BASELINE config 4 at FULL size (1582 teapots = 9 998 240 triangles + 65 spheres, 11.6 M nodes, depth 24): the
    device-built tree and the flattened records equal the host recursion's byte for byte, and the commit is faster."""
    d = demo_builder.instanced_teapots(copies=1582, grid=40).describe()
    dev = GpuScene(d, gpu_device, bvh_builder=abi.CR_BVH_DEVICE)
    info = dev.bvh_info()
    assert info == {"n_nodes": 11608001, "max_depth": 24, "n_visible": 9998305}
    nodes_dev = dev.bvh_nodes()
    recs_dev = [dev.device_records(w) for w in (0, 1)]
    t_dev = dev.commit_info()["ms_total"]
    dev.close()
    host = GpuScene(d, gpu_device, bvh_builder=abi.CR_BVH_HOST)
    assert host.bvh_info() == info
    assert np.array_equal(host.bvh_nodes().view(np.uint8), nodes_dev.view(np.uint8))
    for w in (0, 1):
        assert np.array_equal(host.device_records(w), recs_dev[w])
    assert t_dev < host.commit_info()["ms_total"]
    # preorder invariants, vectorised: skip links bound every subtree, inner nodes point at i + 1
    inner = (nodes_dev["left"] & REF_LEAF) == 0
    idx = np.arange(len(nodes_dev), dtype=np.int64)
    assert np.all(nodes_dev["left"][inner] == idx[inner] + 1)
    assert np.all(nodes_dev["skip"][~inner] == idx[~inner] + 1)
    assert np.all(nodes_dev["skip"][inner] > nodes_dev["right"][inner]) and nodes_dev["skip"][0] == len(nodes_dev)
    host.close()
