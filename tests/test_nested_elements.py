"""nested elements.  Scene::add_element also takes a whole Hittables::HitList or Hittables::BVHWrapper as ONE
element (scene/mod.rs:160-166); the enclosing tree then meets it as a leaf whose hit() scans a list without box tests
(hitlist.rs:52-65) or walks a subtree of its own (bvhwrapper.rs:97-126).  The product flattens the whole tree of
Hittables into its preorder node array (box nodes + "always pass" leaf nodes); the oracle keeps the reference's
recursion.  Closest hits must agree bit for bit."""
import os
import tempfile

import numpy as np
import pytest

from crucible_b200 import abi, demo_builder
from crucible_b200.gpu import GpuScene
from crucible_b200.scene import BVHWrapper, Color, HitList, Lambertian, Metal, Point3, Scene, SolidColor, Sphere, Triangle
from conftest import random_rays
import golden
from scenes_util import compare_hits, nested_scene, random_scene, scene_bounds


def _dedupe(order):
    # the reference stores a span-1 element twice (bvhwrapper.rs:57-59): the flattened walk visits it once
    _, first = np.unique(np.asarray(order), return_index=True)
    return np.asarray(order)[np.sort(first)]


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_flattened_walk_visits_the_oracles_leaves_in_order(crlib, oracle, seed):
    d = nested_scene(seed)
    gs, orc = GpuScene(d, device=-1), oracle.OracleScene(d)
    got, exp = gs.bvh_leaf_order(), _dedupe(orc.bvh_leaf_order())
    hidden = set(d.hidden)
    exp = np.array([p for p in exp if p not in hidden])  # a hidden list member stays in the list but never hits
    assert np.array_equal(got, exp)
    assert gs.bvh_info()["n_visible"] == len(exp)


def test_oracle_nested_hit_equals_the_flat_list(oracle):
    """Property of the reference semantics: nesting changes the ORDER primitives are tested in, not the closest t."""
    d = nested_scene(5, hidden=False)
    orc = oracle.OracleScene(d)
    lo, hi = scene_bounds(d)
    rays = random_rays(20000, lo, hi, 11)
    a, b = orc.trace_batch(rays), orc.trace_batch(rays, brute=True)
    assert np.array_equal(a["prim_index"] >= 0, b["prim_index"] >= 0)
    hit = a["prim_index"] >= 0
    assert np.array_equal(a["t"][hit], b["t"][hit])


def test_oracle_reproduces_the_frozen_nested_fixture(oracle):
    """tests/golden/nested.npz (scripts/make_golden.py nested): an edit to the oracle's nested-element code cannot move the pin silently."""
    z = golden.load("nested")
    orc = oracle.OracleScene(nested_scene(1))
    assert np.array_equal(orc.bvh_leaf_order(), z["leaf_order"])
    got = orc.trace_batch(z["rays"])
    for f in ("prim_index", "obj_id", "front_face", "material", "t", "p", "n", "u", "v"):
        assert np.array_equal(got[f], z["hits"][f]), f


def test_group_call_order_errors(crlib):
    d = random_scene(4, 0, 0, 1)
    gs = GpuScene(d, device=-1)
    assert crlib.cr_scene_end_group(gs.handle) == abi.CR_ERR_STATE
    assert crlib.cr_scene_begin_group(gs.handle, 7) == abi.CR_ERR_INVALID
    assert crlib.cr_scene_begin_group(gs.handle, abi.CR_GROUP_HITLIST) == 0
    assert crlib.cr_scene_commit(gs.handle) == abi.CR_ERR_STATE  # still open
    assert crlib.cr_scene_end_group(gs.handle) == abi.CR_OK
    assert crlib.cr_scene_commit(gs.handle) == abi.CR_OK


def test_save_load_keeps_the_nesting(crlib):
    d = nested_scene(2)
    gs = GpuScene(d, device=-1)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "nested.crscene").encode()
        abi.check(crlib.cr_scene_save(gs.handle, path))
        h = crlib.cr_scene_load(path, -1)
        assert h
        try:
            n = crlib.cr_scene_bvh_leaf_order(h, None, 0)
            out = np.empty(n, np.int32)
            crlib.cr_scene_bvh_leaf_order(h, out.ctypes.data_as(abi.C.c_void_p), n)
            assert np.array_equal(out, gs.bvh_leaf_order())
        finally:
            crlib.cr_scene_destroy(h)


def test_scene_mirror_accepts_lists_and_wrappers(crlib, oracle):
    sc = Scene(16.0 / 9.0, 64, 24, 0.0)
    red = Lambertian(SolidColor(Color(0.7, 0.2, 0.2)), 1.0)
    sc.add_element(Sphere(Point3(0, -100.5, -1), 100.0, red), "ground")
    lst = HitList()
    lst.add(Sphere(Point3(0, 0, -1), 0.5, red))
    lst.add(Sphere(Point3(1, 0, -1), 0.5, Metal(Color(0.8, 0.8, 0.8), 0.0)))
    inner = HitList([Triangle(Point3(-2, 0, -1), Point3(-1, 0, -1), Point3(-1.5, 1, -1), red)])
    lst.add(BVHWrapper.new_wrapper(inner))
    sc.add_element(lst, "unused alias")
    sc.add_element(BVHWrapper.new_wrapper(HitList([Sphere(Point3(-1, 0, -1), 0.5, red)])), "unused alias")
    d = sc.describe()
    assert d.n_prims == 5
    gs, orc = GpuScene(d, device=-1), oracle.OracleScene(d)
    assert np.array_equal(gs.bvh_leaf_order(), _dedupe(orc.bvh_leaf_order()))
    with pytest.raises(TypeError):
        sc.add_element(42, "x")


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_nested_trace_is_bit_exact(gpu_device, oracle, seed):
    d = nested_scene(seed)
    gs, orc = GpuScene(d, gpu_device), oracle.OracleScene(d)
    lo, hi = scene_bounds(d)
    rays = random_rays(200000, lo, hi, 31 + seed)
    rays[:2000, 3:6] = np.array([1.0, 0.0, 0.0])  # irregular rays: the comparison-form box test at every node
    rays[2000:4000, 4] = 0.0
    exp = orc.trace_batch(rays)
    got = gs.trace_batch(rays)
    compare_hits(got, exp)
    if seed == 1:  # the frozen oracle records
        z = golden.load("nested")
        compare_hits(gs.trace_batch(z["rays"]), z["hits"])
    assert gs.last_retried() == 0  # scenes with nested elements stay on the reference-order engine
    assert (exp["prim_index"] >= 0).mean() > 0.05
    for p in d.hidden:
        assert not np.any(got["prim_index"] == p)
    got32 = gs.trace_batch(rays, precision=abi.CR_PRECISION_F32)
    assert (got32["prim_index"] == exp["prim_index"]).mean() > 0.995


@pytest.mark.gpu
def test_nested_render_matches_oracle(gpu_device, oracle):
    d = nested_scene(4)
    cam = demo_builder.book1_end_scene(image_width=128, samples=8).scene_cam
    cam.look_from, cam.look_at = np.array([14.0, 6.0, 18.0]), np.array([0.0, 0.0, 0.0])
    cam = cam.to_abi()
    gs, orc = GpuScene(d, gpu_device), oracle.OracleScene(d)
    rgb, rgb8, st = gs.render(cam, seed=2)
    ref, ref8, ost = orc.render(cam, seed=2)
    assert st["rays"] == ost["rays"] and st["trace_engine"] in (0, 1)
    assert np.abs(rgb - ref).max() <= 1e-11
    rep = gs.replicate(gpu_device) if hasattr(gs, "replicate") else None
    if rep is not None:
        rgb2, _, _ = rep.render(cam, seed=2)
        assert np.array_equal(rgb2, rgb)
