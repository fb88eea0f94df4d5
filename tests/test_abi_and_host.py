"""CPU-side checks of the boundary: the library loads and exports every symbol the header declares,
the ctypes mirror has the header's layout, error behaviour without a GPU, and the host BVH build
(BVHWrapper::new_wrapper, bvhwrapper.rs:15-94) matches the oracle's independent restatement."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest
from scenes_util import random_scene

from crucible_b200 import abi, demo_builder
from crucible_b200.gpu import GpuScene, rows_of_rank
from crucible_b200.scene import SceneDesc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "crucible_gpu.h")


def test_library_exports_every_declared_symbol(crlib):
    text = open(HEADER).read()
    declared = set(re.findall(r"\b(cr_[a-z0-9_]+)\s*\(", text))
    assert declared == set(abi.SIGNATURES), declared ^ set(abi.SIGNATURES)
    for name in declared:
        assert hasattr(crlib, name), f"libcrucible_b200.so does not export {name}"
    assert b"sm_100a" in crlib.cr_version()


def test_integration_doc_binds_every_entry_point():
    """INTEGRATION.md shows the reference-side binding (the Rust `extern "C"` block): it names exactly the header's entry points."""
    declared = set(re.findall(r"\b(cr_[a-z0-9_]+)\s*\(", open(HEADER).read()))
    bound = set(re.findall(r"pub fn (cr_[a-z0-9_]+)\(", open(os.path.join(ROOT, "INTEGRATION.md")).read()))
    assert bound == declared, bound ^ declared


def test_ctypes_layout_matches_header():
    structs = {"CrMaterial": abi.CrMaterial, "CrTexture": abi.CrTexture, "CrKeyframe": abi.CrKeyframe, "CrCamera": abi.CrCamera,
               "CrRenderOpts": abi.CrRenderOpts, "CrStats": abi.CrStats, "CrHit": abi.CrHit}
    lines = []
    for name, cls in structs.items():
        lines.append(f'printf("{name} %zu\\n", sizeof({name}));')
        for f, _ in cls._fields_:
            lines.append(f'printf("{name}.{f} %zu\\n", offsetof({name}, {f}));')
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "crucible_gpu.h"\nint main(void){' + "".join(lines) + "return 0;}"
    with tempfile.TemporaryDirectory() as td:
        c, exe = os.path.join(td, "l.c"), os.path.join(td, "l")
        open(c, "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    got = dict(l.split() for l in out.strip().splitlines())
    for name, cls in structs.items():
        assert int(got[name]) == C.sizeof(cls), name
        for f, _ in cls._fields_:
            assert int(got[f"{name}.{f}"]) == getattr(cls, f).offset, f"{name}.{f}"
    assert np.dtype(abi.HIT_DTYPE).itemsize == C.sizeof(abi.CrHit)


def test_no_cpu_fallback(crlib):
    """Without an sm_100 device every compute entry point fails loudly with CR_ERR_NO_DEVICE."""
    if crlib.cr_device_count() > 0:
        pytest.skip("a GPU is present")
    d = random_scene(10, 0, 0, 1)
    gs = GpuScene(d, device=-1)  # host-only scene: build + introspection work
    assert gs.bvh_info()["n_visible"] == 10
    with pytest.raises(abi.CrucibleError) as e:
        gs.trace_batch(np.zeros((1, 7)))
    assert e.value.code == abi.CR_ERR_NO_DEVICE and "no CPU fallback" in str(e.value)
    cam = demo_builder.book1_end_scene(image_width=16, samples=1).scene_cam.to_abi()
    with pytest.raises(abi.CrucibleError) as e:
        gs.render(cam)
    assert e.value.code == abi.CR_ERR_NO_DEVICE
    with pytest.raises(abi.CrucibleError):
        GpuScene(d, device=0)


def test_error_behaviour_mirrors_reference_panics(crlib):
    d = SceneDesc()
    d.materials = [abi.CrMaterial(kind=abi.CR_MAT_METAL)]
    d.batches = [(abi.CR_PRIM_SPHERE, np.array([[0, 0, 0, -1.0]]), np.zeros(1, np.int32), np.zeros(1, np.int32))]
    with pytest.raises(abi.CrucibleError, match="negative radius"):  # sphere.rs:26
        GpuScene(d, device=-1)
    d = random_scene(3, 0, 0, 1)
    d.materials[1].fuzz = 1.5  # metal.rs:21-25
    with pytest.raises(abi.CrucibleError, match="fuzz"):
        GpuScene(d, device=-1)
    d = random_scene(3, 0, 0, 1)
    d.batches[0][2][:] = 99
    with pytest.raises(abi.CrucibleError, match="material"):
        GpuScene(d, device=-1)
    d = random_scene(3, 0, 0, 1)
    d.textures[2].even = 2  # checker that contains itself: would recurse forever in the reference
    with pytest.raises(abi.CrucibleError, match="nested"):
        GpuScene(d, device=-1)


@pytest.mark.parametrize("n_sph,n_tri,n_quad,seed", [(1, 0, 0, 1), (2, 0, 0, 2), (3, 0, 0, 3), (485, 0, 0, 4), (0, 1000, 0, 5),
                                                     (100, 300, 40, 6), (5, 5, 5, 7)])
def test_host_bvh_equals_oracle_bvh(crlib, oracle, n_sph, n_tri, n_quad, seed):
    d = random_scene(n_sph, n_tri, n_quad, seed)
    gs, o = GpuScene(d, device=-1), oracle.OracleScene(d)
    assert gs.bvh_info() == o.bvh_info()
    assert np.array_equal(gs.bvh_leaf_order(), o.bvh_leaf_order())


def test_host_bvh_book1_teapot_and_hidden(crlib, oracle):
    for sc in (demo_builder.book1_end_scene(seed=3), demo_builder.load_teapot(sky=None), demo_builder.cornell_box()):
        d = sc.describe()
        gs, o = GpuScene(d, device=-1), oracle.OracleScene(d)
        assert gs.bvh_info() == o.bvh_info()
        assert np.array_equal(gs.bvh_leaf_order(), o.bvh_leaf_order())
    sc = demo_builder.book1_end_scene(seed=3)
    # independent anchor: bvhwrapper.rs:57-74 makes one node per span, spans of 1 or 2 are leaves, larger spans split
    # at span / 2  =>  nodes(n) = 1 + nodes(n / 2) + nodes(n - n / 2), depth(n) = 1 + depth(n - n / 2)
    def shape(n):
        if n <= 2:
            return 1, 1
        (a, da), (b, db) = shape(n // 2), shape(n - n // 2)
        return 1 + a + b, 1 + max(da, db)

    assert shape(485) == (511, 9)
    assert GpuScene(sc.describe(), device=-1).bvh_info() == {"n_nodes": 511, "max_depth": 9, "n_visible": 485}
    sc.hide_element("large_metal")
    sc.hide_element("small17")
    d = sc.describe()
    assert len(d.hidden) == 2
    gs, o = GpuScene(d, device=-1), oracle.OracleScene(d)
    assert gs.bvh_info() == o.bvh_info() and gs.bvh_info()["n_visible"] == d.n_prims - 2
    assert np.array_equal(gs.bvh_leaf_order(), o.bvh_leaf_order())
    # teapot: 6320 triangles + ground -> 8191 nodes, depth 13
    info = GpuScene(demo_builder.load_teapot(sky=None).describe(), device=-1).bvh_info()
    assert info["n_nodes"] == 8191 and info["max_depth"] == 13 and info["n_visible"] == 6321


def test_empty_scene_commits(crlib):
    gs = GpuScene(SceneDesc(), device=-1)
    assert gs.bvh_info() == {"n_nodes": 0, "max_depth": 0, "n_visible": 0}


def test_row_sharding_is_a_partition():
    for H, block, world in ((1080, 8, 8), (225, 8, 2), (17, 4, 4), (5, 8, 3), (1080, 16, 1)):
        rows = [rows_of_rank(H, block, r, world) for r in range(world)]
        allr = np.sort(np.concatenate(rows))
        assert np.array_equal(allr, np.arange(H))
        for r in rows:
            assert np.all(np.diff(r) > 0)


def test_scene_mirror_matches_reference_api():
    """Scene::add_element alias collisions panic (scene/mod.rs:170-176); set_samples(0) panics
    (camera/mod.rs:233-240); image height = (w / aspect) as u32 >= 1 (camera/mod.rs:36-47)."""
    from crucible_b200.scene import Color, Dielectric, Metal, Point3, Scene, Sphere

    sc = Scene.new_image(16.0 / 9.0, 400, 24, 180.0, 1)
    assert (sc.scene_cam.image_width, sc.scene_cam.image_height) == (400, 225)
    assert Scene.new_image(16.0 / 9.0, 1920, 24, 180.0, 1).scene_cam.image_height == 1080
    assert Scene.new_image(100.0, 10, 24, 180.0, 1).scene_cam.image_height == 1
    sc.add_element(Sphere(Point3(0, 0, 0), 1.0, Dielectric(1.5)), "a")
    with pytest.raises(ValueError, match="collides"):
        sc.add_element(Sphere(Point3(0, 0, 0), 1.0, Dielectric(1.5)), "a")
    with pytest.raises(ValueError):
        sc.scene_cam.set_samples(0)
    with pytest.raises(ValueError):
        Metal(Color(0.5, 0.5, 0.5), 1.5)
    with pytest.raises(ValueError):
        Sphere(Point3(0, 0, 0), -1.0, Dielectric(1.5))
    sc.scene_cam.set_vfov(20.0)
    sc.scene_cam.set_focus_dist(10.0)
    c = sc.scene_cam.to_abi()
    import math
    assert c.viewport_height == 2.0 * math.tan((20.0 * math.pi / 180.0) / 2.0) * 10.0
    assert c.viewport_width == c.viewport_height * (400 / 225)
    mv = demo_builder.book1_walkthrough(image_width=64, samples=1)
    assert mv.compute_frame_count() == 240 and mv.scene_cam.to_abi().n_from_keys == 12


def test_c_example_links_against_the_abi(crlib):
    """examples/c_abi.c uses the boundary from plain C; without a GPU it reports the missing device and
    exits 0 (no CPU fallback), with one it traces and renders."""
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "c_abi")
        lib_dir = os.path.join(ROOT, "crucible_b200")
        subprocess.run(["gcc", os.path.join(ROOT, "examples", "c_abi.c"), "-I", os.path.join(ROOT, "include"), "-L", lib_dir,
                        "-lcrucible_b200", "-lm", f"-Wl,-rpath,{lib_dir}", "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "BVH: 1 nodes, depth 1, 2 primitives" in out
    assert ("no CPU fallback" in out) or ("hit prim 1 at t = 0.5" in out)


def _build_cpp_example(td):
    exe = os.path.join(td, "render_book1")
    lib_dir = os.path.join(ROOT, "crucible_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "examples", "render_book1.cpp"), "-I", os.path.join(ROOT, "include"),
                    "-L", lib_dir, "-lcrucible_b200", f"-Wl,-rpath,{lib_dir}", "-o", exe], check=True)
    return exe


def test_cpp_host_mirror_builds_the_reference_tree(crlib):
    """include/crucible.hpp: the C++ mirror of Scene / Camera / demo_images::book1_end_scene flattens into the
    C ABI; 22x22 grid minus the exclusion zone + 4 spheres -> 511 nodes, depth 9."""
    with tempfile.TemporaryDirectory() as td:
        exe = _build_cpp_example(td)
        out = subprocess.run([exe, "--describe-only", "--seed", "1"], check=True, capture_output=True, text=True).stdout
        m = re.match(r"prims (\d+) nodes (\d+) depth (\d+) visible (\d+)", out)
        assert m, out
        prims, nodes, depth, vis = map(int, m.groups())
        def shape(n):  # nodes(n) = 1 + nodes(n / 2) + nodes(n - n / 2) (bvhwrapper.rs:57-74)
            if n <= 2:
                return 1, 1
            (a, da), (b, db) = shape(n // 2), shape(n - n // 2)
            return 1 + a + b, 1 + max(da, db)

        # the C++ mirror draws its scene from its own seeded generator: 486 spheres survive the exclusion zone
        assert prims == 486 and vis == prims and (nodes, depth) == shape(prims) == (511, 9)
        if crlib.cr_device_count() == 0:  # no GPU: the render fails loudly, like Camera::render returning Err
            r = subprocess.run([exe, "--width", "32", "--samples", "1"], capture_output=True, text=True)
            assert r.returncode == 1 and "not available" in r.stderr


@pytest.mark.gpu
def test_cpp_host_mirror_renders_a_ppm(crlib, gpu_device):
    with tempfile.TemporaryDirectory() as td:
        exe = _build_cpp_example(td)
        out = os.path.join(td, "img")
        subprocess.run([exe, "--file", out, "--width", "160", "--samples", "8"], check=True, capture_output=True, text=True)
        txt = open(out + ".ppm").read().split("\n")
        assert txt[0] == "P3" and txt[1] == "160 90" and txt[2] == "255"
        px = np.array([[int(v) for v in l.split()] for l in txt[3:3 + 160 * 90]])
        assert px.shape == (160 * 90, 3) and px.min() >= 0 and px.max() <= 255
        lum = ((px / 255.0) ** 2).mean()
        assert 0.28 < lum < 0.48  # same scene family as samples/book1.png (mean linear luminance ~0.355)


def test_rejected_batch_leaves_the_scene_untouched(crlib):
    """A batch is validated before anything is appended (Sphere::new asserts, sphere.rs:26; non-finite coordinates)."""
    lib = crlib
    h = lib.cr_scene_create(-1)
    good = np.array([[0.0, 0, 0, 1], [3.0, 0, 0, 1], [6.0, 0, 0, 1]])
    assert lib.cr_scene_add_spheres(h, good.ctypes.data_as(C.c_void_p), None, None, 3) == 0
    bad = good.copy()
    bad[2, 3] = -1.0
    assert lib.cr_scene_add_spheres(h, bad.ctypes.data_as(C.c_void_p), None, None, 3) == abi.CR_ERR_INVALID
    assert b"negative radius" in lib.cr_last_error()
    tri = np.arange(18, dtype=float).reshape(2, 9)
    tri[1, 4] = np.inf
    assert lib.cr_scene_add_triangles(h, tri.ctypes.data_as(C.c_void_p), None, None, 2) == abi.CR_ERR_INVALID
    tri[1, 4] = 1.0
    assert lib.cr_scene_add_triangles(h, tri.ctypes.data_as(C.c_void_p), None, None, 2) == 3  # first index of the batch
    m = abi.CrMaterial(kind=abi.CR_MAT_METAL)
    assert lib.cr_scene_set_materials(h, C.byref(m), 1) == 0 and lib.cr_scene_commit(h) == 0
    n, d, v = C.c_uint64(), C.c_uint32(), C.c_uint64()
    assert lib.cr_scene_bvh_info(h, C.byref(n), C.byref(d), C.byref(v)) == 0 and v.value == 5
    lib.cr_scene_destroy(h)


def test_reserve_and_cached_staging_blocks_change_nothing(crlib):
    """cr_scene_reserve is only a size hint, and the staging blocks the library caches between scenes (host_pool.h: a
    per-frame rebuild as Scene::render_image does it, scene/mod.rs:332-347) come back clean: a mesh world staged call by
    call, with and without the hint, three scenes in a row, commits to the same tree every time."""
    lib = crlib
    rng = np.random.default_rng(5)
    n_batches, per = 40, 6320  # 40 meshes of teapot size: 18 MB of vertices, 16 MB of elements (pooled blocks)
    batches = [(rng.uniform(-50, 50, (per, 1, 3)) + rng.uniform(-1, 1, (per, 3, 3))).reshape(per, 9) for _ in range(n_batches)]
    mats = [np.full(per, k % 3, np.int32) for k in range(n_batches)]
    sph = np.array([[0.0, -1000.0, 0.0, 1000.0]])
    m = (abi.CrMaterial * 3)(abi.CrMaterial(kind=abi.CR_MAT_LAMBERTIAN, scatter_prob=1.0), abi.CrMaterial(kind=abi.CR_MAT_METAL), abi.CrMaterial(kind=abi.CR_MAT_DIELECTRIC, ior=1.5))
    t = abi.CrTexture(kind=abi.CR_TEX_SOLID)

    def build(hint):
        h = lib.cr_scene_create(-1)
        if hint:
            assert lib.cr_scene_reserve(h, 1, n_batches * per, 0) == 0
        first = 0
        assert lib.cr_scene_add_spheres(h, sph.ctypes.data_as(C.c_void_p), None, None, 1) == 0
        for b, mm in zip(batches, mats):
            first = lib.cr_scene_add_triangles(h, b.ctypes.data_as(C.c_void_p), mm.ctypes.data_as(C.c_void_p), None, per)
        assert first == 1 + (n_batches - 1) * per
        assert lib.cr_scene_set_materials(h, C.cast(m, C.c_void_p), 3) == 0 and lib.cr_scene_set_textures(h, C.byref(t), 1) == 0
        assert lib.cr_scene_set_bvh_builder(h, abi.CR_BVH_HOST) == 0 and lib.cr_scene_commit(h) == 0
        n = lib.cr_scene_bvh_nodes(h, None, 0)
        nodes = np.zeros(n, dtype=abi.BVH_NODE_DTYPE)
        assert lib.cr_scene_bvh_nodes(h, nodes.ctypes.data_as(C.c_void_p), n) == n
        k = lib.cr_scene_bvh_leaf_order(h, None, 0)
        order = np.empty(k, np.int32)
        assert lib.cr_scene_bvh_leaf_order(h, order.ctypes.data_as(C.c_void_p), k) == k
        lib.cr_scene_destroy(h)
        return nodes.tobytes(), order

    ref_nodes, ref_order = build(False)
    assert len(np.unique(ref_order)) == 1 + n_batches * per  # every primitive reached (span-1 nodes list theirs twice)
    for hint in (True, False, True):
        nodes, order = build(hint)
        assert nodes == ref_nodes and np.array_equal(order, ref_order)
    assert lib.cr_device_trim(-1) == 0  # host-only: releases the cached staging blocks
    nodes, order = build(True)
    assert nodes == ref_nodes and np.array_equal(order, ref_order)
    assert lib.cr_scene_reserve(None, 1, 1, 1) == abi.CR_ERR_INVALID


def test_add_batches_equals_the_separate_calls(crlib):
    """cr_scene_add_batches (one call for a run of elements, validated and copied on all host threads) builds the scene the
    separate cr_scene_add_* calls build: same prim indices, same tree, same DFS leaf order; a bad batch leaves the scene untouched."""
    lib = crlib
    rng = np.random.default_rng(9)
    per = 3000
    runs = []  # (kind, data, material, obj_id)
    runs.append((abi.CR_PRIM_SPHERE, np.array([[0.0, -1000.0, 0.0, 1000.0], [2.0, 1.0, 0.0, 1.0]]), np.array([0, 1], np.int32), None))
    for k in range(30):  # 90 000 triangles: enough for the threaded path (>= 64 K primitives)
        tri = (rng.uniform(-40, 40, (per, 1, 3)) + rng.uniform(-1, 1, (per, 3, 3))).reshape(per, 9)
        runs.append((abi.CR_PRIM_TRIANGLE, tri, np.full(per, k % 3, np.int32), np.full(per, 100 + k, np.int32) if k % 2 else None))
    runs.append((abi.CR_PRIM_QUAD, np.array([[0.0, 0, 0, 1, 0, 0, 0, 1, 0]]), None, None))
    runs.append((abi.CR_PRIM_SPHERE, np.array([[5.0, 1.0, 5.0, 0.5]]), np.array([2], np.int32), np.array([7], np.int32)))
    m = (abi.CrMaterial * 3)(abi.CrMaterial(kind=abi.CR_MAT_LAMBERTIAN, scatter_prob=1.0), abi.CrMaterial(kind=abi.CR_MAT_METAL),
                             abi.CrMaterial(kind=abi.CR_MAT_DIELECTRIC, ior=1.5))
    t = abi.CrTexture(kind=abi.CR_TEX_SOLID)
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    add = {abi.CR_PRIM_SPHERE: lib.cr_scene_add_spheres, abi.CR_PRIM_TRIANGLE: lib.cr_scene_add_triangles, abi.CR_PRIM_QUAD: lib.cr_scene_add_quads}

    def finish(h):
        assert lib.cr_scene_set_materials(h, C.cast(m, C.c_void_p), 3) == 0 and lib.cr_scene_set_textures(h, C.byref(t), 1) == 0
        assert lib.cr_scene_set_bvh_builder(h, abi.CR_BVH_HOST) == 0 and lib.cr_scene_commit(h) == 0, lib.cr_last_error()
        n = lib.cr_scene_bvh_nodes(h, None, 0)
        nodes = np.zeros(n, dtype=abi.BVH_NODE_DTYPE)
        assert lib.cr_scene_bvh_nodes(h, nodes.ctypes.data_as(C.c_void_p), n) == n
        k = lib.cr_scene_bvh_leaf_order(h, None, 0)
        order = np.empty(k, np.int32)
        assert lib.cr_scene_bvh_leaf_order(h, order.ctypes.data_as(C.c_void_p), k) == k
        path = None
        import tempfile
        with tempfile.NamedTemporaryFile(suffix=".crscene") as fh:  # the export file holds every staging array (materials, ids)
            assert lib.cr_scene_save(h, fh.name.encode()) == 0
            blob = open(fh.name, "rb").read()
        lib.cr_scene_destroy(h)
        return nodes.tobytes(), order, blob

    h = lib.cr_scene_create(-1)
    for kind, d, mm, oo in runs:
        assert add[kind](h, ptr(d), ptr(mm), ptr(oo), len(d)) >= 0
    ref = finish(h)

    def batched(h, rs):
        n = len(rs)
        kinds = (C.c_int32 * n)(*[r[0] for r in rs])
        counts = (C.c_size_t * n)(*[len(r[1]) for r in rs])
        pd = (C.c_void_p * n)(*[r[1].ctypes.data for r in rs])
        pm = (C.c_void_p * n)(*[None if r[2] is None else r[2].ctypes.data for r in rs])
        po = (C.c_void_p * n)(*[None if r[3] is None else r[3].ctypes.data for r in rs])
        return lib.cr_scene_add_batches(h, n, kinds, pd, pm, po, counts)

    h = lib.cr_scene_create(-1)
    assert batched(h, runs[:5]) == 0           # first index of the call
    assert batched(h, runs[5:]) == 2 + 4 * per
    got = finish(h)
    assert got[0] == ref[0] and np.array_equal(got[1], ref[1]) and got[2] == ref[2]

    h = lib.cr_scene_create(-1)
    assert batched(h, runs[:3]) == 0
    bad = [list(r) for r in runs[3:8]]
    bad[2][1] = bad[2][1].copy()
    bad[2][1][17, 4] = np.nan
    assert batched(h, bad) == abi.CR_ERR_INVALID and b"non-finite" in lib.cr_last_error()
    neg = [(abi.CR_PRIM_SPHERE, np.array([[0.0, 0.0, 0.0, -1.0]]), None, None)]
    assert batched(h, runs[3:4] + neg) == abi.CR_ERR_INVALID and b"negative radius" in lib.cr_last_error()
    assert batched(h, [(7, runs[1][1], None, None)]) == abi.CR_ERR_INVALID
    assert batched(h, runs[3:]) == 2 + 2 * per  # the rejected calls appended nothing
    got = finish(h)
    assert got[0] == ref[0] and np.array_equal(got[1], ref[1]) and got[2] == ref[2]
