"""Object keyframes: TransformTimeline for spheres / triangle vertices, evaluated at the ray's time
by Sphere::hit (sphere.rs:67-70) and Triangle::hit (triangle.rs:91-97).  The timeline KATs are the reference's own
(src/timeline/mod.rs:292-349); the GPU cases compare the CUDA path with the oracle through the C ABI."""
import numpy as np
import pytest

from crucible_b200 import abi, demo_builder
from crucible_b200.scene import (Color, InterpolationType, Lambertian, Metal, Point3, Scene, Sphere, TransformSpace, TransformTimeline,
                                 Triangle)

LERP, NERP = InterpolationType.LERP, InterpolationType.NERP
World, Local = TransformSpace.World, TransformSpace.Local


def _oracle_point_at(oracle, tl, t):
    """The same timeline through the ORACLE's restatement of combine_and_compute (test infrastructure)."""
    keys = tl.anim_keys()
    arr = (abi.CrAnimKey * max(len(keys), 1))(*keys)
    init = (abi.C.c_double * 4)(*tl.start_pos, tl.start_scale)
    out = (abi.C.c_double * 4)()
    oracle.load().orc_combine_and_compute(init, abi.C.cast(arr, abi.C.c_void_p), len(keys), float(t), out)
    return np.array(list(out))


def _evaluators(oracle):
    """(name, f(tl, t)) for the product's host evaluator (cr_anim_point_at) and the oracle's."""
    return [("product", lambda tl, t: tl.combine_and_compute_object(t)), ("oracle", lambda tl, t: _oracle_point_at(oracle, tl, t))]


# ---- the reference's own timeline KATs (src/timeline/mod.rs:292-349), against the PRODUCT's evaluator (the host copy
# of the device routine, cr_anim_point_at) and against the oracle's restatement
def test_check_nerp_scaling(oracle):
    for name, ev in _evaluators(oracle):
        tl = TransformTimeline.new_sphere(Point3(2.0, 3.0, 0.0), Point3(2.0, 1.0, 3.0), 1.0)
        tl.scale_sphere(15.0, 5.0, NERP)
        assert ev(tl, 7.0)[3] == 15.0, name
        assert ev(tl, 3.15)[3] == 1.0, name


def test_check_lerp_scaling(oracle):
    for name, ev in _evaluators(oracle):
        tl = TransformTimeline.new_sphere(Point3(2.0, 3.0, 0.0), Point3(2.0, 1.0, 3.0), 1.0)
        tl.scale_sphere(15.0, 5.0, LERP)
        tl.scale_sphere(5.0, 10.0, LERP)
        assert ev(tl, 5.0)[3] == 15.0, name
        assert abs(ev(tl, 3.15)[3] - 10.0) < 0.2, name
        assert ev(tl, 3.15)[3] == 1.0 + (15.0 - 1.0) * ((3.15 - 0.0) / (5.0 - 0.0)), name
        assert ev(tl, 12.0)[3] == 5.0, name  # past the last keyframe: clamp(proportion) = 1


def test_check_nerp_translate_object(oracle):
    for name, ev in _evaluators(oracle):
        tl = TransformTimeline.new(Point3(2.0, 3.0, 1.0), Point3(0, 0, 0), 1.0)
        tl.translate_x(1.0, 5.0, NERP, Local)
        tl.translate_y(10.0, 3.0, NERP, Local)
        r = ev(tl, 0.0)
        assert r[0] == 2.0 and r[1] == 3.0, name
        r = ev(tl, 5.0)
        assert r[0] == 3.0 and r[1] == 13.0 and r[2] == 1.0 and r[3] == 1.0, name


def test_scale_xyz_matrices(oracle):
    """scale_x / scale_y / scale_z (transform_builder.rs:101-346).  The scale matrix multiplies the TRANSLATED point
    (combined = scale * translate, timeline/mod.rs:260-261), so a vertex is scaled about the world origin; scale_y
    writes its value into row 1, column 0 (:229-246), so y' = v*x + y; only the LAST valid scale key counts (:251-257)."""
    for name, ev in _evaluators(oracle):
        tl = TransformTimeline.new(Point3(2.0, 3.0, 5.0), Point3(0, 0, 0), 1.0)
        tl.scale_x(4.0, 2.0, LERP)
        assert list(ev(tl, 0.0)) == [2.0, 3.0, 5.0, 1.0], name  # s = 0: start + (4 - start) * 0 = 1
        assert list(ev(tl, 1.0)) == [(1.0 + (4.0 - 1.0) * 0.5) * 2.0, 3.0, 5.0, 1.0], name
        assert list(ev(tl, 9.0)) == [8.0, 3.0, 5.0, 1.0], name
        tl = TransformTimeline.new(Point3(2.0, 3.0, 5.0), Point3(0, 0, 0), 1.0)
        tl.scale_y(4.0, 2.0, NERP)
        assert list(ev(tl, 1.0)) == [2.0, 3.0, 5.0, 1.0], name
        assert list(ev(tl, 2.0)) == [2.0, 4.0 * 2.0 + 3.0, 5.0, 1.0], name  # the wrong-slot behaviour, reproduced
        tl = TransformTimeline.new(Point3(2.0, 3.0, 5.0), Point3(0, 0, 0), 1.0)
        tl.scale_z(0.5, 1.0, NERP)
        tl.translate_z(1.0, 1.0, NERP, Local)
        assert list(ev(tl, 1.5)) == [2.0, 3.0, 0.5 * (5.0 + 1.0), 1.0], name  # scale applies after the translation
        # scale_point pushes an x, a y and a z key on the same interval; stable sort keeps x, y, z order, so the last
        # valid one (z) is the only one applied
        tl = TransformTimeline.new(Point3(2.0, 3.0, 5.0), Point3(0, 0, 0), 1.0)
        tl.scale_point(Point3(10.0, 20.0, 3.0), 2.0, NERP)
        assert list(ev(tl, 3.0)) == [2.0, 3.0, 15.0, 1.0], name
        # a later x key takes over from the z key once it starts
        tl.scale_x(2.0, 4.0, NERP)
        assert list(ev(tl, 3.0)) == [2.0, 3.0, 15.0, 1.0], name
        assert list(ev(tl, 4.0)) == [4.0, 3.0, 5.0, 1.0], name
        # LERP chains start from the previous end of the SAME kind (most_recent_matching_transform)
        tl = TransformTimeline.new(Point3(1.0, 1.0, 1.0), Point3(0, 0, 0), 1.0)
        tl.scale_x(3.0, 1.0, LERP)
        tl.scale_z(7.0, 2.0, LERP)   # starts from the init scale 1.0 at t = 0 (no earlier z key)
        tl.scale_x(5.0, 3.0, LERP)   # starts from 3.0 at t = 1
        k = tl.anim_keys()
        assert [(q.kind, q.t0, q.t1, q.a, q.b) for q in k] == [(4, 0.0, 1.0, 1.0, 3.0), (6, 0.0, 2.0, 1.0, 7.0), (4, 1.0, 3.0, 3.0, 5.0)], name
        assert list(ev(tl, 0.5)) == [1.0, 1.0, 1.0 + 6.0 * 0.25, 1.0], name  # z key is later in the list than the x key
        assert list(ev(tl, 2.0)) == [3.0 + 2.0 * 0.5, 1.0, 1.0, 1.0], name   # the second x key is the last valid one


def test_product_evaluator_equals_oracle_on_random_timelines(oracle):
    """cr_anim_point_at (host copy of the device routine) == the oracle's combine_and_compute, bit for bit."""
    rng = np.random.default_rng(5)
    for trial in range(200):
        sphere = trial % 3 == 0
        pos = Point3(*rng.normal(scale=3.0, size=3))
        tl = TransformTimeline.new_sphere(pos, None, float(rng.uniform(0.1, 2.0))) if sphere else TransformTimeline.new(pos, None, 1.0)
        for _ in range(int(rng.integers(0, 7))):
            kf = float(rng.uniform(0.0, 4.0))
            it = LERP if rng.random() < 0.6 else NERP
            op = int(rng.integers(0, 5))
            if op < 3:
                tl._translate(op, float(rng.normal()), kf, it, World if rng.random() < 0.5 else Local)
            elif sphere:
                tl.scale_sphere(float(rng.uniform(0.1, 3.0)), kf, it)
            else:
                tl._scale(int(rng.integers(4, 7)), float(rng.uniform(-2.0, 3.0)), kf, it)
        for t in list(rng.uniform(-0.5, 5.0, 6)) + [0.0, 4.0]:
            a, b = tl.combine_and_compute_object(t), _oracle_point_at(oracle, tl, t)
            assert np.array_equal(a, b), (trial, t, a, b)


def test_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under crucible_b200/ may import, load or call it."""
    import os
    import re

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "crucible_b200")
    bad = []
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"(^|\n)\s*(from|import)\s+oracle\b|liboracle|orc_[a-z_]+\(", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_world_space_deltas_differ_per_vertex():
    """translate_x(World) moves every point of the object TO x (delta = x - that point's coordinate,
    transform_builder.rs:393-403), so the three vertex timelines of a triangle get different keys."""
    sc = Scene.new_image(1.0, 8, 24, 180.0)
    sc.add_element(Triangle(Point3(0, 0, 0), Point3(1, 0, 0), Point3(0, 1, 0), Metal(Color(.5, .5, .5), 0.0)), "t")
    sc.translate_x(5.0, 2.0, LERP, World, "t")
    anim = sc.describe().animation
    assert [(p, pt) for p, pt, _ in anim] == [(0, 0), (0, 1), (0, 2)]
    assert [k[0].a for _, _, k in anim] == [5.0, 4.0, 5.0]
    assert all(k[0].t0 == 0.0 and k[0].t1 == 2.0 and k[0].kind == 0 for _, _, k in anim)


def test_animator_type_checks():
    sc = Scene.new_image(1.0, 8, 24, 180.0)
    sc.add_element(Triangle(Point3(0, 0, 0), Point3(1, 0, 0), Point3(0, 1, 0), Metal(Color(.5, .5, .5), 0.0)), "t")
    with pytest.raises(ValueError, match="ScaleR can only be applied to Spheres"):  # scene_animator.rs:140-150
        sc.scale_r(2.0, 1.0, LERP, "t")
    with pytest.raises(ValueError, match="Could not find an object with the alias"):
        sc.translate_x(1.0, 1.0, LERP, Local, "nope")
    sc.add_element(Sphere(Point3(0, 0, 0), 1.0, Metal(Color(.5, .5, .5), 0.0)), "s")
    for fn, msg in ((sc.scale_x, "ScaleX"), (sc.scale_y, "ScaleY"), (sc.scale_z, "ScaleZ")):  # scene_animator.rs:38-41, 72-75, 106-109
        with pytest.raises(ValueError, match=f"{msg} cannot apply to Spheres"):
            fn(2.0, 1.0, LERP, "s")
    with pytest.raises(ValueError, match="ScaleAll cannot apply to Spheres"):  # :187-190
        sc.scale_all_uniform(2.0, 1.0, LERP, "s")
    lib = abi.load()
    h = lib.cr_scene_create(-1)
    tri = np.zeros((1, 9))
    tri[0, 3] = tri[0, 7] = 1.0
    assert lib.cr_scene_add_triangles(h, tri.ctypes.data_as(abi.C.c_void_p), None, None, 1) == 0
    key = abi.CrAnimKey(0.0, 1.0, 1.0, 2.0, 3, abi.CR_LERP)
    assert lib.cr_scene_set_keyframes(h, 0, 0, abi.C.byref(key), 1) == abi.CR_ERR_INVALID
    assert b"ScaleR can only be applied to Spheres" in lib.cr_last_error()
    sph = np.array([[0.0, 0.0, 0.0, 1.0]])
    assert lib.cr_scene_add_spheres(h, sph.ctypes.data_as(abi.C.c_void_p), None, None, 1) == 1
    for kind, msg in ((4, b"ScaleX"), (5, b"ScaleY"), (6, b"ScaleZ")):
        key.kind = kind
        assert lib.cr_scene_set_keyframes(h, 1, 0, abi.C.byref(key), 1) == abi.CR_ERR_INVALID
        assert msg + b" cannot apply to Spheres" in lib.cr_last_error()
        assert lib.cr_scene_set_keyframes(h, 0, 1, abi.C.byref(key), 1) == abi.CR_OK  # fine on a triangle vertex
    key.kind = 7
    assert lib.cr_scene_set_keyframes(h, 0, 0, abi.C.byref(key), 1) == abi.CR_ERR_INVALID
    key.kind = 0
    assert lib.cr_scene_set_keyframes(h, 0, 3, abi.C.byref(key), 1) == abi.CR_ERR_INVALID
    assert lib.cr_scene_set_keyframes(h, 5, 0, abi.C.byref(key), 1) == abi.CR_ERR_INVALID
    assert lib.cr_scene_set_keyframes(h, 0, 2, abi.C.byref(key), 1) == abi.CR_OK
    lib.cr_scene_destroy(h)


def _moving_scene(image_width=96, samples=4):
    """Three spheres on a checker ground: one slides (LERP, World), one jumps (NERP, Local) and grows (scale_r),
    one is static; a triangle fan drifts (Local).  24 fps, 180 degree shutter: sample times span half a frame."""
    from crucible_b200.scene import CheckerTexture

    sc = Scene.new_movie(16.0 / 9.0, image_width, 24, 180.0, 0, 1.0)
    sc.scene_cam.set_samples(samples)
    sc.scene_cam.set_max_depth(8)
    sc.scene_cam.look_from(Point3(0.0, 2.0, 9.0))
    sc.scene_cam.look_at(Point3(0.0, 0.8, 0.0))
    sc.scene_cam.set_vfov(40.0)
    ground = Lambertian.new_from_texture(CheckerTexture.new_from_color(0.5, Color(.2, .3, .1), Color(.9, .9, .9)), 1.0)
    sc.add_element(Sphere(Point3(0, -1000, 0), 1000.0, ground), "ground")
    sc.add_element(Sphere(Point3(-2.0, 1.0, 0.0), 1.0, Metal(Color(.8, .6, .2), 0.1)), "slider")
    sc.add_element(Sphere(Point3(2.0, 0.6, 0.5), 0.6, Lambertian.new_from_color(Color(.7, .2, .2), 1.0)), "jumper")
    sc.add_element(Sphere(Point3(0.0, 0.5, -2.0), 0.5, Metal(Color(.7, .7, .7), 0.0)), "static")
    for i in range(4):
        a = Point3(-1.0 + 0.5 * i, 0.2, 2.0)
        sc.add_element(Triangle(a, a + Point3(0.5, 0.0, 0.1), a + Point3(0.2, 0.9, 0.0), Metal(Color(.3, .5, .8), 0.2)), f"tri{i}")
    sc.translate_x(-1.2, 1.0, LERP, World, "slider")          # stays inside its construction box for a while
    sc.translate_y(0.3, 0.02, NERP, Local, "jumper")
    sc.scale_r(0.75, 0.5, LERP, "jumper")
    for i in range(4):
        sc.translate_point(Point3(0.3, 0.2, -0.2), 0.5, LERP, Local, f"tri{i}")
    return sc


@pytest.mark.gpu
def test_animated_trace_batch_bit_exact(gpu_device, oracle):
    """Hittables::hit with moving primitives: rays carry a time; ids bit-exact, t / p / n bit-identical."""
    from scenes_util import compare_hits
    from crucible_b200.gpu import GpuScene

    sc = _moving_scene()
    desc, cam = sc.describe(), sc.scene_cam.to_abi()
    assert len(desc.animation) == 2 + 12
    gs, orc = GpuScene(desc, gpu_device), oracle.OracleScene(desc)
    rng = np.random.default_rng(11)
    rays = orc.gen_rays(cam, 0, cam.image_width * cam.image_height)
    rays = np.concatenate([rays, orc.gen_rays(cam, 1, 60000, seed=3)])
    seen = set()
    for t in (0.0, 0.01, 0.02, 0.25, 0.5, 0.9, 3.0):
        rays[:, 6] = t
        got, exp = gs.trace_batch(rays), orc.trace_batch(rays)
        compare_hits(got, exp)
        seen.add(tuple(np.bincount(exp["prim_index"][exp["prim_index"] >= 0], minlength=8)[1:4]))
    assert len(seen) > 3  # the animated spheres are hit by different ray sets at different times
    rays[:, 6] = rng.uniform(0.0, 1.0, len(rays))  # every ray its own time (motion blur)
    compare_hits(gs.trace_batch(rays), orc.trace_batch(rays))
    # static scene + ray times: times are ignored (no keyframes)
    got32 = gs.trace_batch(rays, precision=abi.CR_PRECISION_F32)
    assert (got32["prim_index"] == orc.trace_batch(rays)["prim_index"]).mean() > 0.99


@pytest.mark.gpu
def test_animated_render_matches_oracle(gpu_device, oracle):
    """Frames of a movie with moving objects: same Philox streams, same paths => per-pixel means agree to 1e-11 and
    the number of world.hit calls is identical; the frames differ from each other (objects move, motion blur)."""
    from crucible_b200.gpu import GpuScene

    sc = _moving_scene(image_width=96, samples=4)
    desc, cam = sc.describe(), sc.scene_cam.to_abi()
    gs, orc = GpuScene(desc, gpu_device), oracle.OracleScene(desc)
    frames = []
    for frame in (0, 6, 12, 23):
        cam.frame = frame
        rgb, _, st = gs.render(cam, seed=9)
        ref, _, ost = orc.render(cam, seed=9)
        assert np.abs(rgb - ref).max() < 1e-11, frame
        assert st["rays"] == ost["rays"], frame
        frames.append(rgb)
    assert np.abs(frames[0] - frames[2]).max() > 0.1


def _scaling_scene(image_width=96, samples=4):
    """Three small meshes on a checker ground, animated with the scale bindings of scene_animator.rs:38-219:
    scale_x (LERP), scale_y (NERP, the wrong-slot matrix) and scale_all_uniform (which ends up scaling z only), plus a
    translation on the first mesh so that scale * translate is exercised.  The BVH keeps the construction-time boxes."""
    from crucible_b200.scene import CheckerTexture

    sc = Scene.new_movie(16.0 / 9.0, image_width, 24, 180.0, 0, 1.0)
    sc.scene_cam.set_samples(samples)
    sc.scene_cam.set_max_depth(8)
    sc.scene_cam.look_from(Point3(0.5, 2.5, 7.0))
    sc.scene_cam.look_at(Point3(0.5, 0.6, 0.0))
    sc.scene_cam.set_vfov(40.0)
    ground = Lambertian.new_from_texture(CheckerTexture.new_from_color(0.5, Color(.2, .3, .1), Color(.9, .9, .9)), 1.0)
    sc.add_element(Sphere(Point3(0, -1000, 0), 1000.0, ground), "ground")
    v, f = demo_builder.teapot_mesh()
    f = f[::4]  # 1580 triangles per mesh keeps the oracle quick
    sc.load_mesh(v, f, "pot_x", 0.5, Point3(-2.0, 0.0, 0.5), Metal(Color(.8, .6, .2), 0.1))
    sc.load_mesh(v, f, "pot_y", 0.5, Point3(1.0, 0.0, 0.0), Lambertian.new_from_color(Color(.7, .2, .2), 1.0))
    sc.load_mesh(v, f, "pot_all", 0.5, Point3(3.0, 0.0, 1.0), Metal(Color(.3, .5, .8), 0.2))
    sc.scale_x(1.15, 1.0, LERP, "pot_x")
    sc.translate_x(0.2, 0.5, LERP, Local, "pot_x")
    sc.scale_y(0.05, 0.25, NERP, "pot_y")
    sc.scale_all_uniform(1.1, 0.5, LERP, "pot_all")
    return sc


@pytest.mark.gpu
def test_scaled_meshes_bit_exact(gpu_device, oracle):
    """SURVEY 8f-1 remainder: scale_x / scale_y / scale_z / scale_point / scale_all_uniform on mesh triangles.
    Closest-hit ids bit-exact and t / p / n bit-identical against the oracle at fixed and per-ray times; frames equal
    the oracle's to 1e-11 with the same number of world.hit calls."""
    from scenes_util import compare_hits
    from crucible_b200.gpu import GpuScene

    sc = _scaling_scene()
    desc, cam = sc.describe(), sc.scene_cam.to_abi()
    assert len(desc.animation) == 3 * 1580 * 3
    kinds = {k.kind for _, _, keys in desc.animation for k in keys}
    assert kinds == {0, 4, 5, 6}
    gs, orc = GpuScene(desc, gpu_device), oracle.OracleScene(desc)
    rays = np.concatenate([orc.gen_rays(cam, 0, cam.image_width * cam.image_height), orc.gen_rays(cam, 1, 60000, seed=3)])
    hit_sets = []
    for t in (0.0, 0.2, 0.25, 0.5, 0.75, 2.0):
        rays[:, 6] = t
        got, exp = gs.trace_batch(rays), orc.trace_batch(rays)
        compare_hits(got, exp)
        hit_sets.append(exp["prim_index"].copy())
    assert (hit_sets[0] != hit_sets[3]).mean() > 0.01  # the meshes really change shape
    rays[:, 6] = np.random.default_rng(2).uniform(0.0, 1.0, len(rays))
    compare_hits(gs.trace_batch(rays), orc.trace_batch(rays))
    for frame in (0, 7, 13):
        cam.frame = frame
        rgb, _, st = gs.render(cam, seed=4)
        ref, _, ost = orc.render(cam, seed=4)
        assert np.abs(rgb - ref).max() < 1e-11, frame
        assert st["rays"] == ost["rays"], frame
