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


# ---- src/timeline/mod.rs:292-349, through the mirror of the keyframe builder and the oracle's combine_and_compute
def test_check_nerp_scaling(oracle):
    tl = TransformTimeline.new_sphere(Point3(2.0, 3.0, 0.0), Point3(2.0, 1.0, 3.0), 1.0)
    tl.scale_sphere(15.0, 5.0, NERP)
    assert tl.combine_and_compute_object(7.0)[3] == 15.0
    assert tl.combine_and_compute_object(3.15)[3] == 1.0


def test_check_lerp_scaling(oracle):
    tl = TransformTimeline.new_sphere(Point3(2.0, 3.0, 0.0), Point3(2.0, 1.0, 3.0), 1.0)
    tl.scale_sphere(15.0, 5.0, LERP)
    tl.scale_sphere(5.0, 10.0, LERP)
    assert tl.combine_and_compute_object(5.0)[3] == 15.0
    assert abs(tl.combine_and_compute_object(3.15)[3] - 10.0) < 0.2
    assert tl.combine_and_compute_object(3.15)[3] == 1.0 + (15.0 - 1.0) * ((3.15 - 0.0) / (5.0 - 0.0))
    assert tl.combine_and_compute_object(12.0)[3] == 5.0  # past the last keyframe: clamp(proportion) = 1


def test_check_nerp_translate_object(oracle):
    tl = TransformTimeline.new(Point3(2.0, 3.0, 1.0), Point3(0, 0, 0), 1.0)
    tl.translate_x(1.0, 5.0, NERP, Local)
    tl.translate_y(10.0, 3.0, NERP, Local)
    r = tl.combine_and_compute_object(0.0)
    assert r[0] == 2.0 and r[1] == 3.0
    r = tl.combine_and_compute_object(5.0)
    assert r[0] == 3.0 and r[1] == 13.0 and r[2] == 1.0 and r[3] == 1.0


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
    lib = abi.load()
    h = lib.cr_scene_create(-1)
    tri = np.zeros((1, 9))
    tri[0, 3] = tri[0, 7] = 1.0
    assert lib.cr_scene_add_triangles(h, tri.ctypes.data_as(abi.C.c_void_p), None, None, 1) == 0
    key = abi.CrAnimKey(0.0, 1.0, 1.0, 2.0, 3, abi.CR_LERP)
    assert lib.cr_scene_set_keyframes(h, 0, 0, abi.C.byref(key), 1) == abi.CR_ERR_INVALID
    assert b"ScaleR can only be applied to Spheres" in lib.cr_last_error()
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
