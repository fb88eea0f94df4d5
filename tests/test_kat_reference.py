"""The reference's own unit tests that touch the hot path, ported one for one and run against the
oracle.  These are the golden vectors that PIN the oracle's arithmetic helpers."""
import math

import numpy as np
import pytest


# ---- src/utils.rs:703-772
def test_neg(oracle):
    assert oracle.vec(0, (1.0, 2.0, 3.0)).tolist() == [-1.0, -2.0, -3.0]


def test_plus_assign(oracle):
    assert oracle.vec(1, (1.0, 2.0, 3.0), (2.0, 2.0, 1.0)).tolist() == [3.0, 4.0, 4.0]


def test_dot(oracle):
    assert oracle.vec(2, (1.0, 2.0, 3.0), (2.0, 2.0, 1.0))[0] == 9.0


def test_cross(oracle):
    assert oracle.vec(3, (3.0, -3.0, 1.0), (4.0, 9.0, 2.0)).tolist() == [-15.0, -2.0, 39.0]


def test_length(oracle):
    assert oracle.vec(4, (3.0, 4.0, 0.0))[0] == 5.0


# ---- src/utils.rs:774-778 invalid_color_test (#[should_panic])
def test_invalid_color(oracle):
    assert oracle.load().orc_kat_color_valid(20.0, 30.0, 40.0) == 0
    assert oracle.load().orc_kat_color_valid(0.5, 0.0, 1.0) == 1
    from crucible_b200.scene import Color

    with pytest.raises(ValueError):
        Color(20.0, 30.0, 40.0)


# ---- src/utils.rs:780-785 color_display_test: pins sqrt-gamma + truncation
def test_color_display(oracle):
    assert oracle.color_bytes((0.529, 0.616, 0.730)).tolist() == [185, 200, 217]
    assert oracle.color_bytes((1.0, 0.0, 0.25)).tolist() == [255, 0, 127]


# ---- src/utils.rs:787-805
def test_inv_color(oracle):
    assert oracle.color_neg((1.0, 0.0, 0.0)).tolist() == [0.0, 1.0, 1.0]


def test_add_color(oracle):
    assert oracle.color_add((1.0, 0.0, 0.0), (0.0, 1.0, 0.0)).tolist() == [1.0, 1.0, 0.0]
    assert oracle.color_add((0.75, 0.5, 0.0), (0.75, 0.25, 0.0)).tolist() == [1.0, 0.75, 0.0]  # clamped add


# ---- src/utils.rs:807-831
def test_degrees_convert(oracle):
    assert abs(oracle.load().orc_kat_deg_to_rad(59.2958) - 1.034906943) < 0.0000000005


def test_degrees_circular(oracle):
    lib = oracle.load()
    assert abs(lib.orc_kat_rad_to_deg(lib.orc_kat_deg_to_rad(90.0)) - 90.0) < 0.000000005


# ---- src/utils.rs:833-912
def test_interval(oracle):
    assert oracle.interval(0, 3.0, 20.0) == 17.0
    assert oracle.interval(1, 3.0, 20.0, 3.0) == 1.0
    assert oracle.interval(1, 3.0, 20.0, 21.0) == 0.0
    assert oracle.interval(1, 3.0, 20.0, 15.0) == 1.0
    # surrounds is STRICT
    assert oracle.interval(2, 3.0, 20.0, 3.0) == 0.0
    assert oracle.interval(2, 3.0, 20.0, 21.0) == 0.0
    assert oracle.interval(2, 3.0, 20.0, 15.0) == 1.0
    assert oracle.interval(1, 5.0, 5.0, 5.0) == 1.0  # discrete_contains
    assert oracle.interval(3, 3.0, 10.0, 2.0) == 1.0  # interval_greater
    assert oracle.interval(4, 3.0, 10.0, 11.0) == 1.0  # interval_less
    assert oracle.interval(5, 2.0, 10.0, 4.0) == 0.25  # get_proportion


def test_universe_and_empty(oracle):
    rng = np.random.default_rng(0)
    for x in rng.uniform(-500, 500, 10):
        assert oracle.interval(1, -math.inf, math.inf, x) == 1.0
        assert oracle.interval(1, math.inf, -math.inf, x) == 0.0


# ---- src/camera/mod.rs:382-396
def test_ray_at(oracle):
    assert oracle.ray_at((0, 0, 0), (2.0, -3.0, 1.5), 2.0).tolist() == [4.0, -6.0, 3.0]


def test_average_color(oracle):
    assert oracle.average([[0.0, 1.0, 0.0], [0.5, 0.5, 1.0]]).tolist() == [0.25, 0.75, 0.5]


# ---- src/timeline/mod.rs:329-349 check_nerp_translate, through the mirror of the keyframe builder,
# evaluated by BOTH the oracle and the product's host evaluator
def test_nerp_translate(oracle, crlib):
    from crucible_b200.scene import InterpolationType, Point3, TransformSpace, TransformTimeline

    tl = TransformTimeline.new(Point3(2.0, 3.0, 1.0))
    tl.translate_x(1.0, 5.0, InterpolationType.NERP, TransformSpace.Local)
    tl.translate_y(10.0, 3.0, InterpolationType.NERP, TransformSpace.Local)
    keys, n = tl.keyframes()
    for ev in (lambda t: oracle.point_at((2.0, 3.0, 1.0), keys, n, t), lambda t: tl.combine_and_compute(t)):
        r = ev(0.0)
        assert r[0] == 2.0 and r[1] == 3.0
        r = ev(5.0)
        assert r[0] == 3.0 and r[1] == 13.0


def test_lerp_world_walk(oracle, crlib):
    """Pattern of demo_movies.rs:33-68: World-space LERP keyframes visit the given points exactly."""
    from crucible_b200.scene import InterpolationType, Point3, TransformSpace, TransformTimeline

    tl = TransformTimeline.new(Point3(0.0, 0.0, -12.0))
    pts = [((12.0, 0.0, 0.0), 2.5), ((0.0, 0.0, 12.0), 5.0), ((-12.0, 0.0, 0.0), 7.5), ((0.0, 0.0, -12.0), 10.0)]
    for p, kf in pts:
        tl.translate_point(p, kf, InterpolationType.LERP, TransformSpace.World)
    keys, n = tl.keyframes()
    for p, kf in pts:
        assert np.allclose(oracle.point_at((0.0, 0.0, -12.0), keys, n, kf), p, atol=1e-12)
        assert np.allclose(tl.combine_and_compute(kf)[:3], p, atol=1e-12)
    mid = oracle.point_at((0.0, 0.0, -12.0), keys, n, 1.25)
    assert np.allclose(mid, (6.0, 0.0, -6.0), atol=1e-12)
    assert np.array_equal(mid, tl.combine_and_compute(1.25)[:3])  # host evaluator == oracle, bit for bit


# ---- Philox4x32-10 known answers (Random123 kat_vectors)
@pytest.mark.parametrize("ctr,key,exp", [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
])
def test_philox_kat(oracle, crlib, ctr, key, exp):
    assert tuple(int(x) for x in oracle.philox(ctr, key)) == exp
    out = np.zeros(4, np.uint32)
    c, k = np.array(ctr, np.uint32), np.array(key, np.uint32)
    crlib.cr_philox4x32_10(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    assert tuple(int(x) for x in out) == exp
