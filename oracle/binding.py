"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The struct layouts come from crucible_b200.abi (the public header's PODs); nothing in crucible_b200
imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from crucible_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


class OrcRenderStats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("rays", C.c_uint64), ("node_tests", C.c_uint64), ("sphere_tests", C.c_uint64),
                ("tri_tests", C.c_uint64), ("quad_tests", C.c_uint64), ("seconds", C.c_double), ("threads", C.c_int32),
                ("pad", C.c_int32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "pad"}


def build():
    subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    P = C.c_void_p
    lib.orc_last_error.restype = C.c_char_p
    lib.orc_scene_create.restype = P
    lib.orc_scene_destroy.argtypes = [P]
    for n in ("orc_scene_add_spheres", "orc_scene_add_triangles", "orc_scene_add_quads"):
        getattr(lib, n).restype = C.c_int64
        getattr(lib, n).argtypes = [P, P, P, P, C.c_size_t]
    lib.orc_scene_begin_group.argtypes = [P, C.c_int]
    lib.orc_scene_end_group.argtypes = [P]
    lib.orc_scene_set_hidden.argtypes = [P, C.c_size_t, C.c_int]
    lib.orc_scene_set_keyframes.argtypes = [P, C.c_size_t, C.c_int, P, C.c_size_t]
    lib.orc_combine_and_compute.argtypes = [P, P, C.c_size_t, C.c_double, P]
    lib.orc_combine_and_compute.restype = None
    lib.orc_scene_set_materials.argtypes = [P, P, C.c_size_t]
    lib.orc_scene_set_textures.argtypes = [P, P, C.c_size_t]
    lib.orc_scene_add_image.argtypes = [P, P, C.c_int, C.c_int]
    lib.orc_scene_set_sky.argtypes = [P, C.c_int, C.c_int]
    lib.orc_scene_commit.argtypes = [P]
    lib.orc_scene_bvh_info.argtypes = [P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
    lib.orc_scene_bvh_leaf_order.restype = C.c_int64
    lib.orc_scene_bvh_leaf_order.argtypes = [P, P, C.c_size_t]
    lib.orc_scene_root_bbox.argtypes = [P, P]
    lib.orc_trace_batch.argtypes = [P, P, C.c_size_t, C.c_double, C.c_double, C.c_int, P, P, C.c_int]
    lib.orc_render.argtypes = [P, C.POINTER(abi.CrCamera), C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, P, P,
                               C.POINTER(OrcRenderStats)]
    lib.orc_gen_rays.argtypes = [P, C.POINTER(abi.CrCamera), C.c_uint64, C.c_int, C.c_size_t, P]
    lib.orc_kat_vec.argtypes = [C.c_int, P, P, P]
    lib.orc_kat_color_valid.argtypes = [C.c_double] * 3
    lib.orc_kat_color_bytes.argtypes = [P, P]
    lib.orc_kat_color_neg.argtypes = [P, P]
    lib.orc_kat_color_add.argtypes = [P, P, P]
    lib.orc_kat_color_scale.argtypes = [C.c_double, P, P]
    lib.orc_kat_average.argtypes = [P, C.c_size_t, P]
    lib.orc_kat_interval.restype = C.c_double
    lib.orc_kat_interval.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
    lib.orc_kat_deg_to_rad.restype = C.c_double
    lib.orc_kat_deg_to_rad.argtypes = [C.c_double]
    lib.orc_kat_rad_to_deg.restype = C.c_double
    lib.orc_kat_rad_to_deg.argtypes = [C.c_double]
    lib.orc_kat_ray_at.argtypes = [P, P, C.c_double, P]
    lib.orc_kat_point_at.argtypes = [P, P, C.c_uint32, C.c_double, P]
    lib.orc_philox4x32_10.argtypes = [P, P, P]
    lib.orc_rng_stream.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_size_t, P]
    lib.orc_kat_prim_hit.argtypes = [C.c_int, P, P, C.c_double, C.c_double, P]
    lib.orc_kat_aabb_hit.argtypes = [P, P, C.c_double, C.c_double]
    lib.orc_kat_scatter.argtypes = [P, P, P, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, P, P]
    lib.orc_kat_tex_value.argtypes = [P, C.c_int, C.c_double, C.c_double, P, P]
    lib.orc_kat_sky.argtypes = [P, P, P]
    _lib = lib
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _d(*v):
    return np.array(v, np.float64)


class OracleScene:
    """The reference-order CPU world built from the same flat description the GPU gets."""

    def __init__(self, desc):
        self.lib = load()
        self.handle = self.lib.orc_scene_create()
        self.desc = desc
        desc.apply(self.lib, self.handle, "orc_")

    def close(self):
        if getattr(self, "handle", None):
            self.lib.orc_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        self.close()

    def bvh_info(self):
        n, d, v = C.c_uint64(), C.c_uint32(), C.c_uint64()
        self.lib.orc_scene_bvh_info(self.handle, C.byref(n), C.byref(d), C.byref(v))
        return {"n_nodes": n.value, "max_depth": d.value, "n_visible": v.value}

    def bvh_leaf_order(self):
        n = self.lib.orc_scene_bvh_leaf_order(self.handle, None, 0)
        out = np.empty(n, np.int32)
        self.lib.orc_scene_bvh_leaf_order(self.handle, _p(out), n)
        return out

    def root_bbox(self):
        out = np.zeros(6)
        rc = self.lib.orc_scene_root_bbox(self.handle, _p(out))
        return out if rc == 0 else None

    def trace_batch(self, rays, tmin=0.001, tmax=float("inf"), brute=False, counters=False, threads=0, order_free=False):
        rays = np.ascontiguousarray(rays, np.float64)
        out = np.zeros(len(rays), dtype=abi.HIT_DTYPE)
        cn = np.zeros((len(rays), 4), np.uint32) if counters else None
        threads = threads or (os.cpu_count() or 1)
        rc = self.lib.orc_trace_batch(self.handle, _p(rays), len(rays), tmin, tmax, int(order_free) + 1 if order_free else (1 if brute else 0), _p(out),
                                      _p(cn) if counters else None, threads)
        assert rc == 0, self.lib.orc_last_error().decode()
        return (out, cn) if counters else out

    def render(self, cam, seed=1, rows=None, threads=0, want_rgb8=True):
        H, W = cam.image_height, cam.image_width
        rb, re_, rs = rows if rows else (0, H, 1)
        rgb = np.zeros((H, W, 3), np.float64)
        rgb8 = np.zeros((H, W, 3), np.uint8) if want_rgb8 else None
        st = OrcRenderStats()
        threads = threads or (os.cpu_count() or 1)
        rc = self.lib.orc_render(self.handle, C.byref(cam), seed, rb, re_, rs, threads, _p(rgb),
                                 _p(rgb8) if want_rgb8 else None, C.byref(st))
        assert rc == 0, self.lib.orc_last_error().decode()
        return rgb, rgb8, st.as_dict()

    def gen_rays(self, cam, kind, n, seed=7):
        out = np.zeros((n, 7), np.float64)
        rc = self.lib.orc_gen_rays(self.handle, C.byref(cam), seed, kind, n, _p(out))
        assert rc == 0
        return out

    def scatter(self, ray_in, hit, seed, pixel, sample, bounce):
        att, ray_out = np.zeros(3), np.zeros(7)
        h = np.zeros(1, dtype=abi.HIT_DTYPE)
        h[0] = hit
        ok = self.lib.orc_kat_scatter(self.handle, _p(np.ascontiguousarray(ray_in, np.float64)), _p(h), seed, pixel, sample,
                                      bounce, _p(att), _p(ray_out))
        return bool(ok), att, ray_out

    def tex_value(self, tex, u, v, p):
        out = np.zeros(3)
        self.lib.orc_kat_tex_value(self.handle, tex, u, v, _p(_d(*p)), _p(out))
        return out

    def sky(self, d):
        out = np.zeros(3)
        self.lib.orc_kat_sky(self.handle, _p(_d(*d)), _p(out))
        return out


# ---- thin KAT helpers
def vec(op, a, b=None):
    out = np.zeros(3)
    a = _d(*a)
    bb = _d(*b) if b is not None else None
    load().orc_kat_vec(op, _p(a), _p(bb) if bb is not None else None, _p(out))
    return out


def color_bytes(c):
    out = np.zeros(3, np.uint32)
    load().orc_kat_color_bytes(_p(_d(*c)), _p(out))
    return out


def color_neg(c):
    out = np.zeros(3)
    load().orc_kat_color_neg(_p(_d(*c)), _p(out))
    return out


def color_add(a, b):
    out = np.zeros(3)
    load().orc_kat_color_add(_p(_d(*a)), _p(_d(*b)), _p(out))
    return out


def color_scale(s, a):
    out = np.zeros(3)
    load().orc_kat_color_scale(s, _p(_d(*a)), _p(out))
    return out


def average(cols):
    cols = np.ascontiguousarray(cols, np.float64)
    out = np.zeros(3)
    load().orc_kat_average(_p(cols), len(cols), _p(out))
    return out


def interval(op, lo, hi, x=0.0):
    return load().orc_kat_interval(op, lo, hi, x)


def ray_at(o, d, t):
    out = np.zeros(3)
    load().orc_kat_ray_at(_p(_d(*o)), _p(_d(*d)), t, _p(out))
    return out


def point_at(init, keys, n, t):
    out = np.zeros(3)
    load().orc_kat_point_at(_p(_d(*init)), C.cast(keys, C.c_void_p), n, t, _p(out))
    return out


def philox(ctr, key):
    out = np.zeros(4, np.uint32)
    load().orc_philox4x32_10(_p(np.array(ctr, np.uint32)), _p(np.array(key, np.uint32)), _p(out))
    return out


def rng_stream(seed, pixel, sample, bounce, n):
    out = np.zeros(n)
    load().orc_rng_stream(seed, pixel, sample, bounce, n, _p(out))
    return out


def prim_hit(kind, prim, ray, tmin=0.001, tmax=float("inf")):
    h = np.zeros(1, dtype=abi.HIT_DTYPE)
    got = load().orc_kat_prim_hit(kind, _p(_d(*prim)), _p(_d(*ray)), tmin, tmax, _p(h))
    return bool(got), h[0]


def aabb_hit(box, ray, tmin, tmax):
    return bool(load().orc_kat_aabb_hit(_p(_d(*box)), _p(_d(*ray)), tmin, tmax))
