"""ctypes mirror of include/crucible_gpu.h (POD structs + enum values) and the library loader.

The product path has NO CPU fallback: `load()` raises if libcrucible_b200.so is missing, and every
compute entry point of the library itself fails with CR_ERR_NO_DEVICE when there is no sm_100 GPU.
"""
from __future__ import annotations

import ctypes as C
import os

CR_OK = 0
CR_ERR_INVALID, CR_ERR_NO_DEVICE, CR_ERR_CUDA, CR_ERR_STATE, CR_ERR_LIMIT = -1, -2, -3, -4, -5

CR_MAT_LAMBERTIAN, CR_MAT_METAL, CR_MAT_DIELECTRIC, CR_MAT_EMISSIVE = 0, 1, 2, 3
CR_TEX_SOLID, CR_TEX_CHECKER, CR_TEX_IMAGE = 0, 1, 2
CR_SKY_DEFAULT, CR_SKY_SPHERICAL, CR_SKY_BLACK = 0, 1, 2
CR_PRIM_SPHERE, CR_PRIM_TRIANGLE, CR_PRIM_QUAD = 0, 1, 2
CR_NERP, CR_LERP = 0, 1
CR_PRECISION_F64, CR_PRECISION_F32 = 0, 1
CR_MAX_CAM_KEYS = 32
CR_PPM_P3, CR_PPM_P6, CR_PNG = 0, 1, 2
CR_BVH_AUTO, CR_BVH_HOST, CR_BVH_DEVICE = 0, 1, 2
CR_RENDER_GLOBAL_ROWS = 1
CR_RENDER_REFERENCE_ORDER = 2
CR_TRACE_REFERENCE_ORDER = 0x100


class CrMaterial(C.Structure):
    _fields_ = [("kind", C.c_int32), ("tex", C.c_int32), ("scatter_prob", C.c_double), ("albedo", C.c_double * 3),
                ("fuzz", C.c_double), ("ior", C.c_double), ("emit", C.c_double * 3)]


class CrTexture(C.Structure):
    _fields_ = [("kind", C.c_int32), ("even", C.c_int32), ("odd", C.c_int32), ("image", C.c_int32),
                ("color", C.c_double * 3), ("inv_scale", C.c_double)]


class CrKeyframe(C.Structure):
    _fields_ = [("t0", C.c_double), ("t1", C.c_double), ("delta", C.c_double), ("axis", C.c_int32), ("interp", C.c_int32)]


class CrAnimKey(C.Structure):
    _fields_ = [("t0", C.c_double), ("t1", C.c_double), ("a", C.c_double), ("b", C.c_double), ("kind", C.c_int32), ("interp", C.c_int32)]


class CrCamera(C.Structure):
    _fields_ = [("image_width", C.c_uint32), ("image_height", C.c_uint32),
                ("viewport_width", C.c_double), ("viewport_height", C.c_double),
                ("focus_dist", C.c_double), ("defocus_angle", C.c_double), ("defocus_radius", C.c_double),
                ("vup", C.c_double * 3), ("look_from", C.c_double * 3), ("look_at", C.c_double * 3),
                ("frame_rate", C.c_double), ("frame", C.c_uint32), ("samples", C.c_uint32), ("max_depth", C.c_uint32),
                ("n_from_keys", C.c_uint32), ("n_at_keys", C.c_uint32), ("shutter_angle", C.c_double),
                ("from_keys", CrKeyframe * CR_MAX_CAM_KEYS), ("at_keys", CrKeyframe * CR_MAX_CAM_KEYS)]


class CrRenderOpts(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("precision", C.c_int32), ("pool_paths", C.c_uint32),
                ("row_block", C.c_uint32), ("row_rank", C.c_uint32), ("row_world", C.c_uint32), ("time_kernels", C.c_uint32),
                ("flags", C.c_uint32)]


CR_GROUP_HITLIST, CR_GROUP_BVH = 0, 1
# SceneDesc.batches markers (not ABI values): a nested element opens / closes between primitive batches
GROUP_BEGIN, GROUP_END = -1, -2


class CrStats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("rays", C.c_uint64), ("iterations", C.c_uint64), ("launches", C.c_uint64),
                ("ms_total", C.c_double), ("ms_trace", C.c_double), ("ms_shade", C.c_double), ("ms_raygen", C.c_double),
                ("ms_resolve", C.c_double), ("ms_h2d", C.c_double), ("ms_d2h", C.c_double), ("retried_rays", C.c_uint64),
                ("trace_engine", C.c_uint32), ("reserved", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class CrHit(C.Structure):
    _fields_ = [("prim_index", C.c_int32), ("obj_id", C.c_int32), ("front_face", C.c_int32), ("material", C.c_int32),
                ("t", C.c_double), ("p", C.c_double * 3), ("n", C.c_double * 3), ("u", C.c_double), ("v", C.c_double)]


class CrCommitInfo(C.Structure):
    _fields_ = [("builder", C.c_int32), ("levels", C.c_uint32), ("ms_total", C.c_double), ("ms_build", C.c_double),
                ("ms_pack", C.c_double), ("ms_h2d", C.c_double), ("ms_device", C.c_double), ("ms_d2h", C.c_double),
                ("ms_upload", C.c_double), ("ms_search_tree", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# numpy view of CrBvhNode[]
BVH_NODE_DTYPE = [("lo", "<f8", (3,)), ("hi", "<f8", (3,)), ("left", "<u4"), ("right", "<u4"), ("axis", "<u4"), ("skip", "<u4")]

# numpy view of CrHit[] (same layout, checked in tests)
HIT_DTYPE = [("prim_index", "<i4"), ("obj_id", "<i4"), ("front_face", "<i4"), ("material", "<i4"), ("t", "<f8"),
             ("p", "<f8", (3,)), ("n", "<f8", (3,)), ("u", "<f8"), ("v", "<f8")]

_P = C.c_void_p
_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int32)

# every symbol include/crucible_gpu.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "cr_device_count": (C.c_int, []),
    "cr_scene_create": (_P, [C.c_int]),
    "cr_scene_destroy": (None, [_P]),
    "cr_device_trim": (C.c_int, [C.c_int]),
    "cr_last_error": (C.c_char_p, []),
    "cr_version": (C.c_char_p, []),
    "cr_scene_reserve": (C.c_int, [_P, C.c_size_t, C.c_size_t, C.c_size_t]),
    "cr_scene_add_spheres": (C.c_int64, [_P, _P, _P, _P, C.c_size_t]),
    "cr_scene_add_triangles": (C.c_int64, [_P, _P, _P, _P, C.c_size_t]),
    "cr_scene_add_quads": (C.c_int64, [_P, _P, _P, _P, C.c_size_t]),
    "cr_scene_add_batches": (C.c_int64, [_P, C.c_size_t, _P, _P, _P, _P, _P]),
    "cr_scene_begin_group": (C.c_int, [_P, C.c_int]),
    "cr_scene_end_group": (C.c_int, [_P]),
    "cr_scene_set_hidden": (C.c_int, [_P, C.c_size_t, C.c_int]),
    "cr_scene_set_materials": (C.c_int, [_P, _P, C.c_size_t]),
    "cr_scene_set_textures": (C.c_int, [_P, _P, C.c_size_t]),
    "cr_scene_add_image": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "cr_scene_set_sky": (C.c_int, [_P, C.c_int, C.c_int]),
    "cr_scene_set_keyframes": (C.c_int, [_P, C.c_size_t, C.c_int, _P, C.c_size_t]),
    "cr_scene_set_bvh_builder": (C.c_int, [_P, C.c_int]),
    "cr_scene_commit": (C.c_int, [_P]),
    "cr_scene_commit_info": (C.c_int, [_P, C.POINTER(CrCommitInfo)]),
    "cr_scene_bvh_nodes": (C.c_int64, [_P, _P, C.c_size_t]),
    "cr_scene_device_records": (C.c_int64, [_P, C.c_int, _P, C.c_size_t]),
    "cr_scene_bvh_info": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
    "cr_scene_bvh_leaf_order": (C.c_int64, [_P, _P, C.c_size_t]),
    "cr_trace_batch": (C.c_int, [_P, _P, C.c_size_t, C.c_double, C.c_double, C.c_int, _P]),
    "cr_scene_last_retried": (C.c_int64, [_P]),
    "cr_render": (C.c_int, [_P, C.POINTER(CrCamera), C.POINTER(CrRenderOpts), _P, _P, C.POINTER(CrStats)]),
    "cr_render_device": (C.c_int, [_P, C.POINTER(CrCamera), C.POINTER(CrRenderOpts), _P, _P, _P, C.POINTER(CrStats)]),
    "cr_scene_replicate": (_P, [_P, C.c_int]),
    "cr_render_multi": (C.c_int, [_P, C.c_int, C.POINTER(CrCamera), C.POINTER(CrRenderOpts), _P, _P, _P]),
    "cr_shared_buffer_create": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(C.c_void_p), _P]),
    "cr_shared_buffer_open": (C.c_int, [C.c_int, _P, C.POINTER(C.c_void_p)]),
    "cr_shared_buffer_close": (C.c_int, [C.c_int, _P, C.c_int]),
    "cr_write_ppm": (C.c_int, [C.c_char_p, _P, C.c_uint32, C.c_uint32, C.c_int]),
    "cr_scene_save": (C.c_int, [_P, C.c_char_p]),
    "cr_scene_load": (_P, [C.c_char_p, C.c_int]),
    "cr_render_to_file": (C.c_int, [_P, C.POINTER(CrCamera), C.POINTER(CrRenderOpts), C.c_char_p, C.c_int, C.POINTER(CrStats)]),
    "cr_render_frames": (C.c_int, [_P, C.POINTER(CrCamera), C.POINTER(CrRenderOpts), C.c_uint32, C.c_uint32, C.c_uint32, C.c_char_p,
                                   C.c_uint32, C.c_int, _P]),
    "cr_camera_point_at": (C.c_int, [_PD, _P, C.c_size_t, C.c_double, _PD]),
    "cr_anim_point_at": (C.c_int, [_PD, _P, C.c_size_t, C.c_double, _PD]),
    "cr_measure_fma_peak": (C.c_int, [C.c_int, _PD, _PD]),
    "cr_philox4x32_10": (None, [_P, _P, _P]),
}

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcrucible_b200.so")
_lib = None


class CrucibleError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"crucible_b200 error {code}: {message}")
        self.code = code


def load():
    """Load libcrucible_b200.so.  Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CrucibleError(CR_ERR_STATE, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc < 0:
        raise CrucibleError(rc, load().cr_last_error().decode())
    return rc
