"""The GPU backend behind the reference's `Camera::render` boundary (src/camera/mod.rs:270-317).

`GpuScene` owns one CrScene on one CUDA device and exposes the two hot-path entry points of the
C ABI (`trace_batch`, `render`).  There is no CPU fallback: if the library is missing or no sm_100
device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from .scene import Scene, SceneDesc


class GpuScene:
    def __init__(self, desc: SceneDesc, device: int = 0):
        self.lib = abi.load()
        self.device = device
        self.handle = self.lib.cr_scene_create(device)
        if not self.handle:
            raise abi.CrucibleError(abi.CR_ERR_NO_DEVICE, self.lib.cr_last_error().decode())
        self.desc = desc
        try:
            desc.apply(self.lib, self.handle, "cr_")
        except Exception:
            self.close()
            raise

    def close(self):
        if getattr(self, "handle", None):
            self.lib.cr_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- introspection of the committed BVH
    def bvh_info(self):
        n, d, v = C.c_uint64(), C.c_uint32(), C.c_uint64()
        abi.check(self.lib.cr_scene_bvh_info(self.handle, C.byref(n), C.byref(d), C.byref(v)))
        return {"n_nodes": n.value, "max_depth": d.value, "n_visible": v.value}

    def bvh_leaf_order(self):
        n = abi.check(self.lib.cr_scene_bvh_leaf_order(self.handle, None, 0))
        out = np.empty(n, np.int32)
        abi.check(self.lib.cr_scene_bvh_leaf_order(self.handle, out.ctypes.data_as(C.c_void_p), n))
        return out

    # ---- Hittables::hit on a ray batch
    def trace_batch(self, rays, tmin=0.001, tmax=float("inf"), precision=abi.CR_PRECISION_F64):
        rays = np.ascontiguousarray(rays, np.float64)
        assert rays.ndim == 2 and rays.shape[1] == 7, "rays = [n][7] (origin, direction, time)"
        out = np.zeros(len(rays), dtype=abi.HIT_DTYPE)
        abi.check(self.lib.cr_trace_batch(self.handle, rays.ctypes.data_as(C.c_void_p), len(rays), float(tmin), float(tmax),
                                          int(precision), out.ctypes.data_as(C.c_void_p)))
        return out

    # ---- Camera::render sample loop, host buffers (H2D of the camera, D2H of the framebuffer inside)
    def render(self, cam: abi.CrCamera, seed=1, precision=abi.CR_PRECISION_F64, pool_paths=0, row_block=8, row_rank=0,
               row_world=1, time_kernels=False, want_rgb=True, want_rgb8=True, out_rgb=None, out_rgb8=None):
        H, W = cam.image_height, cam.image_width
        if want_rgb and out_rgb is None:
            out_rgb = np.zeros((H, W, 3), np.float64)
        if want_rgb8 and out_rgb8 is None:
            out_rgb8 = np.zeros((H, W, 3), np.uint8)
        opts = abi.CrRenderOpts(seed, precision, pool_paths, row_block, row_rank, row_world, 1 if time_kernels else 0)
        st = abi.CrStats()
        abi.check(self.lib.cr_render(self.handle, C.byref(cam), C.byref(opts),
                                     out_rgb.ctypes.data_as(C.c_void_p) if out_rgb is not None else None,
                                     out_rgb8.ctypes.data_as(C.c_void_p) if out_rgb8 is not None else None, C.byref(st)))
        return out_rgb, out_rgb8, st.as_dict()

    # ---- device variant: packed rows of this rank into caller-owned device memory, on the caller's stream
    def render_device(self, cam: abi.CrCamera, d_out_rgb: int, d_out_rgb8: int, stream: int = 0, seed=1,
                      precision=abi.CR_PRECISION_F64, pool_paths=0, row_block=8, row_rank=0, row_world=1, time_kernels=False):
        opts = abi.CrRenderOpts(seed, precision, pool_paths, row_block, row_rank, row_world, 1 if time_kernels else 0)
        st = abi.CrStats()
        abi.check(self.lib.cr_render_device(self.handle, C.byref(cam), C.byref(opts), C.c_void_p(d_out_rgb or None),
                                            C.c_void_p(d_out_rgb8 or None), C.c_void_p(stream or None), C.byref(st)))
        return st.as_dict()


def rows_of_rank(height, row_block, rank, world):
    """Global rows rendered by `rank`: (j // row_block) % world == rank, ascending."""
    j = np.arange(height)
    return j[(j // row_block) % max(world, 1) == rank] if world > 1 else j


def write_ppm(fname, rgb8):
    """P3 PPM exactly as Camera::render writes it (camera/mod.rs:286, 306-311): header, one "r g b" line per pixel."""
    h, w, _ = rgb8.shape
    flat = rgb8.reshape(-1, 3)
    with open(fname, "w") as f:
        f.write(f"P3\n{w} {h}\n255\n")
        f.write("\n".join(f"{r} {g} {b}" for r, g, b in flat.tolist()))
        f.write("\n")


def render_scene(scene: Scene, fname: str, device=0, seed=1, precision=abi.CR_PRECISION_F64, gpu_scene=None):
    """Scene::render_scene (scene/mod.rs:283-347): BVH wrap per frame, Camera::render, `<fname>.ppm`;
    movies render `ceil(duration * rate)` frames into `<fname>/artifacts/imageNNN.ppm`."""
    import os

    own = gpu_scene is None
    gs = gpu_scene or GpuScene(scene.describe(), device)
    try:
        if scene.duration is None:
            _, rgb8, st = gs.render(scene.scene_cam.to_abi(), seed=seed, precision=precision, want_rgb=False)
            write_ppm(fname + ".ppm", rgb8)
            return [st]
        os.makedirs(os.path.join(fname, "artifacts"))
        frames = scene.compute_frame_count()
        digits = len(str(frames))
        stats = []
        for frame in range(frames):
            _, rgb8, st = gs.render(scene.scene_cam.to_abi(), seed=seed, precision=precision, want_rgb=False)
            write_ppm(os.path.join(fname, "artifacts", f"image{frame:0{digits}d}.ppm"), rgb8)
            scene.scene_cam.next_frame()
            stats.append(st)
        return stats
    finally:
        if own:
            gs.close()
