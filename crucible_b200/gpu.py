"""The GPU backend behind the reference's `Camera::render` boundary (src/camera/mod.rs:270-317).

`GpuScene` owns one CrScene on one CUDA device and exposes the two hot-path entry points of the
C ABI (`trace_batch`, `render`).  There is no CPU fallback: if the library is missing or no sm_100
device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi
from .scene import Scene, SceneDesc


class GpuScene:
    def __init__(self, desc: SceneDesc, device: int = 0, bvh_builder: int = abi.CR_BVH_AUTO):
        self.lib = abi.load()
        self.device = device
        self.handle = self.lib.cr_scene_create(device)
        if not self.handle:
            raise abi.CrucibleError(abi.CR_ERR_NO_DEVICE, self.lib.cr_last_error().decode())
        self.desc = desc
        try:
            abi.check(self.lib.cr_scene_set_bvh_builder(self.handle, int(bvh_builder)))
            desc.apply(self.lib, self.handle, "cr_")
        except Exception:
            self.close()
            raise

    def close(self):
        if getattr(self, "handle", None):
            self.lib.cr_scene_destroy(self.handle)
            self.handle = None

    def save(self, path):
        """The scene export file (cr_scene_save): everything this scene was given, one binary file."""
        abi.check(self.lib.cr_scene_save(self.handle, os.fsencode(path)))

    @staticmethod
    def load(path, device: int = 0) -> "GpuScene":
        """A committed scene from an export file (cr_scene_load)."""
        lib = abi.load()
        h = lib.cr_scene_load(os.fsencode(path), device)
        if not h:
            raise abi.CrucibleError(abi.CR_ERR_INVALID, lib.cr_last_error().decode())
        other = object.__new__(GpuScene)
        other.lib, other.device, other.handle, other.desc = lib, device, h, None
        return other

    def replicate(self, device: int) -> "GpuScene":
        """A committed copy of this scene on another device (cr_scene_replicate): the replicas of cr_render_multi."""
        h = self.lib.cr_scene_replicate(self.handle, device)
        if not h:
            raise abi.CrucibleError(abi.CR_ERR_CUDA, self.lib.cr_last_error().decode())
        other = object.__new__(GpuScene)
        other.lib, other.device, other.handle, other.desc = self.lib, device, h, self.desc
        return other

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

    def bvh_nodes(self):
        """The committed BVH in preorder as a structured array (abi.BVH_NODE_DTYPE)."""
        n = abi.check(self.lib.cr_scene_bvh_nodes(self.handle, None, 0))
        out = np.zeros(n, dtype=abi.BVH_NODE_DTYPE)
        abi.check(self.lib.cr_scene_bvh_nodes(self.handle, out.ctypes.data_as(C.c_void_p), n))
        return out

    def device_records(self, which):
        """Raw bytes of the flattened device records (0/1 = nodes f64/f32, 2/3 = triangles f64/f32)."""
        n = abi.check(self.lib.cr_scene_device_records(self.handle, which, None, 0))
        out = np.zeros(n, np.uint8)
        abi.check(self.lib.cr_scene_device_records(self.handle, which, out.ctypes.data_as(C.c_void_p), n))
        return out

    def commit_info(self):
        ci = abi.CrCommitInfo()
        abi.check(self.lib.cr_scene_commit_info(self.handle, C.byref(ci)))
        return ci.as_dict()

    # ---- Hittables::hit on a ray batch
    def trace_batch(self, rays, tmin=0.001, tmax=float("inf"), precision=abi.CR_PRECISION_F64, reference_order=False):
        """reference_order: the reference's own DFS throughout (CR_TRACE_REFERENCE_ORDER) instead of the order-free
        engine + reference-order retries.  `last_retried()` tells how many rays the order-free engine handed back."""
        rays = np.ascontiguousarray(rays, np.float64)
        assert rays.ndim == 2 and rays.shape[1] == 7, "rays = [n][7] (origin, direction, time)"
        out = np.zeros(len(rays), dtype=abi.HIT_DTYPE)
        flags = int(precision) | (abi.CR_TRACE_REFERENCE_ORDER if reference_order else 0)
        abi.check(self.lib.cr_trace_batch(self.handle, rays.ctypes.data_as(C.c_void_p), len(rays), float(tmin), float(tmax),
                                          flags, out.ctypes.data_as(C.c_void_p)))
        return out

    def last_retried(self):
        return int(self.lib.cr_scene_last_retried(self.handle))

    # ---- Camera::render sample loop, host buffers (H2D of the camera, D2H of the framebuffer inside)
    def render(self, cam: abi.CrCamera, seed=1, precision=abi.CR_PRECISION_F64, pool_paths=0, row_block=8, row_rank=0,
               row_world=1, time_kernels=False, want_rgb=True, want_rgb8=True, out_rgb=None, out_rgb8=None, reference_order=False):
        H, W = cam.image_height, cam.image_width
        if want_rgb and out_rgb is None:
            out_rgb = np.zeros((H, W, 3), np.float64)
        if want_rgb8 and out_rgb8 is None:
            out_rgb8 = np.zeros((H, W, 3), np.uint8)
        opts = abi.CrRenderOpts(seed, precision, pool_paths, row_block, row_rank, row_world, 1 if time_kernels else 0,
                                abi.CR_RENDER_REFERENCE_ORDER if reference_order else 0)
        st = abi.CrStats()
        abi.check(self.lib.cr_render(self.handle, C.byref(cam), C.byref(opts),
                                     out_rgb.ctypes.data_as(C.c_void_p) if out_rgb is not None else None,
                                     out_rgb8.ctypes.data_as(C.c_void_p) if out_rgb8 is not None else None, C.byref(st)))
        return out_rgb, out_rgb8, st.as_dict()

    # ---- device variant: packed rows of this rank into caller-owned device memory, on the caller's stream
    def render_device(self, cam: abi.CrCamera, d_out_rgb: int, d_out_rgb8: int, stream: int = 0, seed=1,
                      precision=abi.CR_PRECISION_F64, pool_paths=0, row_block=8, row_rank=0, row_world=1, time_kernels=False,
                      global_rows=False, reference_order=False):
        """global_rows: the outputs are full [H][W][3] images (possibly on another device / in another process' buffer)
        and this rank's rows are stored at their global position (CR_RENDER_GLOBAL_ROWS)."""
        opts = abi.CrRenderOpts(seed, precision, pool_paths, row_block, row_rank, row_world, 1 if time_kernels else 0,
                                (abi.CR_RENDER_GLOBAL_ROWS if global_rows else 0) | (abi.CR_RENDER_REFERENCE_ORDER if reference_order else 0))
        st = abi.CrStats()
        abi.check(self.lib.cr_render_device(self.handle, C.byref(cam), C.byref(opts), C.c_void_p(d_out_rgb or None),
                                            C.c_void_p(d_out_rgb8 or None), C.c_void_p(stream or None), C.byref(st)))
        return st.as_dict()


def render_multi(replicas, cam: abi.CrCamera, seed=1, precision=abi.CR_PRECISION_F64, pool_paths=0, row_block=8,
                 time_kernels=False, want_rgb=True, want_rgb8=True, out_rgb=None, out_rgb8=None):
    """Camera::render over several devices of ONE process (cr_render_multi): `replicas` = GpuScene copies of one scene on
    different devices (GpuScene.replicate); rows are sharded in interleaved blocks, every device's resolve kernel stores
    its rows into the image on replicas[0]'s device, one device-to-host copy.  Returns (rgb, rgb8, [stats per replica])."""
    H, W = cam.image_height, cam.image_width
    if want_rgb and out_rgb is None:
        out_rgb = np.zeros((H, W, 3), np.float64)
    if want_rgb8 and out_rgb8 is None:
        out_rgb8 = np.zeros((H, W, 3), np.uint8)
    n = len(replicas)
    handles = (C.c_void_p * n)(*[r.handle for r in replicas])
    opts = abi.CrRenderOpts(seed, precision, pool_paths, row_block, 0, 1, 1 if time_kernels else 0, 0)
    stats = (abi.CrStats * n)()
    abi.check(replicas[0].lib.cr_render_multi(handles, n, C.byref(cam), C.byref(opts),
                                               out_rgb.ctypes.data_as(C.c_void_p) if out_rgb is not None else None,
                                               out_rgb8.ctypes.data_as(C.c_void_p) if out_rgb8 is not None else None,
                                               C.cast(stats, C.c_void_p)))
    return out_rgb, out_rgb8, [stats[i].as_dict() for i in range(n)]


def rows_of_rank(height, row_block, rank, world):
    """Global rows rendered by `rank`: (j // row_block) % world == rank, ascending."""
    j = np.arange(height)
    return j[(j // row_block) % max(world, 1) == rank] if world > 1 else j


def write_ppm(fname, rgb8, fmt=abi.CR_PPM_P3):
    """The file tail of Camera::render (camera/mod.rs:275-311): P3 header, one "r g b" line per pixel
    (`fmt=abi.CR_PPM_P6`: binary PPM, `abi.CR_PNG`: PNG — extensions).  Formatting is done by the library (cr_write_ppm)."""
    rgb8 = np.ascontiguousarray(rgb8, np.uint8)
    h, w, _ = rgb8.shape
    abi.check(abi.load().cr_write_ppm(os.fsencode(fname), rgb8.ctypes.data_as(C.c_void_p), w, h, int(fmt)))


def write_ppm_python(fname, rgb8):
    """Plain-Python statement of the same file format (tests compare cr_write_ppm against it)."""
    h, w, _ = rgb8.shape
    flat = rgb8.reshape(-1, 3)
    with open(fname, "w") as f:
        f.write(f"P3\n{w} {h}\n255\n")
        f.write("\n".join(f"{r} {g} {b}" for r, g, b in flat.tolist()))
        f.write("\n")


def render_scene(scene: Scene, fname: str, device=0, seed=1, precision=abi.CR_PRECISION_F64, gpu_scene=None,
                 fmt=abi.CR_PPM_P3, rank=0, world=1, make_movie=True):
    """Scene::render_scene (scene/mod.rs:283-347).  Still: `Camera::render` into `<fname>.ppm`
    (cr_render_to_file).  Movie: `ceil(duration * rate)` frames into `<fname>/artifacts/imageNNN.ppm`
    through the pipelined frame loop (cr_render_frames; frames f % world == rank on this GPU), then the
    reference's ffmpeg hand-off (movie_maker.rs:6-33) when `ffmpeg` exists.  Returns the per-frame stats."""
    own = gpu_scene is None
    gs = gpu_scene or GpuScene(scene.describe(), device)
    try:
        cam = scene.scene_cam.to_abi()
        opts = abi.CrRenderOpts(seed, precision, 0, 8, 0, 1, 0)
        if scene.duration is None:
            st = abi.CrStats()
            abi.check(gs.lib.cr_render_to_file(gs.handle, C.byref(cam), C.byref(opts), os.fsencode(fname + ".ppm"), int(fmt), C.byref(st)))
            return [st.as_dict()]
        art = os.path.join(fname, "artifacts")
        if rank == 0:
            os.makedirs(art)  # fs::create_dir panics when the directory exists (scene/mod.rs:296-298)
        frames = scene.compute_frame_count()
        digits = len(str(frames))
        stats = render_frames(gs, cam, art, frames, rank, world, seed, precision, fmt)
        for _ in range(frames):
            scene.scene_cam.next_frame()  # camera/mod.rs:160-162: the camera ends up past the last frame
        if make_movie and world == 1:
            make_mp4(scene.frame_rate, digits, fname)
        return stats
    finally:
        if own:
            gs.close()


def render_frames(gs: GpuScene, cam: abi.CrCamera, artifacts_dir: str, frames: int, rank=0, world=1, seed=1,
                  precision=abi.CR_PRECISION_F64, fmt=abi.CR_PPM_P3, pool_paths=0):
    """Frames rank, rank+world, ... (< frames) of a movie through the pipelined loop (cr_render_frames)."""
    opts = abi.CrRenderOpts(seed, precision, pool_paths, 8, 0, 1, 0)
    n = len(range(rank, frames, max(world, 1)))
    stats = (abi.CrStats * max(n, 1))()
    abi.check(gs.lib.cr_render_frames(gs.handle, C.byref(cam), C.byref(opts), rank, max(world, 1), frames, os.fsencode(artifacts_dir),
                                      len(str(frames)), int(fmt), C.cast(stats, C.c_void_p)))
    return [stats[i].as_dict() for i in range(n)]


def make_mp4(frame_rate, padding, fname):
    """movie_maker::make_mp4 (scene/movie_maker.rs:6-33): the same ffmpeg command line; skipped (returns
    False) when no ffmpeg binary is on PATH."""
    import shutil
    import subprocess

    if shutil.which("ffmpeg") is None:
        return False
    subprocess.run(["ffmpeg", "-framerate", str(frame_rate), "-i", f"{fname}/artifacts/image%0{padding}d.ppm", "-vf",
                    "scale=trunc(iw/2)*2:trunc(ih/2)*2", "-c:v", "libx264", "-pix_fmt", "yuv420p", "-crf", "25", f"{fname}/movie.mp4"],
                   check=True)
    return True
