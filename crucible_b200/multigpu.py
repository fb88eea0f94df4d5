"""Multi-GPU sharding of the render: one process per GPU, scene replicated, image rows
sharded in interleaved blocks (stills) or whole frames (timeline animations).  The path has no data-path
collective: every rank traces its own pixels; the only exchange is the framebuffer.  Two exchanges:

  "p2p"  (default on GPUs): rank 0 owns the image in a buffer every rank has mapped (cr_shared_buffer_*: CUDA IPC);
         each rank's resolve kernel stores its rows straight into it over NVLink (CR_RENDER_GLOBAL_ROWS), and a
         barrier closes the frame.  No gather, no row permutation, no second copy.
  "nccl" : packed rows + dist.gather + one row-permutation kernel on rank 0 (also what the gloo CPU test runs).

The RNG is keyed by the GLOBAL pixel index, so the assembled image is bit-identical for any world size.
(A single process driving several devices uses cr_render_multi instead: crucible_b200.gpu.render_multi.)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import abi
from .gpu import GpuScene, rows_of_rank


def frames_of_rank(n_frames: int, rank: int, world: int):
    """Whole-frame sharding for `render_movie` (scene/mod.rs:295-322): frame f -> rank f % world."""
    return list(range(rank, n_frames, max(world, 1)))


_GATHER_MAP_CACHE = {}


def _gather_map(height, row_block, world, max_rows, device):
    """For every global row j: its position in the stacked gather buffer [world][max_rows] (rank r's packed rows
    are the rows with (j // row_block) % world == r in ascending order).  Cached device tensor."""
    key = (height, row_block, world, max_rows, str(device))
    t = _GATHER_MAP_CACHE.get(key)
    if t is None:
        j = np.arange(height)
        r = (j // row_block) % world
        pos = (j // (row_block * world)) * row_block + j % row_block
        t = torch.as_tensor(r * max_rows + pos, device=device, dtype=torch.long)
        _GATHER_MAP_CACHE[key] = t
    return t


def gather_rows(local: torch.Tensor, height: int, row_block: int, rank: int, world: int, group=None, dst: int = 0):
    """Assemble the full [H][W][C] image on `dst` from each rank's packed rows.

    `local` is [rows_local][W][C] holding rows (j // row_block) % world == rank in ascending order.
    Returns the full tensor on dst, None elsewhere.  world == 1 returns `local` unchanged.
    One collective (gather into the slices of one stacked buffer) and ONE row-permutation kernel on dst."""
    if world <= 1:
        return local
    counts = [len(rows_of_rank(height, row_block, r, world)) for r in range(world)]
    assert local.shape[0] == counts[rank], (local.shape, counts, rank)
    max_rows = max(counts)
    send = local
    if local.shape[0] != max_rows:  # NCCL gather wants equal shapes: pad the short ranks
        send = torch.zeros((max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        send[: local.shape[0]] = local
    send = send.contiguous()
    if rank == dst:
        stacked = torch.empty((world, max_rows) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.gather(send, gather_list=list(stacked.unbind(0)), dst=dst, group=group)
        return stacked.view((world * max_rows,) + tuple(local.shape[1:])).index_select(
            0, _gather_map(height, row_block, world, max_rows, local.device))
    dist.gather(send, gather_list=None, dst=dst, group=group)
    return None


class _DevicePointer:
    """A raw device pointer as a __cuda_array_interface__ object (torch.as_tensor turns it into a tensor view)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class SharedImage:
    """[H][W][3] f64 + u8 images on rank `dst`'s GPU that every rank of the group has mapped (CUDA IPC through
    cr_shared_buffer_*).  Collective constructor: every rank of the group calls it."""

    def __init__(self, height, width, device, rank, world, group=None, dst=0):
        import ctypes as C

        self.lib, self.device, self.owner = abi.load(), device, rank == dst
        self.H, self.W = height, width
        self.off8 = (height * width * 3 * 8 + 255) & ~255
        total = self.off8 + height * width * 3
        ptr, handle = C.c_void_p(), (C.c_ubyte * 64)()
        if self.owner:
            abi.check(self.lib.cr_shared_buffer_create(device, total, C.byref(ptr), handle))
        box = [bytes(handle) if self.owner else None]
        dist.broadcast_object_list(box, src=dst, group=group)
        if not self.owner:
            h = (C.c_ubyte * 64).from_buffer_copy(box[0])
            abi.check(self.lib.cr_shared_buffer_open(device, h, C.byref(ptr)))
        self.ptr = int(ptr.value)
        self.ptr64, self.ptr8 = self.ptr, self.ptr + self.off8
        self.rgb = self.rgb8 = None
        if self.owner:
            dev = torch.device("cuda", device)
            self.rgb = torch.as_tensor(_DevicePointer(self.ptr64, (height, width, 3), "<f8"), device=dev)
            self.rgb8 = torch.as_tensor(_DevicePointer(self.ptr8, (height, width, 3), "|u1"), device=dev)

    def close(self):
        if self.ptr:
            self.lib.cr_shared_buffer_close(self.device, abi.C.c_void_p(self.ptr), 1 if self.owner else 0)
            self.ptr = 0


_SHARED = {}


def shared_image(height, width, device, rank, world, group=None):
    key = (height, width, device, rank, world, id(group))
    if key not in _SHARED:
        _SHARED[key] = SharedImage(height, width, device, rank, world, group)
    return _SHARED[key]


def render_sharded(gs: GpuScene, cam: abi.CrCamera, rank: int, world: int, seed=1, precision=abi.CR_PRECISION_F64,
                   row_block=8, pool_paths=0, time_kernels=False, group=None, want_rgb8=True, want_rgb=True, exchange="p2p"):
    """Render this rank's rows on its GPU and assemble the framebuffer on rank 0 (see the module docstring).

    Returns (rgb [H][W][3] f64 device tensor on rank 0 else None, rgb8 likewise, stats dict)."""
    H, W = cam.image_height, cam.image_width
    dev = torch.device("cuda", gs.device)
    if world > 1 and exchange == "p2p":
        img = shared_image(H, W, gs.device, rank, world, group)
        stream = torch.cuda.current_stream(dev).cuda_stream
        st = gs.render_device(cam, img.ptr64 if want_rgb else 0, img.ptr8 if want_rgb8 else 0, stream=stream, seed=seed,
                              precision=precision, pool_paths=pool_paths, row_block=row_block, row_rank=rank, row_world=world,
                              time_kernels=time_kernels, global_rows=True)
        # cr_render_device returns after this rank's stream has drained: its peer stores have landed.  The barrier
        # tells rank 0 that every rank is there.
        dist.barrier(group=group)
        return (img.rgb if want_rgb else None), (img.rgb8 if want_rgb8 else None), st
    rows = rows_of_rank(H, row_block, rank, world)
    local = torch.empty((len(rows), W, 3), dtype=torch.float64, device=dev) if want_rgb else None
    local8 = torch.empty((len(rows), W, 3), dtype=torch.uint8, device=dev) if want_rgb8 else None
    stream = torch.cuda.current_stream(dev).cuda_stream
    st = gs.render_device(cam, local.data_ptr() if want_rgb else 0, local8.data_ptr() if want_rgb8 else 0, stream=stream, seed=seed,
                          precision=precision, pool_paths=pool_paths, row_block=row_block, row_rank=rank,
                          row_world=world, time_kernels=time_kernels)
    full = gather_rows(local, H, row_block, rank, world, group) if want_rgb else None
    full8 = gather_rows(local8, H, row_block, rank, world, group) if want_rgb8 else None
    return full, full8, st


def render_frames_sharded(gs: GpuScene, scene, rank: int, world: int, seed=1, precision=abi.CR_PRECISION_F64,
                          pool_paths=0, on_frame=None):
    """Timeline animation: each rank renders frames rank, rank+world, ...; no inter-GPU traffic.
    `on_frame(frame, rgb8)` receives each finished frame (host array) of this rank."""
    n = scene.compute_frame_count()
    stats = []
    cam = scene.scene_cam.to_abi()
    for f in frames_of_rank(n, rank, world):
        cam.frame = f  # Camera::next_frame advances by one per rendered image (camera/mod.rs:160-162)
        _, rgb8, st = gs.render(cam, seed=seed, precision=precision, pool_paths=pool_paths, want_rgb=False)
        if on_frame is not None:
            on_frame(f, rgb8)
        stats.append(st)
    return stats


def np_rows(full: np.ndarray, row_block: int, rank: int, world: int):
    """Packed rows of `rank` taken from a full image (host helper for tests)."""
    return full[rows_of_rank(full.shape[0], row_block, rank, world)]
