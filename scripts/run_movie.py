"""BASELINE config 5: 240-frame 1080p walk-through of the book1 scene at 64 spp, whole frames sharded over
the ranks (frame f -> rank f % world), no inter-GPU traffic.  Launch with torchrun for N > 1.

  python scripts/run_movie.py [frames] [p3|p6|none]

p3 / p6: Scene::render_movie's frame loop through cr_render_frames (render of frame k+1 overlapped with the
device-to-host copy, formatting and file write of frame k; files under /tmp/crucible_movie/artifacts).
none: frames to host arrays only (the earlier, unpipelined loop).
Prints one JSON line on rank 0 (time = max over ranks, CUDA-synchronised wall clock around the frame loop)."""
import json
import os
import shutil
import sys
import time

sys.path.insert(0, ".")
import torch
import torch.distributed as dist

from crucible_b200 import abi, demo_builder, gpu, multigpu
from crucible_b200.gpu import GpuScene

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 240
mode = sys.argv[2] if len(sys.argv) > 2 else "p3"
sc = demo_builder.book1_walkthrough(image_width=1920, samples=64, duration=frames / 24.0)
assert sc.compute_frame_count() == frames
gs = GpuScene(sc.describe(), local)
cam = sc.scene_cam.to_abi()
gs.render(cam, seed=1, want_rgb=False)  # warm-up (workspace allocation)
out = "/tmp/crucible_movie"
if rank == 0:
    shutil.rmtree(out, ignore_errors=True)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
if mode == "none":
    stats = multigpu.render_frames_sharded(gs, sc, rank, world, seed=1)
else:
    if rank == 0:
        os.makedirs(os.path.join(out, "artifacts"))
    if world > 1:
        dist.barrier()
    fmt = abi.CR_PPM_P3 if mode == "p3" else abi.CR_PPM_P6
    stats = gpu.render_frames(gs, cam, os.path.join(out, "artifacts"), frames, rank, world, seed=1, fmt=fmt)
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
agg = torch.tensor([sum(s["samples"] for s in stats), sum(s["rays"] for s in stats), len(stats), sum(s["ms_total"] for s in stats)],
                   dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dist.all_reduce(agg, op=dist.ReduceOp.SUM)
if rank == 0:
    s, r, n, ms = agg.tolist()
    print(json.dumps({"config": "walkthrough", "output": mode,
                      "note": f"configs[4]: {int(n)} frames 1920x1080 64 spp, frames sharded x{world}, every frame copied to the host"
                              + ("" if mode == "none" else " and written as a PPM file"),
                      "n_gpus": world, "seconds": dt.item(), "gpu_seconds_sum": ms / 1e3, "msamples_per_s": s / dt.item() / 1e6,
                      "mrays_per_s": r / dt.item() / 1e6, "frames_per_s": n / dt.item()}), flush=True)
if world > 1:
    dist.destroy_process_group()
