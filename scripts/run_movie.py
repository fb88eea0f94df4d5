"""BASELINE config 5: 240-frame 1080p walk-through of the book1 scene at 64 spp, whole frames sharded over
the ranks (frame f -> rank f % world), no inter-GPU traffic.  Launch with torchrun for N > 1.
Prints one JSON line on rank 0 (time = max over ranks, CUDA-synchronised wall clock around the frame loop)."""
import json
import os
import sys
import time

sys.path.insert(0, ".")
import torch
import torch.distributed as dist

from crucible_b200 import demo_builder, multigpu
from crucible_b200.gpu import GpuScene

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 240
sc = demo_builder.book1_walkthrough(image_width=1920, samples=64, duration=frames / 24.0)
assert sc.compute_frame_count() == frames
gs = GpuScene(sc.describe(), local)
cam = sc.scene_cam.to_abi()
gs.render(cam, seed=1, want_rgb=False)  # warm-up (workspace allocation)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
lum = []
stats = multigpu.render_frames_sharded(gs, sc, rank, world, seed=1, on_frame=lambda f, img: lum.append(float(img.mean())))
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
agg = torch.tensor([sum(s["samples"] for s in stats), sum(s["rays"] for s in stats), len(stats)], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dist.all_reduce(agg, op=dist.ReduceOp.SUM)
if rank == 0:
    s, r, n = agg.tolist()
    print(json.dumps({"config": "walkthrough", "note": f"configs[4]: {int(n)} frames 1920x1080 64 spp, frames sharded x{world}, D2H of every frame included",
                      "n_gpus": world, "seconds": dt.item(), "msamples_per_s": s / dt.item() / 1e6, "mrays_per_s": r / dt.item() / 1e6,
                      "frames_per_s": n / dt.item()}), flush=True)
if world > 1:
    dist.destroy_process_group()
