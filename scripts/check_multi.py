#!/usr/bin/env python
"""N-GPU checks that need more than one device (run with `gpurun --gpus N`):

  python scripts/check_multi.py                       # one process, cr_render_multi over all devices
  torchrun --nproc-per-node N scripts/check_multi.py  # one process per GPU, shared-buffer P2P exchange vs NCCL gather

Both compare the assembled image with a single-device render bit for bit and print timings."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from crucible_b200 import abi, demo_builder, gpu, multigpu
from crucible_b200.gpu import GpuScene

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
sc = demo_builder.book1_end_scene(image_width=1920, samples=int(os.environ.get("SPP", 16)))
desc, cam = sc.describe(), sc.scene_cam.to_abi()
if world == 1:
    n = abi.load().cr_device_count()
    gs = GpuScene(desc, 0)
    full, full8, st1 = gs.render(cam, seed=1)
    reps = [gs] + [gs.replicate(d) for d in range(1, n)]
    for _ in range(2):
        t0 = time.perf_counter()
        rgb, rgb8, stats = gpu.render_multi(reps, cam, seed=1)
        dt = time.perf_counter() - t0
    print(json.dumps({"mode": "cr_render_multi", "devices": n, "identical": bool(np.array_equal(rgb, full) and np.array_equal(rgb8, full8)),
                      "ms_wall": dt * 1e3, "ms_single": st1["ms_total"], "ms_per_replica": [round(s["ms_total"], 2) for s in stats]}))
else:
    import torch.distributed as dist

    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gs = GpuScene(desc, local)
    out = {}
    for exchange in ("p2p", "nccl"):
        for _ in range(3):
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            full, full8, st = multigpu.render_sharded(gs, cam, rank, world, seed=1, exchange=exchange)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        if rank == 0:
            out[exchange] = (full.cpu().numpy().copy(), full8.cpu().numpy().copy(), dt)
    if rank == 0:
        ref, ref8, st1 = gs.render(cam, seed=1)
        print(json.dumps({"mode": "torchrun", "world": world,
                          "p2p_identical": bool(np.array_equal(out["p2p"][0], ref) and np.array_equal(out["p2p"][1], ref8)),
                          "nccl_identical": bool(np.array_equal(out["nccl"][0], ref) and np.array_equal(out["nccl"][1], ref8)),
                          "ms_p2p": out["p2p"][2] * 1e3, "ms_nccl": out["nccl"][2] * 1e3, "ms_single": st1["ms_total"]}))
    dist.barrier()
    dist.destroy_process_group()
