#!/usr/bin/env python
"""Where does the end-to-end step (scene description -> image on the host) spend its time?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crucible_b200 import abi, demo_builder, multigpu
from crucible_b200.gpu import GpuScene

sc = demo_builder.book1_end_scene(image_width=1920, samples=int(os.environ.get("SPP", "100")), seed=1)
desc, cam = sc.describe(), sc.scene_cam.to_abi()
H, W = cam.image_height, cam.image_width
pinned = torch.empty((H, W, 3), dtype=torch.float64).pin_memory()
pinned8 = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
lib = abi.load()
for it in range(6):
    t = [time.perf_counter()]
    h = lib.cr_scene_create(0); t.append(time.perf_counter())
    lib.cr_scene_destroy(h); t.append(time.perf_counter())
    g = GpuScene(desc, 0); t.append(time.perf_counter())
    full, full8, st = multigpu.render_sharded(g, cam, 0, 1, seed=1); t.append(time.perf_counter())
    pinned.copy_(full, non_blocking=True); pinned8.copy_(full8, non_blocking=True); torch.cuda.synchronize(); t.append(time.perf_counter())
    g.close(); t.append(time.perf_counter())
    names = ["create", "destroy", "GpuScene", "render", "d2h", "close"]
    print(it, " ".join(f"{n}={1e3*(b-a):.2f}" for n, a, b in zip(names, t, t[1:])), f"gpu_ms={st['ms_total']:.2f}", flush=True)
