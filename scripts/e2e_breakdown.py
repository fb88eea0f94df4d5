#!/usr/bin/env python
"""Where does the end-to-end step (scene description -> image on the host) spend its time?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crucible_b200 import abi, demo_builder, multigpu
from crucible_b200.gpu import GpuScene

kw = {"samples": int(os.environ["SPP"])} if "SPP" in os.environ else {}
sc = demo_builder.CONFIGS[os.environ.get("CONFIG", "book1")](**kw)
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
    ci = g.commit_info()
    # the reference-facing call: cr_render with HOST buffers (what bench.py's e2e leg times at N=1)
    _, _, st = g.render(cam, seed=1, out_rgb=pinned.numpy(), out_rgb8=pinned8.numpy()); t.append(time.perf_counter())
    g.close(); t.append(time.perf_counter())
    names = ["create", "destroy", "GpuScene", "cr_render", "close"]
    print(it, " ".join(f"{n}={1e3*(b-a):.2f}" for n, a, b in zip(names, t, t[1:])),
          f"device_ms={st['ms_total']:.2f} h2d_ms={st['ms_h2d']:.2f} d2h_ms={st['ms_d2h']:.2f} launches={st['launches']}",
          {k: round(v, 2) for k, v in ci.items() if k.startswith("ms_")}, flush=True)
