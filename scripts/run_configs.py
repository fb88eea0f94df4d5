"""Times every BASELINE.json configuration at full resolution (spp reduced where stated) on one GPU.
Usage (GPU box): python scripts/run_configs.py [config ...]   -> one JSON line per config/precision."""
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np

from crucible_b200 import abi, demo_builder
from crucible_b200.gpu import GpuScene

PLAN = {  # name: (builder kwargs, note) — every BASELINE.json config at its FULL size
    "book1": (dict(image_width=1920, samples=100), "configs[0] 1920x1080, 100 spp"),
    "cornell": (dict(image_width=1024, samples=1000), "configs[1] 1024x1024, 1000 spp"),
    "teapot": (dict(image_width=1920, samples=256), "configs[2] 1920x1080, 256 spp"),
    "instanced": (dict(image_width=3840, samples=64), "configs[3] 3840x2160, 64 spp, 9 998 240 triangles"),
}
names = sys.argv[1:] or list(PLAN)
for name in names:
    kw, note = PLAN[name]
    t0 = time.time()
    sc = demo_builder.CONFIGS[name](**kw)
    desc = sc.describe()
    t_desc = time.time() - t0
    t0 = time.time()
    gs = GpuScene(desc, 0)
    t_commit = time.time() - t0
    cam = sc.scene_cam.to_abi()
    info = gs.bvh_info()
    for prec, pname in ((abi.CR_PRECISION_F64, "f64"), (abi.CR_PRECISION_F32, "f32")):
        gs.render(cam, seed=1, precision=prec, want_rgb=False, want_rgb8=False)  # warm-up
        rgb, _, st = gs.render(cam, seed=1, precision=prec, time_kernels=True, want_rgb8=False)
        print(json.dumps({"config": name, "note": note, "precision": pname, "prims": desc.n_prims, "bvh": info,
                          "host_describe_s": round(t_desc, 2), "host_commit_s": round(t_commit, 2),
                          "msamples_per_s": st["samples"] / st["ms_total"] / 1e3, "mrays_per_s": st["rays"] / st["ms_total"] / 1e3,
                          "ms_total": st["ms_total"], "ms_trace": st["ms_trace"], "ms_shade": st["ms_shade"], "ms_raygen": st["ms_raygen"],
                          "rays_per_sample": st["rays"] / st["samples"], "mean": float(rgb.mean())}), flush=True)
    gs.close()
