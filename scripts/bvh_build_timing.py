"""Times cr_scene_commit with the host and the device BVH builder and checks the trees are equal.
Usage (GPU box): python scripts/bvh_build_timing.py [copies ...]   (copies of teapot.obj; 1582 = BASELINE config 4)"""
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np

from crucible_b200 import abi, demo_builder
from crucible_b200.gpu import GpuScene

DEVICE_ONLY = "--device-only" in sys.argv  # profiling runs: skip the host builder and the comparison
for copies in [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [30, 1582]:
    grid = max(2, int(np.ceil(np.sqrt(copies))))
    sc = demo_builder.instanced_teapots(copies=copies, grid=grid)
    desc = sc.describe()
    out = {"copies": copies, "prims": desc.n_prims}
    trees = {}
    modes = (("device", abi.CR_BVH_DEVICE), ("host", abi.CR_BVH_HOST), ("device_again", abi.CR_BVH_DEVICE))
    for name, mode in modes[:1] if DEVICE_ONLY else modes:
        t0 = time.time()
        gs = GpuScene(desc, 0, bvh_builder=mode)
        out[name] = {k: round(v, 2) if isinstance(v, float) else v for k, v in gs.commit_info().items()}
        out[name]["wall_s_incl_staging"] = round(time.time() - t0, 2)
        out["bvh"] = gs.bvh_info()
        if not DEVICE_ONLY:
            trees[name] = gs.bvh_nodes()
        gs.close()
    if not DEVICE_ONLY:
        out["equal"] = bool(trees["device"].tobytes() == trees["host"].tobytes() == trees["device_again"].tobytes())
    print(json.dumps(out), flush=True)
