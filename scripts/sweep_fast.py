#!/usr/bin/env python
"""Sweep of the trace engine's scheduling knobs (env: CRB_NODE_SLICE, CRB_MIN_LANES, CRB_REFILL_RT, CRB_FAST_MINB) on the
GPU box: f64 trace milliseconds per scene and setting.   python scripts/sweep_fast.py book1 teapot instanced:200"""
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crucible_b200 import abi, demo_builder
from crucible_b200.gpu import GpuScene

SPP = {"book1": 32, "cornell": 64, "teapot": 64, "instanced": 8}
scenes = {}
for arg in sys.argv[1:]:
    name, _, copies = arg.partition(":")
    kw = {"samples": SPP[name]}
    if copies:
        kw["copies"] = int(copies)
    sc = demo_builder.CONFIGS[name](**kw)
    scenes[arg] = (GpuScene(sc.describe(), 0), sc.scene_cam.to_abi())
grid = {"CRB_NODE_SLICE": os.environ.get("SWEEP_SLICE", "8,32").split(","), "CRB_MIN_LANES": os.environ.get("SWEEP_LANES", "8,16").split(","),
        "CRB_REFILL_RT": os.environ.get("SWEEP_REFILL", "8,16,24").split(","), "CRB_FAST_MINB": os.environ.get("SWEEP_MINB", "8").split(",")}
keys = list(grid)
for combo in itertools.product(*[grid[k] for k in keys]):
    for k, v in zip(keys, combo):
        os.environ[k] = v
    row = {k: v for k, v in zip(keys, combo)}
    for arg, (gs, cam) in scenes.items():
        gs.render(cam, seed=1, want_rgb=False, want_rgb8=False)
        _, _, st = gs.render(cam, seed=1, want_rgb=False, want_rgb8=False, time_kernels=True)
        row[arg] = round(st["ms_trace"], 2)
    print(json.dumps(row), flush=True)
