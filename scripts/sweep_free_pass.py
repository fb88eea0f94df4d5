"""Sweep of DevScene::free_pass_nodes (CRB_FREE_PASS): trace time per scene, f64."""
import os
import sys

sys.path.insert(0, ".")
from crucible_b200 import demo_builder
from crucible_b200.gpu import GpuScene

SCENES = [("book1", dict(image_width=1920, samples=100)), ("teapot", dict(image_width=1920, samples=64)),
          ("cornell", dict(image_width=1024, samples=100)), ("instanced", dict(image_width=3840, samples=8))]
if os.environ.get("SCENES"):
    SCENES = [sc for sc in SCENES if sc[0] in os.environ["SCENES"].split()]
VALUES = [int(v) for v in os.environ.get("VALUES", "0 3 7 31 1073741824").split()]
KS = os.environ.get("KS", "").split()  # when given: sweep CRB_FREE_PASS_K at the first VALUE instead
for name, kw in SCENES:
    sc = demo_builder.CONFIGS[name](**kw)
    gs = GpuScene(sc.describe(), 0)
    cam = sc.scene_cam.to_abi()
    row = []
    for v in (KS or VALUES):
        if KS:
            os.environ["CRB_FREE_PASS"], os.environ["CRB_FREE_PASS_K"] = str(VALUES[0]), v
        else:
            os.environ["CRB_FREE_PASS"] = str(v)
        gs.render(cam, seed=1, want_rgb=False, want_rgb8=False)
        _, _, st = gs.render(cam, seed=1, time_kernels=True, want_rgb=False, want_rgb8=False)
        row.append(f"{v}: trace {st['ms_trace']:.1f} total {st['ms_total']:.1f}")
    print(name, " | ".join(row), flush=True)
    gs.close()
