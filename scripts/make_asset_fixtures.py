"""Turns the reference's two asset files into compact fixtures that travel to the GPU box
(/root/reference does not exist there).  Run in the build container:

    python scripts/make_asset_fixtures.py

  teapot.npz   : vertices [3644][3] f64 and 1-based faces [6320][3] of assets/teapot.obj, parsed with the
                 reference's own rules (asset_loader/obj_loader.rs:64-143: only `v` and `f` lines).
  earthmap.npz : assets/earthmap.jpg decoded ONCE to RGB8 [512][1024][3] (the reference forces every
                 image through to_rgb8, asset_loader/img_loader.rs:28).  Decoder differences (zune-jpeg in
                 the reference vs libjpeg here) may change texels by +-1, which is why the same decoded
                 buffer is handed to the oracle and to the GPU.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crucible_b200.scene import parse_obj  # noqa: E402

REF = "/root/reference/assets"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "crucible_b200", "assets")

v, f = parse_obj(os.path.join(REF, "teapot.obj"))
assert v.shape == (3644, 3) and f.shape == (6320, 3), (v.shape, f.shape)
np.savez_compressed(os.path.join(OUT, "teapot.npz"), vertices=v, faces=f.astype(np.int32))

from PIL import Image  # noqa: E402

im = np.asarray(Image.open(os.path.join(REF, "earthmap.jpg")).convert("RGB"), dtype=np.uint8)
assert im.shape == (512, 1024, 3), im.shape
np.savez_compressed(os.path.join(OUT, "earthmap.npz"), rgb8=im)
print("wrote", os.listdir(OUT))
