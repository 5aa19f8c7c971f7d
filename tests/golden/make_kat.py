#!/usr/bin/env python
"""Turn the reference's only published artifact, assets/example_render.png (a matplotlib export of
render.py's 1800x1800 render, resampled to 1155x1155 plus a 15 px white margin), into a golden
fixture: the cropped RGB image.  Run in the build container only (needs /root/reference)."""
import json
import os

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
src = os.environ.get("GS_REFERENCE", "/root/reference") + "/assets/example_render.png"
im = np.array(Image.open(src).convert("RGB"))
ys, xs = np.where(im.astype(int).sum(axis=2) < 3 * 250)
x0, x1, y0, y1 = xs.min(), xs.max() + 1, ys.min(), ys.max() + 1
crop = im[y0:y1, x0:x1]
Image.fromarray(crop).save(os.path.join(HERE, "example_render_crop.png"), optimize=True)
w, h = x1 - x0, y1 - y0
centres = {}
for name, fx in (("left", 361.72 / 1800), ("middle", 899.5 / 1800), ("right", 1437.28 / 1800)):
    px = int(round((fx * 1800 + 0.5) / 1800 * w - 0.5))
    py = int(round((899.5 + 0.5) / 1800 * h - 0.5))
    centres[name] = [int(v) for v in crop[py, px]]
json.dump({"crop_box": [int(x0), int(y0), int(x1), int(y1)], "size": [int(w), int(h)], "blob_centre_rgb": centres},
          open(os.path.join(HERE, "example_render_kat.json"), "w"), indent=1)
print(crop.shape, centres)
