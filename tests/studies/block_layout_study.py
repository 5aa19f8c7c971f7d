"""CPU study for the tile kernels' lane utilisation.

Runs the CPU oracle's forward on the headline scene (needs no GPU), then counts -- exactly, from the
oracle's own per-Gaussian outputs -- for every (tile entry, pixel) whether the pixel would evaluate the
Gaussian's alpha (exponent <= 0 and alpha >= 1/255; transmittance is ignored, so this is the upper bound
the culling masks work against).  From that it derives, for several pixel-block shapes and hit-list
granularities, the number of warp iterations the hit loops would execute and the live lanes per
iteration:

  warp list     one hit list per warp (today: 4x8 blocks): iterations = entries that touch the block
  half lists    two independent lists per warp, one per 16-lane half: iterations = max over the halves
  quarter lists four lists per warp: iterations = max over the four quarters

    python tests/studies/block_layout_study.py [--tiles 400]

(It lives under tests/ because it drives the CPU oracle, which is test infrastructure.)

prints a table; the numbers quoted in DESIGN.md come from this script.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=400, help="number of tiles sampled (evenly over the image)")
    ap.add_argument("--config", default="C2")
    args = ap.parse_args()
    import gsb200  # noqa: F401
    from gsb200 import scene
    from gsb200.utils.camera_utils import load_nerf_cameras
    import oracle as O

    n, w, h, smin, smax = scene.CONFIGS[args.config]
    params, _cam, _ = scene.synthetic_scene(n, w, h, smin, smax, seed=42, with_target=False)
    cam = load_nerf_cameras(w, h)[0]
    O.set_threads(O.max_threads())
    _img, _dep, buf = O.render_gaussians(**scene.render_kwargs(params, cam))
    xy = np.asarray(buf["points_xy_image"], dtype=np.float32).reshape(-1, 2)
    co = np.asarray(buf["conic_opacity"], dtype=np.float32).reshape(-1, 4)
    pl = np.asarray(buf["point_list"]).reshape(-1)
    ranges = np.asarray(buf["ranges"]).reshape(-1, 2)
    gx = (w + 15) // 16
    num_tiles = ranges.shape[0]
    sample = np.unique(np.linspace(0, num_tiles - 1, args.tiles).astype(np.int64))

    # pixel -> sub-block index for each layout: (rows, cols) of the block a 32-lane warp owns
    layouts = {"4x8 (today)": (4, 8), "2x16": (2, 16), "8x4": (8, 4)}
    py, px = np.meshgrid(np.arange(16), np.arange(16), indexing="ij")
    stats = {}
    entries = live_pairs = 0
    for t in sample:
        a, b = ranges[t]
        if b <= a:
            continue
        ids = pl[a:b]
        tx, ty = (t % gx) * 16, (t // gx) * 16
        dx = xy[ids, 0][:, None, None] - (tx + px)[None].astype(np.float32)
        dy = xy[ids, 1][:, None, None] - (ty + py)[None].astype(np.float32)
        ca, cb, cc, op = (co[ids, k][:, None, None] for k in range(4))
        power = -0.5 * (ca * dx * dx + cc * dy * dy) - cb * dx * dy
        alpha = np.minimum(0.99, op * np.exp(np.minimum(power, 0.0)))
        live = (power <= 0) & (alpha >= 1.0 / 255.0)                   # [entries, 16, 16]
        live &= ((tx + px) < w)[None] & ((ty + py) < h)[None]
        entries += len(ids)
        live_pairs += int(live.sum())
        for name, (br, bc) in layouts.items():
            blocks = live.reshape(len(ids), 16 // br, br, 16 // bc, bc).transpose(0, 1, 3, 2, 4)
            blocks = blocks.reshape(len(ids), -1, br * bc)                # [entries, 8 warps, 32 lanes]
            st = stats.setdefault(name, {"warp": 0, "half": 0, "quarter": 0, "live": 0})
            touch = blocks.any(axis=2)                                    # [entries, 8]
            st["warp"] += int(touch.sum())
            st["live"] += int(blocks.sum())
            # halves / quarters of the warp's lanes in row-major order of the block
            halves = blocks.reshape(len(ids), 8, 2, 16).any(axis=3).sum(axis=0)       # [8 warps, 2]: list lengths
            quarters = blocks.reshape(len(ids), 8, 4, 8).any(axis=3).sum(axis=0)      # [8 warps, 4]
            st["half"] += int(halves.max(axis=1).sum())
            st["quarter"] += int(quarters.max(axis=1).sum())
    print(f"{args.config}: {len(sample)} tiles sampled, {entries} list entries, {live_pairs} live (pixel, Gaussian) pairs "
          f"({live_pairs / max(entries, 1):.1f} per entry)")
    print(f"{'block':14s} {'iterations (warp list)':>24s} {'live lanes/iter':>16s} {'half lists':>12s} {'quarter lists':>14s}")
    base = stats["4x8 (today)"]["warp"]
    for name, st in stats.items():
        print(f"{name:14s} {st['warp']:>16d} ({st['warp'] / base:5.2f}x) {st['live'] / st['warp']:>16.1f} "
              f"{st['half']:>6d} ({st['half'] / base:4.2f}x) {st['quarter']:>8d} ({st['quarter'] / base:4.2f}x)")


if __name__ == "__main__":
    main()
