#!/usr/bin/env python
"""Parity probe (run on the GPU box): element-wise gradient error statistics against the CPU oracle and the
work / decision counters of gsb_selftest_work_counters at the headline sizes.  Prints one JSON object; the
numbers it found are the bounds tests/test_gpu_large.py asserts.

    python tools/parity_probe.py C2 [C3]
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import gsb200  # noqa: E402,F401
from gsb200 import _lib, backward, forward, scene  # noqa: E402
import oracle as O  # noqa: E402

GRAD_KEYS = ["dL_dmean3D", "dL_dcolor", "dL_dshs", "dL_dopacity", "dL_dscale", "dL_drot", "dL_dmean2D", "dL_dconic"]


def work_counters(cam, buf, W, H):
    ctx = _lib.context()
    frame = _lib.make_frame(cam["world_to_camera"], cam["full_proj_matrix"], cam["camera_center"], cam["tan_fovx"],
                            cam["tan_fovy"], W, H, (0.0, 0.0, 0.0), 3, True, 1.0)
    out = torch.zeros(7, dtype=torch.int64, device="cuda")
    p = _lib.ptr
    ctx.check(_lib.lib().gsb_selftest_work_counters(ctx.h, _lib.stream_ptr(ctx.device_index), C.byref(frame),
                                                    p(buf["ranges"]), p(buf["point_list"]), p(buf["points_xy_image"]),
                                                    p(buf["conic_opacity"]), p(buf["n_contrib"]), p(out)))
    torch.cuda.synchronize()
    names = ("K_fwd", "pairs_blended", "K_bwd", "bwd_pairs_evaluated", "bwd_decision_mismatch", "threshold_false_skips",
             "pairs_near_threshold")
    return dict(zip(names, [int(x) for x in out.cpu()]))


def main():
    O.build()
    res = {}
    for cfg in sys.argv[1:] or ["C2"]:
        n, w, h, smin, smax = scene.CONFIGS[cfg]
        params, cam, target = scene.synthetic_scene(n, w, h, smin, smax)
        kw = scene.render_kwargs(params, cam)
        img, depth, buf = forward.render_gaussians(**kw)
        O.set_threads(O.max_threads())
        o_img, o_depth, ob = O.render_gaussians(**kw, return_extra=True)
        dpix = O.compute_image_gradients(o_img, target, lambda_dssim=0)
        g = backward.backward(**scene.backward_kwargs(params, cam, buf, dpix))
        O.set_threads(1)          # serial oracle: one fixed summation order
        og = O.backward(**scene.backward_kwargs(params, cam, ob, dpix), return_extra=True)
        r = {"work": work_counters(cam, buf, w, h), "oracle_pairs_fwd": int(ob.get("_pairs_fwd", -1)),
             "oracle_pairs_bwd": int(og.get("_pairs_bwd", -1)), "grads": {}}
        for k in GRAD_KEYS:
            a = g[k].cpu().numpy().reshape(og[k].shape).astype(np.float64).ravel()
            b = og[k].astype(np.float64).ravel()
            rms = float(np.sqrt(np.mean(b * b)))
            d = np.abs(a - b)
            row = {"n": int(b.size), "rms": rms, "norm_rel": float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)),
                   "max_abs_over_rms": float(d.max() / max(rms, 1e-30))}
            for rt, at in ((1e-3, 1e-3), (1e-3, 1e-4), (1e-3, 1e-5), (1e-4, 1e-4)):
                row[f"viol_r{rt:g}_a{at:g}"] = int((d > rt * np.abs(b) + at * rms).sum())
            r["grads"][k] = row
        res[cfg] = r
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
