#!/usr/bin/env python
"""Wall time of a densify event (train.py:351-713) and of the steps around it on the headline scene."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gsb200  # noqa: E402,F401
from gsb200 import scene, train  # noqa: E402
from gsb200.utils.camera_utils import load_nerf_cameras  # noqa: E402


def main():
    n, w, h, smin, smax = scene.CONFIGS["C2"]
    params, _cam, _ = scene.synthetic_scene(n, w, h, smin, smax, seed=42, with_target=False)
    cams = load_nerf_cameras(w, h)[:4]
    rng = np.random.default_rng(4242)
    targets = [rng.uniform(0, 1, (h, w, 3)).astype(np.float32) for _ in cams]
    lrs = {"lr_pos": 1e-6, "lr_scale": 5e-7, "lr_rot": 5e-7, "lr_sh": 2e-7, "lr_opac": 5e-7, "final_lr_factor": 0.01}
    cfg = {"num_iterations": 7000, "lr_scheduler_config": lrs, "densify_from_iter": 0, "densification_interval": 10,
           "densify_grad_threshold": 0.0002, "percent_dense": 0.003, "cull_opacity_threshold": 0.06}
    T = train.Trainer(cams, targets=targets, params=params, config=cfg)
    for it in range(1, 10):
        T.train_step(it, [it % 4], densify=True)
    torch.cuda.synchronize()
    prof = cProfile.Profile()
    for it in range(10, 32):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if it % 10 == 0:
            prof.enable()
        T.train_step(it, [it % 4], densify=True)
        torch.cuda.synchronize()
        if it % 10 == 0:
            prof.disable()
        print(f"it {it:3d}: {(time.perf_counter() - t0) * 1e3:8.3f} ms  n = {T.num_points}", flush=True)
    pstats.Stats(prof).sort_stats("cumulative").print_stats(18)


if __name__ == "__main__":
    main()
