#!/usr/bin/env python
"""Per-step wall time of a densify-heavy training loop (an event every 10 steps, 320 steps, the scene grows from 300k to
660k Gaussians): prints the steps that take more than 3 ms with the Gaussian count and the buffer capacities -- the
first step, the frames whose point_list capacity is exceeded (one retry) and the events that outgrow the state buffers'
50% headroom (7-18 ms each).  Background for `bench.py --densify-every`: a repetition there occasionally shows a
0.4 s stall that this loop never reproduces -- bench.py samples nvidia-smi during the timed region, and a cudaMalloc /
cudaFree of a growth event then waits for the driver lock the query holds.

    python tools/densify_diag.py
"""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import bench
from gsb200 import train
params, cams, targets, _ = bench.make_scene("C2")
lrs = {k: v * bench.LR_SCALE for k, v in bench.BASE_LRS.items()}; lrs["final_lr_factor"] = bench.FINAL_LR_FACTOR
cfg = {"num_iterations": bench.TOTAL_ITERATIONS, "lr_scheduler_config": lrs}
cfg.update(bench.densify_config(10))
T = train.Trainer(cams, targets=targets, params=params, config=cfg)
it = bench.IT0
ts = []
for s in range(320):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n0 = T.num_points
    T.train_step(it, [it % 16], densify=True); it += 1
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    ts.append(dt)
    if dt > 3.0:
        print(f"step {s} it {it-1}: {dt:.1f} ms  num_points {n0} -> {T.num_points}  fb cap {T.fb.capacity if T.fb else None} params cap {T.params.capacity}", flush=True)
print("median", np.median(ts), "n", T.num_points)
