"""Which part of the end-to-end step costs what: the asynchronous API step with / without the per-step
H2D copy of the target and the per-step loss read-back.  python tools/e2e_variants.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gsb200  # noqa: E402,F401
from gsb200 import backward as gb, forward as gf, loss as gl, optimizer as gopt, scene  # noqa: E402


def main():
    n, w, h = 300000, 800, 800
    params, cam, target = scene.synthetic_scene(n, w, h, 0.004, 0.02)
    dev = torch.device("cuda", 0)
    P = {k: torch.from_numpy(v).to(dev) for k, v in params.items()}
    M = {k: torch.zeros_like(v) for k, v in P.items()}
    V = {k: torch.zeros_like(v) for k, v in P.items()}
    pinned = torch.from_numpy(target).pin_memory()
    resident = torch.from_numpy(target).to(dev)
    bg = np.zeros(3, dtype=np.float32)
    copy_stream = torch.cuda.Stream(device=dev)
    tgt_dev = [torch.empty((h, w, 3), dtype=torch.float32, device=dev) for _ in range(2)]
    tgt_ready = [torch.cuda.Event() for _ in range(2)]
    loss_host = [torch.zeros(1, dtype=torch.float64).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event() for _ in range(2)]

    def run(h2d, readback, steps=30):
        losses = []
        pending = None
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for it in range(steps):
            img, _d, buf = gf.render_gaussians(**scene.render_kwargs(P, cam, background=bg))
            main_s = torch.cuda.current_stream()
            if pending is not None:
                loss_ready[pending].synchronize()
                losses.append(float(loss_host[pending].item()))
            if h2d == "side":
                tgt, ready = tgt_dev[it & 1], tgt_ready[it & 1]
                with torch.cuda.stream(copy_stream):
                    tgt.copy_(pinned, non_blocking=True)
                    ready.record(copy_stream)
                main_s.wait_event(ready)
            elif h2d == "main":
                tgt = tgt_dev[it & 1]
                tgt.copy_(pinned, non_blocking=True)
            else:
                tgt = resident
            if readback == "mapped":   # the loss kernel writes its sum straight into pinned host memory
                loss_sum, dpix = gl.l1_loss_and_gradients(img, tgt, 0.0, out_sum=loss_host[it & 1])
            else:
                loss_sum, dpix = gl.l1_loss_and_gradients(img, tgt, 0.0)
            g = gb.backward(**scene.backward_kwargs(P, cam, buf, dpix, background=bg))
            gopt.adam_update(g["dL_dmean3D"], g["dL_dscale"], g["dL_drot"], g["dL_dopacity"], g["dL_dshs"], n, 1e-6, 5e-7,
                             5e-7, 5e-7, 2e-7, 0.9, 0.999, 1e-8, it, P["positions"], P["scales"], P["rotations"],
                             P["opacities"], P["shs"], M["positions"], M["scales"], M["rotations"], M["opacities"],
                             M["shs"], V["positions"], V["scales"], V["rotations"], V["opacities"], V["shs"])
            if readback == "async":
                loss_host[it & 1].copy_(loss_sum, non_blocking=True)
                loss_ready[it & 1].record(main_s)
                pending = it & 1
            elif readback == "mapped":
                loss_ready[it & 1].record(main_s)
                pending = it & 1
            elif readback == "sync":
                losses.append(float(loss_sum.item()))
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / steps * 1e6

    # the resident trainer loop on the same box, same camera, for comparison
    from gsb200 import train
    T = train.Trainer([cam], targets=[target], params=params, config={"num_iterations": 7000, "lr_scheduler_config": {
        "lr_pos": 1e-6, "lr_scale": 5e-7, "lr_rot": 5e-7, "lr_sh": 2e-7, "lr_opac": 5e-7, "final_lr_factor": 0.01}})
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for it in range(30):
            T.train_step(it + 1, [0], densify=False)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 30 * 1e6
    print(f"trainer (resident state, preallocated buffers)  {dt:8.1f} us/step")
    for h2d in ("none", "side", "main"):
        for rb in ("none", "async", "mapped", "sync"):
            run(h2d, rb, 8)
            print(f"h2d={h2d:5s} readback={rb:5s}  {run(h2d, rb):8.1f} us/step")


if __name__ == "__main__":
    main()
