"""Host-side profile of one end-to-end training step through the reference-facing Python API:
wall time of every call with and without a trailing synchronize (the difference is GPU time the host
did not overlap).  Run on a GPU box: python tools/e2e_profile.py"""
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
    bg = np.zeros(3, dtype=np.float32)
    acc = {}
    copy_stream = torch.cuda.Stream(device=dev)

    def timed(name, fn, sync):
        t0 = time.perf_counter()
        r = fn()
        if sync:
            torch.cuda.synchronize()
        acc.setdefault(name, []).append((time.perf_counter() - t0) * 1e6)
        return r

    for sync in (False, True):
        acc.clear()
        for it in range(12):
            torch.cuda.synchronize()
            t_step = time.perf_counter()
            if sync:
                tgt = timed("h2d", lambda: pinned.to(dev, non_blocking=True), sync)
            else:   # as bench.py does it: on a copy stream, beside the forward
                main = torch.cuda.current_stream()
                with torch.cuda.stream(copy_stream):
                    tgt = timed("h2d", lambda: pinned.to(dev, non_blocking=True), False)
                    ready = copy_stream.record_event()
            kw = timed("kwargs", lambda: scene.render_kwargs(P, cam, background=bg), False)
            img, _d, buf = timed("render_gaussians", lambda: gf.render_gaussians(**kw), sync)
            if not sync:
                main.wait_event(ready)
                tgt.record_stream(main)
            loss_sum, dpix = timed("l1_loss_and_gradients", lambda: gl.l1_loss_and_gradients(img, tgt, 0.0), sync)
            bkw = timed("bkwargs", lambda: scene.backward_kwargs(P, cam, buf, dpix, background=bg), False)
            g = timed("backward", lambda: gb.backward(**bkw), sync)
            timed("adam_update", lambda: gopt.adam_update(
                g["dL_dmean3D"], g["dL_dscale"], g["dL_drot"], g["dL_dopacity"], g["dL_dshs"], n, 1e-6, 5e-7, 5e-7, 5e-7,
                2e-7, 0.9, 0.999, 1e-8, it, P["positions"], P["scales"], P["rotations"], P["opacities"], P["shs"],
                M["positions"], M["scales"], M["rotations"], M["opacities"], M["shs"], V["positions"], V["scales"],
                V["rotations"], V["opacities"], V["shs"]), sync)
            timed("loss.item", lambda: float(loss_sum.item()), False)
            acc.setdefault("STEP", []).append((time.perf_counter() - t_step) * 1e6)
        print(f"--- sync after every call: {sync} (median us over the last 8 of 12 steps)")
        for k, v in acc.items():
            print(f"  {k:24s} {np.median(v[4:]):9.1f}")


def gpu_timeline():
    """GPU-side duration of every call of an asynchronous step (CUDA events on the main stream: the time
    between the event before a call's first kernel and the event after its last, idle gaps included)."""
    n, w, h = 300000, 800, 800
    params, cam, target = scene.synthetic_scene(n, w, h, 0.004, 0.02)
    dev = torch.device("cuda", 0)
    P = {k: torch.from_numpy(v).to(dev) for k, v in params.items()}
    M = {k: torch.zeros_like(v) for k, v in P.items()}
    V = {k: torch.zeros_like(v) for k, v in P.items()}
    tgt = torch.from_numpy(target).to(dev)
    bg = np.zeros(3, dtype=np.float32)
    names = ["render_gaussians", "l1_loss_and_gradients", "backward", "adam_update"]
    rows = []
    for it in range(14):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        img, _d, buf = gf.render_gaussians(**scene.render_kwargs(P, cam, background=bg))
        ev[1].record()
        loss_sum, dpix = gl.l1_loss_and_gradients(img, tgt, 0.0)
        ev[2].record()
        g = gb.backward(**scene.backward_kwargs(P, cam, buf, dpix, background=bg))
        ev[3].record()
        gopt.adam_update(g["dL_dmean3D"], g["dL_dscale"], g["dL_drot"], g["dL_dopacity"], g["dL_dshs"], n, 1e-6, 5e-7,
                         5e-7, 5e-7, 2e-7, 0.9, 0.999, 1e-8, it, P["positions"], P["scales"], P["rotations"],
                         P["opacities"], P["shs"], M["positions"], M["scales"], M["rotations"], M["opacities"], M["shs"],
                         V["positions"], V["scales"], V["rotations"], V["opacities"], V["shs"])
        ev[4].record()
        rows.append(ev)
    torch.cuda.synchronize()
    t = np.array([[r[i].elapsed_time(r[i + 1]) * 1e3 for i in range(4)] + [rows[k - 1][4].elapsed_time(r[0]) * 1e3 if k else 0.0]
                  for k, r in enumerate(rows)])
    print("--- GPU timeline of an asynchronous step (median us over the last 8 of 14 steps)")
    for i, nm in enumerate(names + ["gap to the next step"]):
        print(f"  {nm:24s} {np.median(t[6:, i]):9.1f}")
    print(f"  {'SUM':24s} {np.median(t[6:].sum(axis=1)):9.1f}")


if __name__ == "__main__":
    gpu_timeline()
    main()
