#!/usr/bin/env python
"""Small end-to-end cases for memory / race checking.  compute-sanitizer is CLOSED on this pool ("runs under it have
left GPUs needing a reset ... find a bad access with bounds checks and asserts of your own, small cases, and a
comparison with the CPU reference", gpurun, 2026-10-19), so the cases run on the build with device-side bounds checks
instead (any failed check traps the kernel):

    make -C 3dgs-native_b200/csrc variant TAG=chk EXTRA=-DGSB_DEBUG_CHECKS=1
    GSB200_LIB=3dgs-native_b200/csrc/libgsb200_chk.so python tools/sanitize_case.py

and, where compute-sanitizer is available:  compute-sanitizer --tool memcheck|racecheck python tools/sanitize_case.py
Besides the checks, every forward is run three times and must give the same bits (the per-tile counters, the rank
cursor and the loss accumulator re-zero themselves across frames).

Case 1: render.py's 3-Gaussian scene (forward, both binning paths).  Case 2: 3000 Gaussians at 96x64 -- forward,
L1 loss, backward (tensor-core and shuffle reductions, masks handed on and recomputed), Adam, one densify event,
SSIM / depth loss, the radix sort on its own.  Every kernel of the library runs at least once."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gsb200  # noqa: E402,F401
from gsb200 import _lib, backward, forward, loss, scene, train  # noqa: E402
from gsb200.utils.camera_utils import load_nerf_cameras  # noqa: E402


def main():
    size = int(os.environ.get("GSB_SANITIZE_SIZE", "600"))
    ctx = _lib.context()
    print("library:", _lib.LIB_PATH)
    kw = scene.example_render_kwargs(size, size)
    for binning in (0, 1):
        ctx.set_option("binning", binning)
        img, _, buf = forward.render_gaussians(**kw)
    ctx.set_option("binning", 0)
    torch.cuda.synchronize()
    print("case 1 ok: D =", buf["point_list"].numel())

    n, w, h = 3000, 96, 64
    params, cam, target = scene.synthetic_scene(n, w, h, 0.02, 0.12, seed=11)
    for tile_sort in (0, 1, 3):
        ctx.set_option("tile_sort", tile_sort)
        img, depth, buf = forward.render_gaussians(**scene.render_kwargs(params, cam))
    ctx.set_option("tile_sort", 3)
    ref = None
    for _ in range(3):                     # self-resetting scratch: same bits every time
        i2, d2, b2 = forward.render_gaussians(**scene.render_kwargs(params, cam))
        cur = [i2.clone(), d2.clone()] + [b2[k].clone() for k in sorted(b2) if k != "block_masks"]
        if ref is None:
            ref = cur
        assert all(torch.equal(a, b) for a, b in zip(ref, cur)), "repeated frames differ"
    l1 = [float(loss.l1_loss(img, target)) for _ in range(3)]
    assert l1[0] == l1[1] == l1[2], l1
    dpix = loss.compute_image_gradients(img, target, 0.0)
    for mode in (2, 0):
        ctx.set_option("bwd_reduce", mode)
        for hand_on in (True, False):
            b = dict(buf)
            if not hand_on:
                del b["block_masks"]
            backward.backward(**scene.backward_kwargs(params, cam, b, dpix))
    ctx.set_option("bwd_reduce", 2)
    loss.ssim(img, target)
    loss.depth_loss(depth, depth * 0.9, torch.ones_like(depth))
    cams = load_nerf_cameras(w, h)[:4]
    rng = np.random.default_rng(1)
    targets = [rng.uniform(0, 1, (h, w, 3)).astype(np.float32) for _ in cams]
    T = train.Trainer(cams, targets=targets, params=params,
                      config={"num_iterations": 100, "densify_from_iter": 0, "densification_interval": 2,
                              "densify_grad_threshold": 0.00045, "percent_dense": 0.015, "cull_opacity_threshold": 0.1,
                              "min_valid_points": 100})
    for it in range(1, 4):
        T.train_step(it, [it % 4], densify=True)
    torch.cuda.synchronize()
    print("case 2 ok: num_points", n, "->", T.num_points)
    D = 50000
    g = torch.Generator(device="cuda").manual_seed(1)
    keys = torch.randint(0, 1 << 40, (D,), device="cuda", generator=g, dtype=torch.int64)
    vals = torch.arange(D, device="cuda", dtype=torch.int32)
    for coop in (1, 0):                    # one cooperative launch (9-bit digits: 40 bits = 5 passes) / three kernels per pass
        ctx.set_option("sort_coop", coop)
        k2, v2 = keys.clone(), vals.clone()
        ctx.check(_lib.lib().gsb_sort_pairs64(ctx.h, _lib.stream_ptr(ctx.device_index), _lib.ptr(k2), _lib.ptr(v2), None,
                                              None, D, 0, 40))
        torch.cuda.synchronize()
        assert bool((k2[1:] >= k2[:-1]).all()) and bool(torch.equal(keys[v2.long()], k2))
    ctx.set_option("sort_coop", 1)
    print("radix sort ok")
    # degenerate depths (one plane facing the camera): the bucket sort's bitonic fallback; speculative and waiting frames
    kw = scene.render_kwargs(params, cam)
    view = np.asarray(kw["viewmatrix"], dtype=np.float64).reshape(4, 4)
    m = np.asarray(kw["means3D"], dtype=np.float64)
    m -= np.outer((m - m[0]) @ view[:3, 2], view[:3, 2])
    kw["means3D"] = m.astype(np.float32)
    for spec in (1, 0, 1):
        ctx.set_option("speculate", spec)
        forward.render_gaussians(**kw)
        forward.render_gaussians(**scene.render_kwargs(params, cam))
    ctx.set_option("speculate", 1)
    torch.cuda.synchronize()
    print("degenerate depths / speculation ok")


if __name__ == "__main__":
    main()
