"""GPU parity at BASELINE.json's full sizes.

* config 2 (300k Gaussians, 800x800) and config 3 (1M Gaussians, 1920x1080): forward AND backward
  against the CPU oracle (run on all host cores for the forward; the backward oracle runs
  multi-threaded too -- its float atomics make the summation order vary, which the 1e-3 tolerance
  absorbs).
* config 5 (6M Gaussians, 3840x2160, forward only): the same comparison with the oracle (about ten seconds on the
  box's cores), plus size-independent properties of the domain (partition, sortedness, conservation).
* gradients: norm-wise rel 1e-3 per tensor AND element-wise |d| <= 1e-3 |ref| + 1e-4 rms(ref) for all but a handful
  of elements per tensor, |d| <= 1e-3 |ref| + 1e-3 rms(ref) for EVERY element (measured by tools/parity_probe.py on a
  B200: norm-wise 0.6e-6 .. 1.9e-6, zero or one element beyond the 1e-4 floor at C2 and C3); the backward's alpha >= 1/255 decisions are counted against the
  forward's (measured: the kernel's MUFU test flips 1 of 42M pairs at C2, 2 of 144M at C3; bound 1e-6).
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

INT_KEYS = ["radii", "point_offsets", "point_list", "ranges", "n_contrib"]
GRAD_KEYS = ["dL_dmean3D", "dL_dcolor", "dL_dshs", "dL_dopacity", "dL_dscale", "dL_drot", "dL_dmean2D", "dL_dconic"]
# elements per tensor allowed beyond |d| <= 1e-3 |ref| + 1e-4 rms(ref).  Measured 0 .. 1 (of 0.3 M .. 48 M): the one or two
# (pixel, Gaussian) pairs per frame on which the backward's MUFU alpha >= 1/255 test differs from the forward's shift that
# pixel's replayed T by 0.39%, which can push a Gaussian with few contributing pixels past the 1e-4 floor.
ELEMENTWISE_MAX_VIOLATIONS = 8


@pytest.fixture(scope="module")
def gs():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import types
    import gsb200  # noqa: F401
    from gsb200 import backward, forward, scene
    return types.SimpleNamespace(forward=forward, backward=backward, scene=scene)


@pytest.mark.parametrize("cfg", ["C2", "C3"])
def test_full_size_forward_backward_vs_oracle(gs, oracle, cfg):
    n, w, h, smin, smax = gs.scene.CONFIGS[cfg]
    params, cam, target = gs.scene.synthetic_scene(n, w, h, smin, smax)
    kw = gs.scene.render_kwargs(params, cam)
    img, depth, buf = gs.forward.render_gaussians(**kw)
    oracle.set_threads(oracle.max_threads())
    try:
        o_img, o_depth, ob = oracle.render_gaussians(**kw, return_extra=True)
        for k in INT_KEYS:
            a, b = buf[k].cpu().numpy().reshape(ob[k].shape), ob[k]
            assert np.array_equal(a, b), f"{cfg} {k}: {np.count_nonzero(a != b)} of {b.size} differ"
        assert np.abs(img.cpu().numpy() - o_img).max() <= 1e-4
        assert np.abs(depth.cpu().numpy() - o_depth).max() <= 1e-4 * max(1.0, float(np.abs(o_depth).max()))
        if cfg == "C2":      # the numbers SURVEY.md 8d quotes for the headline scene
            assert int((ob["radii"] > 0).sum()) == 283155 and ob["point_list"].size == 1614352
        dpix = oracle.compute_image_gradients(o_img, target, lambda_dssim=0)
        g = gs.backward.backward(**gs.scene.backward_kwargs(params, cam, buf, dpix))
        og = oracle.backward(**gs.scene.backward_kwargs(params, cam, ob, dpix), return_extra=True)
    finally:
        oracle.set_threads(1)
    for k in GRAD_KEYS:
        a = g[k].cpu().numpy().reshape(og[k].shape).astype(np.float64)
        b = og[k].astype(np.float64)
        rel = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
        assert rel <= 1e-3, (cfg, k, rel)
        # element-wise: north_star's "rel 1e-3 on gradients" per ELEMENT, with an absolute floor of 1e-4 of the
        # tensor's rms (an element that is a cancelling sum of terms far larger than itself cannot be held to a
        # relative bound).  Allowed violations: ELEMENTWISE_MAX_VIOLATIONS per tensor (measured: tools/parity_probe.py)
        rms = np.sqrt(np.mean(b * b))
        bad = int((np.abs(a - b) > 1e-3 * np.abs(b) + 1e-4 * rms).sum())
        assert bad <= ELEMENTWISE_MAX_VIOLATIONS, (cfg, k, bad, b.size)
        assert not (np.abs(a - b) > 1e-3 * np.abs(b) + 1e-3 * rms).any(), (cfg, k)     # ... and none beyond a 1e-3 floor
    assert not g["dL_dcov3D"].any()
    # the backward's one decision per pair (alpha >= 1/255) against the forward's, counted on this frame
    c = work_counters(gs, cam, buf, w, h)
    # work counters of SURVEY 8d: the pairs the reference's loops iterate, as the oracle counted them
    assert c["K_fwd"] == ob["_pairs_fwd"] and c["K_bwd"] == og["_pairs_bwd"] == int(ob["n_contrib"].astype(np.int64).sum())
    assert c["K_fwd"] >= c["K_bwd"] > c["pairs_blended"] > 0
    # measured 1 (C2) and 2 (C3) disagreeing pairs, i.e. ~2e-8 of the evaluated pairs; the bound is 1e-6 (VERDICT r1 asked
    # for 1e-5).  The conservative exponent threshold must never skip a pair the forward blended.
    assert c["bwd_decision_mismatch"] <= 1e-6 * c["bwd_pairs_evaluated"], c
    assert c["threshold_false_skips"] == 0, c


def work_counters(gs, cam, buf, w, h):
    import ctypes as C
    from gsb200 import _lib
    ctx = _lib.context()
    frame = _lib.make_frame(cam["world_to_camera"], cam["full_proj_matrix"], cam["camera_center"], cam["tan_fovx"],
                            cam["tan_fovy"], w, h, (0.0, 0.0, 0.0), 3, True, 1.0)
    out = torch.zeros(7, dtype=torch.int64, device="cuda")
    p = _lib.ptr
    ctx.check(_lib.lib().gsb_selftest_work_counters(ctx.h, _lib.stream_ptr(ctx.device_index), C.byref(frame),
                                                    p(buf["ranges"]), p(buf["point_list"]), p(buf["points_xy_image"]),
                                                    p(buf["conic_opacity"]), p(buf["n_contrib"]), p(out)))
    torch.cuda.synchronize()
    names = ("K_fwd", "pairs_blended", "K_bwd", "bwd_pairs_evaluated", "bwd_decision_mismatch", "threshold_false_skips",
             "pairs_near_threshold")
    return dict(zip(names, [int(x) for x in out.cpu()]))


def test_config5_forward_vs_oracle(gs, oracle):
    """6M Gaussians at 3840x2160 against the oracle (forward.py:384-515, 517-586): radii, offsets, the whole sorted
    point_list (68M entries), tile ranges and n_contrib bit-exact; image and inverse depth within 1e-4."""
    n, w, h, smin, smax = gs.scene.CONFIGS["C5"]
    params, cam, _ = gs.scene.synthetic_scene(n, w, h, smin, smax, with_target=False)
    kw = gs.scene.render_kwargs(params, cam)
    img, depth, buf = gs.forward.render_gaussians(**kw)
    torch.cuda.synchronize()
    oracle.set_threads(oracle.max_threads())
    try:
        o_img, o_depth, ob = oracle.render_gaussians(**kw)
    finally:
        oracle.set_threads(1)
    for k in INT_KEYS:
        a, b = buf[k].cpu().numpy().reshape(ob[k].shape), ob[k]
        assert np.array_equal(a, b), f"C5 {k}: {np.count_nonzero(a != b)} of {b.size} differ"
    assert ob["point_list"].size > 50_000_000
    assert np.abs(img.cpu().numpy() - o_img).max() <= 1e-4
    assert np.abs(depth.cpu().numpy() - o_depth).max() <= 1e-4 * max(1.0, float(np.abs(o_depth).max()))
    assert np.array_equal(buf["final_Ts"].cpu().numpy(), ob["final_Ts"])       # T follows the exact contract


def test_config5_forward_properties(gs):
    """6M Gaussians at 3840x2160 (D ~ 68M): binning and blend invariants."""
    n, w, h, smin, smax = gs.scene.CONFIGS["C5"]
    params, cam, _ = gs.scene.synthetic_scene(n, w, h, smin, smax, with_target=False)
    kw = gs.scene.render_kwargs(params, cam)
    img, depth, buf = gs.forward.render_gaussians(**kw)
    torch.cuda.synchronize()
    D = buf["point_list"].numel()
    radii, offs = buf["radii"], buf["point_offsets"].long()
    tiles = torch.diff(offs, prepend=torch.zeros(1, dtype=torch.long, device=offs.device))
    assert int(offs[-1]) == D and D > 50_000_000                      # conservation: sum tiles_touched == D
    assert torch.equal(tiles > 0, radii > 0)
    rg = buf["ranges"].long()
    lens = rg[:, 1] - rg[:, 0]
    assert int(lens.sum()) == D                                       # the ranges partition [0, D)
    nz = lens > 0
    starts = rg[nz, 0]
    assert torch.equal(starts, torch.cumsum(lens[nz], 0) - lens[nz])  # in tile order, without gaps
    # every Gaussian appears exactly tiles_touched times in point_list
    counts = torch.bincount(buf["point_list"].long(), minlength=n)
    assert torch.equal(counts, tiles)
    # inside every tile the list is sorted by (depth, id): check on a sample of tiles incl. the longest
    depths = buf["depths"]
    sample = torch.cat([torch.topk(lens, 8).indices, torch.randint(0, rg.shape[0], (256,), device=rg.device)])
    for t in sample.tolist():
        s, e = int(rg[t, 0]), int(rg[t, 1])
        if e - s < 2:
            continue
        ids = buf["point_list"][s:e].long()
        key = (depths[ids].view(torch.int32).long() << 32) | ids
        assert bool((key[1:] > key[:-1]).all()), f"tile {t} not sorted"
        gx = (w + 15) // 16                                          # ... and every entry really touches the tile
        tx, ty = t % gx, t // gx
        xy, r = buf["points_xy_image"][ids], radii[ids].float()
        assert bool(((xy[:, 0] + r + 15 >= tx * 16) & (xy[:, 0] - r < tx * 16 + 16 + 1)).all())
        assert bool(((xy[:, 1] + r + 15 >= ty * 16) & (xy[:, 1] - r < ty * 16 + 16 + 1)).all())
    ncon = buf["n_contrib"].long()
    gx, gy = (w + 15) // 16, (h + 15) // 16
    per_tile = lens.view(gy, gx).repeat_interleave(16, 0).repeat_interleave(16, 1)[:h, :w]
    assert bool((ncon <= per_tile).all()) and bool((ncon >= 0).all())  # last contributor lies inside the tile's list
    T = buf["final_Ts"]
    assert bool(torch.isfinite(img).all()) and bool((T >= 1e-4).all()) and bool((T <= 1.0).all())
    assert bool(((ncon == 0) == (T == 1.0)).all())                    # untouched pixels keep T = 1
    assert bool((img.min() >= 0))                                     # clamped colours, black background
