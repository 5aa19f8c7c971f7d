"""GPU parity of the Adam / densify / loss kernels and of the step loop against the golden vectors
(reference source under the Warp shim) and the CPU oracle."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

KEYS = ("positions", "scales", "rotations", "opacities", "shs")


@pytest.fixture(scope="module")
def gs():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import types
    import gsb200  # noqa: F401
    from gsb200 import loss, optimizer, scene, train
    return types.SimpleNamespace(loss=loss, optimizer=optimizer, scene=scene, train=train)


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _adam_call(gs, g, p, m, v, n, it):
    gs.optimizer.adam_update(g["positions"], g["scales"], g["rotations"], g["opacities"], g["shs"], n, 1e-2, 5e-3,
                             5e-3, 5e-3, 2e-3, 0.9, 0.999, 1e-8, it, p["positions"], p["scales"], p["rotations"],
                             p["opacities"], p["shs"], m["positions"], m["scales"], m["rotations"], m["opacities"],
                             m["shs"], v["positions"], v["scales"], v["rotations"], v["opacities"], v["shs"])


def test_adam_matches_reference_source(gs, golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "ref_adam.npz")))
    p = {k: _cuda(g["in_" + k]) for k in KEYS}
    n = p["positions"].shape[0]
    m = {k: torch.zeros_like(p[k]) for k in KEYS}
    v = {k: torch.zeros_like(p[k]) for k in KEYS}
    for step, it in enumerate(g["iterations"]):
        grads = {k: _cuda(g[f"g{step}_{k}"]) for k in KEYS}
        _adam_call(gs, grads, p, m, v, n, int(it))
        for k in KEYS:
            for name, t in (("p", p), ("m", m), ("v", v)):
                want = g[f"{name}{step}_{k}"].reshape(t[k].shape)
                atol = 1e-9 if name == "p" else 1e-30
                assert np.allclose(t[k].cpu().numpy(), want, rtol=3e-6, atol=atol), (name, k, step)


@pytest.mark.parametrize("n", [1, 7, 4096, 100003])
def test_adam_bit_exact_vs_oracle(gs, oracle, n):
    """Same operation order and the same host-computed bias corrections => identical bits."""
    rng = np.random.default_rng(n)
    P = {"positions": rng.normal(size=(n, 3)), "scales": rng.uniform(0.0005, 0.2, (n, 3)),
         "rotations": rng.normal(size=(n, 4)), "opacities": rng.uniform(0, 1, n), "shs": rng.normal(size=(n * 16, 3))}
    P = {k: v.astype(np.float32) for k, v in P.items()}
    p = {k: _cuda(P[k]) for k in KEYS}
    m = {k: torch.zeros_like(p[k]) for k in KEYS}
    v = {k: torch.zeros_like(p[k]) for k in KEYS}
    om, ov = oracle.zeros_like_params(n), oracle.zeros_like_params(n)
    for it in (0, 1, 500):
        G = {k: (rng.normal(size=P[k].shape) * 10.0 ** rng.uniform(-7, 0)).astype(np.float32) for k in KEYS}
        G["shs"][::3] = 0.0
        _adam_call(gs, {k: _cuda(G[k]) for k in KEYS}, p, m, v, n, it)
        oracle.adam_update(G, P, om, ov, n, 1e-2, 5e-3, 5e-3, 5e-3, 2e-3, 0.9, 0.999, 1e-8, it)
        for k in KEYS:
            assert np.array_equal(p[k].cpu().numpy(), P[k]), (k, it)
            assert np.array_equal(m[k].cpu().numpy(), om[k]), (k, it)
            assert np.array_equal(v[k].cpu().numpy(), ov[k]), (k, it)


def test_loss_and_gradient(gs, oracle, golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "ref_loss.npz")))
    assert gs.loss.l1_loss(g["rendered"], g["target"]) == pytest.approx(float(g["l1"]), rel=1e-6)
    assert np.array_equal(gs.loss.compute_image_gradients(g["rendered"], g["target"], 0).cpu().numpy(), g["grad"])
    assert np.array_equal(gs.loss.compute_image_gradients(g["rendered"], g["target"], 0.2).cpu().numpy(),
                          g["grad_dssim02"])
    rng = np.random.default_rng(0)
    r, t = rng.uniform(0, 1, (123, 77, 3)).astype(np.float32), rng.uniform(0, 1, (123, 77, 3)).astype(np.float32)
    assert gs.loss.l1_loss(r, t) == pytest.approx(oracle.l1_loss(r, t), rel=1e-5)
    assert np.array_equal(gs.loss.compute_image_gradients(r, t, 0).cpu().numpy(),
                          oracle.compute_image_gradients(r, t, 0))


def _trainer(gs, params, n_cams=4, w=64, h=48):
    from gsb200.utils.camera_utils import load_nerf_cameras
    cams = load_nerf_cameras(w, h)[:n_cams]
    return gs.train.Trainer(cams, params=params, config={"num_iterations": 100})


def test_densify_matches_reference_source(gs, golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "ref_densify.npz")))
    T = _trainer(gs, {k: g["in_" + k] for k in KEYS})
    T.scene_extent = float(g["scene_extent"])
    T.grads["positions"].copy_(_cuda(g["in_pos_grad"]))
    log = T.densification_and_pruning(600)
    assert T.num_points == int(g["out_num_points"]), log
    for k in KEYS:
        assert np.array_equal(T.params[k].cpu().numpy(), g["out_" + k].reshape(T.params[k].shape)), k
    for part in (T.grads, T.adam_m, T.adam_v):
        assert not part.flat.any()
    # quirk G6: iteration 0 resets opacities (and does nothing else)
    small = {k: (g["in_shs"][: 8 * 16] if k == "shs" else g["in_" + k][:8]) for k in KEYS}
    T2 = _trainer(gs, small)
    T2.params["opacities"].copy_(_cuda(np.linspace(0.1, 0.9, 8).astype(np.float32)))
    assert T2.densification_and_pruning(0)["opacity_reset"]
    assert np.array_equal(T2.params["opacities"].cpu().numpy(), g["reset_opacities_it0"])


def test_densify_vs_oracle_with_last_flag_set(gs, oracle):
    """Quirk G5 (count = last entry of the EXCLUSIVE scan): the last flagged Gaussian is dropped."""
    rng = np.random.default_rng(3)
    n = 5000
    P = {"positions": rng.uniform(-1.3, 1.3, (n, 3)), "scales": np.exp(rng.uniform(np.log(0.01), np.log(0.2), (n, 3))),
         "rotations": rng.normal(size=(n, 4)), "opacities": rng.uniform(0.0, 1.0, n), "shs": rng.normal(size=(n * 16, 3))}
    P = {k: v.astype(np.float32) for k, v in P.items()}
    G = (rng.normal(size=(n, 3)) * 3e-4).astype(np.float32)
    G[-1] = 1.0
    P["scales"][-1] = 0.01
    T = _trainer(gs, P)
    T.grads["positions"].copy_(_cuda(G))
    state = {"params": {k: v.copy() for k, v in P.items()}, "grads": oracle.zeros_like_params(n),
             "adam_m": oracle.zeros_like_params(n), "adam_v": oracle.zeros_like_params(n), "num_points": n,
             "scene_extent": T.scene_extent}
    state["grads"]["positions"] = G.copy()
    log = T.densification_and_pruning(700)
    olog = oracle.densification_and_pruning(state, 700)
    assert log == olog
    assert T.num_points == state["num_points"]
    for k in KEYS:
        assert np.array_equal(T.params[k].cpu().numpy(), state["params"][k]), k


def test_step_loop_vs_oracle(gs, oracle):
    """K iterations of the reference loop (forward, L1 gradient, backward, Adam) on the GPU and in
    the oracle for the same camera sequence: loss curve and parameters agree."""
    n, w, h, K = 3000, 96, 64, 4
    params, _cam0, _ = gs.scene.synthetic_scene(n, w, h, 0.02, 0.12, seed=11)
    from gsb200.utils.camera_utils import load_nerf_cameras
    cams = load_nerf_cameras(w, h)[:6]
    rng = np.random.default_rng(1)
    targets = [rng.uniform(0, 1, (h, w, 3)).astype(np.float32) for _ in cams]
    T = gs.train.Trainer(cams, targets=targets, params=params, config={"num_iterations": 100})
    P = {k: params[k].copy() for k in KEYS}
    om, ov = oracle.zeros_like_params(n), oracle.zeros_like_params(n)
    seq = [0, 3, 5, 1]
    for it in range(K):
        ci = seq[it]
        loss_sum = T.train_step(it, [ci], densify=False)
        kw = gs.scene.render_kwargs(P, cams[ci])
        img, _, buf = oracle.render_gaussians(**kw)
        o_loss = oracle.l1_loss(img, targets[ci])
        dpix = oracle.compute_image_gradients(img, targets[ci], lambda_dssim=0)
        og = oracle.backward(**gs.scene.backward_kwargs(P, cams[ci], buf, dpix))
        grads = {"positions": og["dL_dmean3D"], "scales": og["dL_dscale"], "rotations": og["dL_drot"],
                 "opacities": og["dL_dopacity"], "shs": og["dL_dshs"]}
        lr = T.learning_rates(it)
        oracle.adam_update(grads, P, om, ov, n, lr["lr_pos"], lr["lr_scale"], lr["lr_rot"], lr["lr_opac"], lr["lr_sh"],
                           0.9, 0.999, 1e-8, it)
        assert float(loss_sum.item()) / (3 * h * w) == pytest.approx(o_loss, rel=2e-3), it
    for k in KEYS:
        a, b = T.params[k].cpu().numpy().astype(np.float64), P[k].astype(np.float64)
        rel = np.linalg.norm(a - b) / np.linalg.norm(b)
        # Adam's m/sqrt(v) is sign-like: summation-order noise in near-zero gradients flips a few entries
        assert rel <= 5e-3, (k, rel)


def test_checkpoint_resume(gs, tmp_path):
    """Parameters and Adam state survive a save / load bit-exactly; training continues from there."""
    n, w, h = 2000, 64, 48
    params, _, _ = gs.scene.synthetic_scene(n, w, h, 0.02, 0.12, seed=3)
    from gsb200.utils.camera_utils import load_nerf_cameras
    cams = load_nerf_cameras(w, h)[:4]
    rng = np.random.default_rng(1)
    targets = [rng.uniform(0, 1, (h, w, 3)).astype(np.float32) for _ in cams]
    T = gs.train.Trainer(cams, targets=targets, params=params, config={"num_iterations": 100})
    for it in range(3):
        T.train_step(it, [it % 4], densify=False)
    ckpt = T.save_checkpoint(tmp_path, 2)
    T2 = gs.train.Trainer(cams, targets=targets, num_points=10, config={"num_iterations": 100})
    nxt = T2.load_checkpoint(ckpt)
    assert nxt == 3 and T2.num_points == n
    assert torch.equal(T2.params.flat, T.params.flat)
    assert torch.equal(T2.adam_m.flat, T.adam_m.flat) and torch.equal(T2.adam_v.flat, T.adam_v.flat)
    T2.train_step(nxt, [3], densify=False)
    assert torch.isfinite(T2.params.flat).all()


def test_adam_division_is_bit_identical_to_the_operator(gs):
    """The Adam kernels divide through gs_div_pos (exponent set aside, fast-path sequence, exponent put
    back).  It must equal `a / b` bit for bit: numerators over the whole binary32 range incl. zeros,
    subnormals, huge values, inf and NaN; divisors like Adam's (bias corrections, sqrt(v)+eps) and
    outside the fast window."""
    import ctypes as C
    from gsb200 import _lib
    ctx = _lib.context()
    rng = np.random.default_rng(7)
    n = 1 << 22
    bits = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)       # every exponent, both signs
    a = bits.view(np.float32).copy()
    a[:64] = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 1.17549435e-38, 3e38, -3e38] + [1e-30] * 54,
                      dtype=np.float32)
    b = np.empty(n, dtype=np.float32)
    q = n // 4
    b[:q] = (1.0 - 0.9 ** rng.integers(1, 400, q)).astype(np.float32)            # bc1
    b[q:2 * q] = (1.0 - 0.999 ** rng.integers(1, 7000, q)).astype(np.float32)    # bc2
    b[2 * q:3 * q] = np.exp(rng.uniform(np.log(1e-9), np.log(1e3), q)).astype(np.float32)   # sqrt(v^) + eps
    b[3 * q:] = np.exp(rng.uniform(np.log(1e-30), np.log(1e30), n - 3 * q)).astype(np.float32)  # outside the window too
    b[-8:] = np.array([1.0, 0.99999994, 1.0000001, 2.0 ** -40, 2.0 ** 40, 2.0 ** -41, 1e-45, np.inf], dtype=np.float32)
    # small numerators against each divisor class as well (the case Adam is in all the time)
    a[1::3] = (a[1::3] * np.float32(1e-20)).astype(np.float32)
    ta, tb = _cuda(a), _cuda(b)
    fast, const, ref = torch.empty_like(ta), torch.empty_like(ta), torch.empty_like(ta)
    ctx.check(_lib.lib().gsb_selftest_div(ctx.h, _lib.stream_ptr(ctx.device_index), n, _lib.ptr(ta), _lib.ptr(tb),
                                          _lib.ptr(fast), _lib.ptr(const), _lib.ptr(ref)))
    torch.cuda.synchronize()
    f, r = fast.cpu().numpy().view(np.uint32), ref.cpu().numpy().view(np.uint32)
    both_nan = np.isnan(fast.cpu().numpy()) & np.isnan(ref.cpu().numpy())
    bad = (f != r) & ~both_nan
    assert not bad.any(), f"{bad.sum()} of {n} quotients differ, e.g. a={a[bad][:3]} b={b[bad][:3]}"
    c = const.cpu().numpy().view(np.uint32)
    bad = (c != r) & ~(np.isnan(const.cpu().numpy()) & np.isnan(ref.cpu().numpy()))
    assert not bad.any(), f"{bad.sum()} of {n} reciprocal-based quotients differ, e.g. a={a[bad][:3]} b={b[bad][:3]}"
    # and against the host's IEEE division
    with np.errstate(all="ignore"):
        host = (a / b).view(np.uint32)
    ok = (host == f) | both_nan
    assert ok.all(), f"{(~ok).sum()} quotients differ from the host division"


def test_ssim_and_depth_loss_vs_golden_and_oracle(gs, golden_dir, oracle):
    """loss.ssim / loss.depth_loss against the values the reference's source produced under the Warp shim
    (small image, border windows everywhere) and against the oracle at 800x800 and on a ragged size."""
    g = np.load(os.path.join(golden_dir, "ref_loss2.npz"))
    assert abs(gs.loss.ssim(g["rendered"], g["target"]) - float(g["ssim"])) <= 2e-6
    assert abs(gs.loss.ssim(g["rendered"], g["rendered"]) - 1.0) <= 2e-6
    assert abs(gs.loss.depth_loss(g["rendered_depth"], g["target_depth"], g["depth_mask"]) - float(g["depth_loss"])) <= 2e-6
    rng = np.random.default_rng(5)
    for (h, w) in ((800, 800), (37, 101), (5, 3)):
        r = rng.uniform(0, 1, (h, w, 3)).astype(np.float32)
        t = np.clip(r + rng.normal(0, 0.2, (h, w, 3)), 0, 1).astype(np.float32)
        t[: h // 3] = r[: h // 3]
        want = oracle.ssim(r, t)
        got = gs.loss.ssim(_cuda(r), _cuda(t))
        # the oracle (like the reference) adds the W*H per-pixel values into ONE binary32 accumulator:
        # at 640k pixels that sum alone carries ~2e-5 of rounding; the kernel accumulates in double
        assert abs(got - want) <= 1e-7 * h * w + 5e-6, (h, w, got, want)
        rd = rng.uniform(0.05, 3.0, (h, w)).astype(np.float32)
        td = rng.uniform(0.05, 3.0, (h, w)).astype(np.float32)
        m = (rng.uniform(0, 1, (h, w)) > 0.5).astype(np.float32)
        want = oracle.depth_loss(rd, td, m)
        got = gs.loss.depth_loss(rd, td, m)
        assert abs(got - want) <= 1e-7 * h * w + 5e-6, (h, w, got, want)


@pytest.mark.parametrize("degree", [3, 1])
def test_compact_sh_gradient_expands_to_the_same_bits(gs, degree):
    """gsb_backward_compact_sh publishes the SH gradient as its two rank-1 factors (8 floats per Gaussian);
    gsb_adam_step_peers_compact rebuilds the 48 products on the owner's side.  With a world of one rank
    (no peer mapping needed) parameters and Adam moments must come out bit-identical to the full
    exchange kernel fed with the full gradient -- incl. culled Gaussians (all-zero factors), a degree
    below 3 and a Gaussian count that is not a multiple of 4."""
    import ctypes as C
    from gsb200 import _lib, backward, forward
    n, w, h = 6001, 128, 96
    params, cam, target = gs.scene.synthetic_scene(n, w, h, 0.01, 0.08, seed=11)
    kw = gs.scene.render_kwargs(params, cam)
    kw["degree"] = degree
    img, _d, buf = forward.render_gaussians(**kw)
    dpix = torch.from_numpy(np.random.default_rng(3).normal(size=(h, w, 3)).astype(np.float32)).cuda()
    FG = gs.train.FlatGaussians
    dev = img.device
    bkw = gs.scene.backward_kwargs(params, cam, buf, dpix)
    bkw["degree"] = degree
    # the tile stage once (its atomics make two runs differ in the last bits); both forms of the
    # per-Gaussian stage then start from the same dL_dmean2D / dL_dconic / dL_dcolor
    g = backward.backward(**bkw)
    ctx = _lib.context()
    stream, p = _lib.stream_ptr(ctx.device_index), _lib.ptr
    frame = _lib.make_frame(cam["world_to_camera"], cam["full_proj_matrix"], cam["camera_center"], cam["tan_fovx"],
                            cam["tan_fovy"], w, h, (0.0, 0.0, 0.0), degree, True, 1.0)
    dparams = {k: _lib.to_device(params[k], device=dev) for k in KEYS}
    runs = {}
    for compact in (False, True):
        G, P, M, V = FG(n, dev), FG(n, dev).load(params), FG(n, dev), FG(n, dev)
        G.flat.fill_(123.0)        # the compact form must not depend on what the buffer held before
        M.flat.uniform_(-1e-3, 1e-3, generator=torch.Generator(device=dev).manual_seed(1))
        V.flat.uniform_(0, 1e-5, generator=torch.Generator(device=dev).manual_seed(2))
        fn = _lib.lib().gsb_preprocess_backward_compact_sh if compact else _lib.lib().gsb_preprocess_backward
        ctx.check(fn(ctx.h, stream, C.byref(frame), n, p(dparams["positions"]), p(buf["radii"]), p(dparams["shs"]),
                     p(dparams["scales"]), p(dparams["rotations"]), p(buf["cov3Ds"]), p(buf["clamped_state"]),
                     p(g["dL_dmean2D"]), p(g["dL_dconic"]), p(g["dL_dcolor"]), p(G["positions"]), p(G["shs"]),
                     p(G["scales"]), p(G["rotations"]), p(None)))
        G["opacities"].copy_(g["dL_dopacity"].view_as(G["opacities"]))
        gp, pp = (C.c_uint64 * 1)(G.flat.data_ptr()), (C.c_uint64 * 1)(P.flat.data_ptr())
        lrs = (1e-2, 5e-3, 5e-3, 5e-3, 2e-3)
        if compact:
            scratch = torch.empty(48 * (n + 8), dtype=torch.float32, device=dev)
            ctx.check(_lib.lib().gsb_adam_step_peers_compact(
                ctx.h, stream, n, 1, 0, gp, pp, 0, p(M.flat), p(V.flat), *lrs, 0.9, 0.999, 1e-8, 5, p(scratch),
                scratch.numel(), degree, 0))
        else:
            ctx.check(_lib.lib().gsb_adam_step_peers(
                ctx.h, stream, n, 1, 0, gp, pp, 0, 0, p(M.flat), p(V.flat), *lrs, 0.9, 0.999, 1e-8, 5, 0))
        torch.cuda.synchronize()
        runs[compact] = (P.flat.clone(), M.flat.clone(), V.flat.clone(), G)
    assert torch.equal(runs[False][3]["shs"].reshape(-1), g["dL_dshs"].reshape(-1))   # the stage call = the operator
    for a, b, name in zip(runs[False][:3], runs[True][:3], ("params", "m", "v")):
        assert torch.equal(a, b), name
    # and the factors themselves: dRGB x basis_0 (= SH_C0, exact constant) is the DC row of the full gradient
    full_sh = runs[False][3]["shs"].view(n, 16, 3)
    fac = runs[True][3]["shs"].reshape(-1)[:8 * n].view(n, 8)
    assert torch.equal(full_sh[:, 0, :], fac[:, :3] * np.float32(0.28209479177387814))
    norm = fac[:, 3:6].norm(dim=1)
    live = norm > 0
    assert bool(((norm[live] - 1).abs() < 1e-5).all()) and bool((fac[:, 6:] == 0).all())


def test_init_gaussian_params_matches_reference_source(gs, golden_dir, oracle):
    """the CUDA init kernel against what the reference's init_gaussian_params (train.py:36-92)
    produced under the Warp shim (ref_misc.npz) -- incl. the wp.randf positions -- and against the oracle at a
    size that is not a multiple of the CTA."""
    g = np.load(os.path.join(golden_dir, "ref_misc.npz"))
    dev = torch.device("cuda")
    for n, want in ((32, {k: g["init_" + k] for k in KEYS}), (70001, oracle.init_gaussian_params(70001, 0.1))):
        P = gs.train.FlatGaussians(n, dev, fill=None)
        P.flat.fill_(float("nan"))                  # the kernel must write every element
        gs.optimizer.init_gaussian_params(P["positions"], P["scales"], P["rotations"], P["opacities"], P["shs"], n, 0.1)
        for k in KEYS:
            got = P[k].cpu().numpy()
            assert np.array_equal(got, np.asarray(want[k]).reshape(got.shape)), (n, k)


def test_zero_gradients_matches_reference_source(gs, golden_dir):
    """zero_gradients (train.py:94-115) run by the reference's own kernel under the shim on non-zero
    arrays that are three Gaussians longer than num_points: the first num_points entries become zero, the tail
    keeps its values."""
    g = np.load(os.path.join(golden_dir, "ref_cameras.npz"))
    n = int(g["zg_n"])
    arrs = {k: _cuda(g["zg_in_" + k]) for k in KEYS}
    gs.optimizer.zero_gradients(arrs["positions"], arrs["scales"], arrs["rotations"], arrs["opacities"], arrs["shs"], n)
    for k in KEYS:
        want = g["zg_out_" + k]
        assert np.array_equal(arrs[k].cpu().numpy(), want), k
        assert not want[: n * (16 if k == "shs" else 1)].any() and want[n * (16 if k == "shs" else 1):].all()


def test_step_loop_with_densify_vs_oracle(gs, oracle):
    """eight iterations of the reference loop (train.py:926-1064) crossing two
    densify events (clone + split + prune at iterations 4 and 8), on the GPU and in the oracle, for the same
    camera sequence.  The Gaussian count after every iteration must be the oracle's -- exactly, unless a
    candidate sits so close to a threshold that the 1e-3 gradient tolerance can flip it; the oracle counts those
    (the first event is constructed to have none: the gradient threshold lies in a 2% gap of the norms)."""
    from gsb200.utils.camera_utils import load_nerf_cameras
    n, w, h = 3000, 96, 64
    params, _cam0, _ = gs.scene.synthetic_scene(n, w, h, 0.02, 0.12, seed=11)
    cams = load_nerf_cameras(w, h)[:6]
    rng = np.random.default_rng(1)
    targets = [rng.uniform(0, 1, (h, w, 3)).astype(np.float32) for _ in cams]
    cfg = {"densify_from_iter": 0, "densification_interval": 4, "densify_grad_threshold": 0.00045, "percent_dense": 0.015,
           "cull_opacity_threshold": 0.1, "min_valid_points": 100}
    T = gs.train.Trainer(cams, targets=targets, params=params,
                         config={"num_iterations": 7000, "use_lr_scheduler": False, **cfg})
    st = {"params": {k: params[k].copy() for k in KEYS}, "grads": oracle.zeros_like_params(n),
          "adam_m": oracle.zeros_like_params(n), "adam_v": oracle.zeros_like_params(n), "num_points": n,
          "scene_extent": T.scene_extent}
    seq = [0, 3, 5, 1, 2, 4, 0, 3]
    oracle.set_threads(oracle.max_threads())
    slack, events = 0, 0
    try:
        for it in range(1, 9):
            ci = seq[it - 1]
            loss_sum = T.train_step(it, [ci], densify=True)
            P, m = st["params"], st["num_points"]
            img, _, buf = oracle.render_gaussians(**gs.scene.render_kwargs(P, cams[ci]))
            dpix = oracle.compute_image_gradients(img, targets[ci], lambda_dssim=0)
            og = oracle.backward(**gs.scene.backward_kwargs(P, cams[ci], buf, dpix))
            st["grads"] = {"positions": og["dL_dmean3D"], "scales": og["dL_dscale"], "rotations": og["dL_drot"],
                           "opacities": og["dL_dopacity"], "shs": og["dL_dshs"]}
            oracle.adam_update(st["grads"], P, st["adam_m"], st["adam_v"], m, 1e-2, 5e-3, 5e-3, 5e-3, 2e-3, 0.9, 0.999,
                               1e-8, it)
            assert float(loss_sum.item()) / (3 * h * w) == pytest.approx(oracle.l1_loss(img, targets[ci]), rel=5e-3), it
            if T.densify_due(it):
                # candidates the tolerance could flip: gradient norm within 0.5% of the threshold, largest scale
                # within 0.1% of percent_dense * extent, opacity within 0.1% of the cull threshold
                nr = np.linalg.norm(st["grads"]["positions"], axis=1)
                gt, sthr = cfg["densify_grad_threshold"], cfg["percent_dense"] * T.scene_extent
                near = int((np.abs(nr - gt) < 5e-3 * gt).sum())
                near += int((np.abs(P["scales"].max(axis=1) - sthr) < 1e-3 * sthr).sum())
                near += int((np.abs(P["opacities"] - 0.1) < 1e-4).sum())
                if events == 0:
                    assert near <= 4, "the first event is meant to have (almost) no borderline candidates"
                slack += 3 * near          # a flipped split adds two Gaussians and removes one
                events += 1
            log = oracle.densification_and_pruning(st, it, cfg)
            if T.densify_due(it):
                assert log["cloned"] > 0 and log["split"] > 0 and log["pruned"] > 0, log
            assert abs(T.num_points - st["num_points"]) <= slack, (it, T.num_points, st["num_points"], slack)
            if T.num_points == st["num_points"] and it == 4:
                # same masks => the same Gaussians in the same order: parameters agree like in the plain loop
                for k in KEYS:
                    a, b = T.params[k].cpu().numpy().astype(np.float64), st["params"][k].astype(np.float64)
                    b = b.reshape(a.shape)
                    assert np.linalg.norm(a - b) <= 5e-3 * np.linalg.norm(b), (k, it)
    finally:
        oracle.set_threads(1)
    assert events == 2 and T.num_points != n
