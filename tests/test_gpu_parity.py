"""GPU parity: the sm_100a path (through the C ABI) against the CPU oracle and against the golden
vectors generated from the reference's own source.  Integer / index outputs must be bit-exact;
images, depths and per-Gaussian floats within the stated fp32 tolerances; gradients rel 1e-3."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

INT_KEYS = ["radii", "point_offsets", "point_list", "ranges", "n_contrib"]
FLOAT_KEYS = ["points_xy_image", "depths", "colors", "cov3Ds", "conic_opacity", "final_Ts", "clamped_state"]
GRAD_KEYS = ["dL_dmean3D", "dL_dcolor", "dL_dshs", "dL_dopacity", "dL_dscale", "dL_drot", "dL_dmean2D", "dL_dconic",
             "dL_dcov3D"]
IMAGE_ATOL = 1e-4     # north_star: max abs 1e-4 on images
GRAD_RTOL = 1e-3      # north_star: rel 1e-3 on gradients (norm-wise per tensor)


@pytest.fixture(scope="module")
def gs():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import gsb200  # noqa: F401
    from gsb200 import backward, forward, scene
    import types
    return types.SimpleNamespace(forward=forward, backward=backward, scene=scene)


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _np(t):
    return t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)


def check_forward(got, want, what=""):
    img, depth, buf = got
    o_img, o_depth, o_buf = want
    for k in INT_KEYS:
        a, b = _np(buf[k]).reshape(o_buf[k].shape), o_buf[k]
        assert np.array_equal(a, b), f"{what}{k}: {np.count_nonzero(a != b)} of {b.size} entries differ"
    for k in FLOAT_KEYS:
        a, b = _np(buf[k]).reshape(o_buf[k].shape), o_buf[k]
        assert np.array_equal(a, b) or np.allclose(a, b, rtol=1e-5, atol=1e-6), f"{what}{k}"
    assert np.abs(_np(img) - o_img).max() <= IMAGE_ATOL, what + "image"
    assert np.abs(_np(depth) - o_depth).max() <= IMAGE_ATOL * max(1.0, np.abs(o_depth).max()), what + "depth"


def check_grads(g, og, what=""):
    assert set(g) == set(GRAD_KEYS)
    for k in GRAD_KEYS:
        a = _np(g[k]).reshape(og[k].shape).astype(np.float64)
        b = og[k].astype(np.float64)
        err = np.linalg.norm(a - b)
        ref = np.linalg.norm(b)
        assert err <= GRAD_RTOL * ref + 1e-12, f"{what}{k}: rel err {err / max(ref, 1e-30):.3e}"
        assert np.isfinite(a).all(), k


def _golden_kwargs(g):
    return dict(background=g["in_bg"], means3D=g["in_means"], colors=None, opacity=g["in_opac"],
                scales=g["in_scales"], rotations=g["in_rot"], scale_modifier=float(g["in_scale_modifier"]),
                viewmatrix=g["in_view"], projmatrix=g["in_proj"], tan_fovx=float(g["in_tan_fovx"]),
                tan_fovy=float(g["in_tan_fovy"]), image_height=int(g["in_H"]), image_width=int(g["in_W"]),
                sh=g["in_sh"], degree=int(g["in_degree"]), campos=g["in_campos"], clamped=bool(g["in_clamped"]))


@pytest.mark.parametrize("name", ["ref_lego_small", "ref_lego_deg1", "ref_lego_bg", "ref_example_96"])
def test_golden_from_reference_source(gs, golden_dir, name):
    """Forward + backward on the inputs of tests/golden/*.npz, compared with the arrays the
    reference's own kernels produced under the Warp shim."""
    g = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    img, depth, buf = gs.forward.render_gaussians(**_golden_kwargs(g))
    want = {k[4:]: v for k, v in g.items() if k.startswith("fwd_")}
    # the reference's 12 keys, plus our one documented extra (the culling masks handed on to backward)
    assert set(buf) - {"block_masks"} == set(want)
    check_forward((img, depth, buf), (g["image"], g["depth"], want), name + ": ")
    if int(g["ref_bwd_oob"]):
        return
    grads = gs.backward.backward(
        background=g["in_bg"], means3D=g["in_means"], dL_dpixels=g["dL_dpixels"], opacity=g["in_opac"],
        shs=g["in_sh"], scales=g["in_scales"], rotations=g["in_rot"], scale_modifier=float(g["in_scale_modifier"]),
        viewmatrix=g["in_view"], projmatrix=g["in_proj"], tan_fovx=float(g["in_tan_fovx"]),
        tan_fovy=float(g["in_tan_fovy"]), image_height=int(g["in_H"]), image_width=int(g["in_W"]),
        campos=g["in_campos"], radii=buf["radii"], means2D=buf["points_xy_image"],
        conic_opacity=buf["conic_opacity"], rgb=buf["colors"], cov3Ds=buf["cov3Ds"], clamped=buf["clamped_state"],
        geom_buffer={"radii": buf["radii"], "means2D": buf["points_xy_image"], "conic_opacity": buf["conic_opacity"],
                     "rgb": buf["colors"], "clamped_state": buf["clamped_state"]},
        binning_buffer={"point_list": buf["point_list"]},
        img_buffer={"ranges": buf["ranges"], "final_Ts": buf["final_Ts"], "n_contrib": buf["n_contrib"]},
        degree=int(g["in_degree"]))
    check_grads(grads, {k[4:]: v for k, v in g.items() if k.startswith("bwd_")}, name + ": ")


def test_example_scene_config1(gs, oracle):
    """BASELINE config 1: render.py's 3-Gaussian scene at 1800x1800, pixel-wise vs the oracle."""
    kw = gs.scene.example_render_kwargs()
    got = gs.forward.render_gaussians(**kw)
    want = oracle.render_gaussians(**kw)
    check_forward(got, want, "C1: ")
    assert _np(got[2]["radii"]).tolist() == [542, 485, 542]
    assert got[2]["point_list"].numel() == 11779


@pytest.mark.parametrize("n,w,h,smin,smax,cull", [(3000, 100, 70, 0.01, 0.08, 1), (20000, 320, 240, 0.005, 0.05, 1),
                                                  (20000, 333, 251, 0.005, 0.05, 0), (8000, 256, 256, 0.01, 0.1, 1),
                                                  (2000, 40, 24, 0.3, 2.0, 1)])
def test_synthetic_forward_backward_vs_oracle(gs, oracle, n, w, h, smin, smax, cull):
    """cull = 0 switches the per-block culling masks of the tile kernels off: same results."""
    from gsb200 import _lib
    _lib.context().set_option("blend_cull", cull)
    try:
        params, cam, target = gs.scene.synthetic_scene(n, w, h, smin, smax, seed=n + w)
        kw = gs.scene.render_kwargs(params, cam)
        got = gs.forward.render_gaussians(**kw)
        oracle.set_threads(oracle.max_threads())
        want = oracle.render_gaussians(**kw)
        check_forward(got, want)
        dpix = oracle.compute_image_gradients(want[0], target, lambda_dssim=0)
        grads = gs.backward.backward(**gs.scene.backward_kwargs(params, cam, got[2], dpix))
        oracle.set_threads(1)        # serial: deterministic accumulation order in the checker
        ograds = oracle.backward(**gs.scene.backward_kwargs(params, cam, want[2], dpix))
        check_grads(grads, ograds)
    finally:
        oracle.set_threads(1)
        _lib.context().set_option("blend_cull", 1)


@pytest.mark.parametrize("mode,hand_masks_on", [(0, True), (2, True), (0, False), (2, False)])
@pytest.mark.parametrize("n,w,h,smin,smax,bg", [(12000, 200, 136, 0.005, 0.05, (0.0, 0.0, 0.0)),
                                                (6000, 123, 77, 0.02, 0.3, (0.2, 0.5, 0.9))])
def test_backward_reduction_modes_agree_with_oracle(gs, oracle, mode, hand_masks_on, n, w, h, smin, smax, bg):
    """The two pixel reductions of the backward tile kernel (0: warp-shuffle butterfly with the exact exponential,
    2: tensor-core moments with MUFU) both meet the gradient tolerance, incl. ragged edge tiles,
    Gaussians far larger than a tile and a coloured background (the bg . dL_dpixel term) -- with the
    forward's culling masks handed on and with the masks recomputed."""
    from gsb200 import _lib
    _lib.context().set_option("bwd_reduce", mode)
    try:
        params, cam, target = gs.scene.synthetic_scene(n, w, h, smin, smax, seed=7 * n + w)
        kw = gs.scene.render_kwargs(params, cam, background=bg)
        got = gs.forward.render_gaussians(**kw)
        oracle.set_threads(oracle.max_threads())
        want = oracle.render_gaussians(**kw)
        check_forward(got, want)
        rng = np.random.default_rng(n)
        dpix = rng.normal(size=(h, w, 3)).astype(np.float32)       # arbitrary dL_dpixels, not only +-1/(3HW)
        buffers = dict(got[2])
        assert buffers["block_masks"].numel() == buffers["point_list"].numel()
        if not hand_masks_on:          # the reference's 12 keys only: the backward recomputes the culling masks
            del buffers["block_masks"]
        grads = gs.backward.backward(**gs.scene.backward_kwargs(params, cam, buffers, dpix, background=bg))
        oracle.set_threads(1)
        ograds = oracle.backward(**gs.scene.backward_kwargs(params, cam, want[2], dpix, background=bg))
        check_grads(grads, ograds)
    finally:
        oracle.set_threads(1)
        _lib.context().set_option("bwd_reduce", 2)


@pytest.mark.parametrize("hand_masks_on", [True, False])
def test_backward_scalar_red_path_agrees_with_oracle(gs, oracle, hand_masks_on):
    """bwd_packed = 0: the tensor-core backward adds its nine sums per Gaussian straight into the reference's four
    arrays (scalar REDs) instead of the packed 48-byte records of the default path.  Same tolerance."""
    from gsb200 import _lib
    n, w, h = 9000, 180, 120
    _lib.context().set_option("bwd_packed", 0)
    try:
        params, cam, target = gs.scene.synthetic_scene(n, w, h, 0.005, 0.08, seed=77)
        kw = gs.scene.render_kwargs(params, cam, background=(0.1, 0.3, 0.7))
        got = gs.forward.render_gaussians(**kw)
        oracle.set_threads(oracle.max_threads())
        want = oracle.render_gaussians(**kw)
        check_forward(got, want)
        dpix = np.random.default_rng(5).normal(size=(h, w, 3)).astype(np.float32)
        buffers = dict(got[2])
        if not hand_masks_on:
            del buffers["block_masks"]
        grads = gs.backward.backward(**gs.scene.backward_kwargs(params, cam, buffers, dpix, background=(0.1, 0.3, 0.7)))
        oracle.set_threads(1)
        ograds = oracle.backward(**gs.scene.backward_kwargs(params, cam, want[2], dpix, background=(0.1, 0.3, 0.7)))
        check_grads(grads, ograds)
    finally:
        oracle.set_threads(1)
        _lib.context().set_option("bwd_packed", 1)


def test_backward_writes_into_caller_buffers(gs):
    """backward(out=...) (our extension): results land in the given tensors, the others are allocated."""
    params, cam, target = gs.scene.synthetic_scene(3000, 96, 64, 0.01, 0.1, seed=3)
    img, _d, buf = gs.forward.render_gaussians(**gs.scene.render_kwargs(params, cam))
    dpix = torch.from_numpy(np.random.default_rng(1).normal(size=(64, 96, 3)).astype(np.float32)).cuda()
    kw = gs.scene.backward_kwargs(params, cam, buf, dpix)
    ref = gs.backward.backward(**kw)
    flat = torch.full((3000 * 51 + 16,), 7.0, device="cuda")
    out = {"dL_dmean3D": flat[:9000].view(3000, 3), "dL_dshs": flat[9000:9000 + 144000].view(48000, 3)}
    got = gs.backward.backward(**kw, out=out)
    assert got["dL_dmean3D"].data_ptr() == flat.data_ptr() and got["dL_dshs"].data_ptr() == flat[9000:].data_ptr()
    assert bool((flat[153000:] == 7.0).all())                      # nothing written past the views
    for k in ref:                           # two runs differ by the order of the atomics only
        a, b = got[k].double(), ref[k].double()
        assert float((a - b).norm()) <= 1e-5 * float(b.norm()) + 1e-12, k
    with pytest.raises(ValueError):
        gs.backward.backward(**kw, out={"dL_dscale": torch.empty(5, device="cuda")})


def test_culling_mask_covers_every_contributing_pixel(gs):
    """Property test of gs_block_mask, the conservative ellipse / half-row test both tile kernels cull with:
    for random conics (isotropic to needle-like, sub-pixel to thousands of pixels), centres inside, next
    to and far from the tile, every opacity regime and degenerate inputs, each pixel that passes the
    forward's own tests (power <= 0, alpha >= 1/255, contract arithmetic) must lie in a kept half row."""
    from gsb200 import _lib
    ctx = _lib.context()
    rng = np.random.default_rng(123)
    n = 400_000
    # covariance = R diag(s1^2, s2^2) R^T + 0.3 I  -> conic (what preprocess produces), plus raw extremes
    s1 = np.exp(rng.uniform(np.log(0.05), np.log(3000.0), n))
    s2 = s1 * np.exp(rng.uniform(np.log(1e-3), 0.0, n))
    th = rng.uniform(0, np.pi, n)
    c, s = np.cos(th), np.sin(th)
    xx = c * c * s1 * s1 + s * s * s2 * s2 + 0.3
    yy = s * s * s1 * s1 + c * c * s2 * s2 + 0.3
    xy = c * s * (s1 * s1 - s2 * s2)
    det = xx * yy - xy * xy
    ca, cb, cc = yy / det, -xy / det, xx / det
    x0 = 16.0 * rng.integers(0, 240, n)
    y0 = 16.0 * rng.integers(0, 135, n)
    reach = 3.5 * s1 + 24.0
    gx = x0 + 8.0 + rng.uniform(-1, 1, n) * reach * rng.choice([0.05, 0.5, 1.0, 1.5], n)
    gy = y0 + 8.0 + rng.uniform(-1, 1, n) * reach * rng.choice([0.05, 0.5, 1.0, 1.5], n)
    op = rng.choice([1.0, 0.99, 0.5, 0.05, 0.0040, 1.0 / 255.0, 0.00393, 0.0, 2.0], n)
    op = np.where(rng.uniform(0, 1, n) < 0.5, rng.uniform(0, 1, n), op)
    cases = np.stack([gx, gy, ca, cb, cc, op, x0, y0], axis=1).astype(np.float32)
    # degenerate / hostile inputs: must never crash and must stay conservative
    cases[:64, 2] = np.repeat(np.array([0.0, -1.0, np.nan, np.inf, 1e-30, 1e30, 3.3, 1e-12], dtype=np.float32), 8)
    cases[64:96, 3] = np.float32(np.nan)
    cases[96:128, 5] = np.float32(np.nan)
    cases[128:160, 0] = np.repeat(np.array([1e7, -1e7, np.inf, np.nan], dtype=np.float32), 8)
    tc = _cuda(cases)
    mask = torch.empty(n, dtype=torch.int32, device="cuda")
    active = torch.empty((n, 8), dtype=torch.int32, device="cuda")
    ctx.check(_lib.lib().gsb_selftest_block_mask(ctx.h, _lib.stream_ptr(ctx.device_index), n, _lib.ptr(tc),
                                                 _lib.ptr(mask), _lib.ptr(active)))
    torch.cuda.synchronize()
    m = mask.cpu().numpy().view(np.uint32)
    a = active.cpu().numpy().view(np.uint32)
    act = ((a[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).reshape(n, 16, 2, 8).any(axis=3)  # [n, row, half]
    kept = ((m[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).reshape(n, 16, 2).astype(bool)
    missed = act & ~kept
    assert not missed.any(), f"{missed.any(axis=(1, 2)).sum()} cases lose a contributing half row, e.g. {cases[missed.any(axis=(1, 2))][:3]}"
    # and the test is worth its cost: a good part of the half rows without any contributing pixel is culled
    empty = ~act
    assert kept[empty].mean() < 0.5, kept[empty].mean()


def test_nothing_visible_gives_zero_image(gs, oracle):
    """forward.py:830: when no Gaussian is rendered the image is all ZEROS, not background."""
    params, cam, _ = gs.scene.synthetic_scene(500, 64, 48, 0.01, 0.05)
    params["positions"] = params["positions"] + np.float32(100.0) * cam["camera_center"].astype(np.float32)
    kw = gs.scene.render_kwargs(params, cam, background=(0.3, 0.6, 0.9))
    img, depth, buf = gs.forward.render_gaussians(**kw)
    o = oracle.render_gaussians(**kw)
    assert buf["point_list"].numel() == 0 and o[2]["point_list"].size == 0
    assert not _np(img).any() and not o[0].any()
    check_forward((img, depth, buf), o)


def test_torch_inputs_and_column_opacity(gs, oracle):
    """Accepts torch tensors on any device and (N,1) opacities / (N,16,3) SH like to_warp_array."""
    params, cam, _ = gs.scene.synthetic_scene(2000, 96, 64, 0.01, 0.06, seed=5)
    kw = gs.scene.render_kwargs(params, cam)
    kw_t = dict(kw)
    kw_t["means3D"] = torch.from_numpy(params["positions"]).cuda()
    kw_t["opacity"] = torch.from_numpy(params["opacities"]).reshape(-1, 1)
    kw_t["sh"] = torch.from_numpy(params["shs"]).reshape(-1, 16, 3)
    kw_t["background"] = torch.zeros(3)
    check_forward(gs.forward.render_gaussians(**kw_t), oracle.render_gaussians(**kw))


@pytest.mark.parametrize("degree,clamped,bg", [(0, True, (0.0, 0.0, 0.0)), (1, False, (1.0, 1.0, 1.0)),
                                               (2, True, (0.1, 0.7, 0.3))])
def test_lower_sh_degrees_vs_oracle(gs, oracle, degree, clamped, bg):
    """SH degree < 3 (rows keep stride 16), clamped on/off, non-black backgrounds; the reference's
    own backward is out of bounds here (tests/golden/ref_lego_deg1.npz), the oracle defines it."""
    params, cam, target = gs.scene.synthetic_scene(5000, 128, 96, 0.01, 0.08, seed=degree + 20)
    kw = gs.scene.render_kwargs(params, cam, background=bg, degree=degree, clamped=clamped, scale_modifier=0.7)
    got = gs.forward.render_gaussians(**kw)
    want = oracle.render_gaussians(**kw)
    check_forward(got, want)
    dpix = oracle.compute_image_gradients(want[0], target, lambda_dssim=0)
    bkw = gs.scene.backward_kwargs(params, cam, got[2], dpix, background=bg, degree=degree, scale_modifier=0.7)
    okw = gs.scene.backward_kwargs(params, cam, want[2], dpix, background=bg, degree=degree, scale_modifier=0.7)
    check_grads(gs.backward.backward(**bkw), oracle.backward(**okw))


@pytest.mark.parametrize("n", [0, 1, 2])
def test_tiny_inputs(gs, oracle, n):
    """Empty and near-empty inputs."""
    params, cam, target = gs.scene.synthetic_scene(max(n, 1), 48, 32, 0.2, 0.5, seed=1)
    params = {k: v[: n * (16 if k == "shs" else 1)] for k, v in params.items()}
    kw = gs.scene.render_kwargs(params, cam, background=(0.2, 0.3, 0.4))
    got = gs.forward.render_gaussians(**kw)
    if n == 0:
        assert got[2]["point_list"].numel() == 0 and not _np(got[0]).any()      # forward.py:830: zeros
        assert got[2]["radii"].numel() == 0 and tuple(got[0].shape) == (32, 48, 3)
        return
    want = oracle.render_gaussians(**kw)
    check_forward(got, want)
    dpix = oracle.compute_image_gradients(want[0], target, lambda_dssim=0)
    check_grads(gs.backward.backward(**gs.scene.backward_kwargs(params, cam, got[2], dpix, background=(0.2, 0.3, 0.4))),
                oracle.backward(**gs.scene.backward_kwargs(params, cam, want[2], dpix, background=(0.2, 0.3, 0.4))))


def test_too_many_rendered_raises_value_error(gs):
    """forward.py:765-767: more than 2^30 duplicates -> ValueError (here: 70k splats covering all
    16384 tiles of a 2048x2048 image = 1.15e9 pairs; only the counting pass runs)."""
    n = 70000
    params, cam, _ = gs.scene.synthetic_scene(n, 2048, 2048, 40.0, 50.0, seed=2, with_target=False)
    params["positions"] *= 0.05
    with pytest.raises(ValueError, match="exceeds the maximum"):
        gs.forward.render_gaussians(**gs.scene.render_kwargs(params, cam))
