"""Stage-level GPU parity through the C ABI: prefix sum, duplicate-with-keys, the stable 64-bit radix
sort, tile ranges, and the two binning paths of gsb_forward (bit-exact integer work)."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import gsb200  # noqa: F401
    from gsb200 import _lib
    return _lib


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("n", [1, 2, 255, 2048, 2049, 300000, 1234567])
def test_prefix_sum(L, n):
    rng = np.random.default_rng(n)
    a = rng.integers(0, 30, n).astype(np.int32)
    ctx, s = L.context(), L.stream_ptr()
    d_in, d_out = _cuda(a), torch.empty(n, dtype=torch.int32, device="cuda")
    total = C.c_int64(0)
    ctx.check(L.lib().gsb_scan_tiles(ctx.h, s, n, L.ptr(d_in), L.ptr(d_out), C.byref(total)))
    want = np.cumsum(a, dtype=np.int64)
    assert np.array_equal(d_out.cpu().numpy(), want.astype(np.int32))          # utils/wp_utils.py:46-60 (inclusive)
    assert total.value == int(want[-1])
    last = C.c_int32(0)
    ctx.check(L.lib().gsb_scan_mask(ctx.h, s, n, L.ptr(d_in), L.ptr(d_out), C.byref(last)))
    excl = want - a
    assert np.array_equal(d_out.cpu().numpy(), excl.astype(np.int32))          # array_scan(inclusive=False)
    assert last.value == int(excl[-1])                                         # quirk G5: last flag not counted


@pytest.mark.parametrize("coop", [1, 0])
@pytest.mark.parametrize("n,bits,distinct", [(1, 44, 0), (77, 44, 0), (4096, 44, 0), (4097, 47, 0), (200000, 44, 0),
                                              (200000, 44, 37), (1 << 20, 64, 0), (300001, 40, 2),
                                              (2049, 24, 3), (12289, 33, 0), (148 * 12288, 44, 0), (148 * 12288 + 1, 41, 0)])
def test_radix_sort_pairs_is_stable(L, n, bits, distinct, coop):
    """Ascending, STABLE (ties keep their input order) -- forward.py:799-803 [Warp].  coop = 1: one cooperative
    launch for all passes (inputs of up to num_sms x 12288 pairs); 0: three kernels per pass (any size)."""
    rng = np.random.default_rng(n + bits)
    if distinct:
        keys = rng.integers(0, distinct, n).astype(np.int64) << 20          # massive ties
    else:
        keys = rng.integers(0, 1 << 62, n, dtype=np.int64) >> (62 - min(bits, 62))
    if bits == 64:
        keys = keys | (rng.integers(0, 2, n).astype(np.int64) << 62)
    vals = np.arange(n, dtype=np.int32)
    ctx, s = L.context(), L.stream_ptr()
    dk, dv = _cuda(keys), _cuda(vals)
    ctx.set_option("sort_coop", coop)
    try:
        for _ in range(2):   # twice: the cooperative kernel must leave its grid-barrier counters reusable
            dk, dv = _cuda(keys), _cuda(vals)
            ctx.check(L.lib().gsb_sort_pairs64(ctx.h, s, L.ptr(dk), L.ptr(dv), None, None, n, 0, bits))
            torch.cuda.synchronize()
    finally:
        ctx.set_option("sort_coop", 1)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(dk.cpu().numpy(), keys[order])
    assert np.array_equal(dv.cpu().numpy(), vals[order])


def _scene(n, w, h, smin, smax, seed):
    import gsb200  # noqa: F401
    from gsb200 import scene
    params, cam, _ = scene.synthetic_scene(n, w, h, smin, smax, seed=seed, with_target=False)
    return scene.render_kwargs(params, cam)


def test_duplicate_keys_and_ranges_vs_oracle(L, oracle):
    kw = _scene(15000, 200, 150, 0.005, 0.06, 3)
    o_img, _, ob = oracle.render_gaussians(**kw, return_extra=True)
    ctx, s = L.context(), L.stream_ptr()
    D = ob["_num_rendered"]
    xy, depths = _cuda(ob["points_xy_image"]), _cuda(ob["depths"])
    offs, radii = _cuda(ob["point_offsets"]), _cuda(ob["radii"])
    keys = torch.empty(D, dtype=torch.int64, device="cuda")
    vals = torch.empty(D, dtype=torch.int32, device="cuda")
    ctx.check(L.lib().gsb_duplicate_with_keys(ctx.h, s, 200, 150, 15000, L.ptr(xy), L.ptr(depths), L.ptr(offs),
                                              L.ptr(radii), D, L.ptr(keys), L.ptr(vals)))
    assert np.array_equal(keys.cpu().numpy(), ob["_keys_unsorted"])        # tile|depth keys, bit-exact
    assert np.array_equal(vals.cpu().numpy(), ob["_vals_unsorted"])
    gx, gy = (200 + 15) // 16, (150 + 15) // 16
    bits = 32 + int(np.ceil(np.log2(gx * gy)))
    ctx.check(L.lib().gsb_sort_pairs64(ctx.h, s, L.ptr(keys), L.ptr(vals), None, None, D, 0, bits))
    assert np.array_equal(keys.cpu().numpy(), ob["_keys_sorted"])
    assert np.array_equal(vals.cpu().numpy(), ob["point_list"])
    ranges = torch.empty((gx * gy, 2), dtype=torch.int32, device="cuda")
    ctx.check(L.lib().gsb_tile_ranges(ctx.h, s, D, L.ptr(keys), gx * gy, L.ptr(ranges)))
    assert np.array_equal(ranges.cpu().numpy(), ob["ranges"])


@pytest.mark.parametrize("n,w,h,smin,smax,heavy", [
    (20000, 320, 240, 0.005, 0.05, 0),   # short lists  (<= 1024 per tile)
    (6000, 64, 48, 0.1, 0.4, 0),         # long lists   (> 1024 per tile)
    (9000, 32, 32, 0.5, 1.5, 0),         # every Gaussian in every tile (> 4096)
    (14000, 96, 64, 0.02, 0.6, 0),       # mixed: some tiles below, some above 4096
    (18000, 16, 16, 1.0, 2.0, 0),        # one tile, > 16384 entries: radix fallback
    (6000, 64, 48, 0.1, 0.4, 1),         # a third of the Gaussians at ONE position: hundreds of equal depths per tile
    (3000, 160, 96, 0.02, 0.2, 2)])      # all depths within a few ulp of each other (one plane facing the camera)
def test_both_binning_paths_match_oracle(L, oracle, n, w, h, smin, smax, heavy):
    """gsb_forward's per-tile counting sort + shared-memory sort and the global radix sort give the
    same point_list / ranges / n_contrib as the oracle's stable sort (incl. exact depth ties; the degenerate depth
    distributions send the one-pass bucket sort to its bitonic fallback)."""
    import gsb200  # noqa: F401
    from gsb200 import forward
    kw = _scene(n, w, h, smin, smax, n)
    kw["means3D"] = kw["means3D"].copy()
    kw["means3D"][1::7] = kw["means3D"][0::7][: len(kw["means3D"][1::7])]     # duplicates => depth ties
    if heavy == 1:
        kw["means3D"][::3] = kw["means3D"][0]
    elif heavy == 2:
        # move every centre onto the plane through the first one, perpendicular to the viewing direction
        view = np.asarray(kw["viewmatrix"], dtype=np.float64).reshape(4, 4)
        fwd = view[:3, 2]                     # depth = (p, 1) . view[:, 2]  (row vector times matrix, forward.py:246)
        m = kw["means3D"].astype(np.float64)
        m -= np.outer((m - m[0]) @ fwd, fwd)
        kw["means3D"] = m.astype(np.float32)
    oracle.set_threads(oracle.max_threads())
    try:
        o_img, _, ob = oracle.render_gaussians(**kw)
    finally:
        oracle.set_threads(1)
    longest = int(np.diff(ob["ranges"], axis=1).max())
    if heavy == 2:
        vis = ob["radii"].reshape(-1) > 0
        assert vis.sum() > 100 and np.ptp(ob["depths"].reshape(-1)[vis]) <= 1e-5 * ob["depths"].reshape(-1)[vis].mean()
    ctx = L.context()
    # (binning, tile_sort): bitonic per tile, per-tile radix sort (bitonic for tiles > 4096), the choice between the
    # two by list length, the one-pass bucket sort by depth (default), global radix sort
    for mode, tile_sort in ((0, 0), (0, 1), (0, 2), (0, 3), (1, 2)):
        ctx.set_option("binning", mode)
        ctx.set_option("tile_sort", tile_sort)
        try:
            img, _, buf = forward.render_gaussians(**kw)
        finally:
            ctx.set_option("binning", 0)
            ctx.set_option("tile_sort", 3)
        for k in ("point_offsets", "point_list", "ranges", "n_contrib", "radii"):
            assert np.array_equal(buf[k].cpu().numpy().reshape(ob[k].shape), ob[k]), (mode, tile_sort, k, longest)
        assert np.abs(img.cpu().numpy() - o_img).max() <= 1e-4


def test_speculative_frames_match_waiting_frames(L):
    """gsb_forward queues scatter / sort / blend behind the tile scan without waiting for D, on assumptions taken from
    the previous frame and checked on the device.  A sequence of frames whose D and longest tile list jump up and
    down must give, frame by frame, exactly what the waiting path gives."""
    import gsb200  # noqa: F401
    from gsb200 import forward
    seq = [(4000, 160, 96, 0.01, 0.05), (4000, 160, 96, 0.01, 0.05), (9000, 64, 48, 0.1, 0.4), (9000, 64, 48, 0.1, 0.4),
           (300, 160, 96, 0.01, 0.03), (20000, 320, 240, 0.005, 0.05), (20000, 320, 240, 0.005, 0.05),
           (9000, 32, 32, 0.5, 1.5), (4000, 160, 96, 0.01, 0.05)]
    ctx = L.context()
    res = {}
    for spec in (0, 1):
        ctx.set_option("speculate", spec)
        try:
            out = []
            for (n, w, h, smin, smax) in seq:
                img, dep, buf = forward.render_gaussians(**_scene(n, w, h, smin, smax, n))
                out.append((img.cpu().numpy(), dep.cpu().numpy(),
                            {k: buf[k].cpu().numpy() for k in ("point_list", "ranges", "n_contrib", "final_Ts")}))
        finally:
            ctx.set_option("speculate", 1)
        res[spec] = out
    for (a_img, a_dep, a_buf), (b_img, b_dep, b_buf) in zip(res[0], res[1]):
        assert np.array_equal(a_img, b_img) and np.array_equal(a_dep, b_dep)
        for k in a_buf:
            assert np.array_equal(a_buf[k], b_buf[k]), k


def test_forward_with_a_colour_dependency_gives_the_same_outputs(L):
    """gsb_set_color_dependency: the next gsb_forward runs preprocess without the SH -> RGB evaluation and computes
    rgb / clamped_state with a second kernel behind the event (the multi-GPU trainer lets the SH part of its exchange
    run beside the frame's geometry preprocess and binning).  Every output must be identical; the dependency is one-shot."""
    import ctypes as C
    import gsb200  # noqa: F401
    from gsb200 import forward
    ctx = L.context()
    keys = ("colors", "clamped_state", "radii", "point_list", "ranges", "n_contrib", "final_Ts", "conic_opacity", "depths")
    for (n, w, h, smin, smax) in ((4000, 160, 96, 0.01, 0.05), (9000, 64, 48, 0.1, 0.4), (300, 160, 96, 0.01, 0.03)):
        kw = _scene(n, w, h, smin, smax, n)
        for spec in (1, 0):
            ctx.set_option("speculate", spec)
            try:
                forward.render_gaussians(**kw)                      # (gives the speculative path its previous frame)
                img0, dep0, buf0 = forward.render_gaussians(**kw)
                ev = torch.cuda.Event()
                ev.record()
                ctx.check(L.lib().gsb_set_color_dependency(ctx.h, C.c_void_p(ev.cuda_event)))
                img1, dep1, buf1 = forward.render_gaussians(**kw)
                img2, dep2, buf2 = forward.render_gaussians(**kw)   # dependency consumed: the plain path again
            finally:
                ctx.set_option("speculate", 1)
            for img, dep, buf in ((img1, dep1, buf1), (img2, dep2, buf2)):
                assert torch.equal(img0, img) and torch.equal(dep0, dep)
                for k in keys:
                    assert torch.equal(buf0[k], buf[k]), (k, spec)
