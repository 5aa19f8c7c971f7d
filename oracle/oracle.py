"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE -- never imported by the product).

Mirrors the reference's operator API on numpy arrays so that parity tests read like calls into
the reference:

* ``render_gaussians``  -- forward.py:629-894 (same kwargs, same 12-key dict)
* ``backward``          -- backward.py:955-1196 (same kwargs, same 9-key dict)
* ``adam_update``       -- optimizer.py:6-139 launched as train.py:750-794
* ``densification_and_pruning`` -- train.py:351-713 on a plain dict-of-arrays trainer state
* ``l1_loss`` / ``compute_image_gradients`` -- loss.py:148-176, 217-244

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgs_oracle.so")

TILE = 16
EXP_DET, EXP_LIBM = 0, 1


def build(force: bool = False) -> str:
    """Compile oracle/gs_oracle.c with the committed Makefile (gcc only, a few seconds)."""
    src = os.path.join(_HERE, "gs_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.gso_render.restype = C.c_long
        _lib.gso_render_backward.restype = C.c_long
        _lib.gso_expf_det.restype = C.c_float
        _lib.gso_expf_det.argtypes = [C.c_float]
        _lib.gso_randf.restype = C.c_float
        _lib.gso_randf.argtypes = [C.c_uint32]
        _lib.gso_l1_loss.restype = C.c_double
        _lib.gso_ssim.restype = C.c_double
        _lib.gso_depth_loss.restype = C.c_double
    return _lib


def set_exp_mode(mode: int) -> None:
    lib().gso_set_exp_mode(int(mode))


def set_threads(n: int) -> None:
    lib().gso_set_threads(int(n))


def max_threads() -> int:
    return int(lib().gso_max_threads())


def _f32(a, shape=None):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    if shape is not None:
        a = a.reshape(shape)
    return a


def _i32(a):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _cf(x):
    return C.c_float(float(np.float32(x)))


def expf_det(x: float) -> float:
    return float(lib().gso_expf_det(float(np.float32(x))))


def randf(state: int) -> float:
    return float(lib().gso_randf(int(state) & 0xFFFFFFFF))


# --------------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------------
def render_gaussians(background, means3D, colors=None, opacity=None, scales=None, rotations=None,
                     scale_modifier=1.0, viewmatrix=None, projmatrix=None, tan_fovx=0.5, tan_fovy=0.5,
                     image_height=256, image_width=256, sh=None, degree=3, campos=None, prefiltered=False,
                     antialiasing=False, clamped=True, debug=False, return_extra=False):
    """forward.py:629-894.  ``colors``, ``prefiltered``, ``antialiasing``, ``debug`` are accepted
    and ignored, exactly like the reference.  Returns (image[H,W,3], inv_depth[H,W], dict)."""
    L = lib()
    means = _f32(means3D, (-1, 3))
    N = means.shape[0]
    shs = _f32(sh).reshape(-1, 3)                      # forward.py:687
    assert shs.shape[0] == N * 16, "SH rows are indexed with stride 16 (forward.py:310)"
    opac = _f32(opacity).reshape(-1)                   # utils/wp_utils.py:42-43
    scl = _f32(scales, (-1, 3))
    rot = _f32(rotations, (-1, 4))
    V = _f32(viewmatrix).reshape(16)                   # forward.py:694 (row-major flatten)
    P = _f32(projmatrix).reshape(16)
    cam = _f32([campos[0], campos[1], campos[2]])
    bg = _f32([background[0], background[1], background[2]])
    H, W = int(image_height), int(image_width)
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE

    image = np.zeros((H, W, 3), np.float32)            # forward.py:679-682
    depth_image = np.zeros((H, W), np.float32)
    final_Ts = np.zeros((H, W), np.float32)
    n_contrib = np.zeros((H, W), np.int32)
    radii = np.zeros(N, np.int32)                      # forward.py:703-710
    xy = np.zeros((N, 2), np.float32)
    depths = np.zeros(N, np.float32)
    cov3Ds = np.zeros((N, 6), np.float32)
    rgb = np.zeros((N, 3), np.float32)
    conic_opacity = np.zeros((N, 4), np.float32)
    tiles_touched = np.zeros(N, np.int32)
    clamped_state = np.zeros((N, 3), np.float32)

    L.gso_preprocess(N, _p(means), _p(scl), _cf(scale_modifier), _p(rot), _p(opac), _p(shs), int(degree),
                     int(bool(clamped)), _p(V), _p(P), _p(cam), W, H, _cf(tan_fovx), _cf(tan_fovy),
                     _p(radii), _p(xy), _p(depths), _p(cov3Ds), _p(rgb), _p(conic_opacity), _p(tiles_touched),
                     _p(clamped_state))
    point_offsets = np.zeros(N, np.int32)              # forward.py:755-763
    L.gso_prefix_sum(N, _p(tiles_touched), _p(point_offsets))
    num_rendered = int(point_offsets[-1]) if N > 0 else 0
    if num_rendered > (1 << 30):                       # forward.py:765-767
        raise ValueError("Number of rendered points exceeds the maximum supported by Warp.")

    keys_unsorted = np.zeros(num_rendered, np.int64)
    vals_unsorted = np.zeros(num_rendered, np.int32)
    L.gso_duplicate_with_keys(N, _p(xy), _p(depths), _p(point_offsets), _p(radii), W, H, _p(keys_unsorted),
                              _p(vals_unsorted))
    keys = keys_unsorted.copy()
    point_list = vals_unsorted.copy()
    tk = np.empty_like(keys)
    tv = np.empty_like(point_list)
    L.gso_sort_pairs(_p(keys), _p(point_list), num_rendered, _p(tk), _p(tv))

    ranges = np.zeros((gx * gy, 2), np.int32)          # forward.py:827-828
    pairs = 0
    if num_rendered > 0:                               # forward.py:830 (D == 0 -> image stays ZERO)
        L.gso_identify_tile_ranges(num_rendered, _p(keys), _p(ranges))
        pairs = L.gso_render(W, H, _p(ranges), _p(point_list), _p(xy), _p(rgb), _p(conic_opacity), _p(depths),
                             _p(bg), _p(image), _p(depth_image), _p(final_Ts), _p(n_contrib))
        # forward.py:867-879 track_pixel_stats is provably a no-op; tests assert it.
    out = {
        "radii": radii, "point_offsets": point_offsets, "points_xy_image": xy, "depths": depths,
        "colors": rgb, "cov3Ds": cov3Ds, "conic_opacity": conic_opacity, "point_list": point_list,
        "ranges": ranges, "final_Ts": final_Ts, "n_contrib": n_contrib, "clamped_state": clamped_state,
    }
    if return_extra:  # not part of the reference dict: exposed for parity tests on binning
        out = dict(out)
        out["_tiles_touched"] = tiles_touched
        out["_keys_unsorted"] = keys_unsorted
        out["_vals_unsorted"] = vals_unsorted
        out["_keys_sorted"] = keys
        out["_num_rendered"] = num_rendered
        out["_pairs_fwd"] = int(pairs)
    return image, depth_image, out


def track_pixel_stats(image, background, final_Ts, n_contrib) -> int:
    H, W = final_Ts.shape
    bg = _f32([background[0], background[1], background[2]])
    return int(lib().gso_track_pixel_stats(W, H, _p(_f32(image)), _p(bg), _p(final_Ts), _p(n_contrib)))


# --------------------------------------------------------------------------------------------
# backward
# --------------------------------------------------------------------------------------------
def backward(background, means3D, dL_dpixels, opacity=None, shs=None, scales=None, rotations=None,
             scale_modifier=1.0, viewmatrix=None, projmatrix=None, tan_fovx=0.5, tan_fovy=0.5,
             image_height=256, image_width=256, campos=None, radii=None, means2D=None, conic_opacity=None,
             rgb=None, clamped=None, cov3Ds=None, geom_buffer=None, binning_buffer=None, img_buffer=None,
             degree=3, debug=False, return_extra=False):
    """backward.py:955-1196.  ``scale_modifier`` is accepted and -- like the reference -- NOT
    forwarded to the cov3D backward (quirk G3)."""
    L = lib()
    means = _f32(means3D, (-1, 3))
    N = means.shape[0]
    H, W = int(image_height), int(image_width)
    focal_y = H / (2.0 * tan_fovy)                      # backward.py:1044-1045 (python doubles)
    focal_x = W / (2.0 * tan_fovx)
    ranges = _i32(img_buffer["ranges"])
    final_Ts = _f32(img_buffer["final_Ts"])
    n_contrib = _i32(img_buffer["n_contrib"])
    point_list = _i32(binning_buffer["point_list"])
    if geom_buffer is not None:                         # backward.py:1092-1103
        if radii is None:
            radii = geom_buffer.get("radii")
        if means2D is None:
            means2D = geom_buffer.get("means2D")
        if conic_opacity is None:
            conic_opacity = geom_buffer.get("conic_opacity")
        if rgb is None:
            rgb = geom_buffer.get("rgb")
        if clamped is None:
            clamped = geom_buffer.get("clamped_state")
    radii = _i32(radii)
    xy = _f32(means2D, (-1, 2))
    con_o = _f32(conic_opacity, (-1, 4))
    colors = _f32(rgb, (-1, 3))
    clamped_state = _f32(clamped, (-1, 3))
    cov3 = _f32(cov3Ds, (-1, 6))
    sh = _f32(shs).reshape(-1, 3)
    scl = _f32(scales, (-1, 3))
    rot = _f32(rotations, (-1, 4))
    V = _f32(viewmatrix).reshape(16)
    P = _f32(projmatrix).reshape(16)
    cam = _f32([campos[0], campos[1], campos[2]])
    bg = _f32([background[0], background[1], background[2]])
    dpix = _f32(dL_dpixels, (H, W, 3))

    dL_dmean2D = np.zeros((N, 3), np.float32)           # backward.py:1113-1126
    dL_dconic = np.zeros((N, 4), np.float32)
    dL_dopacity = np.zeros(N, np.float32)
    dL_dcolor = np.zeros((N, 3), np.float32)
    dL_dmean3D = np.zeros((N, 3), np.float32)
    dL_dcov3D = np.zeros((N, 6), np.float32)            # returned untouched (always zeros, T5)
    dL_dsh = np.zeros((N * 16, 3), np.float32)          # always stride 16 (see B4 note)
    dL_dscale = np.zeros((N, 3), np.float32)
    dL_drot = np.zeros((N, 4), np.float32)
    scratch = np.zeros((N, 6), np.float32)              # backward.py:812

    pairs = L.gso_render_backward(W, H, _p(ranges), _p(point_list), _p(bg), _p(xy), _p(con_o), _p(colors),
                                  _p(final_Ts), _p(n_contrib), _p(dpix), _p(dL_dmean2D), _p(dL_dconic),
                                  _p(dL_dopacity), _p(dL_dcolor))
    L.gso_backward_preprocess(N, _p(means), _p(radii), _p(sh), _p(scl), _p(rot), _p(V), _p(P), _cf(tan_fovx),
                              _cf(tan_fovy), _cf(focal_x), _cf(focal_y), _p(cov3), _p(cam), _p(clamped_state),
                              _p(dL_dmean2D), _p(dL_dconic), _p(dL_dcolor), int(degree), _p(dL_dmean3D),
                              _p(dL_dsh), _p(dL_dscale), _p(dL_drot), _p(scratch))
    out = {
        "dL_dmean3D": dL_dmean3D, "dL_dcolor": dL_dcolor, "dL_dshs": dL_dsh, "dL_dopacity": dL_dopacity,
        "dL_dscale": dL_dscale, "dL_drot": dL_drot, "dL_dmean2D": dL_dmean2D, "dL_dconic": dL_dconic,
        "dL_dcov3D": dL_dcov3D,
    }
    if return_extra:
        out["_dL_dcov3D_internal"] = scratch
        out["_pairs_bwd"] = int(pairs)
    return out


# --------------------------------------------------------------------------------------------
# optimizer / densify
# --------------------------------------------------------------------------------------------
PARAM_KEYS = ("positions", "scales", "rotations", "opacities", "shs")
PARAM_WIDTH = {"positions": 3, "scales": 3, "rotations": 4, "opacities": 1, "shs": 48}


def zeros_like_params(n):
    """train.py:216-231 create_gradient_arrays."""
    return {
        "positions": np.zeros((n, 3), np.float32), "scales": np.zeros((n, 3), np.float32),
        "rotations": np.zeros((n, 4), np.float32), "opacities": np.zeros(n, np.float32),
        "shs": np.zeros((n * 16, 3), np.float32),
    }


def init_gaussian_params(n, init_scale=0.1):
    """train.py:36-92 / 193-214."""
    p = zeros_like_params(n)
    lib().gso_init_gaussian_params(n, _cf(init_scale), _p(p["positions"]), _p(p["scales"]), _p(p["rotations"]),
                                   _p(p["opacities"]), _p(p["shs"]))
    return p


def adam_update(grads, params, adam_m, adam_v, num_points, lr_pos, lr_scale, lr_rot, lr_opac, lr_sh, beta1,
                beta2, epsilon, iteration):
    """optimizer.py:6-139 with the argument order of train.py:750-794; in place."""
    g, p, m, v = grads, params, adam_m, adam_v
    lib().gso_adam_update(int(num_points), _p(g["positions"]), _p(g["scales"]), _p(g["rotations"]),
                          _p(g["opacities"]), _p(g["shs"]), _cf(lr_pos), _cf(lr_scale), _cf(lr_rot), _cf(lr_opac),
                          _cf(lr_sh), _cf(beta1), _cf(beta2), _cf(epsilon), int(iteration), _p(p["positions"]),
                          _p(p["scales"]), _p(p["rotations"]), _p(p["opacities"]), _p(p["shs"]),
                          _p(m["positions"]), _p(m["scales"]), _p(m["rotations"]), _p(m["opacities"]), _p(m["shs"]),
                          _p(v["positions"]), _p(v["scales"]), _p(v["rotations"]), _p(v["opacities"]), _p(v["shs"]))


def _scan_excl(mask):
    out = np.zeros_like(mask)
    lib().gso_exclusive_scan(mask.shape[0], _p(mask), _p(out))
    return out


def _alloc(n):
    return zeros_like_params(n)


def _args(p):
    return [_p(p[k]) for k in PARAM_KEYS]


def densification_and_pruning(state, iteration, config=None):
    """train.py:351-713 on ``state`` = {params, grads, adam_m, adam_v, num_points, scene_extent}.
    Mutates ``state`` in place and returns a log dict of what happened (counts)."""
    cfg = dict(densify_from_iter=500, densify_until_iter=15000, densification_interval=100,
               opacity_reset_interval=3000, densify_grad_threshold=0.0002, percent_dense=0.01,
               cull_opacity_threshold=0.005, min_valid_points=1000, max_valid_points=1000000,
               max_allowed_prune_ratio=1.0, background_color=[0.0, 0.0, 0.0])
    cfg.update(config or {})
    L = lib()
    log = {"cloned": 0, "split": 0, "split_removed": 0, "pruned": 0, "opacity_reset": False}
    if (iteration > cfg["densify_from_iter"] and iteration < cfg["densify_until_iter"]
            and iteration % cfg["densification_interval"] == 0):
        n = state["num_points"]
        pos_grads = state["grads"]["positions"]
        n_grads = n
        avg_grads = np.zeros(n, np.float32)
        L.gso_grad_norms(n, _p(pos_grads), _p(avg_grads))                       # train.py:398-408
        gt, pd, ext = _cf(cfg["densify_grad_threshold"]), _cf(cfg["percent_dense"]), _cf(state["scene_extent"])

        clone_mask = np.zeros(n, np.int32)
        L.gso_mark_candidates(n, n_grads, _p(avg_grads), _p(state["params"]["scales"]), gt, ext, pd, 0,
                              _p(clone_mask))
        clone_prefix = _scan_excl(clone_mask)
        total_to_clone = int(clone_prefix[-1]) if n > 0 else 0                    # quirk G5
        if total_to_clone > 0:
            new_n = n + total_to_clone
            out = _alloc(new_n)
            L.gso_clone_gaussians(n, new_n, _p(clone_mask), _p(clone_prefix), *_args(state["params"]),
                                  _cf(0.01), *_args(out))
            state["params"], state["num_points"] = out, new_n
            state["grads"], state["adam_m"], state["adam_v"] = _alloc(new_n), _alloc(new_n), _alloc(new_n)
            log["cloned"] = total_to_clone

        n = state["num_points"]
        split_mask = np.zeros(n, np.int32)
        L.gso_mark_candidates(n, n_grads, _p(avg_grads), _p(state["params"]["scales"]), gt, ext, pd, 1,
                              _p(split_mask))                                      # quirk G4 (stale grads)
        split_prefix = _scan_excl(split_mask)
        total_to_split = int(split_prefix[-1]) if n > 0 else 0
        if total_to_split > 0:
            N0, n_split = n, 2
            new_n = N0 + total_to_split * n_split
            out = _alloc(new_n)
            L.gso_split_gaussians(N0, new_n, _p(split_mask), _p(split_prefix), *_args(state["params"]), n_split,
                                  _cf(0.8), *_args(out))
            state["params"], state["num_points"] = out, new_n
            state["grads"], state["adam_m"], state["adam_v"] = _alloc(new_n), _alloc(new_n), _alloc(new_n)
            log["split"] = total_to_split
            valid = np.zeros(new_n, np.int32)
            L.gso_split_valid_mask(new_n, N0, _p(split_mask), _p(valid))
            prefix = _scan_excl(valid)
            valid_count = int(prefix[-1])
            if valid_count < new_n:
                out = _alloc(valid_count)
                L.gso_compact_gaussians(new_n, valid_count, _p(valid), _p(prefix), *_args(state["params"]),
                                        *_args(out))
                log["split_removed"] = new_n - valid_count
                state["params"], state["num_points"] = out, valid_count
                state["grads"], state["adam_m"], state["adam_v"] = (_alloc(valid_count), _alloc(valid_count),
                                                                    _alloc(valid_count))

        n = state["num_points"]
        valid = np.zeros(n, np.int32)
        L.gso_prune_mask(n, _p(state["params"]["opacities"]), _cf(cfg["cull_opacity_threshold"]), _p(valid))
        prefix = _scan_excl(valid)
        valid_count = int(prefix[-1]) if n > 0 else 0
        prune_count = n - valid_count
        prune_ratio = prune_count / n if n > 0 else 0
        if (valid_count >= cfg["min_valid_points"] and valid_count <= cfg["max_valid_points"]
                and prune_ratio <= cfg["max_allowed_prune_ratio"] and valid_count < n):
            out = _alloc(valid_count)
            L.gso_compact_gaussians(n, valid_count, _p(valid), _p(prefix), *_args(state["params"]), *_args(out))
            state["params"], state["num_points"] = out, valid_count
            state["grads"], state["adam_m"], state["adam_v"] = (_alloc(valid_count), _alloc(valid_count),
                                                                _alloc(valid_count))
            log["pruned"] = prune_count

    background_is_white = all(c == 1.0 for c in cfg["background_color"])
    if (iteration % cfg["opacity_reset_interval"] == 0
            or (background_is_white and iteration == cfg["densify_from_iter"])):  # quirk G6: it 0 too
        state["params"]["opacities"][: state["num_points"]] = np.float32(0.01)
        log["opacity_reset"] = True
    return log


# --------------------------------------------------------------------------------------------
# loss (next row 8f-1) and LR schedule
# --------------------------------------------------------------------------------------------
def l1_loss(rendered, target) -> float:
    r, t = _f32(rendered), _f32(target)
    H, W = r.shape[0], r.shape[1]
    return float(lib().gso_l1_loss(W, H, _p(r), _p(t)))


def compute_image_gradients(rendered, target, lambda_dssim=0.2):
    r, t = _f32(rendered), _f32(target)
    H, W = r.shape[0], r.shape[1]
    g = np.zeros((H, W, 3), np.float32)
    l1_weight = (1.0 - lambda_dssim) / (H * W * 3.0)                 # loss.py:236 (python double)
    lib().gso_l1_grad(W, H, _p(r), _p(t), _cf(l1_weight), _p(g))
    return g


def ssim(rendered, target) -> float:
    """loss.py:178-215."""
    r, t = _f32(rendered), _f32(target)
    H, W = r.shape[0], r.shape[1]
    return float(lib().gso_ssim(W, H, _p(r), _p(t)))


def depth_loss(rendered_depth, target_depth, depth_mask) -> float:
    """loss.py:270-306."""
    r, t, m = _f32(rendered_depth), _f32(target_depth), _f32(depth_mask)
    H, W = r.shape[0], r.shape[1]
    return float(lib().gso_depth_loss(W, H, _p(r), _p(t), _p(m)))


def get_lr(initial_lr, final_lr_factor, iteration, total_iterations):
    """scheduler.py:15-28."""
    if total_iterations <= 1:
        return initial_lr
    progress = min(iteration / (total_iterations - 1), 1.0)
    final_lr = initial_lr * final_lr_factor
    return initial_lr * ((final_lr / initial_lr) ** progress)
