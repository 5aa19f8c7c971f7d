"""``render_gaussians`` with the reference's call surface (forward.py:629-894), executed by the
sm_100a kernels of libgsb200 through the C ABI.  Inputs may be numpy arrays or torch tensors on any
device; outputs are torch CUDA tensors with the reference's shapes and the same 12 dict keys."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_KEYS = ("radii", "point_offsets", "points_xy_image", "depths", "colors", "cov3Ds", "conic_opacity", "point_list",
         "ranges", "final_Ts", "n_contrib", "clamped_state")


def render_gaussians(background, means3D, colors=None, opacity=None, scales=None, rotations=None,
                     scale_modifier=1.0, viewmatrix=None, projmatrix=None, tan_fovx=0.5, tan_fovy=0.5,
                     image_height=256, image_width=256, sh=None, degree=3, campos=None, prefiltered=False,
                     antialiasing=False, clamped=True, debug=False):
    """Render 3D Gaussians.  Same arguments as the reference; ``colors``, ``prefiltered`` and
    ``antialiasing`` are accepted and ignored exactly like there (forward.py:210-211,632).

    Returns ``(rendered_image[H,W,3], depth_image[H,W], intermediate_buffers)``; ``depth_image`` is
    the blended INVERSE depth.  Raises ``ValueError`` when more than 2^30 (tile, Gaussian) pairs
    would be rendered (forward.py:765-767)."""
    ctx = _lib.context()
    dev = torch.device("cuda", ctx.device_index)
    means = _lib.to_device(means3D, device=dev, shape=(-1, 3))
    n = means.shape[0]
    shs = _lib.to_device(sh, device=dev).reshape(-1, 3)          # forward.py:687
    if shs.shape[0] != n * 16:
        raise ValueError("sh must hold 16 coefficients per Gaussian (rows are indexed with stride 16, forward.py:310)")
    opac = _lib.to_device(opacity, device=dev).reshape(-1)       # (N,1) is flattened, utils/wp_utils.py:42-43
    scl = _lib.to_device(scales, device=dev, shape=(-1, 3))
    rot = _lib.to_device(rotations, device=dev, shape=(-1, 4))
    H, W = int(image_height), int(image_width)
    frame = _lib.make_frame(viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, W, H, background, degree, clamped,
                            scale_modifier)
    if debug:  # forward.py:712-715
        print(f"\nGSB200 RENDERING: {W}x{H} image, {n} gaussians")
        print(f"Colors: {'from SH' if colors is None else 'provided'}, SH degree: {degree}")
        print(f"Antialiasing: {antialiasing}, Prefiltered: {prefiltered}")

    gx, gy = (W + _lib.TILE - 1) // _lib.TILE, (H + _lib.TILE - 1) // _lib.TILE
    f32, i32 = torch.float32, torch.int32
    e = lambda *shape, dtype=f32: torch.empty(shape, dtype=dtype, device=dev)  # noqa: E731  (kernels write everything)
    out = {
        "radii": e(n, dtype=i32), "point_offsets": e(n, dtype=i32), "points_xy_image": e(n, 2), "depths": e(n),
        "colors": e(n, 3), "cov3Ds": e(n, 6), "conic_opacity": e(n, 4), "clamped_state": e(n, 3),
        "ranges": e(gx * gy, 2, dtype=i32), "final_Ts": e(H, W), "n_contrib": e(H, W, dtype=i32),
    }
    image, depth = e(H, W, 3), e(H, W)
    L = _lib.lib()
    D = C.c_int64(0)
    cap = max(ctx.capacity_hint, 4 * n, 1024)
    for _attempt in range(2):
        point_list, block_masks = e(cap, dtype=i32), e(cap, dtype=i32)
        rc = L.gsb_forward(ctx.h, _lib.stream_ptr(ctx.device_index), C.byref(frame), n, _lib.ptr(means), _lib.ptr(scl),
                           _lib.ptr(rot), _lib.ptr(opac), _lib.ptr(shs), _lib.ptr(out["radii"]),
                           _lib.ptr(out["point_offsets"]), _lib.ptr(out["points_xy_image"]), _lib.ptr(out["depths"]),
                           _lib.ptr(out["colors"]), _lib.ptr(out["cov3Ds"]), _lib.ptr(out["conic_opacity"]),
                           _lib.ptr(out["clamped_state"]), _lib.ptr(point_list), cap, _lib.ptr(out["ranges"]),
                           _lib.ptr(image), _lib.ptr(depth), _lib.ptr(out["final_Ts"]), _lib.ptr(out["n_contrib"]),
                           C.byref(D), _lib.ptr(block_masks))
        if rc == _lib.GSB_ERR_CAPACITY:
            cap = int(D.value) + int(D.value) // 8 + 1024
            continue
        ctx.check(rc)
        break
    ctx.capacity_hint = max(ctx.capacity_hint, int(D.value) + int(D.value) // 8)
    out["point_list"] = point_list[: int(D.value)]
    res = {k: out[k] for k in _KEYS}
    # not a reference key: the tile kernels' per-entry culling masks; backward() reuses them when the
    # dict is passed on as binning_buffer, and recomputes them when the key is absent.  Entries behind the
    # point where a tile's last pixel terminated are never staged by the forward: their masks stay unwritten,
    # and the backward (which replays at most n_contrib entries per pixel) never reads them
    res["block_masks"] = block_masks[: int(D.value)]
    return image, depth, res
