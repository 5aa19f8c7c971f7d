"""``backward`` with the reference's call surface (backward.py:955-1196) on the sm_100a kernels."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def backward(background, means3D, dL_dpixels, opacity=None, shs=None, scales=None, rotations=None,
             scale_modifier=1.0, viewmatrix=None, projmatrix=None, tan_fovx=0.5, tan_fovy=0.5,
             image_height=256, image_width=256, campos=None, radii=None, means2D=None, conic_opacity=None,
             rgb=None, clamped=None, cov3Ds=None, geom_buffer=None, binning_buffer=None, img_buffer=None,
             degree=3, debug=False, out=None, sh_compact=False):
    """Gradients of all Gaussian parameters.  Same arguments and the same 9-key result dict as the
    reference.  Like there, ``img_buffer`` (ranges, final_Ts, n_contrib) and ``binning_buffer``
    (point_list) are required, ``scale_modifier`` is accepted but not used by the covariance
    backward (quirk G3), and ``dL_dcov3D`` is returned as zeros.  ``dL_dshs`` always has 16 rows per
    Gaussian (the reference under-allocates it for degree < 3 and then writes out of bounds).

    ``out`` (not a reference argument; default None = allocate like the reference does): a dict with
    any of the nine result keys mapping to preallocated contiguous float32 CUDA tensors of the right
    size, which then receive those results -- e.g. views of a flat gradient buffer that a multi-GPU
    exchange reads in place.
    ``sh_compact`` (not a reference argument; default False): ``dL_dshs`` then holds the two rank-1
    factors of the SH gradient in its first 8 n floats -- per Gaussian (dL_dRGB after the clamp mask,
    unit view direction, 0, 0) -- instead of the 48 n products (gsb_backward_compact_sh); the rest of
    the tensor is left untouched.  For ``Trainer.exchange_and_step(compact=True)``."""
    if img_buffer is None or binning_buffer is None:
        raise ValueError("backward() needs img_buffer{ranges,final_Ts,n_contrib} and binning_buffer{point_list} "
                         "(backward.py:1084-1090)")
    ctx = _lib.context()
    dev = torch.device("cuda", ctx.device_index)
    f32, i32 = torch.float32, torch.int32
    td = lambda x, dtype=f32, shape=None: _lib.to_device(x, dtype=dtype, device=dev, shape=shape)  # noqa: E731
    means = td(means3D, shape=(-1, 3))
    n = means.shape[0]
    H, W = int(image_height), int(image_width)
    ranges = td(img_buffer.get("ranges"), i32)
    final_Ts = td(img_buffer.get("final_Ts"))
    n_contrib = td(img_buffer.get("n_contrib"), i32)
    point_list = td(binning_buffer.get("point_list"), i32)
    block_masks = binning_buffer.get("block_masks")   # optional extra of our forward (same point_list)
    block_masks = td(block_masks, i32) if block_masks is not None else None
    if block_masks is not None and block_masks.numel() != point_list.numel():
        raise ValueError("binning_buffer['block_masks'] must have one entry per point_list entry")
    if geom_buffer is not None:  # backward.py:1092-1103
        radii = geom_buffer.get("radii") if radii is None else radii
        means2D = geom_buffer.get("means2D") if means2D is None else means2D
        conic_opacity = geom_buffer.get("conic_opacity") if conic_opacity is None else conic_opacity
        rgb = geom_buffer.get("rgb") if rgb is None else rgb
        clamped = geom_buffer.get("clamped_state") if clamped is None else clamped
    radii_t = td(radii, i32)
    xy = td(means2D, shape=(-1, 2))
    con_o = td(conic_opacity, shape=(-1, 4))
    colors = td(rgb, shape=(-1, 3))
    clamped_state = td(clamped, shape=(-1, 3))
    cov3 = td(cov3Ds, shape=(-1, 6))
    sh = td(shs).reshape(-1, 3)
    if sh.shape[0] != n * 16:
        raise ValueError("shs must hold 16 coefficients per Gaussian (stride-16 indexing, backward.py:101)")
    scl = td(scales, shape=(-1, 3))
    rot = td(rotations, shape=(-1, 4))
    opac = td(opacity).reshape(-1) if opacity is not None else None
    dpix = td(dL_dpixels, shape=(H, W, 3))
    frame = _lib.make_frame(viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, W, H, background, degree, True,
                            scale_modifier)
    shapes = {"dL_dmean3D": (n, 3), "dL_dcolor": (n, 3), "dL_dshs": (n * 16, 3), "dL_dopacity": (n,),
              "dL_dscale": (n, 3), "dL_drot": (n, 4), "dL_dmean2D": (n, 3), "dL_dconic": (n, 4), "dL_dcov3D": (n, 6)}
    g = {}
    for key, shape in shapes.items():
        t = out.get(key) if out else None
        if t is None:
            t = torch.empty(shape, dtype=f32, device=dev)
        else:
            numel = 1
            for d in shape:
                numel *= d
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == f32 and t.is_contiguous()
                    and t.numel() == numel):
                raise ValueError(f"out[{key!r}] must be a contiguous float32 CUDA tensor with {numel} elements")
        g[key] = t
    p = _lib.ptr
    fn = _lib.lib().gsb_backward_compact_sh if sh_compact else _lib.lib().gsb_backward
    rc = fn(ctx.h, _lib.stream_ptr(ctx.device_index), C.byref(frame), n, p(means), p(opac), p(sh),
                                 p(scl), p(rot), p(radii_t), p(xy), p(con_o), p(colors), p(clamped_state), p(cov3),
                                 p(point_list), p(ranges), p(final_Ts), p(n_contrib), p(dpix), p(g["dL_dmean3D"]),
                                 p(g["dL_dcolor"]), p(g["dL_dshs"]), p(g["dL_dopacity"]), p(g["dL_dscale"]),
                                 p(g["dL_drot"]), p(g["dL_dmean2D"]), p(g["dL_dconic"]), p(g["dL_dcov3D"]),
                                 p(block_masks))
    ctx.check(rc)
    return g
