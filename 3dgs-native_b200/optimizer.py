"""The reference's optimizer / densify kernels (optimizer.py:6-415 and the nested kernels of
train.py) as plain functions on torch CUDA tensors.  Each function takes the arguments of the
corresponding ``@wp.kernel`` in the same order, so ``wp.launch(adam_update, dim=n, inputs=[...])``
becomes ``adam_update(*inputs)``.  All updates are in place."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

p = _lib.ptr


def _cs():
    ctx = _lib.context()
    return ctx, _lib.stream_ptr(ctx.device_index)


def adam_update(pos_grads, scale_grads, rot_grads, opacity_grads, sh_grads, num_points, lr_pos, lr_scale, lr_rot,
                lr_opac, lr_sh, beta1, beta2, epsilon, iteration, positions, scales, rotations, opacities, shs,
                m_positions, m_scales, m_rotations, m_opacities, m_shs, v_positions, v_scales, v_rotations,
                v_opacities, v_shs):
    """optimizer.py:6-139 (argument order of the kernel; launched at train.py:750-794)."""
    ctx, s = _cs()
    ctx.check(_lib.lib().gsb_adam_step(
        ctx.h, s, int(num_points), p(pos_grads), p(scale_grads), p(rot_grads), p(opacity_grads), p(sh_grads),
        lr_pos, lr_scale, lr_rot, lr_opac, lr_sh, beta1, beta2, epsilon, int(iteration), p(positions), p(scales),
        p(rotations), p(opacities), p(shs), p(m_positions), p(m_scales), p(m_rotations), p(m_opacities), p(m_shs),
        p(v_positions), p(v_scales), p(v_rotations), p(v_opacities), p(v_shs)))


def reset_opacities(max_opacity, num_points, opacities):
    """optimizer.py:143-158."""
    ctx, s = _cs()
    ctx.check(_lib.lib().gsb_fill_f32(ctx.h, s, p(opacities), int(num_points), float(max_opacity)))


def zero_gradients(pos_grad, scale_grad, rot_grad, opacity_grad, sh_grad, num_points):
    """train.py:94-115."""
    ctx, s = _cs()
    n = int(num_points)
    for t, w in ((pos_grad, 3), (scale_grad, 3), (rot_grad, 4), (opacity_grad, 1), (sh_grad, 48)):
        ctx.check(_lib.lib().gsb_fill_f32(ctx.h, s, p(t), n * w, 0.0))


def init_gaussian_params(positions, scales, rotations, opacities, shs, num_points, init_scale):
    """train.py:36-92."""
    ctx, s = _cs()
    ctx.check(_lib.lib().gsb_init_gaussian_params(ctx.h, s, int(num_points), float(init_scale), p(positions), p(scales),
                                                  p(rotations), p(opacities), p(shs)))


def compute_grad_norms(pos_grad, grad_norms, num_points):
    """train.py:398-405."""
    ctx, s = _cs()
    ctx.check(_lib.lib().gsb_grad_norms(ctx.h, s, int(num_points), p(pos_grad), p(grad_norms)))


def _mark(grads, scales, grad_threshold, scene_extent, percent_dense, num_points, mask, want_split):
    ctx, s = _cs()
    ctx.check(_lib.lib().gsb_mark_candidates(ctx.h, s, int(num_points), int(grads.numel()), p(grads), p(scales),
                                             float(grad_threshold), float(scene_extent), float(percent_dense),
                                             want_split, p(mask)))


def mark_split_candidates(grads, scales, grad_threshold, scene_extent, percent_dense, num_points, split_mask):
    """optimizer.py:180-210.  ``grads`` may be shorter than num_points (the stale array of
    train.py:479-492, quirk G4): missing entries count as 0."""
    _mark(grads, scales, grad_threshold, scene_extent, percent_dense, num_points, split_mask, 1)


def mark_clone_candidates(grads, scales, grad_threshold, scene_extent, percent_dense, num_points, clone_mask):
    """optimizer.py:212-242."""
    _mark(grads, scales, grad_threshold, scene_extent, percent_dense, num_points, clone_mask, 0)


def array_scan(in_array, out_array, inclusive=False) -> int:
    """wp.utils.array_scan(inclusive=False) + ``int(out.numpy()[-1])`` (train.py:431-433, ...).
    Returns the LAST ENTRY of the exclusive scan, i.e. the reference's "total", which does not
    count the last flag (quirk G5)."""
    if inclusive:
        raise NotImplementedError("the reference only uses inclusive=False")
    ctx, s = _cs()
    last = C.c_int32(0)
    ctx.check(_lib.lib().gsb_scan_mask(ctx.h, s, int(in_array.numel()), p(in_array), p(out_array), C.byref(last)))
    return int(last.value)


def split_gaussians(split_mask, prefix_sum, positions, scales, rotations, opacities, shs, N_split, scale_factor,
                    offset, out_positions, out_scales, out_rotations, out_opacities, out_shs):
    """optimizer.py:244-309 (offset == len(positions))."""
    ctx, s = _cs()
    ctx.check(_lib.lib().gsb_split_gaussians(ctx.h, s, int(offset), int(out_positions.shape[0]), p(split_mask),
                                             p(prefix_sum), p(positions), p(scales), p(rotations), p(opacities), p(shs),
                                             int(N_split), float(scale_factor), p(out_positions), p(out_scales),
                                             p(out_rotations), p(out_opacities), p(out_shs)))


def clone_gaussians(clone_mask, prefix_sum, positions, scales, rotations, opacities, shs, noise_scale, offset,
                    out_positions, out_scales, out_rotations, out_opacities, out_shs):
    """optimizer.py:312-362."""
    ctx, s = _cs()
    ctx.check(_lib.lib().gsb_clone_gaussians(ctx.h, s, int(offset), int(out_positions.shape[0]), p(clone_mask),
                                             p(prefix_sum), p(positions), p(scales), p(rotations), p(opacities), p(shs),
                                             float(noise_scale), p(out_positions), p(out_scales), p(out_rotations),
                                             p(out_opacities), p(out_shs)))


def prune_gaussians(opacities, opacity_threshold, num_points, valid_mask):
    """optimizer.py:364-382."""
    ctx, s = _cs()
    ctx.check(_lib.lib().gsb_prune_mask(ctx.h, s, int(num_points), p(opacities), float(opacity_threshold), p(valid_mask)))


def split_valid_mask(split_mask, valid_mask, offset, num_points):
    """mark_split_originals_for_removal + invert_mask (train.py:547-576) in one kernel."""
    ctx, s = _cs()
    ctx.check(_lib.lib().gsb_split_valid_mask(ctx.h, s, int(num_points), int(offset), p(split_mask), p(valid_mask)))


def compact_gaussians(valid_mask, prefix_sum, positions, scales, rotations, opacities, shs, out_positions, out_scales,
                      out_rotations, out_opacities, out_shs):
    """optimizer.py:384-415."""
    ctx, s = _cs()
    ctx.check(_lib.lib().gsb_compact_gaussians(ctx.h, s, int(valid_mask.numel()), int(out_positions.shape[0]),
                                               p(valid_mask), p(prefix_sum), p(positions), p(scales), p(rotations),
                                               p(opacities), p(shs), p(out_positions), p(out_scales), p(out_rotations),
                                               p(out_opacities), p(out_shs)))
