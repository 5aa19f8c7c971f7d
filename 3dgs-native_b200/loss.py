"""L1 loss and its pixel gradient (reference loss.py:11-30, 121-146, 148-176, 217-244) as ONE fused
pass on the GPU: the reference launches W*H threads that all atomic-add into one address, reads
the sum back twice, then runs a second kernel for the sign gradient."""
from __future__ import annotations

import torch

from . import _lib


def l1_loss_and_gradients(rendered, target, lambda_dssim=0.0, out_grad=None, out_sum=None):
    """Returns (loss_sum_device[1] float64, pixel_grad[H,W,3]).  loss = sum / (3*H*W).

    ``out_grad`` / ``out_sum`` (not reference arguments): preallocated outputs.  ``out_sum`` may also be a one-element
    float64 tensor in PINNED HOST memory: the kernel then writes the sum straight there (8 bytes over PCIe; unified
    addressing), and a host that waits for an event recorded behind this call reads it without any copy operation in
    the stream -- a device-to-host cudaMemcpyAsync of 8 bytes costs the stream about 10 us."""
    ctx = _lib.context()
    dev = torch.device("cuda", ctx.device_index)
    r = _lib.to_device(rendered, device=dev)
    t = _lib.to_device(target, device=dev)
    H, W = r.shape[0], r.shape[1]
    grad = out_grad if out_grad is not None else torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    s = out_sum if out_sum is not None else torch.empty(1, dtype=torch.float64, device=dev)
    l1_weight = (1.0 - lambda_dssim) / (H * W * 3.0)                   # loss.py:236 (python double)
    ctx.check(_lib.lib().gsb_l1_loss_grad(ctx.h, _lib.stream_ptr(ctx.device_index), H * W * 3, _lib.ptr(r), _lib.ptr(t),
                                          l1_weight, _lib.ptr(grad), _lib.ptr(s)))
    return s, grad


def l1_loss(rendered, target) -> float:
    """loss.py:148-176 (synchronises: returns a Python float like the reference)."""
    s, _ = l1_loss_and_gradients(rendered, target)
    H, W = rendered.shape[0], rendered.shape[1]
    return float(s.item()) / (W * H * 3)


def compute_image_gradients(rendered, target, lambda_dssim=0.2):
    """loss.py:217-244: d(loss)/d(pixel) = (1 - lambda_dssim)/(3HW) * sign(rendered - target), with
    Warp's sign(0) = +1.  The SSIM term is a TODO in the reference and contributes nothing."""
    return l1_loss_and_gradients(rendered, target, lambda_dssim)[1]


def ssim(rendered, target) -> float:
    """loss.py:178-215: mean over the pixels of the 11x11-window SSIM (mean of the three channels).
    Synchronises: returns a Python float like the reference.  (The reference has no SSIM gradient:
    ``compute_image_gradients`` ignores the term, loss.py:243.)"""
    ctx = _lib.context()
    dev = torch.device("cuda", ctx.device_index)
    r = _lib.to_device(rendered, device=dev)
    t = _lib.to_device(target, device=dev)
    H, W = r.shape[0], r.shape[1]
    s = torch.empty(1, dtype=torch.float64, device=dev)
    ctx.check(_lib.lib().gsb_ssim(ctx.h, _lib.stream_ptr(ctx.device_index), W, H, _lib.ptr(r), _lib.ptr(t), _lib.ptr(s)))
    return float(s.item()) / (W * H)


def depth_loss(rendered_depth, target_depth, depth_mask) -> float:
    """loss.py:270-306: mean of |rendered - target| * mask over the pixels (inverse depths)."""
    ctx = _lib.context()
    dev = torch.device("cuda", ctx.device_index)
    r = _lib.to_device(rendered_depth, device=dev)
    t = _lib.to_device(target_depth, device=dev)
    m = _lib.to_device(depth_mask, device=dev)
    H, W = r.shape[0], r.shape[1]
    s = torch.empty(1, dtype=torch.float64, device=dev)
    ctx.check(_lib.lib().gsb_depth_loss(ctx.h, _lib.stream_ptr(ctx.device_index), H * W, _lib.ptr(r), _lib.ptr(t),
                                        _lib.ptr(m), _lib.ptr(s)))
    return float(s.item()) / (W * H)
