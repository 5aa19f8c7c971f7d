"""ctypes binding of csrc/libgsb200.so (the C ABI declared in include/gsb200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is present every
operator raises.  PyTorch is used only for device memory, streams and H2D/D2H copies."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# GSB200_LIB: an A/B build of the library (csrc/Makefile `variant`), for kernel experiments only
LIB_PATH = os.environ.get("GSB200_LIB") or os.path.join(CSRC, "libgsb200.so")

GSB_OK, GSB_ERR_INVALID, GSB_ERR_CUDA, GSB_ERR_TOO_MANY, GSB_ERR_CAPACITY, GSB_ERR_NOMEM = 0, -1, -2, -3, -4, -5
TILE = 16

vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class Frame(C.Structure):
    """struct gsb_frame (include/gsb200.h)."""
    _fields_ = [("view", f32 * 16), ("proj", f32 * 16), ("campos", f32 * 3), ("tan_fovx", f32), ("tan_fovy", f32),
                ("scale_modifier", f32), ("background", f32 * 3), ("width", i32), ("height", i32), ("degree", i32),
                ("clamped", i32)]


_SIGS = {
    "gsb_version": (C.c_int, []),
    "gsb_create": (C.c_int, [C.POINTER(vp), C.c_int]),
    "gsb_destroy": (C.c_int, [vp]),
    "gsb_last_error_string": (C.c_char_p, [vp]),
    "gsb_reserve": (C.c_int, [vp, vp, i64]),
    "gsb_launch_count": (i64, [vp]),
    "gsb_set_option": (C.c_int, [vp, C.c_char_p, C.c_int]),
    "gsb_preprocess": (C.c_int, [vp, vp, C.POINTER(Frame), i32] + [vp] * 13),
    "gsb_scan_tiles": (C.c_int, [vp, vp, i32, vp, vp, C.POINTER(i64)]),
    "gsb_duplicate_with_keys": (C.c_int, [vp, vp, i32, i32, i32, vp, vp, vp, vp, i64, vp, vp]),
    "gsb_sort_pairs64": (C.c_int, [vp, vp, vp, vp, vp, vp, i64, C.c_int, C.c_int]),
    "gsb_tile_ranges": (C.c_int, [vp, vp, i64, vp, i32, vp]),
    "gsb_bin_by_tile": (C.c_int, [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, i64, vp, C.POINTER(i64), C.POINTER(i32)]),
    "gsb_blend_forward": (C.c_int, [vp, vp, C.POINTER(Frame)] + [vp] * 11),
    "gsb_forward": (C.c_int, [vp, vp, C.POINTER(Frame), i32] + [vp] * 14 + [i64] + [vp] * 5 + [C.POINTER(i64), vp]),
    "gsb_blend_backward": (C.c_int, [vp, vp, C.POINTER(Frame), i32] + [vp] * 13),
    "gsb_preprocess_backward": (C.c_int, [vp, vp, C.POINTER(Frame), i32] + [vp] * 15),
    "gsb_preprocess_backward_compact_sh": (C.c_int, [vp, vp, C.POINTER(Frame), i32] + [vp] * 15),
    "gsb_backward": (C.c_int, [vp, vp, C.POINTER(Frame), i32] + [vp] * 26),
    "gsb_backward_compact_sh": (C.c_int, [vp, vp, C.POINTER(Frame), i32] + [vp] * 26),
    "gsb_adam_step": (C.c_int, [vp, vp, i32] + [vp] * 5 + [f32] * 8 + [i32] + [vp] * 15),
    "gsb_adam_step_phase": (C.c_int, [vp, vp, i32] + [vp] * 5 + [f32] * 8 + [i32] + [vp] * 15 + [i32]),
    "gsb_flat_layout": (C.c_int, [i32, C.POINTER(i64), C.POINTER(i64)]),
    "gsb_adam_step_peers": (C.c_int, [vp, vp, i32, i32, i32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_uint64,
                                      C.c_uint64, vp, vp] + [f32] * 8 + [i32, i32]),
    "gsb_adam_step_peers_compact": (C.c_int, [vp, vp, i32, i32, i32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                              C.c_uint64, vp, vp] + [f32] * 8 + [i32, vp, i64, i32, i32]),
    "gsb_adam_step_peers_phase": (C.c_int, [vp, vp, i32, i32, i32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                            C.c_uint64, C.c_uint64, vp, vp] + [f32] * 8 + [i32, vp, i64, i32, i32, i32]),
    "gsb_set_color_dependency": (C.c_int, [vp, vp]),
    "gsb_fill_f32": (C.c_int, [vp, vp, vp, i64, f32]),
    "gsb_selftest_block_mask": (C.c_int, [vp, vp, i32, vp, vp, vp]),
    "gsb_selftest_work_counters": (C.c_int, [vp, vp, C.POINTER(Frame)] + [vp] * 6),
    "gsb_selftest_div": (C.c_int, [vp, vp, i64, vp, vp, vp, vp, vp]),
    "gsb_accumulate_f32": (C.c_int, [vp, vp, vp, vp, i64]),
    "gsb_init_gaussian_params": (C.c_int, [vp, vp, i32, f32] + [vp] * 5),
    "gsb_grad_norms": (C.c_int, [vp, vp, i32, vp, vp]),
    "gsb_mark_candidates": (C.c_int, [vp, vp, i32, i32, vp, vp, f32, f32, f32, i32, vp]),
    "gsb_scan_mask": (C.c_int, [vp, vp, i32, vp, vp, C.POINTER(i32)]),
    "gsb_clone_gaussians": (C.c_int, [vp, vp, i32, i32, vp, vp] + [vp] * 5 + [f32] + [vp] * 5),
    "gsb_split_gaussians": (C.c_int, [vp, vp, i32, i32, vp, vp] + [vp] * 5 + [i32, f32] + [vp] * 5),
    "gsb_split_valid_mask": (C.c_int, [vp, vp, i32, i32, vp, vp]),
    "gsb_prune_mask": (C.c_int, [vp, vp, i32, vp, f32, vp]),
    "gsb_compact_gaussians": (C.c_int, [vp, vp, i32, i32, vp, vp] + [vp] * 10),
    "gsb_l1_loss_grad": (C.c_int, [vp, vp, i64, vp, vp, f32, vp, vp]),
    "gsb_ssim": (C.c_int, [vp, vp, i32, i32, vp, vp, vp]),
    "gsb_depth_loss": (C.c_int, [vp, vp, i64, vp, vp, vp, vp]),
}

_lib = None
_lock = threading.Lock()


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a with the committed Makefile (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", CSRC, "-j8"], stdout=out)
    return LIB_PATH


def lib():
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                   "(there is no CPU fallback)")
            L = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGS.items():
                fn = getattr(L, name)
                fn.restype = res
                fn.argtypes = args
            _lib = L
    return _lib


class Context:
    """One gsb_ctx per (device); thread-compatible."""

    def __init__(self, device_index: int):
        self.device_index = device_index
        h = vp()
        rc = lib().gsb_create(C.byref(h), device_index)
        if rc != GSB_OK:
            raise RuntimeError(f"gsb_create failed with {rc}: a CUDA device is required (no CPU fallback)")
        self.h = h
        self.capacity_hint = 0

    def check(self, rc: int):
        if rc == GSB_OK:
            return
        msg = (lib().gsb_last_error_string(self.h) or b"").decode()
        if rc == GSB_ERR_TOO_MANY:   # forward.py:765-767
            raise ValueError(msg or "Number of rendered points exceeds the maximum supported by Warp.")
        raise RuntimeError(f"libgsb200 error {rc}: {msg}")

    @property
    def launches(self) -> int:
        return int(lib().gsb_launch_count(self.h))

    def set_option(self, name: str, value: int):
        self.check(lib().gsb_set_option(self.h, name.encode(), int(value)))


_contexts: dict[int, Context] = {}


def context(device=None) -> Context:
    if not torch.cuda.is_available():
        raise RuntimeError("gsb200 needs a CUDA device (sm_100a); there is no CPU fallback")
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    if idx not in _contexts:
        with torch.cuda.device(idx):
            _contexts[idx] = Context(idx)
    return _contexts[idx]


def stream_ptr(device_index=None):
    return vp(torch.cuda.current_stream(device_index).cuda_stream)


def ptr(t):
    return vp(t.data_ptr()) if t is not None else vp(0)


def to_device(x, dtype=torch.float32, device=None, shape=None):
    """utils/wp_utils.py:34-44 ``to_warp_array``: accepts numpy arrays, torch tensors (any device) or
    anything array-like and returns a contiguous CUDA tensor of ``dtype``.  CUDA tensors pass through
    untouched when already contiguous and of the right dtype (like ``wp.array`` inputs do)."""
    if x is None:
        return None
    if isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == dtype and x.is_contiguous() and not x.requires_grad \
            and (device is None or x.device == device):
        return x if shape is None else x.view(shape)       # the common case in a training loop: nothing to do
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    if isinstance(x, torch.Tensor):
        t = x.detach()
        if t.device != device or t.dtype != dtype:
            t = t.to(device=device, dtype=dtype, non_blocking=True)
    else:
        a = np.asarray(x)
        t = torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dtype, non_blocking=True)
    if shape is not None:
        t = t.reshape(shape)
    return t.contiguous()


def host_floats(x, n):
    """A small host vector (background, campos) or matrix from numpy / torch / list."""
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    a = np.asarray(x, dtype=np.float32).reshape(-1)
    if a.size < n:
        raise ValueError(f"expected at least {n} values, got {a.size}")
    return a[:n]


_frame_cache: dict = {}


def make_frame(viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, width, height, background, degree=3, clamped=True,
               scale_modifier=1.0) -> Frame:
    """The per-view constants as a ``struct gsb_frame``.  A training loop passes the same few cameras again and
    again: the struct is cached by the VALUE of its inputs (the bytes of the four small arrays and the scalars), so a
    repeated camera costs four ``tobytes`` and one dict lookup instead of 41 Python-float conversions."""
    vm, pm = host_floats(viewmatrix, 16), host_floats(projmatrix, 16)
    cp, bg = host_floats(campos, 3), host_floats(background, 3)
    key = (vm.tobytes(), pm.tobytes(), cp.tobytes(), bg.tobytes(), float(tan_fovx), float(tan_fovy), int(width),
           int(height), int(degree), bool(clamped), float(scale_modifier))
    f = _frame_cache.get(key)
    if f is not None:
        return f
    if len(_frame_cache) > 4096:
        _frame_cache.clear()
    f = _frame_cache[key] = _build_frame(vm, pm, cp, tan_fovx, tan_fovy, width, height, bg, degree, clamped,
                                         scale_modifier)
    return f


def _build_frame(viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, width, height, background, degree, clamped,
                 scale_modifier) -> Frame:
    f = Frame()
    f.view[:] = host_floats(viewmatrix, 16).tolist()        # forward.py:694: row-major flatten
    f.proj[:] = host_floats(projmatrix, 16).tolist()
    f.campos[:] = host_floats(campos, 3).tolist()
    f.background[:] = host_floats(background, 3).tolist()
    f.tan_fovx, f.tan_fovy = float(tan_fovx), float(tan_fovy)
    f.scale_modifier = float(scale_modifier)
    f.width, f.height = int(width), int(height)
    f.degree, f.clamped = int(degree), int(bool(clamped))
    return f
