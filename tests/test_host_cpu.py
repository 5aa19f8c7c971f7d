"""CPU-only tests of the host-side helpers of the Python mirror (no kernels are called)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")


def test_make_frame_keeps_row_major_matrices():
    import gsb200  # noqa: F401
    from gsb200 import _lib
    v = np.arange(16, dtype=np.float32).reshape(4, 4)
    p = v.T.copy()
    f = _lib.make_frame(v, torch.from_numpy(p), [1, 2, 3], 0.3, 0.4, 80, 60, (0.1, 0.2, 0.3), degree=2, clamped=False,
                        scale_modifier=1.5)
    assert list(f.view) == list(range(16))                      # forward.py:694: row-major flatten
    assert list(f.proj) == p.reshape(-1).tolist()
    assert (f.width, f.height, f.degree, f.clamped) == (80, 60, 2, 0)
    assert abs(f.scale_modifier - 1.5) < 1e-7 and abs(f.tan_fovy - 0.4) < 1e-7
