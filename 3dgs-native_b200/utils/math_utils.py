"""Camera-matrix conventions of the reference (utils/math_utils.py:8-41), restated in numpy.

All matrices the rasterizer consumes are applied as ROW-VECTOR x MATRIX (forward.py:248,253):
the translation sits in the last row of ``world_to_camera``.
"""
import math

import numpy as np


def world_to_view(R, t, translate=np.array([0.0, 0.0, 0.0]), scale=1.0):
    """utils/math_utils.py:8-19.  Returns [[R^T, t], [0, 1]] (re-centred), float32.  NOTE: this is
    the malformed ``view_matrix`` that only render.py:112 feeds to the rasterizer."""
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = np.asarray(R).transpose()
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    c2w = np.linalg.inv(Rt)
    c2w[:3, 3] = (c2w[:3, 3] + translate) * scale
    return np.float32(np.linalg.inv(c2w))


def projection_matrix(fovx, fovy, znear, zfar):
    """utils/math_utils.py:21-41 (OpenGL-style frustum, z_sign = +1), float64."""
    tan_y = math.tan(fovy / 2)
    tan_x = math.tan(fovx / 2)
    top, right = tan_y * znear, tan_x * znear
    bottom, left = -top, -right
    P = np.zeros((4, 4))
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P
