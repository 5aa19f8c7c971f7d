"""Checkpoint I/O ("next" row 8f-3): the reference's PLY point-cloud layout
(utils/point_cloud_utils.py:10-99) written with vectorised numpy instead of a Python loop per point
and without the ``plyfile`` dependency, plus a reader and a resumable training checkpoint
(the reference saves the PLY only: no Adam state, no iteration counter, no loader).

PLY layout (binary little endian, one packed 239-byte record per Gaussian):
x y z, scale_0..2, opacity, rot_x rot_y rot_z rot_w (float32), red green blue (uint8 =
clip(SH_dc + 0.5, 0, 1) * 255 truncated), f_dc_0..2, f_rest_0..44 (float32) where
f_rest_{3(j-1)+c} = SH coefficient j, channel c (coefficient-major interleaved; NOT the
channel-major layout of the INRIA viewer)."""
from __future__ import annotations

import os

import numpy as np

_FLOAT_FIELDS_A = ["x", "y", "z", "scale_0", "scale_1", "scale_2", "opacity", "rot_x", "rot_y", "rot_z", "rot_w"]
_U8_FIELDS = ["red", "green", "blue"]
_FLOAT_FIELDS_B = ["f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{i}" for i in range(45)]
VERTEX_DTYPE = np.dtype([(n, "<f4") for n in _FLOAT_FIELDS_A] + [(n, "u1") for n in _U8_FIELDS]
                        + [(n, "<f4") for n in _FLOAT_FIELDS_B])


def _np(a):
    if hasattr(a, "detach"):
        return a.detach().cpu().numpy()
    if hasattr(a, "numpy") and not isinstance(a, np.ndarray):
        return a.numpy()
    return np.asarray(a)


def vertex_array(params, num_points, colors=None):
    n = int(num_points)
    pos = _np(params["positions"]).reshape(-1, 3)[:n]
    scl = _np(params["scales"]).reshape(-1, 3)[:n]
    rot = _np(params["rotations"]).reshape(-1, 4)[:n]
    opa = _np(params["opacities"]).reshape(-1)[:n]
    shs = _np(params["shs"]).reshape(-1, 16, 3)[:n]
    if colors is None:                                  # utils/point_cloud_utils.py:26-33
        col = np.clip(shs[:, 0, :] + 0.5, 0.0, 1.0)
    else:
        col = _np(colors).reshape(-1, 3)[:n]
    v = np.empty(n, dtype=VERTEX_DTYPE)
    for k, name in enumerate(("x", "y", "z")):
        v[name] = pos[:, k]
        v[f"scale_{k}"] = scl[:, k]
        v[("red", "green", "blue")[k]] = np.clip(col[:, k] * 255, 0, 255).astype(np.int64).astype(np.uint8)
        v[f"f_dc_{k}"] = shs[:, 0, k]
    v["opacity"] = opa
    for k, name in enumerate(("rot_x", "rot_y", "rot_z", "rot_w")):
        v[name] = rot[:, k]
    rest = shs[:, 1:, :].reshape(n, 45)                 # coefficient-major interleaved (lines 58-67)
    for i in range(45):
        v[f"f_rest_{i}"] = rest[:, i]
    return v


def save_ply(params, filepath, num_points, colors=None):
    """Same signature and on-disk layout as the reference's ``save_ply``."""
    v = vertex_array(params, num_points, colors)
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {len(v)}"]
    for name in VERTEX_DTYPE.names:
        header.append(f"property {'uchar' if VERTEX_DTYPE[name] == np.uint8 else 'float'} {name}")
    header.append("end_header")
    d = os.path.dirname(str(filepath))
    if d:
        os.makedirs(d, exist_ok=True)
    with open(filepath, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(v.tobytes())


def load_ply(filepath):
    """Reads a file written by ``save_ply`` back into the parameter dict (numpy, reference shapes)."""
    with open(filepath, "rb") as f:
        names, n = [], None
        while True:
            line = f.readline().decode("ascii").strip()
            if line.startswith("element vertex"):
                n = int(line.split()[-1])
            elif line.startswith("property"):
                names.append(line.split()[-1])
            elif line == "end_header":
                break
        if names != list(VERTEX_DTYPE.names):
            raise ValueError("not a point cloud written with the reference's save_ply layout")
        v = np.frombuffer(f.read(n * VERTEX_DTYPE.itemsize), dtype=VERTEX_DTYPE, count=n)
    shs = np.empty((n, 16, 3), np.float32)
    for k in range(3):
        shs[:, 0, k] = v[f"f_dc_{k}"]
    shs[:, 1:, :] = np.stack([v[f"f_rest_{i}"] for i in range(45)], axis=1).reshape(n, 15, 3)
    return {
        "positions": np.stack([v["x"], v["y"], v["z"]], axis=1).astype(np.float32),
        "scales": np.stack([v["scale_0"], v["scale_1"], v["scale_2"]], axis=1).astype(np.float32),
        "rotations": np.stack([v["rot_x"], v["rot_y"], v["rot_z"], v["rot_w"]], axis=1).astype(np.float32),
        "opacities": np.array(v["opacity"], dtype=np.float32),
        "shs": shs.reshape(n * 16, 3),
    }
