"""The reference's camera struct (utils/camera_utils.py:8-91), consumed unchanged by
``forward.render_gaussians`` / ``backward.backward``.

``load_camera`` takes the same ``camera_info`` dict (camera_id, camera_to_world, width, height,
focal) and returns a dict with the same keys: ``world_to_camera`` (= (W2C)^T, translation in the
last row), ``full_proj_matrix`` (= world_to_camera @ P^T), ``tan_fovx/y``, ``camera_center``, ...
"""
import json
import os

import numpy as np

from .math_utils import projection_matrix, world_to_view

_DISTORTION_KEYS = ("k1", "k2", "p1", "p2", "k3", "k4")


def load_camera(camera_info):
    c2w = np.asarray(camera_info["camera_to_world"], dtype=np.float64).copy()
    c2w[:3, 1:3] *= -1                      # Blender/OpenGL axes -> COLMAP (y down, z forward), line 15
    w2c = np.linalg.inv(c2w).astype(np.float32)
    R = w2c[:3, :3]                          # views into w2c, exactly like the reference (lines 22-23)
    T = w2c[:3, 3]
    w2c[3, 3] = 1.0
    w2c = w2c.T                              # row-vector convention (line 27)

    width, height = camera_info.get("width"), camera_info.get("height")
    fx = fy = camera_info.get("focal")
    fovx = 2 * np.arctan(width / (2 * fx))
    fovy = 2 * np.arctan(height / (2 * fy))
    view_matrix = world_to_view(R=R, t=T)
    proj = projection_matrix(fovx=fovx, fovy=fovy, znear=0.01, zfar=100.0).T
    full_proj = w2c @ proj                   # float32 @ float64 -> float64 (line 48)

    model = camera_info.get("camera_model", "OPENCV")
    if model == "OPENCV" or model is None:
        camera_type = 0
    elif model == "OPENCV_FISHEYE":
        camera_type = 1
    else:
        raise ValueError(f"Unsupported camera_model '{model}'")
    return {
        "R": R, "T": T,
        "camera_center": np.linalg.inv(w2c)[3, :3],
        "view_matrix": view_matrix, "proj_matrix": proj, "full_proj_matrix": full_proj,
        "tan_fovx": np.tan(fovx * 0.5), "tan_fovy": np.tan(fovy * 0.5),
        "fx": fx, "fy": fy, "cx": width / 2, "cy": height / 2, "width": width, "height": height,
        "camera_to_world": c2w, "world_to_camera": w2c, "camera_type": camera_type,
        "distortion_params": np.array([camera_info.get(k, 0.0) for k in _DISTORTION_KEYS], dtype=np.float32),
    }


_POSES = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "lego_train_poses.json")


def load_nerf_cameras(width=800, height=800, transforms_path=None):
    """train.py:265-321 load_nerf_data without the PNG probing: focal from ``camera_angle_x``,
    one ``load_camera`` struct per frame.  Defaults to the bundled Lego train poses."""
    with open(transforms_path or _POSES) as f:
        tr = json.load(f)
    focal = 0.5 * width / np.tan(0.5 * tr["camera_angle_x"])      # train.py:297
    cams = []
    for i, frame in enumerate(tr["frames"]):
        cams.append(load_camera({"camera_id": i, "camera_to_world": frame["transform_matrix"], "width": width,
                                 "height": height, "focal": focal}))
    return cams


def scene_extent(cameras, camera_extent_factor=1.0):
    """train.py:233-257: max distance of a camera centre to the centroid of all centres, >= 1."""
    if not cameras:
        return 1.0
    pos = np.array([c["camera_center"] for c in cameras])
    centre = np.mean(pos, axis=0)
    d = max(float(np.linalg.norm(p - centre)) for p in pos)
    return max(d * camera_extent_factor, 1.0)
