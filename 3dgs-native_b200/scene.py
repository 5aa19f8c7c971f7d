"""Synthetic inputs for parity tests and benchmarks.

* ``example_scene``  -- the 3-Gaussian fixture of the reference's render.py:11-82 (config 1).
* ``synthetic_scene`` -- N Gaussians in the reference's init cube, drawn from
  ``numpy.random.default_rng(42)`` in a fixed order, plus a target image (configs 2-5).
"""
import math

import numpy as np

from .utils.camera_utils import load_nerf_cameras
from .utils.math_utils import projection_matrix, world_to_view

# (N, W, H, s_min, s_max) for BASELINE.json's configs 2, 3 and 5
CONFIGS = {
    "C2": (300_000, 800, 800, 0.004, 0.02),
    "C3": (1_000_000, 1920, 1080, 0.003, 0.015),
    "C5": (6_000_000, 3840, 2160, 0.002, 0.01),
}

_EXAMPLE_SH = [
    [0.71734341, 0.91905449, 0.49961076], [0.08068483, 0.82132256, 0.01301602],
    [0.8335743, 0.31798138, 0.19709007], [0.82589597, 0.28206231, 0.790489],
    [0.24008527, 0.21312673, 0.53132892], [0.19493135, 0.37989934, 0.61886235],
    [0.98106522, 0.28960672, 0.57313965], [0.92623716, 0.46034381, 0.5485369],
    [0.81660616, 0.7801104, 0.27813915], [0.96114063, 0.69872817, 0.68313804],
    [0.95464185, 0.21984855, 0.92912192], [0.23503135, 0.29786121, 0.24999751],
    [0.29844887, 0.6327788, 0.05423596], [0.08934335, 0.11851827, 0.04186001],
    [0.59331831, 0.919777, 0.71364335], [0.83377388, 0.40242542, 0.8792624],
]


def example_scene(image_width=1800, image_height=1800, fovx=45.0, fovy=45.0, znear=0.01, zfar=100.0):
    """render.py:11-82, including its quirks: fov in *radians* = 45.0, and ``view_matrix`` (not
    ``world_to_camera``) handed to the rasterizer by render.py:112."""
    T = np.array([0, 0, 5], dtype=np.float32)
    R = np.array([[1, 0, 0], [0, 1, 0], [0, 0, -1]], dtype=np.float32)
    w2c = np.eye(4, dtype=np.float32)
    w2c[:3, :3] = R
    w2c[:3, 3] = T
    w2c = w2c.T
    view = world_to_view(R=R, t=T)
    proj = projection_matrix(fovx=fovx, fovy=fovy, znear=znear, zfar=zfar).T
    cam = {
        "R": R, "T": T, "camera_center": np.linalg.inv(w2c)[3, :3], "view_matrix": view, "proj_matrix": proj,
        "world_to_camera": w2c, "full_proj_matrix": w2c @ proj, "tan_fovx": math.tan(fovx * 0.5),
        "tan_fovy": math.tan(fovy * 0.5), "width": image_width, "height": image_height,
    }
    cam["focal_x"] = image_width / (2 * cam["tan_fovx"])
    cam["focal_y"] = image_height / (2 * cam["tan_fovy"])
    pts = np.array([[-5, 0, -10], [0, 0, -10], [5, 0, -10]], dtype=np.float32)
    n = len(pts)
    shs = np.array(_EXAMPLE_SH * n).reshape(n, 16, 3)
    opacities = np.ones((n, 1), dtype=np.float32)
    scales = np.ones((n, 3), dtype=np.float32)
    rotations = np.zeros((n, 4), dtype=np.float32)
    rotations[:, 3] = 1.0
    colors = np.ones((n, 3), dtype=np.float32)
    return pts, shs, scales, colors, rotations, opacities, cam


def example_render_kwargs(image_width=1800, image_height=1800):
    """The exact call of render.py:104-125 as a kwargs dict."""
    pts, shs, scales, colors, rotations, opacities, cam = example_scene(image_width, image_height)
    return dict(background=np.array([0.0, 0.0, 0.0], dtype=np.float32), means3D=pts, colors=colors,
                opacity=opacities, scales=scales, rotations=rotations, scale_modifier=1.0,
                viewmatrix=cam["view_matrix"], projmatrix=cam["full_proj_matrix"], tan_fovx=cam["tan_fovx"],
                tan_fovy=cam["tan_fovy"], image_height=image_height, image_width=image_width, sh=shs, degree=3,
                campos=cam["camera_center"], prefiltered=False, antialiasing=False, clamped=True, debug=False)


def synthetic_scene(n, width, height, s_min, s_max, seed=42, camera_index=0, with_target=True):
    """SURVEY 8d generator.  Draw order is part of the definition: means, scales, quaternions,
    opacities, SH, target.  Returns (params dict of float32 arrays, camera struct, target)."""
    rng = np.random.default_rng(seed)
    means = rng.uniform(-1.3, 1.3, (n, 3)).astype(np.float32)
    scales = np.exp(rng.uniform(math.log(s_min), math.log(s_max), (n, 3))).astype(np.float32)
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    rotations = q.astype(np.float32)                       # (x, y, z, w)
    opacities = rng.uniform(0.05, 0.95, n).astype(np.float32)
    sh = rng.normal(size=(n, 16, 3))
    sh[:, 0, :] *= 2.0
    sh[:, 1:, :] *= 0.3
    sh = sh.astype(np.float32)
    target = rng.uniform(0, 1, (height, width, 3)).astype(np.float32) if with_target else None
    cam = load_nerf_cameras(width, height)[camera_index]
    params = {"positions": means, "scales": scales, "rotations": rotations, "opacities": opacities,
              "shs": sh.reshape(n * 16, 3)}
    return params, cam, target


def render_kwargs(params, cam, background=(0.0, 0.0, 0.0), degree=3, clamped=True, scale_modifier=1.0):
    """The forward call of train.py:935-955 as kwargs."""
    return dict(background=np.array(background, dtype=np.float32), means3D=params["positions"], colors=None,
                opacity=params["opacities"], scales=params["scales"], rotations=params["rotations"],
                scale_modifier=scale_modifier, viewmatrix=cam["world_to_camera"],
                projmatrix=cam["full_proj_matrix"], tan_fovx=cam["tan_fovx"], tan_fovy=cam["tan_fovy"],
                image_height=cam["height"], image_width=cam["width"], sh=params["shs"], degree=degree,
                campos=cam["camera_center"], prefiltered=False, antialiasing=False, clamped=clamped)


def backward_kwargs(params, cam, buffers, dL_dpixels, background=(0.0, 0.0, 0.0), degree=3, scale_modifier=1.0):
    """The backward call of train.py:1006-1044 as kwargs (``buffers`` = forward's dict)."""
    geom = {"radii": buffers["radii"], "means2D": buffers["points_xy_image"],
            "conic_opacity": buffers["conic_opacity"], "rgb": buffers["colors"],
            "clamped": buffers["clamped_state"]}
    return dict(background=np.array(background, dtype=np.float32), means3D=params["positions"],
                dL_dpixels=dL_dpixels, opacity=params["opacities"], shs=params["shs"], scales=params["scales"],
                rotations=params["rotations"], scale_modifier=scale_modifier, viewmatrix=cam["world_to_camera"],
                projmatrix=cam["full_proj_matrix"], tan_fovx=cam["tan_fovx"], tan_fovy=cam["tan_fovy"],
                image_height=cam["height"], image_width=cam["width"], campos=cam["camera_center"],
                radii=buffers["radii"], means2D=buffers["points_xy_image"],
                conic_opacity=buffers["conic_opacity"], rgb=buffers["colors"], cov3Ds=buffers["cov3Ds"],
                clamped=buffers["clamped_state"], geom_buffer=geom,
                binning_buffer={k: buffers[k] for k in ("point_list", "block_masks") if k in buffers},
                img_buffer={"ranges": buffers["ranges"], "final_Ts": buffers["final_Ts"],
                            "n_contrib": buffers["n_contrib"]}, degree=degree, debug=False)
