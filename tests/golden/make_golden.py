#!/usr/bin/env python
"""Generate golden vectors by executing the UNMODIFIED reference sources under the Warp shim.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Imports /root/reference/{forward,backward,optimizer,loss,train}.py with ``tests/warp_shim`` first
on sys.path (so ``import warp`` resolves to the shim) and with inert stubs for the plotting / IO
modules train.py imports (matplotlib, imageio, tqdm, plyfile).  Every array written below is an
output of the reference's own Python source; nothing from this repository's oracle or CUDA code is
involved.  Outputs: tests/golden/*.npz (small; committed).
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GS_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "..", "warp_shim"))
sys.path.insert(0, REF)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


mpl = _stub("matplotlib")
mpl.pyplot = _stub("matplotlib.pyplot")
_stub("imageio", imread=None, imwrite=None)
_stub("tqdm", tqdm=lambda *a, **k: None)
_stub("plyfile", PlyData=None, PlyElement=None)

import warp as wp  # noqa: E402  (the shim)

import backward as ref_backward  # noqa: E402
import forward as ref_forward  # noqa: E402
import loss as ref_loss  # noqa: E402
import optimizer as ref_optimizer  # noqa: E402
import train as ref_train  # noqa: E402
from utils.camera_utils import load_camera  # noqa: E402
from utils.math_utils import projection_matrix, world_to_view  # noqa: E402

import json  # noqa: E402


def npy(a):
    return a.numpy() if isinstance(a, wp.array) else np.asarray(a)


def lego_camera(width, height, frame=0):
    tr = json.load(open(os.path.join(REF, "data/lego/transforms_train.json")))
    focal = 0.5 * width / np.tan(0.5 * tr["camera_angle_x"])        # train.py:297
    info = {"camera_id": frame, "camera_to_world": tr["frames"][frame]["transform_matrix"], "width": width,
            "height": height, "focal": focal}
    return load_camera(info)


def random_scene(rng, n, s_lo, s_hi):
    means = rng.uniform(-1.3, 1.3, (n, 3)).astype(np.float32)
    scales = np.exp(rng.uniform(np.log(s_lo), np.log(s_hi), (n, 3))).astype(np.float32)
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    rot = q.astype(np.float32)
    opac = rng.uniform(0.05, 0.95, n).astype(np.float32)
    sh = rng.normal(size=(n, 16, 3))
    sh[:, 0] *= 2.0
    sh[:, 1:] *= 0.3
    return means, scales, rot, opac, sh.astype(np.float32)


def run_case(name, means, scales, rot, opac, sh, viewmatrix, projmatrix, tan_fovx, tan_fovy, W, H, campos, bg,
             degree, clamped, scale_modifier, target, expect_bwd_oob=False):
    """forward.render_gaussians -> loss.compute_image_gradients -> backward.backward, exactly the
    call sequence of train.py:935-1044."""
    bg = np.array(bg, dtype=np.float32)
    img, depth, buf = ref_forward.render_gaussians(
        background=bg, means3D=means, colors=None, opacity=opac, scales=scales, rotations=rot,
        scale_modifier=scale_modifier, viewmatrix=viewmatrix, projmatrix=projmatrix, tan_fovx=tan_fovx,
        tan_fovy=tan_fovy, image_height=H, image_width=W, sh=sh, degree=degree, campos=campos,
        prefiltered=False, antialiasing=False, clamped=clamped)
    l1 = ref_loss.l1_loss(img, target)
    dL_dpix = ref_loss.compute_image_gradients(img, target, lambda_dssim=0)
    view_w = wp.mat44(np.asarray(viewmatrix).flatten())
    proj_w = wp.mat44(np.asarray(projmatrix).flatten())
    campos_w = wp.vec3(campos[0], campos[1], campos[2])
    geom = {"radii": buf["radii"], "means2D": buf["points_xy_image"], "conic_opacity": buf["conic_opacity"],
            "rgb": buf["colors"], "clamped": buf["clamped_state"]}
    params = {
        "positions": wp.array(means, dtype=wp.vec3), "scales": wp.array(scales, dtype=wp.vec3),
        "rotations": wp.array(rot, dtype=wp.vec4), "opacities": wp.array(opac.reshape(-1), dtype=float),
        "shs": wp.array(sh.reshape(-1, 3), dtype=wp.vec3),
    }
    grads = {}
    try:
      grads = ref_backward.backward(
        background=bg, means3D=params["positions"], dL_dpixels=dL_dpix, opacity=params["opacities"],
        shs=params["shs"], scales=params["scales"], rotations=params["rotations"], scale_modifier=scale_modifier,
        viewmatrix=view_w, projmatrix=proj_w, tan_fovx=tan_fovx, tan_fovy=tan_fovy, image_height=H,
        image_width=W, campos=campos_w, radii=buf["radii"], means2D=buf["points_xy_image"],
        conic_opacity=buf["conic_opacity"], rgb=buf["colors"], cov3Ds=buf["cov3Ds"], clamped=buf["clamped_state"],
        geom_buffer=geom, binning_buffer={"point_list": buf["point_list"]},
        img_buffer={"ranges": buf["ranges"], "final_Ts": buf["final_Ts"], "n_contrib": buf["n_contrib"]},
        degree=degree, debug=False)
      assert not expect_bwd_oob
    except IndexError as e:
        # The reference sizes dL_dsh as N*(degree+1)^2 (backward.py:1122-1123) but indexes it with
        # stride 16 (backward.py:101): for degree < 3 its own backward writes out of bounds.
        assert expect_bwd_oob, e
        print(f"{name}: reference backward is out-of-bounds for degree={degree}: {e}")
    out = dict(
        in_means=means, in_scales=scales, in_rot=rot, in_opac=opac, in_sh=sh, in_view=np.asarray(viewmatrix),
        in_proj=np.asarray(projmatrix), in_tan_fovx=np.float64(tan_fovx), in_tan_fovy=np.float64(tan_fovy),
        in_W=W, in_H=H, in_campos=np.asarray(campos), in_bg=bg, in_degree=degree, in_clamped=int(clamped),
        in_scale_modifier=np.float64(scale_modifier), in_target=target, ref_bwd_oob=int(expect_bwd_oob),
        image=npy(img), depth=npy(depth), l1=np.float64(l1), dL_dpixels=npy(dL_dpix),
    )
    for k, v in buf.items():
        out["fwd_" + k] = npy(v)
    for k, v in grads.items():
        out["bwd_" + k] = npy(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: N={len(means)} {W}x{H} D={len(out['fwd_point_list'])} visible={(out['fwd_radii'] > 0).sum()} "
          f"l1={l1:.5f} max n_contrib={out['fwd_n_contrib'].max()}")


def case_lego_small():
    rng = np.random.default_rng(7)
    W, H = 40, 36                                    # ragged in both directions (3x3 tiles)
    cam = lego_camera(W, H, frame=0)
    means, scales, rot, opac, sh = random_scene(rng, 56, 0.05, 0.45)
    means[5] = means[4]                              # exact depth tie -> stable-sort order matters
    scales[5] = scales[4] * 1.5
    means[9] = cam["camera_center"] * 1.5            # behind the camera -> near-culled
    opac[11] = 0.0                                   # never contributes
    scales[13] = 3.0                                 # huge splat: covers every tile
    opac[13] = 0.3
    target = rng.uniform(0, 1, (H, W, 3)).astype(np.float32)
    target[:4] = 0.0                                 # exact zeros: sign(0) = +1 where render is 0 too
    run_case("ref_lego_small", means, scales, rot, opac, sh, cam["world_to_camera"], cam["full_proj_matrix"],
             cam["tan_fovx"], cam["tan_fovy"], W, H, cam["camera_center"], (0.0, 0.0, 0.0), 3, True, 1.0, target)


def case_lego_deg1():
    rng = np.random.default_rng(11)
    W, H = 33, 17
    cam = lego_camera(W, H, frame=17)
    means, scales, rot, opac, sh = random_scene(rng, 40, 0.08, 0.5)
    rot *= 1.1                                       # non-unit quaternions: forward does not normalise (F4)
    target = rng.uniform(0, 1, (H, W, 3)).astype(np.float32)
    run_case("ref_lego_deg1", means, scales, rot, opac, sh, cam["world_to_camera"], cam["full_proj_matrix"],
             cam["tan_fovx"], cam["tan_fovy"], W, H, cam["camera_center"], (0.2, 0.5, 0.9), 1, False, 1.3, target,
             expect_bwd_oob=True)


def case_lego_bg():
    """degree 3, coloured background, clamped=False, scale_modifier != 1, non-unit quaternions."""
    rng = np.random.default_rng(13)
    W, H = 50, 20
    cam = lego_camera(W, H, frame=42)
    means, scales, rot, opac, sh = random_scene(rng, 44, 0.08, 0.5)
    rot *= 0.9
    target = rng.uniform(0, 1, (H, W, 3)).astype(np.float32)
    run_case("ref_lego_bg", means, scales, rot, opac, sh, cam["world_to_camera"], cam["full_proj_matrix"],
             cam["tan_fovx"], cam["tan_fovy"], W, H, cam["camera_center"], (0.2, 0.5, 0.9), 3, False, 1.3, target)


def case_example_scene():
    """render.py:11-125 at 96x96 (same fov/points/SH; malformed view_matrix exercised in backward)."""
    W = H = 96
    sys.modules["matplotlib.pyplot"].figure = lambda *a, **k: None
    import render as ref_render
    pts, shs, scales, colors, rotations, opacities, cam = ref_render.setup_example_scene(W, H)
    rng = np.random.default_rng(3)
    target = rng.uniform(0, 1, (H, W, 3)).astype(np.float32)
    run_case("ref_example_96", pts, scales, rotations, opacities, shs.astype(np.float32), cam["view_matrix"],
             cam["full_proj_matrix"], cam["tan_fovx"], cam["tan_fovy"], W, H, cam["camera_center"],
             (0.0, 0.0, 0.0), 3, True, 1.0, target)


def case_adam():
    rng = np.random.default_rng(5)
    n = 24
    P = {"positions": rng.normal(size=(n, 3)), "scales": rng.uniform(0.0005, 0.2, (n, 3)),
         "rotations": rng.normal(size=(n, 4)), "opacities": rng.uniform(0, 1, n),
         "shs": rng.normal(size=(n * 16, 3))}
    P = {k: v.astype(np.float32) for k, v in P.items()}
    dt = {"positions": wp.vec3, "scales": wp.vec3, "rotations": wp.vec4, "opacities": float, "shs": wp.vec3}
    params = {k: wp.array(v, dtype=dt[k]) for k, v in P.items()}
    m = {k: wp.zeros(params[k].shape, dtype=dt[k]) for k in P}
    v = {k: wp.zeros(params[k].shape, dtype=dt[k]) for k in P}
    out = {"in_" + k: val for k, val in P.items()}
    lrs = dict(lr_pos=1e-2, lr_scale=5e-3, lr_rot=5e-3, lr_opac=5e-3, lr_sh=2e-3)
    its = [0, 1, 7]
    for step, it in enumerate(its):
        G = {k: (rng.normal(size=val.shape) * (10.0 ** rng.uniform(-6, 0))).astype(np.float32)
             for k, val in P.items()}
        G["opacities"][::5] = 0.0
        grads = {k: wp.array(val, dtype=dt[k]) for k, val in G.items()}
        wp.launch(ref_optimizer.adam_update, dim=n, inputs=[          # argument order of train.py:750-794
            grads["positions"], grads["scales"], grads["rotations"], grads["opacities"], grads["shs"], n,
            lrs["lr_pos"], lrs["lr_scale"], lrs["lr_rot"], lrs["lr_opac"], lrs["lr_sh"], 0.9, 0.999, 1e-8, it,
            params["positions"], params["scales"], params["rotations"], params["opacities"], params["shs"],
            m["positions"], m["scales"], m["rotations"], m["opacities"], m["shs"],
            v["positions"], v["scales"], v["rotations"], v["opacities"], v["shs"]])
        for k in P:
            out[f"g{step}_{k}"] = G[k]
            out[f"p{step}_{k}"] = npy(params[k])
            out[f"m{step}_{k}"] = npy(m[k])
            out[f"v{step}_{k}"] = npy(v[k])
    out["iterations"] = np.array(its)
    np.savez_compressed(os.path.join(HERE, "ref_adam.npz"), **out)
    print("ref_adam: 3 steps on", n, "Gaussians")


def case_densify():
    """train.py:351-713 driven on a trainer object built without __init__ (no dataset needed)."""
    rng = np.random.default_rng(9)
    n = 1300
    T = object.__new__(ref_train.NeRFGaussianSplattingTrainer)
    T.config = ref_train.GaussianParams.get_config_dict()
    T.scene_extent = 4.729
    T.num_points = n
    P = {"positions": rng.uniform(-1.3, 1.3, (n, 3)), "scales": np.exp(rng.uniform(np.log(0.01), np.log(0.2), (n, 3))),
         "rotations": rng.normal(size=(n, 4)), "opacities": rng.uniform(0.0, 0.05, n) ** 1.0,
         "shs": rng.normal(size=(n * 16, 3))}
    P = {k: v.astype(np.float32) for k, v in P.items()}
    P["opacities"][::3] = rng.uniform(0.2, 0.9, len(P["opacities"][::3])).astype(np.float32)
    G = (rng.normal(size=(n, 3)) * 3e-4).astype(np.float32)
    G[-1] = 0.0                                       # last flag clear (see the G5 probe below)
    dt = {"positions": wp.vec3, "scales": wp.vec3, "rotations": wp.vec4, "opacities": float, "shs": wp.vec3}
    T.params = {k: wp.array(v, dtype=dt[k]) for k, v in P.items()}
    T.grads = T.create_gradient_arrays()
    T.grads["positions"] = wp.array(G, dtype=wp.vec3)
    T.adam_m = T.create_gradient_arrays()
    T.adam_v = T.create_gradient_arrays()
    out = {"in_" + k: v for k, v in P.items()}
    out["in_pos_grad"] = G
    out["scene_extent"] = np.float64(T.scene_extent)
    import io
    import contextlib
    wp.OOB_SCALAR_READ_IS_ZERO = True                 # quirk G4: stale avg_grads read past its end -> 0
    wp.OOB_WRITE_IS_DROPPED = True                    # quirk G5: writes one past the end -> dropped
    with contextlib.redirect_stdout(io.StringIO()) as log:
        T.densification_and_pruning(600)
    out["log"] = np.array(log.getvalue())
    out["out_num_points"] = T.num_points
    for k in P:
        out["out_" + k] = npy(T.params[k])
    # iteration 0: nothing but the opacity reset (quirk G6)
    T2 = object.__new__(ref_train.NeRFGaussianSplattingTrainer)
    T2.config = T.config
    T2.scene_extent = 4.729
    T2.num_points = 8
    T2.params = {"opacities": wp.array(np.linspace(0.1, 0.9, 8).astype(np.float32), dtype=float)}
    with contextlib.redirect_stdout(io.StringIO()):
        T2.densification_and_pruning(0)
    out["reset_opacities_it0"] = npy(T2.params["opacities"])
    # quirk G5 probe: when the LAST Gaussian is a clone candidate the exclusive-scan total misses
    # it and clone_gaussians writes one element past the end of the new arrays (optimizer.py:356).
    T3 = object.__new__(ref_train.NeRFGaussianSplattingTrainer)
    T3.config, T3.scene_extent, T3.num_points = T.config, 4.729, n
    P3 = {k: v.copy() for k, v in P.items()}
    P3["scales"][-1] = 0.01
    G3 = G.copy()
    G3[-1] = 1.0
    T3.params = {k: wp.array(v, dtype=dt[k]) for k, v in P3.items()}
    T3.grads = T3.create_gradient_arrays()
    T3.grads["positions"] = wp.array(G3, dtype=wp.vec3)
    T3.adam_m = T3.create_gradient_arrays()
    T3.adam_v = T3.create_gradient_arrays()
    n_before = len(wp.OOB_WRITES)
    with contextlib.redirect_stdout(io.StringIO()):
        T3.densification_and_pruning(600)
    out["g5_probe_num_points"] = T3.num_points
    out["g5_probe_positions"] = npy(T3.params["positions"])
    print("ref_densify: G5 probe ->", T3.num_points, "points;", len(wp.OOB_WRITES) - n_before, "dropped OOB writes")
    wp.OOB_SCALAR_READ_IS_ZERO = False
    wp.OOB_WRITE_IS_DROPPED = False
    out["oob_writes"] = np.array(sorted(set(wp.OOB_WRITES)), dtype=np.int64).reshape(-1, 2)
    np.savez_compressed(os.path.join(HERE, "ref_densify.npz"), **out)
    print("ref_densify:", n, "->", T.num_points, "|", log.getvalue().replace("\n", " | ")[:300])


def case_loss():
    """loss.py l1_loss + compute_image_gradients incl. exactly-equal pixels ([Warp] sign(0) = +1)."""
    rng = np.random.default_rng(21)
    H, W = 13, 21
    r = rng.uniform(0, 1, (H, W, 3)).astype(np.float32)
    t = rng.uniform(0, 1, (H, W, 3)).astype(np.float32)
    t[2:5] = r[2:5]                                   # equal pixels
    r[7], t[7] = 0.0, 0.0                             # black render on black target
    l1 = ref_loss.l1_loss(r, t)
    g = ref_loss.compute_image_gradients(r, t, lambda_dssim=0)
    g2 = ref_loss.compute_image_gradients(r, t, lambda_dssim=0.2)
    np.savez_compressed(os.path.join(HERE, "ref_loss.npz"), rendered=r, target=t, l1=np.float64(l1), grad=npy(g),
                        grad_dssim02=npy(g2))
    print("ref_loss written", l1)


def case_loss2():
    """loss.py ssim (11x11 Gaussian window, sigma 1.5, truncated at the borders) and depth_loss -- dead code in
    train.py (967-974) but part of loss.py's surface."""
    rng = np.random.default_rng(22)
    H, W = 17, 23
    r = rng.uniform(0, 1, (H, W, 3)).astype(np.float32)
    t = np.clip(r + rng.normal(0, 0.1, (H, W, 3)), 0, 1).astype(np.float32)
    t[3:6, 4:9] = r[3:6, 4:9]
    s1 = ref_loss.ssim(r, t)
    s_same = ref_loss.ssim(r, r)
    rd = rng.uniform(0.1, 2.0, (H, W)).astype(np.float32)
    td = rng.uniform(0.1, 2.0, (H, W)).astype(np.float32)
    mask = (rng.uniform(0, 1, (H, W)) > 0.3).astype(np.float32)
    dl = ref_loss.depth_loss(rd, td, mask)
    np.savez_compressed(os.path.join(HERE, "ref_loss2.npz"), rendered=r, target=t, ssim=np.float64(s1),
                        ssim_same=np.float64(s_same), rendered_depth=rd, target_depth=td, depth_mask=mask,
                        depth_loss=np.float64(dl))
    print("ref_loss2 written", s1, s_same, dl)


def case_misc():
    """LR schedule, init kernel, exclusive scan + prefix sum helpers."""
    from scheduler import LRScheduler
    s = LRScheduler(1e-2, 0.01)
    its = np.array([0, 1, 100, 3500, 6999, 8000])
    lrs = np.array([s.get_lr(int(i), 7000) for i in its])
    n = 32
    dt = {"positions": wp.vec3, "scales": wp.vec3, "rotations": wp.vec4, "opacities": float, "shs": wp.vec3}
    arrs = {k: wp.zeros(n * (16 if k == "shs" else 1), dtype=dt[k]) for k in dt}
    wp.launch(ref_train.init_gaussian_params, dim=n, inputs=[arrs["positions"], arrs["scales"], arrs["rotations"],
                                                             arrs["opacities"], arrs["shs"], n, 0.1])
    out = {"lr_its": its, "lr_vals": lrs}
    for k in dt:
        out["init_" + k] = npy(arrs[k])
    np.savez_compressed(os.path.join(HERE, "ref_misc.npz"), **out)
    print("ref_misc written")


def case_cameras():
    """utils/camera_utils.py:8-91 load_camera on all 100 Lego train poses (800x800, the headline size) and on a
    non-square frame, plus NeRFGaussianSplattingTrainer.calculate_scene_extent (train.py:233-257) called unbound on a
    stand-in object; and the zero_gradients kernel (train.py:94-115) run on non-zero arrays."""
    tr = json.load(open(os.path.join(REF, "data/lego/transforms_train.json")))
    out = {}
    for tag, (W, H) in (("sq", (800, 800)), ("hd", (1920, 1080))):
        focal = 0.5 * W / np.tan(0.5 * tr["camera_angle_x"])        # train.py:297
        cams = [load_camera({"camera_id": i, "camera_to_world": f["transform_matrix"], "width": W, "height": H,
                             "focal": focal}) for i, f in enumerate(tr["frames"])]
        for key in ("world_to_camera", "full_proj_matrix", "camera_center", "view_matrix", "proj_matrix", "R", "T"):
            out[f"{tag}_{key}"] = np.stack([np.asarray(c[key]) for c in cams])
        for key in ("tan_fovx", "tan_fovy", "fx", "fy", "cx", "cy"):
            out[f"{tag}_{key}"] = np.array([c[key] for c in cams], dtype=np.float64)
        holder = types.SimpleNamespace(cameras=cams, config={})
        out[f"{tag}_scene_extent"] = np.float64(ref_train.NeRFGaussianSplattingTrainer.calculate_scene_extent(holder))
    rng = np.random.default_rng(12)
    n = 37
    dt = {"positions": wp.vec3, "scales": wp.vec3, "rotations": wp.vec4, "opacities": float, "shs": wp.vec3}
    width = {"positions": 3, "scales": 3, "rotations": 4, "opacities": 1, "shs": 3}
    arrs = {}
    for k in dt:
        rows = (n + 3) * (16 if k == "shs" else 1)      # three Gaussians beyond num_points must stay untouched
        a = rng.normal(size=(rows, width[k]) if width[k] > 1 else (rows,)).astype(np.float32)
        out["zg_in_" + k] = a
        arrs[k] = wp.array(a, dtype=dt[k])
    wp.launch(ref_train.zero_gradients, dim=n + 3, inputs=[arrs["positions"], arrs["scales"], arrs["rotations"],
                                                           arrs["opacities"], arrs["shs"], n])
    for k in dt:
        out["zg_out_" + k] = npy(arrs[k])
    out["zg_n"] = np.int64(n)
    np.savez_compressed(os.path.join(HERE, "ref_cameras.npz"), **out)
    print("ref_cameras written", out["sq_scene_extent"])


if __name__ == "__main__":
    which = sys.argv[1:] or ["lego_small", "lego_deg1", "lego_bg", "example_scene", "adam", "densify", "loss", "loss2", "misc",
                             "cameras"]
    for w in which:
        globals()["case_" + w]()
