#!/usr/bin/env python
"""Kernel micro-benchmark for A/B work on the tile kernels (run on the GPU box):

    python tools/kbench.py [--cfg C2] [--libs default,x,y] [--iters 30]

For every library variant (csrc/libgsb200_<tag>.so built by `make variant TAG=<tag> EXTRA=...`; `default` = the
shipped libgsb200.so) a fresh process renders the headline view and times, with CUDA events and the L2 flushed
before every call: the forward tile kernel, the backward tile kernel (stage entry points), the whole forward and
the whole backward operator; then the steady-state loop forward + loss + backward + Adam without flushes.
Prints one line per variant; checks n_contrib / final_T (bit-exact) and the gradients (1e-6-ish) against the
default library's results so a fast but wrong variant is caught at once."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child_forward_only(cfg, iters, opts):
    """Forward operator only (config 5 is forward-only): whole-operator time, L2 flushed before every call."""
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    import gsb200  # noqa: F401
    from gsb200 import scene, train
    from gsb200.utils.camera_utils import load_nerf_cameras
    n, w, h, smin, smax = scene.CONFIGS[cfg]
    params, _cam, _ = scene.synthetic_scene(n, w, h, smin, smax, seed=42, with_target=False)
    cams = load_nerf_cameras(w, h)[:1]
    T = train.Trainer(cams, params=params, config={"num_iterations": 7000})
    for kv in opts:
        k, v = kv.split("=")
        T.ctx.set_option(k, int(v))
    for _ in range(3):
        fb = T.forward(0)
    torch.cuda.synchronize()
    flush = torch.empty((256 << 20) // 4, dtype=torch.float32, device=T.device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = []
    for i in range(iters):
        flush.fill_(float(i))
        e0.record(); T.forward(0); e1.record(); e1.synchronize()
        t.append(e0.elapsed_time(e1))
    import hashlib
    hsh = hashlib.sha1(fb.point_list[:fb.num_rendered].cpu().numpy().tobytes()).hexdigest()[:12]
    print(f"FWD {cfg} {' '.join(opts) or 'defaults'}: forward {np.median(t) * 1e3:9.1f} us (min {min(t) * 1e3:9.1f}), "
          f"D={fb.num_rendered}, point_list sha1 {hsh}", flush=True)


def child(cfg, iters, out_path, opts=()):
    import ctypes as C
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    import gsb200  # noqa: F401
    from gsb200 import _lib, scene, train
    from gsb200.utils.camera_utils import load_nerf_cameras
    n, w, h, smin, smax = scene.CONFIGS[cfg]
    params, _cam, _ = scene.synthetic_scene(n, w, h, smin, smax, seed=42, with_target=False)
    cams = load_nerf_cameras(w, h)[:4]
    rng = np.random.default_rng(4242)
    targets = [rng.uniform(0, 1, (h, w, 3)).astype(np.float32) for _ in cams]
    lrs = {"lr_pos": 1e-6, "lr_scale": 5e-7, "lr_rot": 5e-7, "lr_sh": 2e-7, "lr_opac": 5e-7, "final_lr_factor": 0.01}
    T = train.Trainer(cams, targets=targets, params=params, config={"num_iterations": 7000, "lr_scheduler_config": lrs},
                      overlap_sh=int(os.environ.get("GSB_OVERLAP_SH", "1")))   # 3: single-GPU Adam overlap
    L, ctx, p = _lib.lib(), T.ctx, _lib.ptr
    for kv in opts:
        k_, v_ = kv.split("=")
        ctx.set_option(k_, int(v_))
    s = lambda: _lib.stream_ptr(ctx.device_index)  # noqa: E731
    fb = T.forward(0)
    T.loss_and_pixel_gradients(fb, T.targets[0])
    T.backward(0, fb, T.grads)
    torch.cuda.synchronize()
    frame, g, N = T.frames[0], T.grads, T.num_points
    calls = {
        "blend_fwd": lambda: L.gsb_blend_forward(ctx.h, s(), C.byref(frame), p(fb.ranges), p(fb.point_list), p(fb.xy),
                                                 p(fb.colors), p(fb.conic_opacity), p(fb.depths), p(fb.image), p(fb.depth),
                                                 p(fb.final_T), p(fb.n_contrib), p(fb.block_masks)),
        "blend_bwd": lambda: L.gsb_blend_backward(ctx.h, s(), C.byref(frame), N, p(fb.ranges), p(fb.point_list), p(fb.xy),
                                                  p(fb.conic_opacity), p(fb.colors), p(fb.final_T), p(fb.n_contrib),
                                                  p(fb.dpix), p(fb.dL_dmean2D), p(fb.dL_dconic), p(g["opacities"]),
                                                  p(fb.dL_dcolor), p(fb.block_masks)),
        "fwd_whole": lambda: (T.forward(0), 0)[1],
        "bwd_whole": lambda: (T.backward(0, fb, T.grads), 0)[1],
    }
    flush = torch.empty((256 << 20) // 4, dtype=torch.float32, device=T.device)
    res = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, fn in calls.items():
        for _ in range(3):
            ctx.check(fn())
        cold, warm = [], []
        for i in range(iters):
            flush.fill_(float(i))
            e0.record(); ctx.check(fn()); e1.record(); e1.synchronize()
            cold.append(e0.elapsed_time(e1))
        for i in range(iters):
            e0.record(); ctx.check(fn()); e1.record(); e1.synchronize()
            warm.append(e0.elapsed_time(e1))
        res[name] = {"cold_us": round(float(np.median(cold)) * 1e3, 1), "warm_us": round(float(np.median(warm)) * 1e3, 1)}
    # steady-state training step
    for it in range(5):
        T.train_step(1 + it, [it % 4], densify=False)
    torch.cuda.synchronize()
    reps = []
    for r in range(5):
        e0.record()
        for it in range(20):
            T.train_step(10 + 20 * r + it, [it % 4], densify=False)
        T.join_exchange()      # (overlap_sh: the last step's SH phase on the side stream belongs to the timed region)
        e1.record(); e1.synchronize()
        reps.append(e0.elapsed_time(e1) / 20)
    res["step_us"] = round(float(np.median(reps)) * 1e3, 1)
    # results for the cross-variant check (fresh frame: the training steps moved the parameters a little)
    T2 = train.Trainer(cams, targets=targets, params=params, config={"num_iterations": 7000, "lr_scheduler_config": lrs})
    fb = T2.forward(0)
    T2.loss_and_pixel_gradients(fb, T2.targets[0])
    T2.backward(0, fb, T2.grads)
    torch.cuda.synchronize()
    np.savez(out_path, n_contrib=fb.n_contrib.cpu().numpy(), final_T=fb.final_T.cpu().numpy(),
             image=fb.image.cpu().numpy(), grads=T2.grads.flat.cpu().numpy(), D=fb.num_rendered)
    print("KBENCH " + json.dumps(res))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="C2")
    ap.add_argument("--libs", default="default")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--child", default="")
    ap.add_argument("--fwd-only", action="store_true", help="time the forward operator only (e.g. --cfg C5)")
    ap.add_argument("--opt", action="append", default=[], help="gsb_set_option name=value (with --fwd-only)")
    args = ap.parse_args()
    if args.fwd_only:
        return child_forward_only(args.cfg, args.iters, args.opt)
    if args.child:
        return child(args.cfg, args.iters, args.child, args.opt)
    import numpy as np
    csrc = os.path.join(ROOT, "3dgs-native_b200", "csrc")
    base = None
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    # a variant is a library tag, optionally followed by options:  default  default:fwd_dyn=0  x:bwd_reduce=1:fwd_dyn=0
    for spec in args.libs.split(","):
        tag, *vopts = spec.split(":")
        env = dict(os.environ)
        if tag != "default":
            env["GSB200_LIB"] = os.path.join(csrc, f"libgsb200_{tag}.so")
        out = os.path.join("/tmp", f"kbench_{spec.replace(':', '_').replace('=', '')}.npz")
        cmd = [sys.executable, os.path.abspath(__file__), "--cfg", args.cfg, "--iters", str(args.iters), "--child", out]
        for o in list(args.opt) + vopts:
            cmd += ["--opt", o]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True)
        tag = spec
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("KBENCH ")]
        if r.returncode != 0 or not line:
            print(f"{tag}: FAILED rc={r.returncode}\n{r.stderr[-1500:]}")
            continue
        res = json.loads(line[0][7:])
        d = np.load(out)
        check = ""
        if base is None:
            base = {k: d[k] for k in d.files}
        else:
            ok_int = np.array_equal(d["n_contrib"], base["n_contrib"]) and np.array_equal(d["final_T"], base["final_T"])
            img = float(np.abs(d["image"] - base["image"]).max())
            a, b = d["grads"].astype(np.float64), base["grads"].astype(np.float64)
            check = f" | vs first: n_contrib/final_T {'identical' if ok_int else 'DIFFER'}, image maxdiff {img:.2e}, " \
                    f"grads rel {np.linalg.norm(a - b) / np.linalg.norm(b):.2e}"
        print(f"{tag:>12}: fwd {res['blend_fwd']['cold_us']:7.1f}/{res['blend_fwd']['warm_us']:7.1f}  "
              f"bwd {res['blend_bwd']['cold_us']:7.1f}/{res['blend_bwd']['warm_us']:7.1f}  "
              f"fwd_whole {res['fwd_whole']['cold_us']:7.1f}/{res['fwd_whole']['warm_us']:7.1f}  "
              f"bwd_whole {res['bwd_whole']['cold_us']:7.1f}/{res['bwd_whole']['warm_us']:7.1f}  "
              f"step {res['step_us']:7.1f} us (cold/warm L2){check}", flush=True)


if __name__ == "__main__":
    main()
