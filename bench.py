#!/usr/bin/env python
"""Benchmark of the rasterizer hot path on BASELINE.json's headline configuration.

Workload (config 2 / config 4 of BASELINE.json): 300k synthetic Gaussians (SH degree 3) in the
reference's init cube, NeRF-synthetic Lego camera poses at 800x800, random target images.
One STEP = one training view per rank: forward, fused L1 loss + pixel gradient, backward, the
gradient all-reduce over NCCL when N > 1, and the fused Adam update.  `value` = views/s summed over
ranks with everything resident in HBM; `fwd_bwd_frames_per_s` is the pure forward+backward rate of
the single-view headline; `e2e` drives the same step through the reference-facing Python API with
the target image coming from pinned host memory and the loss read back every step.

The K-step timed loop is repeated REPEATS times (each repetition bracketed by barrier + synchronize and
timed with CUDA events, max over ranks); `value` / `ms_per_step` are the MEDIAN repetition and `repeats`
carries every repetition with min / max.  After the timed phases the run checks -- and fails (rc != 0)
otherwise -- that the replicas of a multi-GPU run are bit-identical, compares one fused-exchange step with
the NCCL all-reduce + replicated Adam on the same gradients, and runs densify events (clone + split +
prune) under view sharding.  `--densify-every K` puts densify events INTO the timed loop (both arms).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nproc-per-node 8 ... bench.py --gpus 8 --steps 20 --warmup 3
    python bench.py --impl reference --steps 3 --warmup 1      # the CPU oracle on all host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "fwd+bwd frames/s at 800² w/ 300k Gaussians; train views/s at 1/2/4/8 B200"
UNIT = "views/s"
N_CAMERAS = 16  # distinct Lego poses cycled through by the steps
REPEATS = 5     # repetitions of the K-step timed loop (median reported, min / max beside it)
# Adam runs for real every step, but with the reference's learning rates (1e-2 on positions, 5e-3 on raw
# scales) and RANDOM targets the synthetic scene inflates within ten steps (D doubles) and every step count
# would measure a different workload.  BOTH arms scale the rates by LR_SCALE so the scene keeps the named
# shape (D ~ 1.6M); the kernels' work per step is unchanged.
LR_SCALE = 1e-4
BASE_LRS = {"lr_pos": 1e-2, "lr_scale": 5e-3, "lr_rot": 5e-3, "lr_sh": 2e-3, "lr_opac": 5e-3}   # config.py:33-41
FINAL_LR_FACTOR, TOTAL_ITERATIONS = 0.01, 7000
# densify events of the post-timing check / of --densify-every: the reference's thresholds, from iteration 0
# (percent_dense and the cull threshold are set so that the synthetic scene -- scales 0.004..0.02, opacities
# 0.05..0.95 -- has clone AND split AND prune candidates; the gradient threshold is the reference's)
DENSIFY_CFG = {"densify_from_iter": 0, "densify_grad_threshold": 0.0002, "percent_dense": 0.003,
               "cull_opacity_threshold": 0.06}
IT0 = 1                   # first iteration number of every loop (iteration 0 resets all opacities, quirk G6)
FLUSH_BYTES = 256 << 20   # > 126 MB L2: written between the iterations of the per-stage table


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--cull", type=int, default=1)
    ap.add_argument("--tile-sort", type=int, default=3,
                    help="A/B: 3 (default) one-pass bucket sort by depth, 0 bitonic per tile, 1 per-tile radix sort, "
                         "2 radix when the longest list > 2048")
    ap.add_argument("--bwd-reduce", type=int, default=2, help="A/B: 0 shuffle butterfly, 1 / 2 tensor-core moments")
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "peers", "multimem", "hybrid"])
    ap.add_argument("--bwd-packed", type=int, default=1,
                    help="A/B: 1 = backward tile kernel accumulates into packed records with vector REDs, 0 = nine scalar REDs")
    ap.add_argument("--opt", action="append", default=[], help="A/B: gsb_set_option name=value (repeatable)")
    ap.add_argument("--overlap-sh", type=int, default=1,
                    help="A/B (fused exchange modes): 1 = the SH part of the exchange runs on a side stream beside the next "
                         "step's geometry preprocess and binning (both phases start behind the opening barrier), 2 = the "
                         "same with the SH phase started behind the first phase's barrier, 0 = one exchange between two barriers")
    ap.add_argument("--sh-compact", type=int, default=1,
                    help="A/B (peers exchange): 1 = SH gradients cross NVLink as their rank-1 factors, 0 = in full")
    ap.add_argument("--densify-every", type=int, default=0,
                    help="K > 0: densify / prune (train.py:351-713) every K steps INSIDE the timed loops of both arms "
                         "(the Gaussian count then changes over the run); 0 = fixed N in the timed loops, densify "
                         "events run and are checked after them")
    ap.add_argument("--repeats", type=int, default=REPEATS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-stages", action="store_true")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def nvlink_bytes(index):
    """(tx, rx) data bytes of all NVLinks of one GPU from NVML's throughput counters, or None where the driver does not
    expose them (nvidia-smi nvlink -gt d prints N/A on this pool)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        vals = pynvml.nvmlDeviceGetFieldValues(h, [(pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, 0xFFFFFFFF),
                                                   (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, 0xFFFFFFFF)])
        out = []
        for v in vals:
            if v.nvmlReturn != 0:
                return None
            out.append(int(v.value.ullVal) * 1024)      # KiB
        return tuple(out)
    except Exception:
        return None


def make_scene(cfg_name):
    import gsb200  # noqa: F401
    from gsb200 import scene
    from gsb200.utils.camera_utils import load_nerf_cameras
    n, w, h, smin, smax = scene.CONFIGS[cfg_name]
    params, _cam, _ = scene.synthetic_scene(n, w, h, smin, smax, seed=42, with_target=False)
    cams = load_nerf_cameras(w, h)[:N_CAMERAS]
    rng = np.random.default_rng(4242)
    targets = [rng.uniform(0, 1, (h, w, 3)).astype(np.float32) for _ in range(N_CAMERAS)]
    return params, cams, targets, (n, w, h)


def workload_config(cfg_name, dims, densify_every):
    """The `config` object of the JSON line: the workload only, identical in both arms."""
    n, w, h = dims
    return {
        "workload": f"{cfg_name}: {n} synthetic Gaussians (SH degree 3), {w}x{h}, {N_CAMERAS} Lego train poses, random "
                    "targets; one train view per rank per step = forward + L1 loss/gradient + backward + gradient sum "
                    "over the ranks + Adam",
        "views_per_rank_per_step": 1,
        "learning_rates": f"reference values x {LR_SCALE:g} with the reference's exponential schedule (keeps the synthetic "
                          "scene at the named shape)",
        "densify": (f"every {densify_every} steps inside the timed loop" if densify_every > 0
                    else "off inside the timed loop (fixed N)"),
        "l2": "per-step working set ~0.5 GB (params, grads, Adam state, binning buffers) > 126 MB L2, no explicit flush "
              "between steps; the per-stage table flushes L2 (256 MB write) before every timed stage call",
    }


def densify_config(densify_every):
    cfg = dict(DENSIFY_CFG)
    cfg["densification_interval"] = max(int(densify_every), 1)
    return cfg


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle (restatement of the reference) on the host cores
# ------------------------------------------------------------------------------------------------
def oracle_step(O, scene, state, cam, target, it, densify_every=0):
    """One iteration of train.py:926-1064 in the oracle, with the benchmark's learning rates."""
    P, n = state["params"], state["num_points"]
    kw = scene.render_kwargs(P, cam)
    img, _, buf = O.render_gaussians(**kw)
    dpix = O.compute_image_gradients(img, target, lambda_dssim=0)
    g = O.backward(**scene.backward_kwargs(P, cam, buf, dpix))
    state["grads"] = {"positions": g["dL_dmean3D"], "scales": g["dL_dscale"], "rotations": g["dL_drot"],
                      "opacities": g["dL_dopacity"], "shs": g["dL_dshs"]}
    lr = {k: O.get_lr(v * LR_SCALE, FINAL_LR_FACTOR, it, TOTAL_ITERATIONS) for k, v in BASE_LRS.items()}
    O.adam_update(state["grads"], P, state["adam_m"], state["adam_v"], n, lr["lr_pos"], lr["lr_scale"], lr["lr_rot"],
                  lr["lr_opac"], lr["lr_sh"], 0.9, 0.999, 1e-8, it)
    if densify_every > 0:       # train.py:1060: called every iteration, decides by itself
        O.densification_and_pruning(state, it, densify_config(densify_every))


def run_oracle(cfg_name, steps, warmup, threads=None, densify_every=0):
    import gsb200  # noqa: F401
    from gsb200 import scene
    from gsb200.utils.camera_utils import scene_extent
    import oracle as O
    O.build()
    cores = threads or O.max_threads()
    O.set_threads(cores)
    params, cams, targets, dims = make_scene(cfg_name)
    n = dims[0]
    state = {"params": {k: np.array(v, copy=True) for k, v in params.items()}, "grads": O.zeros_like_params(n),
             "adam_m": O.zeros_like_params(n), "adam_v": O.zeros_like_params(n), "num_points": n,
             "scene_extent": scene_extent(cams, 1.0)}
    for it in range(IT0, IT0 + warmup):
        oracle_step(O, scene, state, cams[it % N_CAMERAS], targets[it % N_CAMERAS], it, densify_every)
    t0 = time.perf_counter()
    for it in range(IT0 + warmup, IT0 + warmup + steps):
        oracle_step(O, scene, state, cams[it % N_CAMERAS], targets[it % N_CAMERAS], it, densify_every)
    dt = time.perf_counter() - t0
    O.set_threads(1)
    return steps / dt, dt / steps * 1e3, cores, dims


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vps, ms, cores, dims = run_oracle(args.config, args.steps, args.warmup, densify_every=args.densify_every)
    sample = (f"{args.steps} full {args.config} train views (forward + L1 gradient + backward + Adam, N={dims[0]}, "
              f"{dims[1]}x{dims[2]}), CPU restatement of the reference (Warp is not installable), {cores} host threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": vps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.config, dims, args.densify_every),
        "cpu_baseline": {"value": vps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "host_cores": os.cpu_count()},
        "e2e": {"value": vps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def stage_table(torch, T, cam_index, target, iters=10):
    """Per-stage device times (CUDA events on the launch stream, L2 flushed before every timed call)
    through the stage-level C ABI, the algorithmic bytes of SURVEY.md 8d for each stage, and the work
    counters K_fwd / K_bwd of the rendered frame."""
    import ctypes as C
    from gsb200 import _lib, train
    L, ctx, p = _lib.lib(), T.ctx, _lib.ptr
    s = lambda: _lib.stream_ptr(ctx.device_index)  # noqa: E731
    fb = T.forward(cam_index)
    T.loss_and_pixel_gradients(fb, target)
    N, D, Pn, Tg = T.num_points, fb.num_rendered, fb.W * fb.H, fb.ranges.shape[0]
    frame, P = T.frames[cam_index], T.params
    dev = T.device
    tiles = torch.empty(N, dtype=torch.int32, device=dev)
    keys = torch.empty(D, dtype=torch.int64, device=dev)
    vals = torch.empty(D, dtype=torch.int32, device=dev)
    keys_s = torch.empty(D, dtype=torch.int64, device=dev)
    vals_s = torch.empty(D, dtype=torch.int32, device=dev)
    tk, tv = torch.empty_like(keys), torch.empty_like(vals)
    g = T.grads
    bits = 32 + max(1, int(np.ceil(np.log2(Tg))))
    Dh = C.c_int64(0)
    # Adam runs on copies: the trainer's parameters must stay what the other ranks hold
    FG = train.FlatGaussians
    aP, aG, aM, aV = (FG(N, dev, fill=None) for _ in range(4))
    aP.flat.copy_(T.params.flat)
    aG.flat.copy_(T.grads.flat)
    aM.flat.uniform_(-1e-6, 1e-6)
    aV.flat.uniform_(0, 1e-10)
    lr = T.learning_rates(1000)

    def adam():
        from gsb200 import optimizer as gopt
        gopt.adam_update(aG["positions"], aG["scales"], aG["rotations"], aG["opacities"], aG["shs"], N, lr["lr_pos"],
                         lr["lr_scale"], lr["lr_rot"], lr["lr_opac"], lr["lr_sh"], 0.9, 0.999, 1e-8, 1000,
                         aP["positions"], aP["scales"], aP["rotations"], aP["opacities"], aP["shs"],
                         aM["positions"], aM["scales"], aM["rotations"], aM["opacities"], aM["shs"],
                         aV["positions"], aV["scales"], aV["rotations"], aV["opacities"], aV["shs"])
        return 0

    def stage_sort_inputs():
        keys_s.copy_(keys)
        vals_s.copy_(vals)

    # name -> (untimed preparation or None, timed call, algorithmic bytes)
    stages = {
        "preprocess": (None, lambda: L.gsb_preprocess(ctx.h, s(), C.byref(frame), N, p(P["positions"]), p(P["scales"]),
                                                      p(P["rotations"]), p(P["opacities"]), p(P["shs"]), p(fb.radii), p(fb.xy),
                                                      p(fb.depths), p(fb.cov3Ds), p(fb.colors), p(fb.conic_opacity), p(tiles),
                                                      p(fb.clamped_state)), 320 * N),
        "scan": (None, lambda: L.gsb_scan_tiles(ctx.h, s(), N, p(tiles), p(fb.point_offsets), None), 8 * N),
        "duplicate": (None, lambda: L.gsb_duplicate_with_keys(ctx.h, s(), fb.W, fb.H, N, p(fb.xy), p(fb.depths),
                                                              p(fb.point_offsets), p(fb.radii), D, p(keys), p(vals)),
                      20 * N + 12 * D),
        "sort": (stage_sort_inputs, lambda: L.gsb_sort_pairs64(ctx.h, s(), p(keys_s), p(vals_s), p(tk), p(tv), D, 0, bits),
                 24 * D),
        "bin_by_tile": (None, lambda: L.gsb_bin_by_tile(ctx.h, s(), fb.W, fb.H, N, p(fb.xy), p(fb.depths), p(fb.radii),
                                                        p(fb.point_offsets), p(fb.point_list), fb.capacity, p(fb.ranges),
                                                        C.byref(Dh), None), 20 * N + 20 * D + 8 * Tg),
        "tile_ranges": (None, lambda: L.gsb_tile_ranges(ctx.h, s(), D, p(keys_s), Tg, p(fb.ranges)), 8 * D + 8 * Tg),
        "blend_forward": (None, lambda: L.gsb_blend_forward(ctx.h, s(), C.byref(frame), p(fb.ranges), p(fb.point_list),
                                                            p(fb.xy), p(fb.colors), p(fb.conic_opacity), p(fb.depths),
                                                            p(fb.image), p(fb.depth), p(fb.final_T), p(fb.n_contrib),
                                                            p(fb.block_masks)), 44 * D + 24 * Pn),
        "l1_loss_grad": (None, lambda: L.gsb_l1_loss_grad(ctx.h, s(), 3 * Pn, p(fb.image), p(target), 1.0 / (3 * Pn),
                                                          p(fb.dpix), p(fb.loss_sum)), 36 * Pn),
        "blend_backward": (None, lambda: L.gsb_blend_backward(ctx.h, s(), C.byref(frame), N, p(fb.ranges), p(fb.point_list),
                                                              p(fb.xy), p(fb.conic_opacity), p(fb.colors), p(fb.final_T),
                                                              p(fb.n_contrib), p(fb.dpix), p(fb.dL_dmean2D), p(fb.dL_dconic),
                                                              p(g["opacities"]), p(fb.dL_dcolor), p(fb.block_masks)),
                           40 * D + 20 * Pn + 44 * N),
        "preprocess_backward": (None, lambda: L.gsb_preprocess_backward(
            ctx.h, s(), C.byref(frame), N, p(P["positions"]), p(fb.radii), p(P["shs"]), p(P["scales"]), p(P["rotations"]),
            p(fb.cov3Ds), p(fb.clamped_state), p(fb.dL_dmean2D), p(fb.dL_dconic), p(fb.dL_dcolor), p(g["positions"]),
            p(g["shs"]), p(g["scales"]), p(g["rotations"]), None), 544 * N),
        "adam": (None, adam, 1652 * N),
    }
    # the two whole operators as the trainer calls them (gsb_forward fuses the binning's counting pass
    # into preprocess and hides the scan behind the read-back of D, so it is faster than its stages)
    stages["forward_whole"] = (None, lambda: (T.forward(cam_index), 0)[1], 348 * N + 88 * D + 24 * Pn + 8 * Tg)
    stages["backward_whole"] = (None, lambda: (T.backward(cam_index, fb, T.grads), 0)[1], 588 * N + 40 * D + 20 * Pn)
    flush = torch.empty(FLUSH_BYTES // 4, dtype=torch.float32, device=dev)
    out = {}
    for name, (prep, fn, nbytes) in stages.items():
        for _ in range(2):
            if prep:
                prep()
            ctx.check(fn())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        times = []
        for i in range(iters):
            if prep:
                prep()
            flush.fill_(float(i))          # 256 MB written: nothing of the stage's inputs is left in L2
            e0.record()
            ctx.check(fn())
            e1.record()
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        out[name] = {"ms": round(ms, 4), "ms_min": round(min(times), 4), "ms_max": round(max(times), 4),
                     "alg_bytes": int(nbytes), "gbps": round(nbytes / (ms * 1e-3) / 1e9, 1)}
    out["sort"]["note"] = ("the stable 64-bit radix sort entry point gsb_sort_pairs64 (a real multi-pass radix moves "
                           "several times the 24 B/pair counted here); gsb_forward bins with bin_by_tile instead of "
                           "duplicate + sort + tile_ranges")
    # work counters of SURVEY 8d on this frame (restore the frame first: the stages above re-ran parts of it)
    fb = T.forward(cam_index)
    counters = torch.zeros(7, dtype=torch.int64, device=dev)
    ctx.check(L.gsb_selftest_work_counters(ctx.h, s(), C.byref(frame), p(fb.ranges), p(fb.point_list), p(fb.xy),
                                           p(fb.conic_opacity), p(fb.n_contrib), p(counters)))
    torch.cuda.synchronize()
    names = ("K_fwd", "pairs_blended", "K_bwd", "bwd_pairs_evaluated", "bwd_decision_mismatch", "threshold_false_skips",
             "pairs_near_threshold")
    work = dict(zip(names, [int(x) for x in counters.cpu()]))
    return out, (N, D, Pn, Tg), work


def params_checksum(torch, flat):
    """Two 64-bit checksums of a float32 buffer's BITS (order-sensitive): equal on two ranks <=> identical replicas
    (up to a 2^-64-ish collision)."""
    bits = flat.view(torch.int32).to(torch.int64)
    idx = torch.arange(bits.numel(), device=flat.device, dtype=torch.int64)
    return torch.stack([bits.sum(), (bits * (idx % 1000003 + 1)).sum(), torch.tensor(flat.numel(), device=flat.device)])


def replicas_identical(torch, dist, T):
    """All ranks hold the same num_points and the same parameter bits."""
    T.join_exchange()
    c = params_checksum(torch, T.params.flat)
    got = [torch.empty_like(c) for _ in range(T.world_size)]
    dist.all_gather(got, c)
    return all(bool(torch.equal(g, got[0])) for g in got)


def peers_vs_nccl_step(torch, dist, T, it, cam_index):
    """ONE step on the same per-rank gradients through the trainer's exchange (fused peer kernel) and through
    NCCL all-reduce + the replicated Adam kernel, from the same parameters and moments: relative difference of
    the resulting parameters (0 when the summation order coincides, as it does at two ranks)."""
    from gsb200 import train
    n, dev = T.num_points, T.device
    fb = T.forward(cam_index)
    T.loss_and_pixel_gradients(fb, T.targets[cam_index])
    T._compact_step = False
    T.backward(cam_index, fb, T.grads)                     # this rank's gradients, full form
    m_full, v_full = T.gathered_moments()
    ref = {"params": train.FlatGaussians(n, dev, fill=None), "grads": train.FlatGaussians(n, dev, fill=None),
           "adam_m": train.FlatGaussians(n, dev, fill=None), "adam_v": train.FlatGaussians(n, dev, fill=None)}
    ref["params"].flat.copy_(T.params.flat)
    ref["grads"].flat.copy_(T.grads.flat)
    ref["adam_m"].flat.copy_(m_full)
    ref["adam_v"].flat.copy_(v_full)
    torch.cuda.synchronize()
    dist.barrier()
    T.exchange_and_step(it, compact=False)                 # the trainer's own exchange mode
    dist.all_reduce(ref["grads"].flat, op=dist.ReduceOp.SUM)
    saved = (T.params, T.grads, T.adam_m, T.adam_v)
    T.params, T.grads, T.adam_m, T.adam_v = ref["params"], ref["grads"], ref["adam_m"], ref["adam_v"]
    try:
        T.optimizer_step(it)                               # the single-GPU Adam kernel on the all-reduced gradients
    finally:
        T.params, T.grads, T.adam_m, T.adam_v = saved
    torch.cuda.synchronize()
    a, b = T.params.flat.double(), ref["params"].flat.double()
    return float((a - b).norm() / b.norm())


def densify_check(torch, dist, T, it0, batch, events=2):
    """Densify / prune under view sharding (BASELINE config 4): `events` densify events two steps apart with the
    reference's thresholds; num_points and the parameter bits must stay identical on all ranks."""
    saved = dict(T.config)
    T.config.update(densify_config(2))
    n_before, hist, t0 = T.num_points, [], time.perf_counter()
    it = it0 + (it0 % 2)                       # start on an even iteration: events at it, it + 2, ...
    ok = True
    for k in range(2 * events):
        T.train_step(it + k, batch(it + k), densify=True)
        hist.append(T.num_points)
        if dist is not None:
            ok = ok and replicas_identical(torch, dist, T)
    torch.cuda.synchronize()
    T.config.clear()
    T.config.update(saved)
    return {"events": events, "steps": 2 * events, "num_points_before": n_before, "num_points_per_step": hist,
            "replicas_identical": ok if dist is not None else None, "wall_ms": round((time.perf_counter() - t0) * 1e3, 1),
            "thresholds": f"{DENSIFY_CFG}, every 2 steps"}


def ours(args):
    import torch
    import gsb200  # noqa: F401
    from gsb200 import _lib, backward as gb, forward as gf, loss as gl, optimizer as gopt, scene, train

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # keep stdout to the one JSON line: some images export NCCL_DEBUG=VERSION, which prints there
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    params, cams, targets, (n, w, h) = make_scene(args.config)
    lrs = {k: v * LR_SCALE for k, v in BASE_LRS.items()}
    lrs["final_lr_factor"] = FINAL_LR_FACTOR
    cfg = {"num_iterations": TOTAL_ITERATIONS, "lr_scheduler_config": lrs}
    densify_on = args.densify_every > 0
    if densify_on:
        cfg.update(densify_config(args.densify_every))
    T = train.Trainer(cams, targets=targets, params=params, rank=rank, world_size=world, exchange=args.exchange,
                      sh_compact=bool(args.sh_compact), overlap_sh=args.overlap_sh, config=cfg)
    T.ctx.set_option("blend_cull", args.cull)
    T.ctx.set_option("bwd_reduce", args.bwd_reduce)
    T.ctx.set_option("tile_sort", args.tile_sort)
    T.ctx.set_option("bwd_packed", args.bwd_packed)
    for kv in args.opt:                      # any other A/B option of gsb_set_option, e.g. --opt speculate=0 --opt pdl=0
        name, value = kv.split("=")
        T.ctx.set_option(name, int(value))

    def batch(it):   # one view per rank per step, cycling through the poses
        return [(it * world + r) % N_CAMERAS for r in range(world)]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    K, W, R = args.steps, max(args.warmup, 0), max(args.repeats, 1)
    # ---- value: resident inputs, device-timed; the K-step loop R times, median reported ------------
    it = IT0
    for _ in range(W):
        T.train_step(it, batch(it), densify=densify_on)
        it += 1
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    rep_ms, rep_launches = [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nvl0 = nvlink_bytes(local) if (rank == 0 and world > 1) else None
    for _ in range(R):
        barrier()
        launches0 = T.ctx.launches
        e0.record()
        for _ in range(K):
            T.train_step(it, batch(it), densify=densify_on)
            it += 1
        T.join_exchange()     # overlap_sh: the last step's SH phase (side stream) belongs to the timed region
        e1.record()
        barrier()
        rep_ms.append(max_over_ranks(e0.elapsed_time(e1)))
        rep_launches.append(T.ctx.launches - launches0)
    nvl1 = nvlink_bytes(local) if (rank == 0 and world > 1) else None
    nvlink = None
    if nvl0 is not None and nvl1 is not None:
        nvlink = {"tx_bytes_per_step": (nvl1[0] - nvl0[0]) / (R * K), "rx_bytes_per_step": (nvl1[1] - nvl0[1]) / (R * K),
                  "source": "NVML NVLINK_THROUGHPUT_DATA_TX/RX of rank 0's GPU, all links, over the timed repetitions"}
    order = sorted(range(R), key=lambda i: rep_ms[i])
    mid = order[(R - 1) // 2]                   # the median repetition (lower median for even R)
    ms_total, launches = rep_ms[mid], rep_launches[mid]
    value = K * world / (ms_total * 1e-3)
    repeats = {"R": R, "steps_each": K, "ms_per_step": [round(x / K, 5) for x in rep_ms],
               "median": round(ms_total / K, 5), "min": round(min(rep_ms) / K, 5), "max": round(max(rep_ms) / K, 5),
               "views_per_s_min": K * world / (max(rep_ms) * 1e-3), "views_per_s_max": K * world / (min(rep_ms) * 1e-3)}
    num_points_timed = T.num_points

    # ---- where the exchange + Adam part of the step goes (10 extra, untimed steps with CUDA events) -----
    # device time this rank spends between the end of its backward and the end of the exchange + Adam
    # (incl. waiting for the slowest rank at the first barrier); max over ranks of the per-step median
    T.exchange_events, T.exchange_parts = [], []
    for _ in range(10):
        T.train_step(it, batch(it), densify=False)
        it += 1
    T.join_exchange()
    barrier()
    # (medians: the first of these steps can wait a long time for a rank that is still reading its NVML counters)
    exchange_ms = max_over_ranks(float(np.median([a.elapsed_time(b) for a, b in T.exchange_events])))
    parts = None
    if T.exchange_parts:
        arr = np.median(np.array([[e[i].elapsed_time(e[i + 1]) for i in range(3)] for e in T.exchange_parts]), axis=0)
        parts = {"wait_and_barrier_ms": round(max_over_ranks(float(arr[0])), 4),
                 "fused_kernel_ms": round(max_over_ranks(float(arr[1])), 4),
                 "final_barrier_ms": round(max_over_ranks(float(arr[2])), 4)}
    T.exchange_events = None

    # ---- pure forward+backward of the single-view headline (camera 0) ---------------------------
    tgt0 = T.targets[0]
    for _ in range(3):
        fb = T.forward(0)
        T.loss_and_pixel_gradients(fb, tgt0)
        T.backward(0, fb, T.grads)
    fb_ms = []
    for _ in range(R):
        barrier()
        e0.record()
        for _ in range(K):
            fb = T.forward(0)
            T.loss_and_pixel_gradients(fb, tgt0)
            T.backward(0, fb, T.grads)
        e1.record()
        barrier()
        fb_ms.append(max_over_ranks(e0.elapsed_time(e1)))
    fwd_bwd_fps = K / (float(np.median(fb_ms)) * 1e-3)
    num_rendered = fb.num_rendered

    # ---- e2e: reference-facing Python API, host target in, loss out, every step --------------------
    e2e = None
    if not args.no_e2e:
        pinned = [torch.from_numpy(t).pin_memory() for t in targets]
        copy_stream = torch.cuda.Stream(device=dev)
        tgt_dev = [torch.empty((h, w, 3), dtype=torch.float32, device=dev) for _ in range(2)]
        tgt_ready = [torch.cuda.Event() for _ in range(2)]
        bg = np.zeros(3, dtype=np.float32)

        def e2e_step(it, pending=None):
            P, G, M, V = T.params, T.grads, T.adam_m, T.adam_v
            ci = batch(it)[rank]
            cam = cams[ci]
            # H2D of the step's input from pinned memory, on a copy stream, issued FIRST: the forward does not read
            # the target, so the 7.7 MB transfer (150-300 us over PCIe, depending on the box) runs beside the tail
            # of the previous step and this step's forward; the loss kernel waits for it.  Two device buffers,
            # reused: buffer it & 1 was last read by the loss kernel of step it - 2, and the forward call of step
            # it - 1 has returned, i.e. its host wait (behind preprocess of it - 1 in the stream, hence behind all
            # of step it - 2) is over.
            main = torch.cuda.current_stream()
            tgt, ready = tgt_dev[it & 1], tgt_ready[it & 1]
            with torch.cuda.stream(copy_stream):
                tgt.copy_(pinned[ci], non_blocking=True)
                ready.record(copy_stream)
            img, _depth, buf = gf.render_gaussians(**scene.render_kwargs(P.as_dict(), cam, background=bg))
            if pending is not None:
                consume(pending)                     # the previous step's loss (its copy finished long ago)
            main.wait_event(ready)
            # D2H of the step's result, every step: the loss kernel writes its sum STRAIGHT into pinned host memory
            # (out_sum may be a pinned host tensor: 8 bytes over PCIe, no copy operation in the stream -- a D2H
            # cudaMemcpyAsync of 8 bytes costs the stream ~10 us, tools/e2e_variants.py); the host consumes the
            # value one step later (after the next forward call), behind an event recorded at the end of this step,
            # so preparing the next step overlaps this step's Adam.  All K values are read before the clock stops.
            slot = it & 1
            _, dpix = gl.l1_loss_and_gradients(img, tgt, 0.0, out_sum=loss_host[slot])
            if world > 1:
                # multi-GPU: the gradients land directly in the symmetric flat buffer (backward's `out`),
                # and the public exchange step (fused NVLink reduction + Adam + parameter broadcast, or
                # NCCL all-reduce + Adam) replaces adam_update -- the reference has no multi-GPU API
                gb.backward(**scene.backward_kwargs(P.as_dict(), cam, buf, dpix, background=bg),
                            out={"dL_dmean3D": G["positions"], "dL_dscale": G["scales"], "dL_drot": G["rotations"],
                                 "dL_dopacity": G["opacities"], "dL_dshs": G["shs"]}, sh_compact=T.sh_compact)
                T.exchange_and_step(it, compact=T.sh_compact)
            else:
                g = gb.backward(**scene.backward_kwargs(P.as_dict(), cam, buf, dpix, background=bg))
                lr = T.learning_rates(it)
                gopt.adam_update(g["dL_dmean3D"], g["dL_dscale"], g["dL_drot"], g["dL_dopacity"], g["dL_dshs"], T.num_points,
                                 lr["lr_pos"], lr["lr_scale"], lr["lr_rot"], lr["lr_opac"], lr["lr_sh"], 0.9, 0.999, 1e-8,
                                 it, P["positions"], P["scales"], P["rotations"], P["opacities"], P["shs"],
                                 M["positions"], M["scales"], M["rotations"], M["opacities"], M["shs"],
                                 V["positions"], V["scales"], V["rotations"], V["opacities"], V["shs"])
            loss_ready[slot].record(main)
            return slot

        def consume(slot):
            loss_ready[slot].synchronize()
            losses.append(float(loss_host[slot].item()) / (3 * h * w))

        loss_host = [torch.zeros(1, dtype=torch.float64).pin_memory() for _ in range(2)]
        loss_ready = [torch.cuda.Event() for _ in range(2)]
        losses = []
        for _ in range(max(W, 1)):
            consume(e2e_step(it))
            it += 1
        e2e_dt = []
        for _ in range(R):
            barrier()
            losses.clear()
            t0 = time.perf_counter()
            pending = None
            for _ in range(K):
                pending = e2e_step(it, pending)
                it += 1
            consume(pending)
            torch.cuda.synchronize()
            e2e_dt.append(max_over_ranks(time.perf_counter() - t0))
            assert len(losses) == K and all(np.isfinite(losses)), "every step's loss must have been read back"
        barrier()
        dt = float(np.median(e2e_dt))
        e2e = {"value": K * world / dt, "unit": UNIT, "h2d_bytes_per_step": int(h * w * 3 * 4 + 2 * 64 + 24),
               "d2h_bytes_per_step": 8 + 8, "ms_per_step": dt / K * 1e3,
               "ms_per_step_min": min(e2e_dt) / K * 1e3, "ms_per_step_max": max(e2e_dt) / K * 1e3, "repeats": R,
               "api": ("forward.render_gaussians + loss.l1_loss_and_gradients + backward.backward + optimizer.adam_update"
                       if world == 1 else "forward.render_gaussians + loss.l1_loss_and_gradients + backward.backward(out="
                       "flat gradient buffer) + Trainer.exchange_and_step"),
               "loss_readback": "every step the loss kernel writes its sum straight into pinned host memory (8 bytes over "
                                "PCIe, no copy operation); consumed by the host one step later behind an event; all K read "
                                "inside the timed region", "last_loss": losses[-1]}

    # ---- per-stage table + roofline of the dominant kernel (rank 0; the other ranks wait at the next barrier) ----
    stages, roofline, work = None, None, None
    peak, peak_src = peaks()
    if rank == 0 and not args.no_stages:
        stages, (N_, D_, P_, Tg_), work = stage_table(torch, T, 0, tgt0)
    clocks = sampler.stop() if rank == 0 else None   # sampled across all timed loops and the stage table
    if stages is not None:
        top = max((k for k in stages if k not in ("sort", "duplicate", "tile_ranges", "forward_whole", "backward_whole")),
                  key=lambda k: stages[k]["ms"])
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                tj = json.load(f)
            traffic, traffic_src = tj.get(top), tj.get("_source")
        except Exception:
            pass
        st = stages[top]
        # FP32 roofline of the two tile kernels: the reference's loops evaluate K_fwd / K_bwd (pixel, Gaussian)
        # pairs at ~31 / ~70 flops each; peak = 148 SMs x 128 lanes x 2 flop x the SM clock sampled during this run
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        fwd_tf = work["K_fwd"] * 31 / (stages["blend_forward"]["ms"] * 1e-3) / 1e12
        bwd_tf = work["K_bwd"] * 70 / (stages["blend_backward"]["ms"] * 1e-3) / 1e12
        fp32 = {"peak_tflops": round(fp32_peak, 2), "sm_mhz": sm_mhz, "flops_per_pair": {"forward": 31, "backward": 70},
                "blend_forward": {"pairs": work["K_fwd"], "tflops": round(fwd_tf, 2), "frac": round(fwd_tf / fp32_peak, 4),
                                  "blended_pairs_only_tflops": round(work["pairs_blended"] * 31 / (stages["blend_forward"]["ms"] * 1e-3) / 1e12, 3)},
                "blend_backward": {"pairs": work["K_bwd"], "tflops": round(bwd_tf, 2), "frac": round(bwd_tf / fp32_peak, 4),
                                   "blended_pairs_only_tflops": round(work["pairs_blended"] * 70 / (stages["blend_backward"]["ms"] * 1e-3) / 1e12, 3)},
                "note": "reference-equivalent work: the pairs the reference's per-pixel loops iterate (K_fwd, K_bwd, counted on "
                        "this frame by gsb_selftest_work_counters = the oracle's counters) at the survey's flop counts, over the "
                        "kernel time of the stage table; the kernels cull most of those pairs before any arithmetic, so this is "
                        "throughput in the reference's units, not issued FP32 instructions (ncu: profiles/)"}
        roofline = {"bound": "hbm", "kernel": top, "achieved": st["gbps"], "peak": peak, "unit": "GB/s",
                    "frac": round(st["gbps"] / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": peak_src, "alg_bytes_per_launch": st["alg_bytes"], "ms_per_launch": st["ms"],
                    "fp32": fp32, "work_counters": work,
                    "note": "the tile kernels are instruction-issue bound and their working set sits in L2 (DRAM traffic "
                            "below the algorithmic bytes), so the HBM fraction of the kernel with the largest share "
                            "of the step is small by nature; the fp32 object is the roofline that applies to them"}

    # ---- multi-GPU correctness, inside the run the driver records (collective; after all timing) -------------
    checks = None
    if dist is not None:
        barrier()
        same = replicas_identical(torch, dist, T)
        rel = peers_vs_nccl_step(torch, dist, T, it, batch(it)[rank])
        it += 1
        same_after = replicas_identical(torch, dist, T)
        checks = {"replicas_identical": bool(same and same_after), "exchange_vs_nccl_one_step_rel_diff": rel,
                  "exchange_vs_nccl_note": "same per-rank gradients, parameters and moments through Trainer.exchange_and_step "
                                           "and through all_reduce + the single-GPU Adam kernel"}
    densify = densify_check(torch, dist, T, it, batch)
    if checks is not None:
        checks["replicas_identical_after_densify"] = densify["replicas_identical"]

    # ---- CPU baseline: the oracle on the host cores, bounded sample (rank 0, N=1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        vps, ms, cores, _ = run_oracle(args.config, 2, 1, densify_every=args.densify_every)
        cpu = {"value": vps, "unit": UNIT, "cores": cores, "kind": "port", "ms_per_view": ms,
               "sample": f"2 full {args.config} train views after 1 warm-up (forward + L1 gradient + backward + Adam) on "
                         f"the CPU restatement of the reference, {cores} threads of {os.cpu_count()} host CPUs"}

    failed = checks is not None and not (checks["replicas_identical"] and checks["replicas_identical_after_densify"])
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.config, (n, w, h), args.densify_every),
            "repeats": repeats,
            "details": {"exchange": T.exchange + ("+sh_compact (SH gradients cross NVLink as their rank-1 factors: "
                                                   "76 instead of 236 B per Gaussian and peer)" if T.sh_compact else "")
                        + ("+overlap_sh (two phases: 11 of 59 floats between the barriers of the main stream, the SH "
                           "coefficients on a side stream beside the next step's geometry preprocess and binning; "
                           "exchange_plus_adam_ms / exchange_parts then describe the first phase)" if T.overlap_sh else ""),
                        "step": {"nccl": "NCCL all-reduce of 59*N gradient floats + replicated Adam",
                                 "peers": "fused NVLink peer-load gradient reduction/Adam/parameter broadcast kernel",
                                 "multimem": "fused NVSwitch multimem gradient reduction/Adam/parameter broadcast kernel",
                                 "hybrid": "fused kernel: NVLink peer-load gradient reduction, Adam, multimem.st parameter broadcast",
                                 "none": "single GPU: Adam"}[T.exchange],
                        "exchange_plus_adam_ms": round(exchange_ms, 4), "exchange_parts": parts, "views_per_step": world,
                        "nvlink_counters": nvlink,
                        "nvlink_model_bytes_per_direction_per_step": (None if world == 1 else int(
                            (world - 1) / world * ((76 if T.sh_compact else 236) + 236) * n)),
                        "num_rendered_view0": int(num_rendered), "num_points_after_timed_loops": int(num_points_timed),
                        "host_cores": os.cpu_count()},
            "multi_gpu_checks": checks, "densify_check": densify,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "fwd_bwd_frames_per_s": fwd_bwd_fps, "roofline": roofline, "cpu_baseline": cpu, "stages": stages,
        }
        emit(line)
    if dist is not None:
        dist.destroy_process_group()
    if failed:
        raise SystemExit("bench.py: replicas differ between ranks (see multi_gpu_checks in the JSON line)")


_JSON_FD = None


def emit(line: dict):
    """The ONE JSON line of the contract, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    args = parse()
    # Libraries (NCCL prints its version banner, OpenMP warnings, ...) write to file descriptor 1; the
    # contract is one JSON line on stdout, so everything else is routed to stderr.
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
