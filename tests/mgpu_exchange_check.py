"""Run under torchrun with >= 2 ranks (one GPU each): the three gradient-exchange modes of the
trainer -- NCCL all-reduce + replicated Adam, fused NVLink peer loads, fused NVSwitch multimem --
must give the same parameters (up to fp32 summation order), identical on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gsb200  # noqa: E402,F401
from gsb200 import scene, train  # noqa: E402
from gsb200.utils.camera_utils import load_nerf_cameras  # noqa: E402


def compact_exchange_is_bit_identical(rank, world, cams, targets, params):
    """Two training runs differ in the last bits (the tile kernel's atomics), so the compact SH exchange
    is compared with the full one on ONE set of per-view gradients: the per-Gaussian backward stage is run
    twice from the same tile-stage outputs (full SH gradient / rank-1 factors), the fused exchange + Adam
    once on each, from the same parameters and moments.  Parameters must match bit for bit on every rank."""
    import ctypes as C
    from gsb200 import _lib
    T = train.Trainer(cams, targets=targets, params=params, rank=rank, world_size=world, exchange="peers",
                      sh_compact=True, config={"num_iterations": 100})
    ci = rank % len(cams)
    fb = T.forward(ci)
    T.loss_and_pixel_gradients(fb, T.targets[ci])
    T._compact_step = False
    T.backward(ci, fb, T.grads)                       # full form; leaves the tile-stage outputs in fb
    G_full, G_comp = T.grads, T._new_flat(T.num_points)
    G_comp.flat.copy_(G_full.flat)
    P, p = T.params, _lib.ptr
    T.ctx.check(_lib.lib().gsb_preprocess_backward_compact_sh(
        T.ctx.h, _lib.stream_ptr(T.ctx.device_index), C.byref(T.frames[ci]), T.num_points, p(P["positions"]),
        p(fb.radii), p(P["shs"]), p(P["scales"]), p(P["rotations"]), p(fb.cov3Ds), p(fb.clamped_state),
        p(fb.dL_dmean2D), p(fb.dL_dconic), p(fb.dL_dcolor), p(G_comp["positions"]), p(G_comp["shs"]),
        p(G_comp["scales"]), p(G_comp["rotations"]), p(None)))
    for k in ("positions", "scales", "rotations"):      # the stage is deterministic: same bits as the full call
        assert torch.equal(G_comp[k], G_full[k]), k
    torch.cuda.synchronize()
    dist.barrier()
    start = (T.params.flat.clone(), T.adam_m.flat.clone(), T.adam_v.flat.clone())
    out = {}
    for compact, G in ((False, G_full), (True, G_comp)):
        T.params.flat.copy_(start[0]); T.adam_m.flat.copy_(start[1]); T.adam_v.flat.copy_(start[2])
        torch.cuda.synchronize()
        dist.barrier()                                  # nobody pushes into a buffer that is being restored
        T.grads = G
        T.exchange_and_step(3, compact=compact)
        torch.cuda.synchronize()
        dist.barrier()
        out[compact] = T.params.flat.clone()
    assert not torch.equal(out[False], start[0])
    assert torch.equal(out[False], out[True]), "compact SH exchange differs from the full one"
    if rank == 0:
        print("compact SH exchange: parameters bit-identical to the full exchange")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, w, h, steps = 20001, 160, 112, 3          # n deliberately not a multiple of 4 * world
    params, _, _ = scene.synthetic_scene(n, w, h, 0.01, 0.06, seed=5, with_target=False)
    cams = load_nerf_cameras(w, h)[:8]
    rng = np.random.default_rng(7)
    targets = [rng.uniform(0, 1, (h, w, 3)).astype(np.float32) for _ in cams]
    results = {}
    modes = ["nccl", "peers", "peers_full"]    # "peers" = default (compact SH exchange), "peers_full" = sh_compact off
    probe = train.Trainer(cams, targets=targets, params=params, rank=rank, world_size=world, exchange="auto")
    assert probe.exchange == "peers"
    if int(getattr(probe.params.symm, "multicast_ptr", 0) or 0) != 0:
        modes.append("multimem")
    for mode in modes:
        T = train.Trainer(cams, targets=targets, params=params, rank=rank, world_size=world,
                          exchange=mode.split("_")[0], sh_compact=(mode == "peers"), config={"num_iterations": 100})
        assert T.exchange == mode.split("_")[0] and T.sh_compact == (mode == "peers")
        for it in range(steps):
            T.train_step(it, [(it * world + r) % len(cams) for r in range(world)], densify=False)
        torch.cuda.synchronize()
        results[mode] = T.params.flat.clone()
        # replicas identical on all ranks
        gathered = [torch.empty_like(results[mode]) for _ in range(world)]
        dist.all_gather(gathered, results[mode])
        for g in gathered:
            assert torch.equal(g, gathered[0]), f"{mode}: replicas differ between ranks"
    ref = results["nccl"].double()
    for mode in modes[1:]:
        d = (results[mode].double() - ref).norm() / ref.norm()
        # Adam's m/sqrt(v) is sign-like for near-zero gradients, so different summation orders can flip a
        # few entries by 2*lr; the norm-wise difference stays tiny
        assert d < 2e-3, (mode, float(d))
        if rank == 0:
            print(f"exchange {mode}: rel diff vs nccl = {float(d):.3e}")
    compact_exchange_is_bit_identical(rank, world, cams, targets, params)
    if rank == 0:
        print("MGPU_EXCHANGE_OK modes=" + ",".join(modes))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
