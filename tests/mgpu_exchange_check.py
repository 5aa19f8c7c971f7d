"""Run under torchrun with >= 2 ranks (one GPU each): the three gradient-exchange modes of the
trainer -- NCCL all-reduce + replicated Adam, fused NVLink peer loads, fused NVSwitch multimem --
must give the same parameters (up to fp32 summation order), identical on every rank; the step loop
with densify / prune keeps num_points and every parameter bit-identical across the ranks; checkpoints
written in the fused modes hold every shard's Adam moments."""
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
    # (compact, gradients, two-phase exchange with the SH phase on a side stream)
    for compact, G, overlap in ((False, G_full, False), (True, G_comp, False), (False, G_full, True), (True, G_comp, True)):
        T.params.flat.copy_(start[0]); T.adam_m.flat.copy_(start[1]); T.adam_v.flat.copy_(start[2])
        torch.cuda.synchronize()
        dist.barrier()                                  # nobody pushes into a buffer that is being restored
        T.grads = G
        T.exchange_and_step(3, compact=compact, overlap=overlap)
        assert (T._sh_event is not None) == overlap
        T.join_exchange()
        torch.cuda.synchronize()
        dist.barrier()
        out[(compact, overlap)] = (T.params.flat.clone(), T.adam_m.flat.clone(), T.adam_v.flat.clone())
    assert not torch.equal(out[(False, False)][0], start[0])
    for key in ((True, False), (False, True), (True, True)):
        for a, b, what in zip(out[(False, False)], out[key], ("parameters", "first moments", "second moments")):
            assert torch.equal(a, b), f"exchange compact={key[0]} overlap={key[1]}: {what} differ from the one-piece full exchange"
    if rank == 0:
        print("compact SH exchange and the two-phase exchange: parameters and moments bit-identical to the full one-piece exchange")


HAS_MULTICAST = False
DENSIFY_CFG = {"num_iterations": 100, "densify_from_iter": 0, "densification_interval": 2,
               "densify_grad_threshold": 0.0002, "percent_dense": 0.01, "cull_opacity_threshold": 0.1,
               "min_valid_points": 100}


def assert_replicas_identical(T, what):
    """num_points and every parameter bit-identical on all ranks."""
    world = dist.get_world_size()
    cnt = torch.tensor([T.num_points], dtype=torch.int64, device=T.device)
    cnts = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    assert all(int(c) == int(cnts[0]) for c in cnts), f"{what}: num_points differ between ranks: {[int(c) for c in cnts]}"
    flat = T.params.flat.clone()
    got = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(got, flat)
    for g in got:
        assert torch.equal(g, got[0]), f"{what}: parameters differ between ranks"


def densify_keeps_replicas_identical(rank, world, cams, targets, params):
    """BASELINE config 4 as north_star names it: the step loop WITH densify / prune under view sharding.  Iterations
    1..6 (0 would reset every opacity, quirk G6) with densify events at 2, 4 and 6 in every exchange mode:
    num_points and all parameters must stay bit-identical across the ranks (reference train.py:351-713 run
    identically everywhere, masks from 398-433), and the Gaussian counts must agree with the NCCL run up to the
    few candidates whose summed gradient norm sits within atomics noise of the threshold."""
    counts = {}
    dmodes = ["nccl", "peers", "peers_full"] + (["hybrid"] if HAS_MULTICAST else [])
    for mode in dmodes:
        T = train.Trainer(cams, targets=targets, params=params, rank=rank, world_size=world,
                          exchange=mode.split("_")[0], sh_compact=(mode in ("peers", "hybrid")), config=dict(DENSIFY_CFG))
        hist = [T.num_points]                # hist[it] = Gaussians after iteration it
        for it in range(1, 7):
            batch = [(it * world + r) % len(cams) for r in range(world)]
            assert T.densify_due(it) == (it in (2, 4, 6))
            T.train_step(it, batch, densify=True)
            torch.cuda.synchronize()
            assert_replicas_identical(T, f"{mode} after iteration {it}")
            hist.append(T.num_points)
        assert hist[1] == params["positions"].shape[0] and hist[2] != hist[1] and hist[4] != hist[3], (mode, hist)
        counts[mode] = hist
        if rank == 0:
            print(f"densify {mode}: num_points per iteration {hist}")
    n0 = params["positions"].shape[0]
    for mode in dmodes[1:]:
        for a, b in zip(counts[mode], counts["nccl"]):
            assert abs(a - b) <= max(8, n0 // 500), (mode, counts[mode], counts["nccl"])

    # The same per-rank gradients through both exchanges, then one densify call: with two ranks a + b is the
    # same float whoever adds it, so num_points AND every parameter must match the NCCL path bit for bit.
    Tn = train.Trainer(cams, targets=targets, params=params, rank=rank, world_size=world, exchange="nccl",
                       config=dict(DENSIFY_CFG))
    Tp = train.Trainer(cams, targets=targets, params=params, rank=rank, world_size=world, exchange="peers",
                       sh_compact=False, config=dict(DENSIFY_CFG))
    ci = rank % len(cams)
    fb = Tp.forward(ci)
    Tp.loss_and_pixel_gradients(fb, Tp.targets[ci])
    Tp.backward(ci, fb, Tp.grads)
    Tn.grads.flat.copy_(Tp.grads.flat)
    torch.cuda.synchronize()
    dist.barrier()
    it = 2
    Tn.exchange_and_step(it, publish_position_grad=True)
    Tp.exchange_and_step(it, publish_position_grad=True)
    torch.cuda.synchronize()
    if world == 2:
        assert torch.equal(Tn.grads["positions"], Tp.grads["positions"]), "published position gradient != all-reduce"
        assert torch.equal(Tn.params.flat, Tp.params.flat)
    else:
        assert torch.allclose(Tn.grads["positions"], Tp.grads["positions"], rtol=1e-5, atol=1e-9)
    ln, lp = Tn.densification_and_pruning(it), Tp.densification_and_pruning(it)
    assert ln["cloned"] > 0 and ln["split"] > 0 and ln["pruned"] > 0, ln
    if world == 2:
        assert ln == lp and Tn.num_points == Tp.num_points, (ln, lp)
        assert torch.equal(Tn.params.flat, Tp.params.flat)
    assert_replicas_identical(Tp, "peers after densify on shared gradients")
    if rank == 0:
        print(f"densify on shared gradients: {lp} -> {Tp.num_points} Gaussians, identical to the NCCL path")


def checkpoint_gathers_moment_shards(rank, world, cams, targets, params, tmp):
    """In the fused modes each rank holds the Adam moments of its shard only: the checkpoint must contain all
    of them (compared with an NCCL-mode run from the same gradients is not bit-stable, so the check is
    structural: every shard's moments are non-zero in the file, and a resumed trainer continues bit-identically
    to the uninterrupted one on every rank)."""
    T = train.Trainer(cams, targets=targets, params=params, rank=rank, world_size=world, exchange="peers",
                      config={"num_iterations": 100})
    for it in range(2):
        T.train_step(it, [(it * world + r) % len(cams) for r in range(world)], densify=False)
    ckpt = T.save_checkpoint(tmp, 1)
    st = np.load(os.path.join(ckpt, "state.npz"))
    n = T.num_points
    for r in range(world):
        g0, g1 = train.shard_range(n, r, world)
        assert np.abs(st["adam_m"][3 * g0: 3 * g1]).max() > 0, f"moments of rank {r}'s shard missing in the checkpoint"
    T2 = train.Trainer(cams, targets=targets, num_points=64, rank=rank, world_size=world, exchange="peers",
                       config={"num_iterations": 100})
    nxt = T2.load_checkpoint(ckpt)
    assert nxt == 2 and T2.num_points == n and torch.equal(T2.params.flat, T.params.flat)
    # continue both from the same per-rank gradients: identical parameters afterwards
    batch = [(2 * world + r) % len(cams) for r in range(world)]
    ci = batch[rank]
    fb = T.forward(ci)
    T.loss_and_pixel_gradients(fb, T.targets[ci])
    T._compact_step = False
    T.backward(ci, fb, T.grads)
    T2.grads.flat.copy_(T.grads.flat)
    torch.cuda.synchronize()
    dist.barrier()
    T.exchange_and_step(nxt, compact=False)
    T2.exchange_and_step(nxt, compact=False)
    torch.cuda.synchronize()
    assert torch.equal(T2.params.flat, T.params.flat), "resumed run diverges: Adam moments were not restored"
    if rank == 0:
        print("checkpoint: moment shards gathered, resume bit-identical")


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
    global HAS_MULTICAST
    HAS_MULTICAST = int(getattr(probe.params.symm, "multicast_ptr", 0) or 0) != 0
    if HAS_MULTICAST:
        modes += ["multimem", "hybrid", "hybrid_full"]   # hybrid: peer-load gradients + multimem.st parameters
    for mode in modes:
        compact = mode in ("peers", "hybrid")
        T = train.Trainer(cams, targets=targets, params=params, rank=rank, world_size=world,
                          exchange=mode.split("_")[0], sh_compact=compact, config={"num_iterations": 100})
        assert T.exchange == mode.split("_")[0] and T.sh_compact == compact
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
    densify_keeps_replicas_identical(rank, world, cams, targets, params)
    import tempfile
    tmp = [tempfile.mkdtemp(prefix="gsb_ckpt_") if rank == 0 else None]
    dist.broadcast_object_list(tmp, src=0)
    checkpoint_gathers_moment_shards(rank, world, cams, targets, params, tmp[0])
    if rank == 0:
        print("MGPU_EXCHANGE_OK modes=" + ",".join(modes))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
