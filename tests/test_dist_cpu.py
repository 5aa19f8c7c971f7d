"""CPU (gloo, world_size 2): the view-sharded data-parallel step -- contiguous view split, one
all-reduce(SUM) of the flat 59*N gradient buffer, identical Adam on every rank -- equals the
single-process step on the whole batch.  Per-view gradients come from the CPU oracle (this is a
test of the host-side sharding logic, not of the kernels)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("positions", "scales", "rotations", "opacities", "shs")
GKEYS = {"positions": "dL_dmean3D", "scales": "dL_dscale", "rotations": "dL_drot", "opacities": "dL_dopacity",
         "shs": "dL_dshs"}


def _setup_paths():
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)


def _view_grads_flat(O, scene, train, params, cam, target):
    n = params["positions"].shape[0]
    img, _, buf = O.render_gaussians(**scene.render_kwargs(params, cam))
    dpix = O.compute_image_gradients(img, target, lambda_dssim=0)
    g = O.backward(**scene.backward_kwargs(params, cam, buf, dpix))
    offs, total = train.flat_layout(n)
    flat = np.zeros(total, np.float32)
    for k in KEYS:
        v = g[GKEYS[k]].reshape(-1)
        flat[offs[k]:offs[k] + v.size] = v
    return flat


def _worker(rank, world, port, batch, out_dir):
    _setup_paths()
    import gsb200  # noqa: F401
    from gsb200 import scene, train
    from gsb200.utils.camera_utils import load_nerf_cameras
    import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, w, h = 300, 48, 32
    params, _, _ = scene.synthetic_scene(n, w, h, 0.05, 0.3, seed=9, with_target=False)
    cams = load_nerf_cameras(w, h)
    rng = np.random.default_rng(2)
    targets = {ci: rng.uniform(0, 1, (h, w, 3)).astype(np.float32) for ci in sorted(set(batch))}
    mine = train.shard_views(batch, rank, world)
    flat = None
    for ci in mine:                                   # sum over this rank's views
        f = _view_grads_flat(O, scene, train, params, cams[ci], targets[ci])
        flat = f if flat is None else flat + f
    t = torch.from_numpy(flat)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)          # the ONE exchange step
    np.save(os.path.join(out_dir, f"grads_rank{rank}.npy"), t.numpy())
    np.save(os.path.join(out_dir, f"views_rank{rank}.npy"), np.array(mine))
    dist.destroy_process_group()


def test_view_sharded_step_equals_single_process(tmp_path):
    _setup_paths()
    import gsb200  # noqa: F401
    from gsb200 import scene, train
    from gsb200.utils.camera_utils import load_nerf_cameras
    import oracle as O
    O.build()
    batch = [3, 17, 42, 8]
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, batch, str(tmp_path)), nprocs=world, join=True)
    g0 = np.load(tmp_path / "grads_rank0.npy")
    g1 = np.load(tmp_path / "grads_rank1.npy")
    assert np.array_equal(g0, g1)                     # every rank holds the same reduced gradients
    assert np.load(tmp_path / "views_rank0.npy").tolist() == [3, 17]
    assert np.load(tmp_path / "views_rank1.npy").tolist() == [42, 8]
    # single process: sum over the whole batch
    n, w, h = 300, 48, 32
    params, _, _ = scene.synthetic_scene(n, w, h, 0.05, 0.3, seed=9, with_target=False)
    cams = load_nerf_cameras(w, h)
    rng = np.random.default_rng(2)
    targets = {ci: rng.uniform(0, 1, (h, w, 3)).astype(np.float32) for ci in sorted(set(batch))}
    ref = None
    for ci in batch:
        f = _view_grads_flat(O, scene, train, params, cams[ci], targets[ci])
        ref = f if ref is None else ref + f
    assert np.allclose(g0, ref, rtol=1e-5, atol=1e-9)  # float summation order differs (2+2 vs 4)
    assert np.linalg.norm(ref) > 0


def test_shard_views_rejects_ragged_batches():
    _setup_paths()
    import gsb200  # noqa: F401
    from gsb200 import train
    assert train.shard_views([5, 6, 7, 8], 1, 2) == [7, 8]
    assert train.shard_views([9], 0, 1) == [9]
    with pytest.raises(ValueError):
        train.shard_views([1, 2, 3], 0, 2)
    offs, total = train.flat_layout(5)
    assert offs["positions"] == 0 and all(o % 4 == 0 for o in offs.values()) and total >= 59 * 5


def test_compact_sh_exchange_only_with_one_view_per_rank():
    """The rank-1 form of the SH gradient is valid for ONE view: every rank must take the same decision from
    global quantities (batch size, world size)."""
    sys.path.insert(0, ROOT) if ROOT not in sys.path else None
    import gsb200  # noqa: F401
    from gsb200 import train
    assert train.compact_sh_step(True, 4, 4) and train.compact_sh_step(True, 2, 2)
    assert not train.compact_sh_step(True, 8, 4)       # two views per rank: factors do not add
    assert not train.compact_sh_step(True, 1, 1)       # nobody to exchange with
    assert not train.compact_sh_step(False, 4, 4)


def _densify_worker(rank, world, port, out_dir):
    _setup_paths()
    import gsb200  # noqa: F401
    from gsb200 import scene, train
    from gsb200.utils.camera_utils import load_nerf_cameras, scene_extent
    import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, w, h = 600, 48, 32
    params, _, _ = scene.synthetic_scene(n, w, h, 0.01, 0.08, seed=4, with_target=False)
    cams = load_nerf_cameras(w, h)
    target = np.random.default_rng(20 + rank).uniform(0, 1, (h, w, 3)).astype(np.float32)
    flat = _view_grads_flat(O, scene, train, params, cams[5 + 7 * rank], target)   # this rank's view
    np.save(os.path.join(out_dir, f"local_rank{rank}.npy"), flat)
    t = torch.from_numpy(flat.copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    offs, _ = train.flat_layout(n)
    state = {"params": {k: v.copy() for k, v in params.items()}, "grads": O.zeros_like_params(n),
             "adam_m": O.zeros_like_params(n), "adam_v": O.zeros_like_params(n), "num_points": n,
             "scene_extent": scene_extent(cams, 1.0)}
    state["grads"]["positions"] = t.numpy()[offs["positions"]: offs["positions"] + 3 * n].reshape(n, 3).copy()
    cfg = {"densify_from_iter": 0, "densification_interval": 2, "densify_grad_threshold": 2e-5, "percent_dense": 0.01,
           "cull_opacity_threshold": 0.1, "min_valid_points": 10}
    assert train.densify_due({**cfg, "densify_until_iter": 15000}, 2)
    log = O.densification_and_pruning(state, 2, cfg)
    np.savez(os.path.join(out_dir, f"densify_rank{rank}.npz"), num_points=state["num_points"], cloned=log["cloned"],
             split=log["split"], pruned=log["pruned"], **state["params"])
    dist.destroy_process_group()


def test_densify_after_the_exchange_is_identical_on_every_rank(tmp_path):
    """every rank runs the identical densify and stays in lock-step.  That holds exactly when the
    candidate masks are computed from the SUMMED position gradient (train.py:398-433) -- and not when a rank
    uses its own view's gradient, which is what a fused exchange that never publishes the sum would leave."""
    _setup_paths()
    import gsb200  # noqa: F401
    from gsb200 import scene, train
    from gsb200.utils.camera_utils import load_nerf_cameras, scene_extent
    import oracle as O
    O.build()
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_densify_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "densify_rank0.npz"), np.load(tmp_path / "densify_rank1.npz")
    assert int(r0["num_points"]) == int(r1["num_points"]) != 600
    assert int(r0["cloned"]) > 0 and int(r0["split"]) > 0 and int(r0["pruned"]) > 0
    for k in KEYS:
        assert np.array_equal(r0[k], r1[k]), k
    # the failure mode of round 1: masks from the LOCAL gradient diverge
    n, w, h = 600, 48, 32
    params, _, _ = scene.synthetic_scene(n, w, h, 0.01, 0.08, seed=4, with_target=False)
    cams = load_nerf_cameras(w, h)
    offs, _ = train.flat_layout(n)
    counts = []
    for rank in range(world):
        local = np.load(tmp_path / f"local_rank{rank}.npy")
        state = {"params": {k: v.copy() for k, v in params.items()}, "grads": O.zeros_like_params(n),
                 "adam_m": O.zeros_like_params(n), "adam_v": O.zeros_like_params(n), "num_points": n,
                 "scene_extent": scene_extent(cams, 1.0)}
        state["grads"]["positions"] = local[offs["positions"]: offs["positions"] + 3 * n].reshape(n, 3).copy()
        O.densification_and_pruning(state, 2, {"densify_from_iter": 0, "densification_interval": 2,
                                               "densify_grad_threshold": 2e-5, "percent_dense": 0.01,
                                               "cull_opacity_threshold": 0.1, "min_valid_points": 10})
        counts.append(state["num_points"])
    assert counts[0] != counts[1], "local gradients happened to give equal counts: the test scene is too easy"


def test_shard_range_matches_the_kernel_rule():
    """train.shard_range restates adam_step_peers_impl's ownership rule (csrc/optimizer.cu): contiguous, covering,
    boundaries on multiples of four Gaussians (so every segment's shard starts 16-byte aligned)."""
    _setup_paths()
    import gsb200  # noqa: F401
    from gsb200 import train
    for n in (1, 3, 4, 5, 1000, 20001, 300000):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                g0, g1 = train.shard_range(n, r, world)
                assert g0 == prev and g0 % 4 == 0 and g1 >= g0
                prev = g1
            assert prev == n
