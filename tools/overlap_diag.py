#!/usr/bin/env python
"""Timeline of one multi-GPU training step with and without the two-phase exchange (Trainer(overlap_sh=...)): forward,
loss + backward, the main stream's share of the exchange, the NEXT forward (which waits for the SH phase in front of its
colour kernel) and how long the SH phase runs past the main stream's exchange.  Run under torchrun with >= 2 ranks:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/overlap_diag.py
"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import gsb200
from gsb200 import scene, train
from gsb200.utils.camera_utils import load_nerf_cameras
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, w, h = 300000, 800, 800
import bench
params, cams, targets, _ = bench.make_scene("C2")
lrs = {k: v * bench.LR_SCALE for k, v in bench.BASE_LRS.items()}; lrs["final_lr_factor"] = bench.FINAL_LR_FACTOR
for ovl in (True, False):
    T = train.Trainer(cams, targets=targets, params=params, rank=rank, world_size=world, exchange="auto", overlap_sh=ovl,
                      config={"num_iterations": 7000, "lr_scheduler_config": lrs})
    ev = []
    it = 1
    for s in range(25):
        b = [(it * world + r) % 16 for r in range(world)]
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        T.train_step(it, b, densify=False); it += 1
        e[1].record()                      # end of step on main (after phase 1 + barrier)
        side_done = T._sh_event
        # next forward only (to time it)
        ev.append((e, side_done))
    torch.cuda.synchronize()
    steps = [a[0][0].elapsed_time(a[0][1]) for a in ev[5:]]
    print(f"rank {rank} overlap={ovl}: step(main-stream) median {np.median(steps)*1e3:.1f} us", flush=True)
    # timeline of one step with explicit events
    for rep in range(3):
        b = [(it * world + r) % 16 for r in range(world)]
        mine = rank
        t = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        T.join_exchange(); torch.cuda.synchronize(); dist.barrier()
        t[0].record()
        fb = T.forward(b[mine]); t[1].record()
        T.loss_and_pixel_gradients(fb, T.targets[b[mine]])
        T._compact_step = train.compact_sh_step(T.sh_compact, len(b), world)
        T.backward(b[mine], fb, T.grads); t[2].record()
        T.exchange_and_step(it, overlap=ovl); it += 1
        t[3].record()
        if T._sh_event is not None:
            side_ev = torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(T._side_stream):
                side_ev.record()
        else:
            side_ev = None
        fb = T.forward(b[mine]); t[4].record()
        torch.cuda.synchronize()
        msg = f"rank {rank} ovl={ovl}: fwd {t[0].elapsed_time(t[1])*1e3:.0f} loss+bwd {t[1].elapsed_time(t[2])*1e3:.0f} exchange(main) {t[2].elapsed_time(t[3])*1e3:.0f} next fwd {t[3].elapsed_time(t[4])*1e3:.0f}"
        if side_ev is not None:
            msg += f" side phase after main-exchange-end {t[3].elapsed_time(side_ev)*1e3:.0f}"
        print(msg, flush=True)
    del T
dist.destroy_process_group()
