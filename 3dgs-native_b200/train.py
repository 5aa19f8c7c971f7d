"""The reference's step loop (train.py:920-1066) on GPU-resident state, plus view-batched data
parallelism (BASELINE config 4).

What the reference does per iteration and what happens here instead:

* picks one random camera, decodes its PNG from disk, copies all parameters device->host->device
  (train.py:928-955)              -> cameras are pre-packed ``gsb_frame`` structs, targets and
                                     parameters stay in HBM
* forward, L1 loss (2 read-backs), sign gradient, backward, 5 gradient copies (965-1051)
                                  -> gsb_forward, fused L1 loss+gradient, gsb_backward writing
                                     straight into the flat gradient buffer
* adam_update(iteration) (1054)   -> one fused kernel over the flat buffers
* densification_and_pruning(iteration) (1060, 351-713) -> same orchestration, same quirks

Multi-GPU (not in the reference): a step's batch of views is split contiguously over the ranks,
each rank sums the gradients of its views into one flat [59*N] buffer, the buffers are summed
(NCCL all-reduce, or the fused peer-memory exchange + Adam kernel), and every rank applies the
identical densify step: replicas stay bit-identical because every parameter is computed once (or
NCCL's result is the same on all ranks), the summed position gradient the candidate masks are
computed from is published to every rank on densify steps, and clone/split noise is an index hash."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, optimizer
from .config import GaussianParams
from .scheduler import LRScheduler

KEYS = ("positions", "scales", "rotations", "opacities", "shs")
# exchange modes in which ONE kernel sums a rank's shard of the gradients out of the peers' buffers, applies Adam to
# that shard and stores the parameters into every rank's buffer (gsb_adam_step_peers): gradient loads / parameter
# stores are peer accesses ("peers"), multimem.ld_reduce / multimem.st ("multimem"), or peer loads + multimem.st ("hybrid")
FUSED_EXCHANGES = ("peers", "multimem", "hybrid")
WIDTH = {"positions": 3, "scales": 3, "rotations": 4, "opacities": 1, "shs": 48}
SHAPE = {"positions": lambda n: (n, 3), "scales": lambda n: (n, 3), "rotations": lambda n: (n, 4),
         "opacities": lambda n: (n,), "shs": lambda n: (n * 16, 3)}


def shard_views(batch, rank: int, world_size: int):
    """Contiguous split of a step's global view batch over the ranks: rank r takes
    batch[r*B/G : (r+1)*B/G].  The batch must be a multiple of the world size."""
    B = len(batch)
    per = B // world_size
    if per * world_size != B or per < 1:
        raise ValueError(f"batch of {B} views cannot be split evenly over {world_size} ranks")
    return list(batch[rank * per:(rank + 1) * per])


def compact_sh_step(enabled: bool, num_views: int, world_size: int) -> bool:
    """Whether a step's SH gradients can cross NVLink as their rank-1 factors (gsb_backward_compact_sh +
    gsb_adam_step_peers_compact): only when every rank renders exactly ONE view -- the factors of two views
    do not add -- and there is somebody to exchange with.  A function of global quantities only, so all
    ranks decide alike."""
    return bool(enabled) and world_size > 1 and num_views == world_size


def shard_range(n: int, rank: int, world_size: int):
    """The Gaussians [g0, g1) whose Adam state rank ``rank`` owns in the fused exchange modes (same rule
    as adam_step_peers_impl, csrc/optimizer.cu: boundaries on multiples of 4 Gaussians)."""
    g0 = (n * rank // world_size) // 4 * 4
    g1 = n if rank == world_size - 1 else (n * (rank + 1) // world_size) // 4 * 4
    return g0, g1


def densify_due(config: dict, iteration: int) -> bool:
    """train.py:385-392: the iterations on which clone / split / prune candidates are marked.  A function of
    the iteration and the config only, so every rank of a data-parallel run decides alike."""
    return (iteration > config["densify_from_iter"] and iteration < config["densify_until_iter"]
            and iteration % config["densification_interval"] == 0)


def flat_layout(n: int):
    """Offsets (in floats) of the five tensors inside a FlatGaussians buffer and its total length:
    every segment starts on a 16-byte boundary."""
    offs, total = {}, 0
    for k in KEYS:
        offs[k] = total
        total += (n * WIDTH[k] + 3) // 4 * 4
    return offs, max(total, 4)


def headroom(n: int) -> int:
    """Capacity (in Gaussians) allocated for a state that currently holds ``n``: densify grows the set by a few per
    cent per event, and every reallocation costs cudaMalloc / cudaFree calls -- and, in the fused exchange modes, a
    symmetric-memory rendezvous across the ranks -- so buffers are sized once with 50% to spare."""
    return max(n + n // 2, n + 4096)


class FlatGaussians:
    """One contiguous fp32 buffer holding the five per-Gaussian tensors back to back (each segment
    starts on a 16-byte boundary), with views of the reference's shapes.  The gradient instance of
    this class is what gets all-reduced: one message of 59*N floats.

    The buffer is allocated for ``capacity`` >= n Gaussians; ``resize(n)`` re-lays the five segments out for another
    count inside the same memory (the layout depends on n, the base address does not: peer mappings stay valid)."""

    def __init__(self, n: int, device, fill: float | None = 0.0, symmetric_group=None, capacity: int | None = None):
        self.capacity = max(int(capacity or 0), n)
        _offs, total = flat_layout(self.capacity)
        self.symm = None
        if symmetric_group is not None:
            # symmetric memory: every rank can address every other rank's copy (NVLink peer mapping,
            # plus an NVSwitch multicast address when the fabric supports it)
            import torch.distributed._symmetric_memory as symm_mem
            self.store = symm_mem.empty(total, dtype=torch.float32, device=device)
            self.store.zero_()
            self.symm = symm_mem.rendezvous(self.store, symmetric_group)
        else:
            self.store = (torch.zeros if fill == 0.0 else torch.empty)(total, dtype=torch.float32, device=device)
        self.resize(n)

    def resize(self, n: int):
        """The same memory laid out for ``n`` Gaussians (contents are NOT moved)."""
        if n > self.capacity:
            raise ValueError(f"{n} Gaussians do not fit a buffer of capacity {self.capacity}")
        self.n = n
        offs, total = flat_layout(n)
        self.flat = self.store[:total]
        self.views = {k: self.flat[offs[k]: offs[k] + n * WIDTH[k]].view(SHAPE[k](n)) for k in KEYS}
        return self

    def __getitem__(self, k):
        return self.views[k]

    def zero_(self):
        self.flat.zero_()

    def load(self, params: dict):
        for k in KEYS:
            self.views[k].copy_(_lib.to_device(params[k], device=self.flat.device).reshape(self.views[k].shape))
        return self

    def as_dict(self):
        return dict(self.views)


class FrameBuffers:
    """Every per-view output of forward/backward, allocated once and reused across steps.  The per-Gaussian arrays are
    allocated for ``n_capacity`` Gaussians (the kernels take pointers and a count), so a densify event that stays
    inside the capacity reuses them."""

    def __init__(self, n: int, width: int, height: int, device, capacity: int, n_capacity: int | None = None):
        f32, i32 = torch.float32, torch.int32
        e = lambda *s, dtype=f32: torch.empty(s, dtype=dtype, device=device)  # noqa: E731
        gx, gy = (width + 15) // 16, (height + 15) // 16
        self.n, self.W, self.H = n, width, height
        self.n_capacity = n = max(int(n_capacity or 0), n)
        self.radii, self.point_offsets = e(n, dtype=i32), e(n, dtype=i32)
        self.xy, self.depths, self.colors, self.cov3Ds = e(n, 2), e(n), e(n, 3), e(n, 6)
        self.conic_opacity, self.clamped_state = e(n, 4), e(n, 3)
        self.ranges = e(gx * gy, 2, dtype=i32)
        self.final_T, self.n_contrib = e(height, width), e(height, width, dtype=i32)
        self.image, self.depth = e(height, width, 3), e(height, width)
        self.point_list, self.block_masks = e(capacity, dtype=i32), e(capacity, dtype=i32)
        self.capacity = capacity
        self.num_rendered = 0
        self.dpix = e(height, width, 3)
        self.loss_sum = torch.zeros(1, dtype=torch.float64, device=device)
        # backward scratch that is not a parameter gradient
        self.dL_dcolor, self.dL_dmean2D, self.dL_dconic = e(n, 3), e(n, 3), e(n, 4)

    def as_reference_dict(self):
        """The 12-key dict of forward.py:881-894 (views of the first n rows of the per-Gaussian arrays);
        point_offsets is valid after Trainer.forward(..., point_offsets=True)."""
        n = self.n
        return {"radii": self.radii[:n], "point_offsets": self.point_offsets[:n], "points_xy_image": self.xy[:n],
                "depths": self.depths[:n], "colors": self.colors[:n], "cov3Ds": self.cov3Ds[:n],
                "conic_opacity": self.conic_opacity[:n], "point_list": self.point_list[: self.num_rendered],
                "ranges": self.ranges, "final_Ts": self.final_T, "n_contrib": self.n_contrib,
                "clamped_state": self.clamped_state[:n]}


class Trainer:
    def __init__(self, cameras, targets=None, num_points=None, params=None, config=None, device=None,
                 rank=0, world_size=1, process_group=None, exchange="auto", sh_compact=True, overlap_sh=True):
        """``exchange`` selects how the gradient sum and the Adam step are done when world_size > 1:
        "nccl" = all_reduce + replicated Adam; "peers" / "multimem" / "hybrid" = the fused kernel of
        gsb_adam_step_peers over symmetric memory (NVLink loads and stores / NVSwitch in-fabric reduction and
        multicast / NVLink gradient loads + multicast parameter stores);
        "auto" = peers when symmetric memory is available (measured fastest at 2 and 8 B200s:
        4470 vs 4349 multimem vs 3986 nccl views/s at 8 GPUs), else nccl.
        ``sh_compact`` (peers only): on steps where every rank renders exactly one view, the ranks
        publish the SH gradient as its two rank-1 factors (8 instead of 48 floats per Gaussian,
        gsb_backward_compact_sh) and the owner of a Gaussian expands them -- same bits, a third of
        the gradient bytes over NVLink.
        ``overlap_sh`` (fused exchange modes, ``train_step`` only): the exchange runs in two phases -- positions,
        scales, rotations and opacities (11 of the 59 floats per Gaussian) between two barriers on the main stream, the
        SH coefficients (48 of 59) on a side stream BESIDE the next step's geometry preprocess and binning, which do not
        read them; the next forward evaluates the colours behind that stream's event (gsb_set_color_dependency).
        Same kernels, same bits; steps with a densify event exchange in one piece."""
        self.config = GaussianParams.get_config_dict()
        if config:
            self.config.update(config)
        self.ctx = _lib.context(device)
        self.device = torch.device("cuda", self.ctx.device_index)
        self.rank, self.world_size, self.pg = rank, world_size, process_group
        self.exchange = self._pick_exchange(exchange)
        self.sh_compact = bool(sh_compact) and self.exchange in ("peers", "hybrid")
        # fused exchange modes: on by default.  One rank / nccl (``optimizer_step(overlap=True)``, the SH part of Adam
        # beside the next frame's geometry preprocess and binning): only on request (``overlap_sh=3``) -- measured a
        # wash on one B200, 676 against 681 us per step in tools/kbench.py but 685 against 680 us in bench.py: the two
        # streams' kernels mostly take turns on the SMs instead of sharing them
        self.overlap_sh = bool(overlap_sh) and (self.exchange in FUSED_EXCHANGES or int(overlap_sh) == 3)
        # 1: both phases start behind the opening barrier; 2: the SH phase starts behind the first phase's barrier
        self.overlap_order = int(overlap_sh) if self.overlap_sh else 0
        self._side_stream = None     # second phase of the exchange (overlap_sh)
        self._sh_event = None        # recorded behind it: the SH coefficients of every rank are final
        self._compact_step = False
        self.cameras = cameras
        bg = self.config["background_color"]
        self.frames = [_lib.make_frame(c["world_to_camera"], c["full_proj_matrix"], c["camera_center"], c["tan_fovx"],
                                       c["tan_fovy"], c["width"], c["height"], bg, self.config["sh_degree"], True,
                                       self.config["scale_modifier"]) for c in cameras]
        self.targets = None if targets is None else [_lib.to_device(t, device=self.device) for t in targets]
        from .utils.camera_utils import scene_extent
        self.scene_extent = scene_extent(cameras, self.config.get("camera_extent_factor", 1.0))
        self._scratch_bufs, self._densify_tmp, self.fb = {}, [None, None], None
        if params is not None:
            n = int(np.asarray(params["positions"].shape)[0]) if not isinstance(params["positions"], torch.Tensor) \
                else params["positions"].shape[0]
            self.num_points = n
            self.params = self._new_flat(n, capacity=headroom(n)).load(params)
        else:
            self.num_points = n = int(num_points or self.config["num_points"])
            self.params = self._new_flat(n, capacity=headroom(n))
            optimizer.init_gaussian_params(self.params["positions"], self.params["scales"], self.params["rotations"],
                                           self.params["opacities"], self.params["shs"], n, self.config["initial_scale"])
        self._alloc_state()
        sc = self.config["lr_scheduler_config"]
        ff = sc["final_lr_factor"]
        self.lr_scheduler = None
        if self.config["use_lr_scheduler"]:
            self.lr_scheduler = {k: LRScheduler(sc[k], ff) for k in ("lr_pos", "lr_scale", "lr_rot", "lr_sh", "lr_opac")}
        self._last_hw = (1, 1)
        self.losses = []

    # ---- state -------------------------------------------------------------------------------
    exchange_events = None   # set to a list to collect (start, end) CUDA events around exchange_and_step
    exchange_parts = []      # with it: (before barrier, after barrier, after kernel, after barrier) per step

    def _pick_exchange(self, want):
        if self.world_size <= 1:
            return "none"
        if want == "nccl":
            return "nccl"
        try:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem
            group = self.pg if self.pg is not None else dist.group.WORLD
            probe = symm_mem.empty(4, dtype=torch.float32, device=self.device)
            hdl = symm_mem.rendezvous(probe, group)
            has_mc = int(getattr(hdl, "multicast_ptr", 0) or 0) != 0
            self._symm_group = group
        except Exception as e:  # no peer access / symmetric memory unavailable
            if want in ("peers", "multimem", "hybrid"):
                raise RuntimeError(f"exchange={want!r} needs symmetric memory: {e}") from e
            return "nccl"
        if want in ("multimem", "hybrid") and not has_mc:
            raise RuntimeError(f"exchange={want!r} requested but the fabric offers no multicast mapping")
        return want if want in ("multimem", "hybrid") else "peers"

    def _new_flat(self, n, symmetric=True, capacity=None):
        if symmetric and self.exchange in FUSED_EXCHANGES:
            return FlatGaussians(n, self.device, symmetric_group=self._symm_group, capacity=capacity)
        return FlatGaussians(n, self.device, capacity=capacity)

    def _alloc_state(self):
        """Gradients and Adam moments for self.num_points Gaussians, all zeros (train.py:160-164, 474-476).  The
        buffers are reused while the count fits their capacity: a densify event then costs three memsets instead of
        three allocations (and, in the fused exchange modes, no symmetric-memory rendezvous)."""
        n = self.num_points
        old = getattr(self, "grads", None)
        if old is not None and old.capacity >= n and self.adam_m.capacity >= n and self.adam_v.capacity >= n:
            for buf in (self.grads, self.adam_m, self.adam_v):
                buf.resize(n).zero_()
        else:
            cap = headroom(n)
            self.grads = self._new_flat(n, capacity=cap)
            self.adam_m = FlatGaussians(n, self.device, capacity=cap)   # only this rank's shard is used in the fused modes
            self.adam_v = FlatGaussians(n, self.device, capacity=cap)
        self.grads_tmp = None
        if self.sh_compact:       # the expanded SH gradient of this rank's shard (local scratch)
            shard = -(-self.grads.capacity // self.world_size) + 8
            if getattr(self, "sh_local", None) is None or self.sh_local.numel() < 48 * shard:
                self.sh_local = torch.empty(48 * shard, dtype=torch.float32, device=self.device)
        else:
            self.sh_local = None

    def _frame_buffers(self, cam):
        W, H = cam["width"], cam["height"]
        fb = self.fb
        if fb is None or fb.n_capacity < self.num_points or fb.W != W or fb.H != H:
            cap = max(self.ctx.capacity_hint, 4 * self.num_points, 1024)
            fb = self.fb = FrameBuffers(self.num_points, W, H, self.device, cap, n_capacity=headroom(self.num_points))
        fb.n = self.num_points
        return fb

    def _scratch(self, name, count, dtype):
        """Reusable device scratch of the densify step (masks, prefixes, gradient norms); contents undefined."""
        t = self._scratch_bufs.get(name)
        if t is None or t.numel() < count or t.dtype != dtype:
            t = self._scratch_bufs[name] = torch.empty(headroom(count), dtype=dtype, device=self.device)
        return t[:count]

    # ---- forward / backward on preallocated buffers ---------------------------------------------
    def forward(self, cam_index: int, point_offsets: bool = False) -> FrameBuffers:
        """render_gaussians as called at train.py:935-955.  ``point_offsets``: also compute that output (the
        inclusive scan of tiles_touched, forward.py:755-763) -- nothing in the step loop reads it, so by default the
        scan is left out of the frame."""
        cam, frame = self.cameras[cam_index], self.frames[cam_index]
        fb = self._frame_buffers(cam)
        P, p, L = self._params, _lib.ptr, _lib.lib()      # (not the property: see ``params``)
        D = C.c_int64(0)
        fb.has_point_offsets = bool(point_offsets)
        sh_event, self._sh_event = self._sh_event, None
        if sh_event is not None:
            # the previous step's SH exchange may still be running on the side stream: this frame's geometry preprocess
            # and binning do not read the SH coefficients; the colours are evaluated behind the event (one-shot; an
            # error return hands it back to the context, so the retry below keeps it)
            self.ctx.check(L.gsb_set_color_dependency(self.ctx.h, C.c_void_p(sh_event.cuda_event)))
        for _ in range(2):
            rc = L.gsb_forward(self.ctx.h, _lib.stream_ptr(self.ctx.device_index), C.byref(frame), self.num_points,
                               p(P["positions"]), p(P["scales"]), p(P["rotations"]), p(P["opacities"]), p(P["shs"]),
                               p(fb.radii), p(fb.point_offsets if point_offsets else None), p(fb.xy), p(fb.depths), p(fb.colors), p(fb.cov3Ds),
                               p(fb.conic_opacity), p(fb.clamped_state), p(fb.point_list), fb.capacity, p(fb.ranges),
                               p(fb.image), p(fb.depth), p(fb.final_T), p(fb.n_contrib), C.byref(D),
                               p(fb.block_masks))
            if rc == _lib.GSB_ERR_CAPACITY:
                fb.capacity = int(D.value) + int(D.value) // 2 + 1024     # 50% to spare: the scene grows by densify events
                fb.point_list = torch.empty(fb.capacity, dtype=torch.int32, device=self.device)
                fb.block_masks = torch.empty(fb.capacity, dtype=torch.int32, device=self.device)
                continue
            self.ctx.check(rc)
            break
        fb.num_rendered = int(D.value)
        self.ctx.capacity_hint = max(self.ctx.capacity_hint, fb.num_rendered + fb.num_rendered // 4)
        return fb

    def loss_and_pixel_gradients(self, fb: FrameBuffers, target):
        """l1_loss + compute_image_gradients (train.py:965-979), fused; the loss stays on the device."""
        H, W = fb.H, fb.W
        l1_weight = (1.0 - 0.0) / (H * W * 3.0)          # train.py:978 passes lambda_dssim=0
        self.ctx.check(_lib.lib().gsb_l1_loss_grad(self.ctx.h, _lib.stream_ptr(self.ctx.device_index), H * W * 3,
                                                   _lib.ptr(fb.image), _lib.ptr(target), l1_weight, _lib.ptr(fb.dpix),
                                                   _lib.ptr(fb.loss_sum)))

    def backward(self, cam_index: int, fb: FrameBuffers, out: FlatGaussians):
        """backward() as called at train.py:1006-1044; parameter gradients land in ``out``."""
        frame, p = self.frames[cam_index], _lib.ptr
        P = self.params
        fn = _lib.lib().gsb_backward_compact_sh if self._compact_step else _lib.lib().gsb_backward
        rc = fn(
            self.ctx.h, _lib.stream_ptr(self.ctx.device_index), C.byref(frame), self.num_points, p(P["positions"]),
            p(P["opacities"]), p(P["shs"]), p(P["scales"]), p(P["rotations"]), p(fb.radii), p(fb.xy),
            p(fb.conic_opacity), p(fb.colors), p(fb.clamped_state), p(fb.cov3Ds), p(fb.point_list), p(fb.ranges),
            p(fb.final_T), p(fb.n_contrib), p(fb.dpix), p(out["positions"]), p(fb.dL_dcolor), p(out["shs"]),
            p(out["opacities"]), p(out["scales"]), p(out["rotations"]), p(fb.dL_dmean2D), p(fb.dL_dconic), p(None),
            p(fb.block_masks))
        self.ctx.check(rc)

    # ---- one optimisation step -------------------------------------------------------------------
    def learning_rates(self, iteration):
        sc = self.config["lr_scheduler_config"]
        if self.lr_scheduler:       # train.py:720-725
            T = self.config["num_iterations"]
            return {k: s.get_lr(iteration, T) for k, s in self.lr_scheduler.items()}
        return {k: sc[k] for k in ("lr_pos", "lr_scale", "lr_rot", "lr_sh", "lr_opac")}

    def optimizer_step(self, iteration, overlap=False):
        """train.py:716-794.  ``overlap`` (``train_step`` passes ``self.overlap_sh``): the update of positions, scales,
        rotations and opacities runs on the main stream, the update of the SH coefficients -- 81% of Adam's bytes -- on
        a side stream BESIDE the next frame's geometry preprocess and binning, which do not read SH (they are bound
        by atomics and latencies, Adam by HBM bandwidth); the next ``forward`` evaluates the colours behind that
        stream's event.  Same kernel, same arithmetic, same bits."""
        lr, G, P, M, V = self.learning_rates(iteration), self.grads, self.params, self.adam_m, self.adam_v
        if overlap:
            L, cfg, p = _lib.lib(), self.config, _lib.ptr

            def phase(k):
                self.ctx.check(L.gsb_adam_step_phase(
                    self.ctx.h, _lib.stream_ptr(self.ctx.device_index), self.num_points, p(G["positions"]), p(G["scales"]),
                    p(G["rotations"]), p(G["opacities"]), p(G["shs"]), lr["lr_pos"], lr["lr_scale"], lr["lr_rot"],
                    lr["lr_opac"], lr["lr_sh"], cfg["adam_beta1"], cfg["adam_beta2"], cfg["adam_epsilon"], iteration,
                    p(P["positions"]), p(P["scales"]), p(P["rotations"]), p(P["opacities"]), p(P["shs"]),
                    p(M["positions"]), p(M["scales"]), p(M["rotations"]), p(M["opacities"]), p(M["shs"]),
                    p(V["positions"]), p(V["scales"]), p(V["rotations"]), p(V["opacities"]), p(V["shs"]), k))

            main = torch.cuda.current_stream(self.device)
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream(device=self.device)
            side = self._side_stream
            side.wait_stream(main)          # the gradients are complete
            with torch.cuda.stream(side):
                phase(2)
                done = torch.cuda.Event()
                done.record(side)
            phase(1)
            self._sh_event = done
            return
        optimizer.adam_update(G["positions"], G["scales"], G["rotations"], G["opacities"], G["shs"], self.num_points,
                              lr["lr_pos"], lr["lr_scale"], lr["lr_rot"], lr["lr_opac"], lr["lr_sh"],
                              self.config["adam_beta1"], self.config["adam_beta2"], self.config["adam_epsilon"], iteration,
                              P["positions"], P["scales"], P["rotations"], P["opacities"], P["shs"],
                              M["positions"], M["scales"], M["rotations"], M["opacities"], M["shs"],
                              V["positions"], V["scales"], V["rotations"], V["opacities"], V["shs"])

    def accumulate_view(self, cam_index: int, target, first: bool):
        """forward + loss gradient + backward of one view; gradients are written (first view) or
        added (further views of the batch) into self.grads."""
        fb = self.forward(cam_index)
        self.loss_and_pixel_gradients(fb, target)
        if first:
            self.backward(cam_index, fb, self.grads)
        else:
            if self.grads_tmp is None or self.grads_tmp.n != self.num_points:
                self.grads_tmp = FlatGaussians(self.num_points, self.device, fill=None)
            self.backward(cam_index, fb, self.grads_tmp)
            self.ctx.check(_lib.lib().gsb_accumulate_f32(self.ctx.h, _lib.stream_ptr(self.ctx.device_index),
                                                         _lib.ptr(self.grads.flat), _lib.ptr(self.grads_tmp.flat),
                                                         self.grads.flat.numel()))
        return fb

    def all_reduce_gradients(self):
        if self.world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(self.grads.flat, op=dist.ReduceOp.SUM, group=self.pg)

    # The parameters and the Adam moments as everybody outside the step's hot path sees them: an access first joins a
    # second phase (SH coefficients / their moments) that may still be running on the side stream (overlap_sh).
    # ``forward`` alone uses ``_params``: its kernels read the SH coefficients only behind that stream's event.
    @property
    def params(self):
        self.join_exchange()
        return self._params

    @params.setter
    def params(self, value):
        self._params = value

    @property
    def adam_m(self):
        self.join_exchange()
        return self._adam_m

    @adam_m.setter
    def adam_m(self, value):
        self._adam_m = value

    @property
    def adam_v(self):
        self.join_exchange()
        return self._adam_v

    @adam_v.setter
    def adam_v(self, value):
        self._adam_v = value

    def join_exchange(self):
        """Make the current stream wait for a second exchange phase that is still in flight (overlap_sh): call before
        anything other than ``forward`` reads the parameters or the SH moments."""
        ev, self._sh_event = self._sh_event, None
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)

    def exchange_and_step(self, iteration, compact=None, publish_position_grad=False, overlap=False):
        """``compact``: the SH segment of every rank's gradient buffer holds the rank-1 factors
        written by gsb_backward_compact_sh (None = whatever train_step decided for this step).
        ``publish_position_grad`` (fused modes; nccl leaves the sum in ``self.grads`` anyway): after
        the call every rank's ``self.grads["positions"]`` holds the gradient summed over all ranks --
        what densification_and_pruning needs (train.py:398-433).

        Gradient sum over the ranks + Adam.  nccl: all_reduce then the replicated fused Adam.
        peers / multimem: ONE kernel between two cross-rank barriers -- each rank reduces its shard
        of the gradients straight out of its peers' buffers, updates that shard, and writes the new
        parameters into every rank's buffer (gsb_adam_step_peers).
        ``overlap`` (fused modes; ``train_step`` passes ``self.overlap_sh``): two phases, see ``__init__``; the caller
        must not read SH coefficients or SH moments before ``forward`` / ``join_exchange``."""
        self.join_exchange()
        if compact is not None:
            if compact and not self.sh_compact:
                raise ValueError("compact SH exchange needs exchange='peers' or 'hybrid' and sh_compact=True")
            self._compact_step = bool(compact)
        if self.exchange in ("none", "nccl"):
            self.all_reduce_gradients()
            self.optimizer_step(iteration, overlap=overlap)
            return
        G, P = self.grads.symm, self.params.symm
        lr = self.learning_rates(iteration)
        W = self.world_size
        gp = (C.c_uint64 * W)(*[int(x) for x in G.buffer_ptrs])
        pp = (C.c_uint64 * W)(*[int(x) for x in P.buffer_ptrs])
        g_mc = int(G.multicast_ptr) if self.exchange == "multimem" else 0
        p_mc = int(P.multicast_ptr) if self.exchange in ("multimem", "hybrid") else 0
        ev = None
        if self.exchange_events is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
        G.barrier(channel=0)        # every rank's backward has finished writing its gradients
        if ev:
            ev[1].record()
        if overlap:
            L, cfg = _lib.lib(), self.config
            sh_local = self.sh_local if self._compact_step else None

            def phase(k):
                self.ctx.check(L.gsb_adam_step_peers_phase(
                    self.ctx.h, _lib.stream_ptr(self.ctx.device_index), self.num_points, W, self.rank, gp, pp,
                    0 if self._compact_step else g_mc, p_mc, _lib.ptr(self.adam_m.flat), _lib.ptr(self.adam_v.flat),
                    lr["lr_pos"], lr["lr_scale"], lr["lr_rot"], lr["lr_opac"], lr["lr_sh"], cfg["adam_beta1"],
                    cfg["adam_beta2"], cfg["adam_epsilon"], iteration, _lib.ptr(sh_local),
                    sh_local.numel() if sh_local is not None else 0, cfg["sh_degree"], int(bool(publish_position_grad)), k))

            main = torch.cuda.current_stream(self.device)
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream(device=self.device)
            side = self._side_stream
            # both phases start behind the opening barrier (they touch disjoint segments of the gradient, moment and
            # parameter buffers); the SH phase is the longer one and gets the side stream
            def sh_phase():
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    phase(2)
                    P.barrier(channel=2)    # every rank's SH coefficients have landed everywhere (and nobody reads
                    done = torch.cuda.Event()   # the SH gradients any more)
                    done.record(side)
                return done

            if self.overlap_order != 2:
                done = sh_phase()
            phase(1)
            if ev:
                ev[2].record()
            P.barrier(channel=1)    # every rank's positions / scales / rotations / opacities have landed everywhere
            if ev:
                ev[3].record()
                self.exchange_parts.append(ev)
            if self.overlap_order == 2:
                done = sh_phase()
            self._sh_event = done
            return
        if self._compact_step:
            self.ctx.check(_lib.lib().gsb_adam_step_peers_compact(
                self.ctx.h, _lib.stream_ptr(self.ctx.device_index), self.num_points, W, self.rank, gp, pp, p_mc,
                _lib.ptr(self.adam_m.flat), _lib.ptr(self.adam_v.flat), lr["lr_pos"], lr["lr_scale"], lr["lr_rot"],
                lr["lr_opac"], lr["lr_sh"], self.config["adam_beta1"], self.config["adam_beta2"],
                self.config["adam_epsilon"], iteration, _lib.ptr(self.sh_local), self.sh_local.numel(),
                self.config["sh_degree"], int(bool(publish_position_grad))))
        else:
            self.ctx.check(_lib.lib().gsb_adam_step_peers(
                self.ctx.h, _lib.stream_ptr(self.ctx.device_index), self.num_points, W, self.rank, gp, pp, g_mc, p_mc,
                _lib.ptr(self.adam_m.flat), _lib.ptr(self.adam_v.flat), lr["lr_pos"], lr["lr_scale"], lr["lr_rot"],
                lr["lr_opac"], lr["lr_sh"], self.config["adam_beta1"], self.config["adam_beta2"],
                self.config["adam_epsilon"], iteration, int(bool(publish_position_grad))))
        if ev:
            ev[2].record()
        P.barrier(channel=1)        # every rank's parameter shard has landed everywhere
        if ev:
            ev[3].record()
            self.exchange_parts.append(ev)

    def train_step(self, iteration: int, cam_indices, targets=None, densify=True):
        """One step on a batch of views.  ``cam_indices`` is the GLOBAL batch; this rank takes the
        contiguous slice [rank*B/G, (rank+1)*B/G).  With one view and one rank this is exactly one
        iteration of the reference loop.  Returns the device tensor holding this rank's last
        sum|render - target| (divide by 3HW for the reference's loss)."""
        mine = shard_views(list(range(len(cam_indices))), self.rank, self.world_size)
        # one view on every rank: the SH gradient can travel as its rank-1 factors (same on all ranks)
        self._compact_step = compact_sh_step(self.sh_compact, len(cam_indices), self.world_size)
        fb = None
        for j, b in enumerate(mine):
            ci = cam_indices[b]
            tgt = targets[b] if targets is not None else self.targets[ci]
            fb = self.accumulate_view(ci, tgt, first=(j == 0))
        self._last_hw = (fb.H, fb.W)
        loss_sum = fb.loss_sum           # densify may drop the frame buffers; the scalar tensor stays alive
        # a function of the iteration and the config only: every rank decides alike
        publish = bool(densify) and self.densify_due(iteration)
        overlap = self.overlap_sh and not publish     # a densify event needs every parameter at once
        if self.exchange_events is not None:     # bench.py: device time of the exchange + Adam part
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self.exchange_and_step(iteration, publish_position_grad=publish, overlap=overlap)
            e1.record()
            self.exchange_events.append((e0, e1))
        else:
            self.exchange_and_step(iteration, publish_position_grad=publish, overlap=overlap)
        if densify:
            self.densification_and_pruning(iteration)
        return loss_sum

    # ---- densify / prune, train.py:351-713 ------------------------------------------------------
    def densify_due(self, iteration) -> bool:
        return densify_due(self.config, iteration)

    def _alloc_like(self, n):
        """The wp.zeros outputs of train.py:441-447 etc.  The intermediate Gaussian sets of one densify call (clone ->
        split -> compact -> prune) alternate between two reusable plain buffers (every element of an output set is
        written by the kernel that produces it); the set that survives is copied into the parameter buffer."""
        self._tmp_turn = 1 - getattr(self, "_tmp_turn", 1)
        t = self._densify_tmp[self._tmp_turn]
        if t is None or t.capacity < n:
            t = self._densify_tmp[self._tmp_turn] = FlatGaussians(n, self.device, fill=None, capacity=headroom(n))
        return t.resize(n)

    def _replace(self, new: FlatGaussians):
        """train.py:474-476 and its siblings: new parameter set, gradients and Adam moments reset to zeros.  The
        parameter buffer (symmetric memory in the fused exchange modes) is kept while the new count fits its capacity:
        the new set is copied in; otherwise a larger one is allocated (collective: every rank arrives here with the
        same count)."""
        if new is not self.params:
            if self.params.capacity >= new.n:
                self.params.resize(new.n)
            else:
                self.params = self._new_flat(new.n, capacity=headroom(new.n))
            self.params.flat.copy_(new.flat)
        self.num_points = new.n
        self._alloc_state()

    def densification_and_pruning(self, iteration):
        """train.py:351-713.  Multi-GPU: every rank runs this on identical inputs (parameters are replicas,
        ``self.grads["positions"]`` holds the all-rank sum: NCCL's result, or published by the fused exchange
        kernel), so masks, counts and the new parameter sets are identical without any further exchange."""
        cfg = self.config
        log = {"cloned": 0, "split": 0, "split_removed": 0, "pruned": 0, "opacity_reset": False}
        i32 = torch.int32
        if self.densify_due(iteration):
            self.join_exchange()     # (train_step exchanges in one piece on these iterations; other callers may not)
            n = self.num_points
            avg_grads = self._scratch("avg_grads", n, torch.float32)     # every kernel below writes all of its output
            optimizer.compute_grad_norms(self.grads["positions"], avg_grads, n)
            gt, pd, ext = cfg["densify_grad_threshold"], cfg["percent_dense"], self.scene_extent
            P = cur = self.params                # `cur`: the Gaussian set as the reference's self.params sees it
            clone_mask = self._scratch("clone_mask", n, i32)
            optimizer.mark_clone_candidates(avg_grads, P["scales"], gt, ext, pd, n, clone_mask)
            clone_prefix = self._scratch("clone_prefix", n, i32)
            total_to_clone = optimizer.array_scan(clone_mask, clone_prefix, inclusive=False)
            if total_to_clone > 0:
                out = self._alloc_like(n + total_to_clone)
                optimizer.clone_gaussians(clone_mask, clone_prefix, P["positions"], P["scales"], P["rotations"],
                                          P["opacities"], P["shs"], 0.01, n, out["positions"], out["scales"],
                                          out["rotations"], out["opacities"], out["shs"])
                cur = out
                log["cloned"] = total_to_clone
            n, P = cur.n, cur
            split_mask = self._scratch("split_mask", n, i32)
            # quirk G4: avg_grads still has the pre-clone length
            optimizer.mark_split_candidates(avg_grads, P["scales"], gt, ext, pd, n, split_mask)
            split_prefix = self._scratch("split_prefix", n, i32)
            total_to_split = optimizer.array_scan(split_mask, split_prefix, inclusive=False)
            if total_to_split > 0:
                N0, n_split = n, 2
                new_n = N0 + total_to_split * n_split
                out = self._alloc_like(new_n)
                optimizer.split_gaussians(split_mask, split_prefix, P["positions"], P["scales"], P["rotations"],
                                          P["opacities"], P["shs"], n_split, 0.8, N0, out["positions"], out["scales"],
                                          out["rotations"], out["opacities"], out["shs"])
                P = cur = out
                log["split"] = total_to_split
                valid = self._scratch("valid", new_n, i32)
                optimizer.split_valid_mask(split_mask, valid, N0, new_n)
                prefix = self._scratch("prefix", new_n, i32)
                valid_count = optimizer.array_scan(valid, prefix, inclusive=False)
                if valid_count < new_n:
                    out = self._alloc_like(valid_count)
                    optimizer.compact_gaussians(valid, prefix, P["positions"], P["scales"], P["rotations"],
                                                P["opacities"], P["shs"], out["positions"], out["scales"],
                                                out["rotations"], out["opacities"], out["shs"])
                    log["split_removed"] = new_n - valid_count
                    cur = out
            n, P = cur.n, cur
            valid = self._scratch("valid", n, i32)
            optimizer.prune_gaussians(P["opacities"], cfg["cull_opacity_threshold"], n, valid)
            prefix = self._scratch("prefix", n, i32)
            valid_count = optimizer.array_scan(valid, prefix, inclusive=False)
            prune_count = n - valid_count
            prune_ratio = prune_count / n if n > 0 else 0
            if (valid_count >= cfg["min_valid_points"] and valid_count <= cfg["max_valid_points"]
                    and prune_ratio <= cfg["max_allowed_prune_ratio"] and valid_count < n):
                out = self._alloc_like(valid_count)
                optimizer.compact_gaussians(valid, prefix, P["positions"], P["scales"], P["rotations"], P["opacities"],
                                            P["shs"], out["positions"], out["scales"], out["rotations"],
                                            out["opacities"], out["shs"])
                cur = out
                log["pruned"] = prune_count
            if cur is not self.params:
                self._replace(cur)               # gradients and Adam moments restart from zeros (train.py:474-476)
        background_is_white = all(c == 1.0 for c in cfg["background_color"])
        if (iteration % cfg["opacity_reset_interval"] == 0
                or (background_is_white and iteration == cfg["densify_from_iter"])):   # quirk G6: iteration 0 too
            optimizer.reset_opacities(0.01, self.num_points, self.params["opacities"])
            log["opacity_reset"] = True
        return log

    # ---- checkpoints (train.py:796-849 saves the PLY only; here the run can also be resumed) ------
    def gathered_moments(self):
        """The full Adam moments as (m, v) flat CUDA tensors.  In the fused exchange modes a rank only ever
        updates the moments of its own shard of Gaussians (shard_range): the shards are put together with one
        all-reduce of buffers that are zero outside the owner's shard.  Collective when world_size > 1."""
        self.join_exchange()
        if self.exchange not in FUSED_EXCHANGES:
            return self.adam_m.flat, self.adam_v.flat
        import torch.distributed as dist
        g0, g1 = shard_range(self.num_points, self.rank, self.world_size)
        offs, _total = flat_layout(self.num_points)
        out = []
        for src in (self.adam_m.flat, self.adam_v.flat):
            t = torch.zeros_like(src)
            for k in KEYS:
                a, b = offs[k] + g0 * WIDTH[k], offs[k] + g1 * WIDTH[k]
                t[a:b] = src[a:b]
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg)   # x + 0 + ... + 0: exact
            out.append(t)
        return out[0], out[1]

    def save_checkpoint(self, directory, iteration):
        """point_cloud/iteration_N/point_cloud.ply in the reference's layout (save_ply) + loss.txt,
        plus state.npz (Adam moments, iteration) which the reference does not have.  Collective when
        world_size > 1 (the moment shards are gathered); rank 0 alone writes the files."""
        import os
        from .utils.point_cloud_utils import save_ply
        ckpt = os.path.join(str(directory), "point_cloud", f"iteration_{iteration}")
        m, v = self.gathered_moments()
        if self.rank == 0:
            os.makedirs(ckpt, exist_ok=True)
            save_ply(self.params.as_dict(), os.path.join(ckpt, "point_cloud.ply"), self.num_points)
            with open(os.path.join(str(directory), "loss.txt"), "w") as f:
                for row in self.losses:
                    f.write(f"{row[1] if isinstance(row, tuple) else row}\n")
            np.savez(os.path.join(ckpt, "state.npz"), iteration=iteration, num_points=self.num_points,
                     adam_m=m.cpu().numpy(), adam_v=v.cpu().numpy())
        if self.world_size > 1:
            import torch.distributed as dist
            dist.barrier(group=self.pg)          # the files exist when any rank returns
        return ckpt

    def load_checkpoint(self, ckpt_dir):
        """Restores parameters (bit-exact: the PLY stores raw float32) and the Adam state (every rank loads
        the full moments; in the fused exchange modes it goes on using its own shard of them); returns the
        iteration to continue from -- pass it to ``train(start_iteration=...)``."""
        import os
        from .utils.point_cloud_utils import load_ply
        params = load_ply(os.path.join(str(ckpt_dir), "point_cloud.ply"))
        st = np.load(os.path.join(str(ckpt_dir), "state.npz"))
        self.join_exchange()
        self._replace(FlatGaussians(int(st["num_points"]), self.device).load(params))     # copied into the state buffers
        self.adam_m.flat.copy_(torch.from_numpy(st["adam_m"]))
        self.adam_v.flat.copy_(torch.from_numpy(st["adam_v"]))
        return int(st["iteration"]) + 1

    # ---- the reference's loop --------------------------------------------------------------------
    def train(self, num_iterations=None, batch_size=1, seed=42, log_every=0, start_iteration=0):
        """train.py:920-1066 with a seeded camera sampler; losses are read back every ``log_every``
        steps only (the reference blocks on the loss every iteration).  ``start_iteration``: continue a
        run restored by load_checkpoint -- the sampler is advanced past the iterations already done, so
        the camera sequence is the one an uninterrupted run would have seen."""
        T = num_iterations or self.config["num_iterations"]
        rng = np.random.default_rng(seed)
        for it in range(T):
            cams = [int(rng.integers(0, len(self.cameras))) for _ in range(batch_size)]
            if it < start_iteration:
                continue
            loss_sum = self.train_step(it, cams)
            if log_every and it % log_every == 0:
                H, W = self._last_hw
                self.losses.append((it, float(loss_sum.item()) / (3 * H * W), self.num_points))
        return self.losses
