#!/usr/bin/env python
"""gsb_sort_pairs64 (the stable 64-bit radix sort, forward.py:791-824) on tile|depth keys of the headline size and of
config 5: device time per call and per 8-bit pass, checked against torch.sort (stable).

    python tools/sortbench.py [D ...]        (default 1614352 68406031)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gsb200  # noqa: E402,F401
from gsb200 import _lib  # noqa: E402


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [1614352, 68406031]
    ctx, L, p = _lib.context(), _lib.lib(), _lib.ptr
    s = lambda: _lib.stream_ptr(ctx.device_index)  # noqa: E731
    dev = torch.device("cuda")
    for D in sizes:
        tiles = 2500 if D < 10_000_000 else 32400
        bits = 32 + int(np.ceil(np.log2(tiles)))
        g = torch.Generator(device=dev).manual_seed(D)
        tile = torch.randint(0, tiles, (D,), device=dev, generator=g, dtype=torch.int64)
        depth = (torch.rand(D, device=dev, generator=g) * 6.0 + 1.0).view(torch.int32).to(torch.int64)
        keys = (tile << 32) | depth
        vals = torch.arange(D, device=dev, dtype=torch.int32)
        ks, vs = torch.empty_like(keys), torch.empty_like(vals)
        tk, tv = torch.empty_like(keys), torch.empty_like(vals)
        passes = (bits + 7) // 8
        coop = int(os.environ.get("GSB_SORT_COOP", "1")) and D <= 148 * 12288
        if coop and (bits + 8) // 9 < passes:
            passes = (bits + 8) // 9      # the cooperative kernel widens the digit to 9 bits when that saves a pass
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        times = []
        ctx.set_option("sort_coop", int(os.environ.get("GSB_SORT_COOP", "1")))
        for it in range(8):
            ks.copy_(keys)
            vs.copy_(vals)
            if os.environ.get("GSB_SORT_BUSY", "1") == "1":
                torch.cuda._sleep(600000)   # ~0.3 ms of GPU work in front: the host's launch latency is not timed
            e0.record()
            ctx.check(L.gsb_sort_pairs64(ctx.h, s(), p(ks), p(vs), p(tk), p(tv), D, 0, bits))
            e1.record()
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
        ref_k, order = torch.sort(keys, stable=True)
        ok = bool(torch.equal(ks, ref_k)) and bool(torch.equal(vs.long(), order))
        ms = float(np.median(times[2:]))
        print(f"D={D}: {ms * 1e3:9.1f} us per sort, {passes} passes ({'one cooperative launch' if coop else 'three kernels each'}) -> {ms * 1e3 / passes:7.1f} us per pass "
              f"({(24 if coop else 32) * D / (ms * 1e-3 / passes) / 1e9:7.1f} GB/s of {24 if coop else 32} B per pair and pass), stable and sorted: {ok}", flush=True)
        assert ok


if __name__ == "__main__":
    main()
