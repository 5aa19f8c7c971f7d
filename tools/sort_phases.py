#!/usr/bin/env python
"""Phase clocks of the cooperative radix sort (pass 1 of a sort), from a -DGSB_SORT_TIMING build:

    GSB200_LIB=3dgs-native_b200/csrc/libgsb200_st.so python tools/sort_phases.py [D ...]
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gsb200  # noqa: E402,F401
from gsb200 import _lib  # noqa: E402

NAMES = ["load keys + zero", "rank", "warp scan + publish", "barrier 1", "table read", "scan", "reorder", "write out",
         "barrier 2"]


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [1614352, 100000]
    ctx, L, p = _lib.context(), _lib.lib(), _lib.ptr
    raw = C.CDLL(os.environ["GSB200_LIB"])
    dev = torch.device("cuda")
    for D in sizes:
        g = torch.Generator(device=dev).manual_seed(D)
        keys = (torch.randint(0, 2500, (D,), device=dev, generator=g, dtype=torch.int64) << 32) | \
            (torch.rand(D, device=dev, generator=g) * 6.0 + 1.0).view(torch.int32).to(torch.int64)
        vals = torch.arange(D, device=dev, dtype=torch.int32)
        tk, tv = torch.empty_like(keys), torch.empty_like(vals)
        for _ in range(3):
            ks, vs = keys.clone(), vals.clone()
            ctx.check(L.gsb_sort_pairs64(ctx.h, _lib.stream_ptr(ctx.device_index), p(ks), p(vs), p(tk), p(tv), D, 0, 44))
            torch.cuda.synchronize()
        G = (D + 12287) // 12288 if D > 148 * 2048 else (D + 2047) // 2048
        G = min(G, 148)
        buf = (C.c_longlong * (160 * 16))()
        assert raw.gsb_debug_sort_clocks(buf, 160 * 16) == 0
        clk = np.array(buf, dtype=np.int64).reshape(160, 16)[:G, :10]
        d = np.diff(clk, axis=1) / 1.965e3   # us at 1965 MHz
        print(f"D={D}, G={G}: per-phase us of pass 1 (median over CTAs / max / CTA 0 / last CTA)")
        for i, nme in enumerate(NAMES):
            print(f"  {nme:22s} {np.median(d[:, i]):7.2f} {d[:, i].max():7.2f} {d[0, i]:7.2f} {d[-1, i]:7.2f}")
        print(f"  {'pass':22s} {np.median(clk[:, 9] - clk[:, 0]) / 1.965e3:7.2f}")


if __name__ == "__main__":
    main()
