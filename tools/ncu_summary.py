#!/usr/bin/env python
"""Markdown summary of an `ncu --set full` report: one column per captured kernel, the metrics DESIGN.md and
bench.py refer to.   python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "live lanes per instruction (of 32)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots used (active cycles)"),
    ("sm__cycles_active.avg", "SM active cycles"),
    ("sm__cycles_elapsed.avg", "SM elapsed cycles"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__waves_per_multiprocessor", "waves per SM"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU data-pipe wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
    ("lts__t_bytes.sum", "L2 bytes"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    kernels = rows[2:]
    names = [r[idx["Kernel Name"]].replace("void ", "").replace("<unnamed>::", "").split("(")[0] for r in kernels]
    print("| metric | " + " | ".join(f"`{n}`" for n in names) + " |")
    print("|---|" + "---|" * len(names))
    for key, label in METRICS:
        if key not in idx:
            continue
        cells = []
        for r in kernels:
            v = r[idx[key]]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:,.0f}" if abs(f) >= 1000 else f"{f:.2f}"
            except ValueError:
                pass
            cells.append(f"{v} {units[idx[key]]}".strip())
        print(f"| {label} (`{key}`) | " + " | ".join(cells) + " |")
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
             and "not_issued" not in h]
    cells = []
    for r in kernels:
        top = sorted(((float(r[idx[h]] or 0), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")])
                      for h in stall), reverse=True)[:5]
        cells.append(", ".join(f"{n} {v:.2f}" for v, n in top if n != "selected"))
    print("| top stalls (cycles per issued instruction) | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
