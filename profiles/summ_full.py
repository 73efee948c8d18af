#!/usr/bin/env python
"""Summarise an ncu --set full report: per kernel duration, DRAM bytes, pipe utilisation, occupancy, top stalls.
usage: python profiles/summ_full.py gpurun_out/x.ncu-rep"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__inst_executed.sum", "inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "fmaH%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__cycles_active.avg", "cycles"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    if not stall_cols:
        stall_cols = [h for h in hdr if "warp_issue_stalled" in h and h.endswith(".ratio")]
    for r in data:
        name = r[col["Kernel Name"]]
        print("=== %s" % name[:110])
        parts = []
        for k, short in KEYS:
            if k in col:
                parts.append("%s=%s%s" % (short, r[col[k]], units[col[k]] if units[col[k]] not in ("", "%") else ""))
        print("   " + "  ".join(parts))
        st = []
        for h in stall_cols:
            try:
                st.append((float(r[col[h]].replace(",", "")), h.split("issue_stalled_")[1].split("_per")[0]))
            except Exception:
                pass
        st.sort(reverse=True)
        print("   stalls/issue: " + ", ".join("%s=%.2f" % (n, v) for v, n in st[:6]))


if __name__ == "__main__":
    main(sys.argv[1])
