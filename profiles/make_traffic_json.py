#!/usr/bin/env python
"""Extract dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu --set full reports and write
profiles/ncu_traffic_<tag>.json (read by fqss_b200/roofline.py for bench.py's roofline.traffic).
usage: python profiles/make_traffic_json.py <tag> <per_gpu_batch> rep1.ncu-rep [rep2 ...]"""
import csv
import io
import json
import os
import re
import subprocess
import sys

CLASS = [  # (regex on the demangled kernel name, kernel class used by csrc/prof.cu)
    (r"tcn_dw_fwd_kernel<\(?(bool\))?(1|true)", "tcn_dw_fwd"), (r"tcn_dw_fwd_kernel<\(?(bool\))?(0|false)", "tcn_dw_fwd(float)"),
    (r"tcn_hidden_fq_kernel<\(?(bool\))?(1|true)", "tcn_hidden_fq"), (r"tcn_hidden_fq_kernel<\(?(bool\))?(0|false)", "tcn_hidden_fq(float)"),
    (r"tcn_tail_bwd_kernel", "tcn_tail_bwd"), (r"tcn_gln2_bwd_kernel<\(?(int\))?1", "tcn_gln2_bwd<1>"),
    (r"tcn_gln2_sums_codes_kernel", "tcn_gln2_bwd<1>"), (r"tcn_gln2_sums_lean_kernel", "tcn_gln2_sums"),
    (r"tcn_gln2_dw_bwd(_lean)?_kernel", "tcn_gln2_dw_bwd"),
    (r"tcn_gln2_bwd_kernel<\(?(int\))?2", "tcn_gln2_bwd<2>"), (r"tcn_dw_bwd_kernel", "tcn_dw_bwd"), (r"tcn_gln1_bwd_kernel", "tcn_gln1_bwd"),
    (r"pw_gemm_kernel<\(?(int\))?\d+, \(?(int\))?1[,>]", "gemm_expand"), (r"pw_gemm_kernel<\(?(int\))?\d+, \(?(int\))?2[,>]", "gemm_resskip"),
    (r"pw_gemm_kernel<\(?(int\))?\d+, \(?(int\))?3[,>]", "gemm_dgrad_bf16"), (r"pw_gemm_kernel<\(?(int\))?\d+, \(?(int\))?4[,>]", "gemm_dgrad_add"),
    (r"wgrad_kernel", "wgrad"),
]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main(tag, batch, reps):
    agg = {}
    for rep in reps:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        col = {h: i for i, h in enumerate(hdr)}
        for r in data:
            name = r[col["Kernel Name"]]
            cls = next((c for rx, c in CLASS if re.search(rx, name)), None)
            if cls is None:
                continue
            rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
            wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            dur = r[col["gpu__time_duration.sum"]] + units[col["gpu__time_duration.sum"]]
            e = agg.setdefault(cls, {"n": 0, "rd": 0.0, "wr": 0.0, "dur": []})
            e["n"] += 1
            e["rd"] += rd
            e["wr"] += wr
            e["dur"].append(dur)
    kernels = {k: {"dram_bytes_per_launch": (e["rd"] + e["wr"]) / e["n"], "dram_read": e["rd"] / e["n"], "dram_write": e["wr"] / e["n"],
                   "launches_captured": e["n"], "ncu_durations": e["dur"]} for k, e in sorted(agg.items())}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic_%s.json" % tag)
    with open(path, "w") as f:
        json.dump({"per_gpu_batch": int(batch), "source": [os.path.basename(r) for r in reps],
                   "note": "ncu --set full --clock-control none, one steady-state QAT step (bench.py --profile-step); bytes per launch",
                   "kernels": kernels}, f, indent=1)
    print(path)
    for k, v in kernels.items():
        print("%-22s %8.1f MB/launch (%d launches)" % (k, v["dram_bytes_per_launch"] / 1e6, v["launches_captured"]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3:])
