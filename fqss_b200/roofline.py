"""Live per-kernel attribution of a QAT step and the roofline of its dominant kernel (bench.py).

Timing comes from the library's own launch scopes (csrc/prof.cu): with collection on, every kernel
launch is bracketed by CUDA events recorded on the launching stream, so the numbers are taken inside a
real step (warm caches, real producer/consumer order), not under a profiler.  ALGORITHMIC bytes per
launch follow SURVEY.md 8(d): every distinct input tensor read once + every output written once, at
the dtype actually stored (DESIGN.md "Kernels" lists the same formulas).
"""
import ctypes as C
import json
import os

from . import _native as N

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md, used only without MEASURED_PEAKS.json


def measured_peaks():
    path = os.path.join(_ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def _prof_api():
    return N.lib()


def launch_count():
    """Kernels launched by libfqss_sm100 in this process so far."""
    return int(_prof_api().fqss_launch_count())


def profile(fn, repeats=1):
    """Run fn() `repeats` times with per-kernel event timing; -> {class: {"ms": per call of fn, "scopes", "kernels"}}."""
    import torch
    L = _prof_api()
    torch.cuda.synchronize()
    L.fqss_prof_reset()
    prev = L.fqss_set_wgrad_overlap(0)      # per-kernel timing wants every kernel alone on the stream
    from . import losses
    prev_t = losses.set_teacher_overlap(False)
    L.fqss_prof_enable(1)
    try:
        for _ in range(repeats):
            fn()
        torch.cuda.synchronize()
    finally:
        L.fqss_prof_enable(0)
        L.fqss_set_wgrad_overlap(prev)
        losses.set_teacher_overlap(prev_t)
    out = {}
    name = C.create_string_buffer(64)
    for i in range(L.fqss_prof_nslots()):
        ms, sc, kn = C.c_double(), C.c_int64(), C.c_int64()
        if L.fqss_prof_read(i, name, 64, C.byref(ms), C.byref(sc), C.byref(kn)) != 0:
            raise N.FqssError(L.fqss_last_error().decode())
        out[name.value.decode()] = {"ms": ms.value / repeats, "scopes": sc.value / repeats, "kernels": kn.value / repeats}
    L.fqss_prof_reset()
    return out


def algorithmic_bytes(B, M, Cio=128, Chid=512):
    """Bytes per launch of each TCN kernel class (student, steady state).  H/I = hidden / io elements."""
    H, I = B * Chid * M, B * Cio * M
    w1, w2 = Chid * Cio * 2, 2 * Cio * Chid * 2
    return {
        # forward
        "gemm_expand": 2 * I + w1 + 4 * H + H,                  # x_op bf16 in, y1 fp32 + code1 (u8) out
        "tcn_dw_fwd": H + 4 * H + H,                            # code1 (u8) in, y3 + code3 (u8) out
        "tcn_hidden_fq": H + 2 * H,                             # code3 (u8) in, a4 operand (bf16 codes) out
        "gemm_resskip": 2 * H + w2 + 4 * I * 2 + 4 * I * 4 + 2 * I,   # a4, x_in, skip_in -> res_y, skip_y, x_out, skip_out, x_out_op
        "tcn_dw_fwd(float)": 8 * H,
        "tcn_hidden_fq(float)": 4 * H + 4 * H,                  # [hi ; lo] bf16 pair out
        # backward
        "tcn_tail_bwd": 4 * I * 6 + 4 * I * 2 + 2 * I * 2,      # g_x, g_skip, res_y, skip_y, x_in, skip_in -> g_xd, g_skip_in, dY2
        "gemm_dgrad_bf16": 2 * 2 * I + w2 + 2 * H,              # dY2 in, g_a4 bf16 out
        "tcn_gln2_bwd<1>": H + 2 * H,                           # code3 (u8), g_a4 (bf16) in (sums out)
        "tcn_gln2_sums": H + 2 * H,                             # same pass, instruction-lean kernel (tcn_rows.cuh)
        "tcn_gln2_bwd<2>": 4 * H + 2 * H + 2 * H,               # y3, g_a4 in, g_y3 bf16 out
        "tcn_gln2_dw_bwd": 4 * H + 2 * H + H + H + 2 * H,       # y3, g_a4 (bf16), code3, code1 (u8) in, g_n1 bf16 out (fused P2 + D)
        "tcn_dw_bwd": H + 2 * H + 2 * H,                        # code1 (u8), g_y3 in, g_n1 bf16 out
        "tcn_gln1_bwd": 4 * H + 2 * H + 2 * H,                  # y1, g_n1 in, dY1 bf16 out
        "gemm_dgrad_add": 2 * H + w1 + 4 * I + 4 * I,           # dY1, g_xd in, g_x_in out
        "tcn_hid_bwd_a": 4 * H + 2 * H,
        "tcn_hid_bwd_b": 4 * H + 2 * H + 2 * H,
    }


def _ncu_traffic(kernel, B):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture, if one matches."""
    pdir = os.path.join(_ROOT, "profiles")
    try:
        names = sorted(f for f in os.listdir(pdir) if f.startswith("ncu_traffic_") and f.endswith(".json"))
    except OSError:
        return None
    for fn in reversed(names):
        try:
            with open(os.path.join(pdir, fn)) as f:
                d = json.load(f)
            e = d.get("kernels", {}).get(kernel)
            if e and int(d.get("per_gpu_batch", -1)) == int(B):
                return float(e["dram_bytes_per_launch"])
        except Exception:
            continue
    return None


def step_roofline(step_fn, B, M, ms_per_step, repeats=2, top=10):
    """-> (roofline object for the dominant kernel, breakdown list) for bench.py's JSON line."""
    prof = profile(step_fn, repeats)
    total = sum(v["ms"] for v in prof.values())
    table = algorithmic_bytes(B, M)
    ranked = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])
    breakdown = [{"kernel": k, "ms_per_step": round(v["ms"], 3), "launches_per_step": v["kernels"],
                  "share": round(v["ms"] / total, 4) if total else None} for k, v in ranked[:top]]
    peak, how = measured_peaks()
    for k, v in ranked:
        if k in table and v["scopes"] > 0:
            avg_ms = v["ms"] / v["scopes"]
            ach = table[k] / (avg_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": k, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": _ncu_traffic(k, B), "peak_source": how,
                    "algorithmic_bytes_per_launch": table[k], "avg_launch_us": round(avg_ms * 1e3, 2),
                    "launches_per_step": v["scopes"], "share_of_step": round(v["ms"] / ms_per_step, 4),
                    "timing": "CUDA events on the launching stream around every launch of this kernel inside %d full QAT "
                              "steps (csrc/prof.cu)" % repeats,
                    "kernel_time_sum_ms_per_step": round(total, 3)}
            return roof, breakdown
    return {"error": "no known kernel class in the profile"}, breakdown


def all_kernel_fractions(step_fn, B, M, repeats=2):
    """Text table: every timed kernel class with its HBM fraction where a byte formula exists (profiles/)."""
    prof = profile(step_fn, repeats)
    table = algorithmic_bytes(B, M)
    peak, how = measured_peaks()
    total = sum(v["ms"] for v in prof.values())
    lines = ["# per-kernel attribution of one QAT step (B=%d/GPU, M=%d), CUDA events on the launching stream" % (B, M),
             "# HBM peak %.1f GB/s, %s; kernel time sum %.2f ms/step" % (peak, how, total),
             "%-28s %8s %9s %10s %7s %9s %6s" % ("kernel", "n/step", "ms/step", "avg us", "share", "GB/s", "frac")]
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        avg = v["ms"] / max(v["scopes"], 1e-9)
        if k in table:
            gbs = table[k] / (avg * 1e-3) / 1e9
            tail = "%9.0f %6.3f" % (gbs, gbs / peak)
        else:
            tail = "%9s %6s" % ("-", "-")
        lines.append("%-28s %8.1f %9.3f %10.1f %6.1f%% %s" % (k, v["scopes"], v["ms"], avg * 1e3, 100 * v["ms"] / total, tail))
    return "\n".join(lines)
