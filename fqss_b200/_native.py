"""ctypes binding of libfqss_sm100.so (the C ABI declared in include/fqss.h).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised.
Everything here is plumbing (pointers, pitches, streams); the arithmetic lives in csrc/*.cu.
"""
import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# FQSS_LIB_PATH: development knob -- A/B runs of two builds of the library inside one GPU session
LIB_PATH = os.environ.get("FQSS_LIB_PATH") or os.path.join(_HERE, "_lib", "libfqss_sm100.so")

_lib = None
_lock = threading.Lock()

c_f32p = C.c_void_p
i64 = C.c_int64
i32 = C.c_int
vp = C.c_void_p
sz = C.c_size_t
f32 = C.c_float
f64 = C.c_double


class PwDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("quant", C.c_int32), ("n_bits", C.c_int32), ("C", C.c_int32),
                ("bcast", C.c_int32), ("_pad", C.c_int32), ("rows", i64), ("cols", i64),
                ("x1", vp), ("ld1", i64), ("x2", vp), ("ld2", i64), ("y", vp), ("ldy", i64),
                ("slope", vp), ("gamma", vp), ("beta", vp), ("stats", vp), ("eps", f32), ("_pad2", f32),
                ("rmin", vp), ("rmax", vp)]


class WqItem(C.Structure):
    _fields_ = [("g", vp), ("w", vp), ("out", vp), ("g_rmin", vp), ("g_rmax", vp), ("rmin", vp), ("rmax", vp),
                ("outer", C.c_int32), ("ch", C.c_int32), ("inner", C.c_int32), ("n_bits", C.c_int32)]


class PrepItem(C.Structure):
    _fields_ = [("W", vp), ("wmin", vp), ("wmax", vp), ("bias", vp), ("amin", vp), ("amax", vp),
                ("Wc", vp), ("WcT", vp), ("s1", vp), ("s0", vp), ("dws", vp),
                ("N", C.c_int32), ("K", C.c_int32), ("Ntot", C.c_int32), ("n_off", C.c_int32), ("split", C.c_int32),
                ("_pad", C.c_int32)]


class GatherItem(C.Structure):
    _fields_ = [("src", vp), ("offset", i64), ("numel", i64)]


class PwGrads(C.Structure):
    _fields_ = [("g", vp), ("ldg", i64), ("gx1", vp), ("ldg1", i64), ("gx2", vp), ("ldg2", i64),
                ("g_rmin", vp), ("g_rmax", vp), ("g_slope", vp), ("g_gamma", vp), ("g_beta", vp)]


PW_IDENT, PW_PRELU, PW_RELU, PW_ADD, PW_SUB, PW_MUL, PW_GLN = range(7)

_SIGS = {
    "fqss_abi_version": (i32, []),
    "fqss_last_error": (C.c_char_p, []),
    "fqss_ws_bytes": (sz, [i64]),
    "fqss_fq_act_fwd": (i32, [vp, vp, vp, i64, vp, vp, i32, vp]),
    "fqss_fq_act_bwd": (i32, [vp, vp, vp, vp, vp, i64, vp, vp, i32, vp, sz, vp]),
    "fqss_fq_weight_fwd": (i32, [vp, vp, vp, i32, i32, i32, vp, vp, i32, vp]),
    "fqss_fq_weight_bwd": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, i32, vp]),
    "fqss_weight_observe": (i32, [vp, i32, i32, i32, vp, vp, vp]),
    "fqss_act_observe": (i32, [vp, i64, i64, i64, vp, vp, f64, vp, sz, vp]),
    "fqss_pw_fwd": (i32, [C.POINTER(PwDesc), vp]),
    "fqss_pw_bwd": (i32, [C.POINTER(PwDesc), C.POINTER(PwGrads), vp, sz, vp]),
    "fqss_gln_stats": (i32, [vp, i64, i64, i64, i32, vp, vp]),
    "fqss_conv1x1_fwd": (i32, [vp, i64, vp, vp, vp, i64, i32, i32, i32, i32, vp]),
    "fqss_conv1x1_dgrad": (i32, [vp, i64, vp, vp, i64, i32, i32, i32, i32, vp]),
    "fqss_conv1x1_wgrad": (i32, [vp, i64, vp, i64, vp, vp, i32, i32, i32, i32, vp, sz, vp]),
    "fqss_dwconv_fwd": (i32, [vp, i64, vp, vp, vp, i64, i32, i32, i32, i32, i32, vp]),
    "fqss_dwconv_bwd": (i32, [vp, i64, vp, i64, vp, vp, i64, vp, vp, i32, i32, i32, i32, i32, vp, sz, vp]),
    "fqss_sconv_fwd": (i32, [vp, i64, vp, vp, i64, i32, i32, i32, i32, i32, i32, vp]),
    "fqss_sconv_bwd": (i32, [vp, i64, vp, i64, vp, vp, i64, vp, i32, i32, i32, i32, i32, i32, vp, sz, vp]),
    "fqss_tconv_fwd": (i32, [vp, i64, vp, vp, i64, i32, i32, i32, i32, i32, vp]),
    "fqss_tconv_bwd": (i32, [vp, i64, vp, i64, vp, vp, i64, vp, i32, i32, i32, i32, i32, vp, sz, vp]),
    "fqss_pw_gemm": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i64, i32, vp]),
    "fqss_pw_gemm_ex": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i64, i32, vp]),
    "fqss_split_bf16": (i32, [vp, i64, vp, i64, i64, i32, i32, i32, vp]),
    "fqss_tcn_prep": (i32, [vp] * 11 + [i32] * 5 + [vp]),
    "fqss_tcn_prep_fold": (i32, [vp] * 7 + [i32] * 4 + [vp]),
    "fqss_fq_weight_fwd_batch": (i32, [C.POINTER(WqItem), i32, vp]),
    "fqss_fq_weight_bwd_batch": (i32, [C.POINTER(WqItem), i32, vp]),
    "fqss_tcn_prep_batch": (i32, [C.POINTER(PrepItem), i32, vp]),
    "fqss_rowscale_bf16": (i32, [vp, i64, vp, i64, i64, i32, i32, vp, vp, vp]),
    "fqss_wgrad_codes_ws_bytes": (sz, [i32, i32, i32, i32]),
    "fqss_wgrad_codes": (i32, [vp, vp, i32, i32, i64, i32, i32, vp, vp, vp, vp, vp, vp, sz, vp]),
    "fqss_tcn_encode": (i32, [vp, i64, vp, i64, i64, i32, vp, vp, vp]),
    "fqss_tcn_block_fwd": (i32, [vp, vp]),
    "fqss_tcn_block_bwd": (i32, [vp, vp, vp]),
    "fqss_set_wgrad_overlap": (i32, [i32]),
    "fqss_tcn_ws_bytes": (sz, [i32, i32, i32]),
    "fqss_absmax": (i32, [vp, i64, i64, i64, vp, vp, sz, vp]),
    "fqss_split": (i32, [vp, i64, vp, vp, i64, i32, i32, i32, i32, vp]),
    "fqss_combine": (i32, [vp, i64, i64, vp, i64, i64, i32, i32, i32, vp]),
    "fqss_kd_loss": (i32, [vp, i64, vp, i64, vp, i64, i32, i32, f32, vp, vp, i64, vp, sz, vp]),
    "fqss_kd_loss_dp": (i32, [vp, i64, vp, i64, vp, i64, i32, i32, f32, vp, vp, i64, vp, sz, vp, i32, vp]),
    "fqss_loss_stats": (i32, [vp, i64, vp, i64, vp, i64, i32, i32, vp, vp]),
    "fqss_loss_grad_apply": (i32, [vp, i64, vp, i64, vp, i64, i32, i32, vp, vp, i64, vp]),
    "fqss_arena_gather": (i32, [C.POINTER(GatherItem), i32, vp, vp]),
    "fqss_arena_sumsq": (i32, [vp, i64, vp, vp, sz, vp]),
    "fqss_arena_scale_clip": (i32, [vp, i64, vp, f32, f32, vp]),
    "fqss_arena_adam": (i32, [vp, vp, vp, vp, i64, f32, f32, f32, f32, i32, vp]),
    "fqss_arena_adam_dev": (i32, [vp, vp, vp, vp, i64, f32, f32, f32, f32, vp, vp, vp]),
    "fqss_music_loss_ws_bytes": (sz, [i32]),
    "fqss_music_kd_loss": (i32, [vp, i64, vp, i64, vp, i64, i32, i32, i32, f32, vp, vp, i64, vp, sz, vp]),
    "fqss_mask_head_fwd": (i32, [vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i64, vp]),
    "fqss_pw_gemm_nstore": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i64, vp]),
    "fqss_frames_split": (i32, [vp, i64, vp, i64, i64, i32, i32, i32, i32, vp, vp]),
    "fqss_frames_encode": (i32, [vp, i64, vp, i64, i64, i32, i32, i32, i32, i32, vp, vp, vp]),
    "fqss_set_bwd_tail_side": (i32, [i32]),
    "fqss_tcn_bwd_join": (i32, [vp]),
    "fqss_attn_smem_bytes": (sz, [i32, i32, i32]),
    "fqss_attn_fwd": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "fqss_attn_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "fqss_lstm_rec_fwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "fqss_lstm_rec_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "fqss_edge_dec_prep": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp]),
    "fqss_edge_enc_prep": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "fqss_sub_fq_codes": (i32, [vp, i64, vp, i64, vp, i64, i64, i32, vp, vp, vp]),
    "fqss_dec_wgrad_fold": (i32, [vp, vp, i32, i32, vp]),
    "fqss_mask_head_ws_bytes": (sz, [i32]),
    "fqss_mask_head_bwd": (i32, [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i64, vp, sz, vp]),
    "fqss_fq_affine_tensor": (i32, [vp, vp, vp, vp, i64, f32, i32, i32, i32, vp]),
    "fqss_fq_affine_channel": (i32, [vp, vp, vp, vp, i32, i32, i32, vp, i32, i32, vp]),
    "fqss_fq_affine_bwd": (i32, [vp, vp, vp, i64, vp]),
    "fqss_prof_enable": (i32, [i32]),
    "fqss_prof_reset": (i32, []),
    "fqss_prof_nslots": (i32, []),
    "fqss_split_ex": (i32, [vp, i64, vp, vp, i64, i32, i32, i32, i32, i32, i32, vp]),
    "fqss_cln_fwd": (i32, [vp, i64, vp, vp, f32, vp, i64, vp, vp, i32, i32, i32, vp]),
    "fqss_cln_bwd": (i32, [vp, i64, vp, i64, vp, vp, vp, vp, i64, vp, vp, i32, i32, i32, vp, sz, vp]),
    "fqss_ola_fwd": (i32, [vp, i64, vp, i64, i64, i32, i32, i32, i32, vp]),
    "fqss_ola_bwd": (i32, [vp, i64, vp, i64, i64, i32, i32, i32, i32, vp]),
    "fqss_prof_read": (i32, [i32, C.c_char_p, i32, C.POINTER(f64), C.POINTER(i64), C.POINTER(i64)]),
    "fqss_launch_count": (i64, []),
}

EXPORTED = tuple(_SIGS.keys())


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        "fqss_b200: %s not found -- run `python -m fqss_b200.build` (there is no CPU or "
                        "PyTorch fallback for the CUDA path)" % LIB_PATH)
                L = C.CDLL(LIB_PATH)
                for name, (res, args) in _SIGS.items():
                    fn = getattr(L, name)
                    fn.restype = res
                    fn.argtypes = args
                if L.fqss_abi_version() != 21:
                    raise RuntimeError("fqss_b200: ABI version mismatch (%d)" % L.fqss_abi_version())
                _lib = L
    return _lib


class FqssError(RuntimeError):
    pass


launch_count = 0      # kernels-launching C-ABI calls made (bench.py reports it)


def check(rc):
    global launch_count
    launch_count += 1
    if rc != 0:
        raise FqssError("libfqss_sm100: error %d: %s" % (rc, lib().fqss_last_error().decode()))


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise FqssError("fqss_b200 ops need CUDA tensors (no CPU fallback); got a %s tensor" % t.device)


# ---------------------------------------------------------------------------------------------
# workspace: one zero-initialised scratch buffer per (device, stream); calls on a stream are ordered
# ---------------------------------------------------------------------------------------------
_ws = {}


def workspace(rows, device):
    need = int(lib().fqss_ws_bytes(int(rows)))
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.zeros(max(need, 1 << 20), dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf


def ptr(t):
    return 0 if t is None else t.data_ptr()
