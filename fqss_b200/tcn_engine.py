"""Fused TCN engine (host side): drives the tensor-core / fused kernels of libfqss_sm100 for the
ConvBlocks of the separator (`MaskGenerator.TCN` + `MaskGenerator.adds`).  Plumbing only -- buffer
allocation, pointer tables, launch order; all arithmetic is in csrc/tcn_fwd.cu, tcn_bwd.cu,
gemm_tc.cu, wgrad_tc.cu.

The whole stack is ONE autograd node: forward launches 4 kernels (+1 one-warp constants launch) per block,
backward 12, and the residual-stream / skip-sum gradients never leave fp32.  Forward-only calls (no_grad)
take `_fused_tcn_infer`: same kernels without the stores only backward needs.  Used when the quantised
model is in steady state (observers off); the per-layer wrappers in qat_layers.py remain the general
path (observer calibration, foreign compositions) and the definition of the drop-in API.
"""
import ctypes as C
import os

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _native as N
from ._native import check, i64, lib, ptr, stream_ptr, vp

f32p = vp


class QRange(C.Structure):
    _fields_ = [("rmin", vp), ("rmax", vp)]


class TcnBlock(C.Structure):
    _fields_ = [("B", C.c_int32), ("M", C.c_int32), ("dil", C.c_int32), ("quant", C.c_int32), ("first_block", C.c_int32),
                ("has_res", C.c_int32), ("Cio", C.c_int32), ("Chid", C.c_int32), ("split", C.c_int32), ("no_skip", C.c_int32),
                ("ld", i64),
                ("Wc1", vp), ("Wc1T", vp), ("s1_1", vp), ("s0_1", vp), ("dws1", vp),
                ("Wc2", vp), ("Wc2T", vp), ("s1_2", vp), ("s0_2", vp), ("dws2", vp),
                ("wdw", vp), ("bdw", vp),
                ("slope1", vp), ("slope3", vp), ("gn1_w", vp), ("gn1_b", vp), ("gn2_w", vp), ("gn2_b", vp),
                ("q_in", QRange), ("q1", QRange), ("q2", QRange), ("q3", QRange), ("q4", QRange), ("qres", QRange),
                ("qskip", QRange), ("qadd", QRange), ("qadds", QRange),
                ("x_op", vp), ("x_in", vp), ("skip_in", vp),
                ("y1", vp), ("stats1", vp), ("y3", vp), ("stats3", vp), ("a4_op", vp),
                ("res_y", vp), ("skip_y", vp), ("x_out", vp), ("x_out_op", vp), ("skip_out", vp), ("rc1", vp), ("rc3", vp), ("code1", vp), ("code3", vp)]


class TcnBlockGrads(C.Structure):
    _fields_ = [("g_x_out", vp), ("g_skip_out", vp), ("g_x_in", vp), ("g_skip_in", vp),
                ("dY2", vp), ("g_hid_a", vp), ("g_hid_b", vp), ("dY1", vp), ("g_xd", vp),
                ("dW1q", vp), ("db1", vp), ("dW2q", vp), ("db2", vp), ("dwdw", vp), ("dbdw", vp),
                ("g_gn1_w", vp), ("g_gn1_b", vp), ("g_gn2_w", vp), ("g_gn2_b", vp), ("g_slope1", vp), ("g_slope3", vp),
                ("g_q", vp), ("ws", vp), ("ws_bytes", C.c_size_t)]


_L = None


BWD_TAIL_SIDE = os.environ.get("FQSS_BWD_TAIL_SIDE", "1") not in ("", "0")     # see fqss_set_bwd_tail_side (include/fqss.h)


def _libx():
    global _L
    if _L is None:
        L = lib()
        L.fqss_set_bwd_tail_side(1 if BWD_TAIL_SIDE else 0)
        L.fqss_tcn_block_fwd.argtypes = [C.POINTER(TcnBlock), vp]
        L.fqss_tcn_block_bwd.argtypes = [C.POINTER(TcnBlock), C.POINTER(TcnBlockGrads), vp]
        _L = L
    return _L


def pw_gemm(act_bf16, w_bf16, s1, s0, M, addend=None, out_dtype=torch.float32, mul=None):
    """out[b,o,m] = s1[o] * sum_k act[b, k % rows, m] w[o,k] + s0[o] (+ addend) on tcgen05; act [B,rows,ld], w [N,K]
    bf16 (rows < K: split [hi ; lo] activations against [hi | hi | lo] weights, see include/fqss.h)."""
    N.require_cuda(act_bf16, w_bf16, s1, s0, addend)
    B, a_rows, ld = act_bf16.shape
    Nn, K = w_bf16.shape
    assert act_bf16.is_contiguous() and w_bf16.is_contiguous() and a_rows <= K
    out = torch.empty((B, Nn, ld), device=act_bf16.device, dtype=out_dtype)
    if mul is not None:          # out = relu(.) * mul[b, o % C, m]; mul: fp32 [B,C,M] with row pitch ld
        Cm = mul.shape[1]
        assert mul.stride(2) == 1 and mul.stride(1) == ld and mul.stride(0) == Cm * ld and out_dtype == torch.float32
        check(lib().fqss_pw_gemm_ex(ptr(act_bf16), ptr(w_bf16), ptr(s1), ptr(s0), ptr(out), None, ptr(mul), Cm, B, K, Nn, M, ld,
                                    a_rows, stream_ptr()))
        return out
    f32 = out if out_dtype == torch.float32 else None
    b16 = out if out_dtype == torch.bfloat16 else None
    check(lib().fqss_pw_gemm(ptr(act_bf16), ptr(w_bf16), ptr(s1), ptr(s0), ptr(f32) or None, ptr(b16) or None,
                             ptr(addend) or None, B, K, Nn, M, ld, a_rows, stream_ptr()))
    return out


def split_bf16_weights(w):
    """fp32 [N,K] -> bf16 [N,3K] = [hi | hi | lo]."""
    w = w.detach().reshape(w.shape[0], -1).contiguous()
    out = torch.empty((w.shape[0], 3 * w.shape[1]), dtype=torch.bfloat16, device=w.device)
    check(lib().fqss_split_bf16(ptr(w), w.shape[1], ptr(out), 0, w.shape[0], w.shape[1], 1, 0, stream_ptr()))
    return out


def split_bf16_acts(x, ld):
    """fp32 [B,C,M] (any row pitch) -> bf16 [B,2C,ld] = per sample [hi rows ; lo rows]."""
    B, Cc, M = x.shape
    if x.stride(2) != 1 or x.stride(0) != Cc * x.stride(1):
        x = x.contiguous()
    out = torch.empty((B, 2 * Cc, ld), dtype=torch.bfloat16, device=x.device)
    check(lib().fqss_split_bf16(ptr(x), x.stride(1), ptr(out), ld, B * Cc, M, Cc, 1, stream_ptr()))
    return out


# ---------------------------------------------------------------------------------------------
# parameter layout of one ConvBlock as seen by the engine
# ---------------------------------------------------------------------------------------------
# order of the per-block tensors handed to the autograd node (and of the gradients it returns)
_BLOCK_SLOTS = ("W1", "b1", "w1min", "w1max", "slope1", "q1min", "q1max", "g1w", "g1b", "q2min", "q2max",
                "Wdw", "bdw", "wdmin", "wdmax", "slope3", "q3min", "q3max", "g2w", "g2b", "q4min", "q4max",
                "Wres", "bres", "wrmin", "wrmax", "qresmin", "qresmax", "Wskip", "bskip", "wsmin", "wsmax",
                "qskipmin", "qskipmax", "qaddmin", "qaddmax", "qaddsmin", "qaddsmax")


_ZERO_BIAS = {}


def zero_bias(n, dev):
    """Zero vector standing in for the bias of a bias-free depthwise conv (the music model's): the row kernels read bdw
    unconditionally (a NULL test in the kernel cost the speech model's float kernel 2 us per launch)."""
    key = (n, dev)
    z = _ZERO_BIAS.get(key)
    if z is None:
        z = _ZERO_BIAS[key] = torch.zeros(n, device=dev)
    return z


def _aq(mod):
    q = getattr(mod, "activation_fake_quantize", None)
    if q is None or isinstance(q, torch.nn.Identity):
        return None, None
    return q.min_range, q.max_range


def _wq(mod):
    q = getattr(mod, "weight_fake_quantize", None)
    if q is None or isinstance(q, torch.nn.Identity):
        return None, None
    return q.min_range, q.max_range


def block_tensors_noskip(block):
    """dict slot -> tensor (or None) for a quantised skip-less ConvBlock (ConvTasNetMusicQ, convtasnetq_music.py:141-199:
    net = [Conv1dNlQ, -, GroupNormQ, DepthwiseSeparableConv(net = [Conv1dNlQ, -, GroupNormQ, Conv1dQ])], add = AddQ)."""
    c1, n1 = block.net[0], block.net[2]
    ds = block.net[3].net
    dw, n2, res = ds[0], ds[2], ds[3]
    d = dict(W1=c1.conv1d.weight, b1=c1.conv1d.bias, slope1=c1.nl.weight, g1w=n1.groupnorm.weight, g1b=n1.groupnorm.bias,
             Wdw=dw.conv1d.weight, bdw=dw.conv1d.bias, slope3=dw.nl.weight, g2w=n2.groupnorm.weight, g2b=n2.groupnorm.bias,
             Wres=res.conv1d.weight, bres=res.conv1d.bias)
    d["w1min"], d["w1max"] = _wq(c1)
    d["wdmin"], d["wdmax"] = _wq(dw)
    d["wrmin"], d["wrmax"] = _wq(res)
    d["q1min"], d["q1max"] = _aq(c1)
    d["q2min"], d["q2max"] = _aq(n1)
    d["q3min"], d["q3max"] = _aq(dw)
    d["q4min"], d["q4max"] = _aq(n2)
    d["qresmin"], d["qresmax"] = _aq(res)
    d["qaddmin"], d["qaddmax"] = _aq(block.add)
    return {k: d.get(k) for k in _BLOCK_SLOTS}, dw.conv1d.dilation[0]


def block_tensors(block, adds_mod, quant):
    """dict slot -> tensor (or None) for a quantised (`quant`) or float ConvBlock."""
    if not hasattr(block, "shared_block"):
        return block_tensors_noskip(block)
    sb = block.shared_block
    if quant:
        c1, n1, dw, n2 = sb[0], sb[2], sb[3], sb[5]
        d = dict(W1=c1.conv1d.weight, b1=c1.conv1d.bias, slope1=c1.nl.weight, g1w=n1.groupnorm.weight, g1b=n1.groupnorm.bias,
                 Wdw=dw.conv1d.weight, bdw=dw.conv1d.bias, slope3=dw.nl.weight, g2w=n2.groupnorm.weight, g2b=n2.groupnorm.bias,
                 Wres=block.res_conv.conv1d.weight, bres=block.res_conv.conv1d.bias,
                 Wskip=block.skip_conv.conv1d.weight, bskip=block.skip_conv.conv1d.bias)
        d["w1min"], d["w1max"] = _wq(c1)
        d["wdmin"], d["wdmax"] = _wq(dw)
        d["wrmin"], d["wrmax"] = _wq(block.res_conv)
        d["wsmin"], d["wsmax"] = _wq(block.skip_conv)
        d["q1min"], d["q1max"] = _aq(c1)
        d["q2min"], d["q2max"] = _aq(n1)
        d["q3min"], d["q3max"] = _aq(dw)
        d["q4min"], d["q4max"] = _aq(n2)
        d["qresmin"], d["qresmax"] = _aq(block.res_conv)
        d["qskipmin"], d["qskipmax"] = _aq(block.skip_conv)
        d["qaddmin"], d["qaddmax"] = _aq(block.add)
        d["qaddsmin"], d["qaddsmax"] = _aq(adds_mod) if adds_mod is not None else (None, None)
        dil = dw.conv1d.dilation[0]
    else:
        d = dict(W1=sb[0].weight, b1=sb[0].bias, slope1=sb[1].weight, g1w=sb[2].weight, g1b=sb[2].bias,
                 Wdw=sb[3].weight, bdw=sb[3].bias, slope3=sb[4].weight, g2w=sb[5].weight, g2b=sb[5].bias,
                 Wres=block.res_conv.weight, bres=block.res_conv.bias, Wskip=block.skip_conv.weight, bskip=block.skip_conv.bias)
        dil = sb[3].dilation[0]
    return {k: d.get(k) for k in _BLOCK_SLOTS}, dil


class BlockState:
    """Prepared weights + saved activations of one block for one forward/backward."""
    __slots__ = ("t", "dil", "first", "has_res", "prep", "act", "blk")


def _prep_items(t, quant, has_res, q_in, dev, prep_items, wq_items):
    """Allocate the prepared-weight buffers of one block and append its weight-preparation work to the batch lists
    (three 1x1 convs -> fqss_tcn_prep_batch, the depthwise weight -> fqss_fq_weight_fwd_batch)."""
    Chid, Cio = t["W1"].shape[0], t["W1"].shape[1]
    has_skip = t["Wskip"] is not None
    n2 = (Cio if has_res else 0) + (Cio if has_skip else 0)
    bf = torch.bfloat16
    P = dict(Wc1=torch.empty((Chid, Cio), dtype=bf, device=dev), Wc1T=torch.empty((Cio, Chid), dtype=bf, device=dev),
             s1_1=torch.empty(Chid, device=dev), s0_1=torch.empty(Chid, device=dev), dws1=torch.empty(Chid, device=dev),
             Wc2=torch.empty((n2, Chid), dtype=bf, device=dev), Wc2T=torch.empty((Chid, n2), dtype=bf, device=dev),
             s1_2=torch.empty(n2, device=dev), s0_2=torch.empty(n2, device=dev), dws2=torch.empty(n2, device=dev))
    qi = q_in if quant else (None, None)
    q4 = (t["q4min"], t["q4max"]) if quant else (None, None)

    def item(W, wmin, wmax, bias, amin, amax, Wc, WcT, s1, s0, dws, Nn, K, Ntot, off):
        it = N.PrepItem()
        it.W, it.wmin, it.wmax, it.bias = ptr(W), ptr(wmin) or None, ptr(wmax) or None, ptr(bias) or None
        it.amin, it.amax = ptr(amin) or None, ptr(amax) or None
        it.Wc, it.WcT, it.s1, it.s0, it.dws = ptr(Wc), ptr(WcT), ptr(s1), ptr(s0), ptr(dws)
        it.N, it.K, it.Ntot, it.n_off, it.split = Nn, K, Ntot, off, 0
        prep_items.append(it)
    item(t["W1"], t["w1min"], t["w1max"], t["b1"], qi[0], qi[1], P["Wc1"], P["Wc1T"], P["s1_1"], P["s0_1"], P["dws1"], Chid, Cio, Chid, 0)
    off = 0
    if has_res:
        item(t["Wres"], t["wrmin"], t["wrmax"], t["bres"], q4[0], q4[1], P["Wc2"], P["Wc2T"], P["s1_2"], P["s0_2"], P["dws2"], Cio, Chid, n2, 0)
        off = Cio
    if has_skip:
        item(t["Wskip"], t["wsmin"], t["wsmax"], t["bskip"], q4[0], q4[1], P["Wc2"], P["Wc2T"], P["s1_2"], P["s0_2"], P["dws2"], Cio, Chid, n2, off)
    if quant:
        wdw = torch.empty_like(t["Wdw"], memory_format=torch.contiguous_format)
        it = N.WqItem()
        it.w, it.out, it.rmin, it.rmax = ptr(t["Wdw"]), ptr(wdw), ptr(t["wdmin"]), ptr(t["wdmax"])
        it.outer, it.ch, it.inner, it.n_bits = 1, Chid, t["Wdw"].shape[-1], 8
        wq_items.append(it)
    else:
        wdw = t["Wdw"]
    P["wdw"] = wdw
    return P


def _run_batches(prep_items, wq_items, wq_bwd=False):
    L = _libx()
    s = stream_ptr()
    if prep_items:
        arr = (N.PrepItem * len(prep_items))(*prep_items)
        check(L.fqss_tcn_prep_batch(arr, len(prep_items), s))
    if wq_items:
        arr = (N.WqItem * len(wq_items))(*wq_items)
        fn = L.fqss_fq_weight_bwd_batch if wq_bwd else L.fqss_fq_weight_fwd_batch
        check(fn(arr, len(wq_items), s))


def _fill_block(blk, t, P, quant, first, has_res, dil, B, M, ld, q_in):
    blk.B, blk.M, blk.dil, blk.quant, blk.first_block, blk.has_res = B, M, dil, int(quant), int(first), int(has_res)
    blk.Chid, blk.Cio, blk.ld = t["W1"].shape[0], t["W1"].shape[1], ld
    blk.no_skip = int(t["Wskip"] is None)
    for k in ("Wc1", "Wc1T", "s1_1", "s0_1", "dws1", "Wc2", "Wc2T", "s1_2", "s0_2", "dws2", "wdw"):
        setattr(blk, k, ptr(P[k]))
    blk.bdw = ptr(t["bdw"]) if t["bdw"] is not None else ptr(zero_bias(t["W1"].shape[0], t["W1"].device))
    blk.slope1, blk.slope3 = ptr(t["slope1"]), ptr(t["slope3"])
    blk.gn1_w, blk.gn1_b, blk.gn2_w, blk.gn2_b = ptr(t["g1w"]), ptr(t["g1b"]), ptr(t["g2w"]), ptr(t["g2b"])

    def qr(a, b):
        r = QRange()
        r.rmin, r.rmax = (ptr(a) or None), (ptr(b) or None)
        return r
    blk.q_in = qr(*q_in) if quant else qr(None, None)
    for name, lo, hi in (("q1", "q1min", "q1max"), ("q2", "q2min", "q2max"), ("q3", "q3min", "q3max"), ("q4", "q4min", "q4max"),
                         ("qres", "qresmin", "qresmax"), ("qskip", "qskipmin", "qskipmax"), ("qadd", "qaddmin", "qaddmax"),
                         ("qadds", "qaddsmin", "qaddsmax")):
        setattr(blk, name, qr(t[lo], t[hi]))


def _alloc_acts(B, Cio, Chid, ld, has_res, dev, quant=True, has_skip=True):
    bf = torch.bfloat16
    A = dict(y1=torch.empty((B, Chid, ld), device=dev), y3=torch.empty((B, Chid, ld), device=dev),
             stats1=torch.empty(2 * B + 1, dtype=torch.float64, device=dev), stats3=torch.empty(2 * B + 1, dtype=torch.float64, device=dev),
             a4_op=torch.empty((B, Chid, ld), dtype=bf, device=dev), rc1=torch.empty(16 + 2 * B, device=dev),
             rc3=torch.empty(16 + 2 * B, device=dev), skip_y=None, skip_out=None)
    if has_skip:
        A.update(skip_y=torch.empty((B, Cio, ld), device=dev), skip_out=torch.empty((B, Cio, ld), device=dev))
    if quant:
        A.update(code1=torch.empty((B, Chid, ld), dtype=torch.uint8, device=dev),
                 code3=torch.empty((B, Chid, ld), dtype=torch.uint8, device=dev))
    if has_res:
        A.update(res_y=torch.empty((B, Cio, ld), device=dev), x_out=torch.empty((B, Cio, ld), device=dev),
                 x_out_op=torch.empty((B, Cio, ld), dtype=bf, device=dev))
    return A


def _as_pitched(x, ld):
    """[B,C,M] tensor with row pitch exactly `ld` (view when possible, else one repack)."""
    B, Cc, M = x.shape
    if x.stride(2) == 1 and x.stride(1) == ld and x.stride(0) == Cc * ld and x.data_ptr() % 16 == 0:
        return x
    buf = torch.empty((B, Cc, ld), device=x.device, dtype=x.dtype)
    buf[:, :, :M].copy_(x)
    return buf[:, :, :M]


class CodeConv1x1(Function):
    """1x1 conv whose input is the output of an 8-bit activation quantiser and whose weight is fake-quantised per output
    channel (the bottleneck and mask convs, qat_layers.py:137-141, 202-207): both operands become integer codes on the
    tcgen05 GEMM, exactly as inside the ConvBlocks.  Returns the pre-activation y = conv(x, FQ(W)) + bias (fp32).
    Channel counts the tiles do not cover (the music model's 40-row Linear decoder / RQB, convtasnetq_music.py:260) are
    zero-padded to the next multiple of 128: a zero weight has code 0, so the padding adds exact
    zeros to the integer sums."""

    @staticmethod
    def forward(ctx, x, qmin, qmax, W, wmin, wmax, bias):
        N.require_cuda(x, qmin, qmax, W, wmin, wmax, bias)
        L = _libx()
        B, Ci, M = x.shape
        Co = W.shape[0]
        Cip, Cop = (Ci + 127) // 128 * 128, (Co + 127) // 128 * 128      # Cip is also the dgrad GEMM's N
        ld = (M + 7) // 8 * 8
        dev = x.device
        s = stream_ptr()
        xv = _as_pitched(x.detach(), ld)
        bf = torch.bfloat16
        if Cip == Ci:
            x_op = torch.empty((B, Ci, ld), dtype=bf, device=dev)
            check(L.fqss_tcn_encode(ptr(xv), ld, ptr(x_op), ld, B * Ci, M, ptr(qmin), ptr(qmax), s))
        else:
            tmp = torch.empty((B, Ci, ld), dtype=bf, device=dev)
            check(L.fqss_tcn_encode(ptr(xv), ld, ptr(tmp), ld, B * Ci, M, ptr(qmin), ptr(qmax), s))
            x_op = torch.zeros((B, Cip, ld), dtype=bf, device=dev)
            x_op[:, :Ci, :M].copy_(tmp[:, :, :M])
        Wf = W.detach().reshape(Co, Ci)
        wmn, wmx, bs = wmin, wmax, bias
        if (Cip, Cop) != (Ci, Co):
            Wp = torch.zeros((Cop, Cip), device=dev)
            Wp[:Co, :Ci].copy_(Wf)
            Wf = Wp
            wmn, wmx = torch.full((Cop,), -1.0, device=dev), torch.full((Cop,), 1.0, device=dev)
            wmn[:Co].copy_(wmin.detach().reshape(-1))
            wmx[:Co].copy_(wmax.detach().reshape(-1))
            if bias is not None:
                bs = torch.zeros(Cop, device=dev)
                bs[:Co].copy_(bias.detach())
        Wc, WcT = torch.empty((Cop, Cip), dtype=bf, device=dev), torch.empty((Cip, Cop), dtype=bf, device=dev)
        s1, s0, dws = torch.empty(Cop, device=dev), torch.empty(Cop, device=dev), torch.empty(Cop, device=dev)
        check(L.fqss_tcn_prep(ptr(Wf), ptr(wmn), ptr(wmx), ptr(bs) or None, ptr(qmin), ptr(qmax), ptr(Wc), ptr(WcT), ptr(s1),
                              ptr(s0), ptr(dws), Cop, Cip, Cop, 0, 0, s))
        y = pw_gemm(x_op, Wc, s1, s0, M)
        ctx.save_for_backward(x_op, WcT, dws, W, wmin, wmax, qmin, qmax)
        ctx.meta = (B, Ci, Co, M, ld, bias is not None)
        return y[:, :Co, :M]

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        L = _libx()
        x_op, WcT, dws, W, wmin, wmax, qmin, qmax = ctx.saved_tensors
        B, Ci, Co, M, ld, has_bias = ctx.meta
        Cip, Cop = (Ci + 127) // 128 * 128, (Co + 127) // 128 * 128      # Cip is also the dgrad GEMM's N
        dev = gy.device
        s = stream_ptr()
        if Cop == Co:
            gy = _as_pitched(gy, ld)
        else:
            gp = torch.zeros((B, Cop, ld), device=dev)
            gp[:, :Co, :M].copy_(gy)
            gy = gp
        dY = torch.empty((B, Cop, ld), dtype=torch.bfloat16, device=dev)
        db = torch.empty(Cop, dtype=torch.float64, device=dev)
        check(L.fqss_rowscale_bf16(ptr(gy), ld, ptr(dY), ld, B * Cop, M, Cop, ptr(dws), ptr(db), s))
        gx = None
        if ctx.needs_input_grad[0]:
            gx = pw_gemm(dY, WcT, torch.ones(Cip, device=dev), torch.zeros(Cip, device=dev), M)[:, :Ci, :M]
        dWq = torch.empty((Cop, Cip), device=dev)
        ws = torch.empty(int(L.fqss_wgrad_codes_ws_bytes(B, M, Cop, Cip)), dtype=torch.uint8, device=dev)
        check(L.fqss_wgrad_codes(ptr(dY), ptr(x_op), B, M, ld, Cop, Cip, ptr(qmin), ptr(qmax), ptr(dws), ptr(db), ptr(dWq), ptr(ws),
                                 ws.numel(), s))
        if (Cip, Cop) != (Ci, Co):
            dWq = dWq[:Co, :Ci].contiguous()
            db = db[:Co]
        gW = torch.empty_like(W, memory_format=torch.contiguous_format)
        gwmin, gwmax = torch.empty_like(wmin), torch.empty_like(wmax)
        check(lib().fqss_fq_weight_bwd(ptr(dWq), ptr(W), ptr(gW), ptr(gwmin), ptr(gwmax), 1, Co, Ci, ptr(wmin), ptr(wmax), 8, s))
        return gx, None, None, gW, gwmin, gwmax, (db.float() if has_bias else None)


def code_conv_eligible(layer, q_in, x):
    """True when a Conv1dQ / Conv1dNlQ fed by activation quantiser `q_in` can run as a code-operand tcgen05 GEMM."""
    from .qat import qat_layers as QL
    from .qat.qat_quant import GradientActivationFakeQuantize as AQ, GradientWeightFakeQuantize as WQm
    if not isinstance(layer, (QL.Conv1dQ, QL.Conv1dNlQ)) or not x.is_cuda or x.dtype != torch.float32 or x.dim() != 3:
        return False
    conv, wq = layer.conv1d, layer.weight_fake_quantize
    if conv.kernel_size[0] != 1 or conv.stride[0] != 1 or conv.padding[0] != 0 or conv.groups != 1:
        return False
    if not isinstance(wq, WQm) or wq.observer_mode or wq.n_bits != 8 or wq.min_range.numel() != conv.out_channels:
        return False
    if not isinstance(q_in, AQ) or q_in.observing() or q_in.n_bits != 8:
        return False
    # channel counts that are not tile multiples are zero-padded by CodeConv1x1; worth it only where the tensor is large
    # enough for the SIMT kernel to hurt (the music model's 40-row decoder / RQB convs over 256 K frames)
    aligned = conv.in_channels % 64 == 0 and conv.out_channels % 128 == 0
    padded = min(conv.in_channels, conv.out_channels) >= 32 and x.shape[0] * x.shape[2] >= (1 << 16)
    return (aligned or padded) and conv.out_channels <= 1024 and W_contig(conv.weight)


def code_linear_eligible(weight, wq, q_in, x):
    """Same test as code_conv_eligible for a bare [out, in] weight with its quantiser (the music model's Linear decoder / RQB)."""
    from .qat.qat_quant import GradientActivationFakeQuantize as AQ, GradientWeightFakeQuantize as WQm
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 3 or weight.dim() != 2 or x.shape[1] != weight.shape[1]:
        return False
    if not isinstance(wq, WQm) or wq.observer_mode or wq.n_bits != 8 or wq.min_range.numel() != weight.shape[0]:
        return False
    if not isinstance(q_in, AQ) or q_in.observing() or q_in.n_bits != 8:
        return False
    Co, Ci = weight.shape
    aligned = Ci % 64 == 0 and Co % 128 == 0
    padded = min(Ci, Co) >= 32 and x.shape[0] * x.shape[2] >= (1 << 16)
    return (aligned or padded) and Co <= 1024 and weight.is_contiguous()


def W_contig(w):
    return w.is_contiguous()


def code_conv(layer, q_in, x):
    """conv part of `layer.forward(x)` on the code-operand GEMM (caller applies the layer's nl + activation quantiser)."""
    conv, wq = layer.conv1d, layer.weight_fake_quantize
    return CodeConv1x1.apply(x, q_in.min_range, q_in.max_range, conv.weight, wq.min_range, wq.max_range, conv.bias)


def float_conv_eligible(conv, x):
    """Un-quantised 1x1 conv, forward only (no gradient wanted anywhere): the split-bf16 tcgen05 GEMM applies."""
    if torch.is_grad_enabled() and (x.requires_grad or conv.weight.requires_grad):
        return False
    if conv.kernel_size[0] != 1 or conv.stride[0] != 1 or conv.padding[0] != 0 or conv.groups != 1 or x.dtype != torch.float32:
        return False
    padded = conv.out_channels >= 32 and x.shape[0] * x.shape[2] >= (1 << 16)          # output rows zero-padded to 128 (float_conv)
    return conv.in_channels % 64 == 0 and (conv.out_channels % 128 == 0 or padded) and conv.out_channels <= 1024 and conv.weight.is_contiguous()


_FLOAT_W = {}


def float_conv(conv, x):
    """y = conv1x1(x) with fp32-grade accuracy on the bf16 tensor pipe: x = hi + lo, three-term product (float_engine's scheme)."""
    B, Ci, M = x.shape
    ld = (M + 7) // 8 * 8
    w = conv.weight
    from . import parallel
    gen = parallel.param_generation() if getattr(w, "_fqss_in_arena", False) else 0
    key = (w.data_ptr(), w._version, gen, conv.bias.data_ptr() if conv.bias is not None else 0)
    ent = _FLOAT_W.get(id(conv))
    if ent is None or ent[0] != key:
        Co = w.shape[0]
        Cop = (Co + 127) // 128 * 128
        wf = w.detach().reshape(Co, -1)
        if Cop != Co:
            wp = torch.zeros((Cop, wf.shape[1]), device=w.device)
            wp[:Co].copy_(wf)
            wf = wp
        ones = torch.ones(Cop, device=w.device)
        b = torch.zeros(Cop, device=w.device)
        if conv.bias is not None:
            b[:Co].copy_(conv.bias.detach())
        ent = (key, split_bf16_weights(wf), ones, b)
        _FLOAT_W[id(conv)] = ent
    y = pw_gemm(split_bf16_acts(x.detach(), ld), ent[1], ent[2], ent[3], M)
    return y[:, :w.shape[0], :M]


def float_linear_eligible(weight, x):
    """Un-quantised per-frame Linear layer as a channels-first 1x1 conv, forward only (the music teacher's decoder)."""
    if torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad):
        return False
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 3 or weight.dim() != 2 or x.shape[1] != weight.shape[1]:
        return False
    Co, Ci = weight.shape
    padded = Co >= 32 and x.shape[0] * x.shape[2] >= (1 << 16)
    return Ci % 64 == 0 and (Co % 128 == 0 or padded) and Co <= 1024 and weight.is_contiguous()


class _BareConv:
    """Adapter: float_conv() keys its split-weight cache on a module with .weight / .bias."""
    __slots__ = ("weight", "bias")

    def __init__(self, weight):
        self.weight, self.bias = weight, None


_BARE = {}


def float_linear(weight, x):
    ent = _BARE.get(id(weight))
    if ent is None or ent.weight is not weight:
        ent = _BARE[id(weight)] = _BareConv(weight)
    return float_conv(ent, x)


class MaskHead(Function):
    """Mask head of the quantised separator (convtasnetq.py:97-99, :203): 1x1 conv bn -> S*F on integer-code operands, ReLU +
    FQ_m and `* feats` + FQ_p in the GEMM epilogue (fqss_mask_head_fwd); backward of the whole elementwise tail in one pass
    (fqss_mask_head_bwd) feeding the dgrad / wgrad GEMMs.  x: output of the preceding 8-bit quantiser (qmin, qmax)."""

    @staticmethod
    def forward(ctx, x, qmin, qmax, W, wmin, wmax, bias, qm_min, qm_max, feats, qp_min, qp_max, want_codes=False):
        N.require_cuda(x, qmin, qmax, W, wmin, wmax, bias, qm_min, qm_max, feats, qp_min, qp_max)
        L = _libx()
        B, Ci, M = x.shape
        Co, Cf = W.shape[0], feats.shape[1]
        ld = (M + 7) // 8 * 8
        dev = x.device
        s = stream_ptr()
        xv = _as_pitched(x.detach(), ld)
        fv = _as_pitched(feats.detach(), ld)
        x_op = torch.empty((B, Ci, ld), dtype=torch.bfloat16, device=dev)
        check(L.fqss_tcn_encode(ptr(xv), ld, ptr(x_op), ld, B * Ci, M, ptr(qmin), ptr(qmax), s))
        bf = torch.bfloat16
        Wc, WcT = torch.empty((Co, Ci), dtype=bf, device=dev), torch.empty((Ci, Co), dtype=bf, device=dev)
        s1, s0, dws = torch.empty(Co, device=dev), torch.empty(Co, device=dev), torch.empty(Co, device=dev)
        check(L.fqss_tcn_prep(ptr(W.detach().reshape(Co, Ci)), ptr(wmin), ptr(wmax), ptr(bias) or None, ptr(qmin), ptr(qmax), ptr(Wc),
                              ptr(WcT), ptr(s1), ptr(s0), ptr(dws), Co, Ci, Co, 0, 0, s))
        train = any(ctx.needs_input_grad)        # (grad mode is always off inside Function.forward)
        y = torch.empty((B, Co, ld), device=dev) if train else None
        out = torch.empty((B, Co, ld), device=dev)
        # the FQ_p codes of the masked features as a bf16 GEMM operand: what the decoder's tensor-core path reads
        codes = torch.empty((B, Co, ld), dtype=bf, device=dev) if want_codes else torch.empty(0, dtype=bf, device=dev)
        check(L.fqss_mask_head_fwd(ptr(x_op), ptr(Wc), ptr(s1), ptr(s0), ptr(fv), Cf, ptr(qm_min), ptr(qm_max), ptr(qp_min),
                                   ptr(qp_max), ptr(y) or None, ptr(out), ptr(codes) if want_codes else None, B, Ci, Co, M, ld, s))
        ctx.mark_non_differentiable(codes)
        ctx.save_for_backward(x_op, WcT, dws, W, wmin, wmax, qmin, qmax, y, fv, qm_min, qm_max, qp_min, qp_max)
        ctx.meta = (B, Ci, Co, Cf, M, ld, bias is not None)
        return out[:, :, :M], codes

    @staticmethod
    @once_differentiable
    def backward(ctx, g, _g_codes=None):
        L = _libx()
        x_op, WcT, dws, W, wmin, wmax, qmin, qmax, y, fv, qm_min, qm_max, qp_min, qp_max = ctx.saved_tensors
        B, Ci, Co, Cf, M, ld, has_bias = ctx.meta
        dev = g.device
        s = stream_ptr()
        g = _as_pitched(g, ld)
        dY = torch.empty((B, Co, ld), dtype=torch.bfloat16, device=dev)
        g_feats = torch.empty((B, Cf, ld), device=dev)
        g_q = torch.empty(4, device=dev)
        g_bias = torch.empty(Co, device=dev) if has_bias else None
        db = torch.empty(Co, dtype=torch.float64, device=dev)
        ws = torch.empty(int(L.fqss_mask_head_ws_bytes(Co)), dtype=torch.uint8, device=dev)
        check(L.fqss_mask_head_bwd(ptr(g), ld, ptr(y), ptr(fv), ptr(dws), ptr(qm_min), ptr(qm_max), ptr(qp_min), ptr(qp_max), ptr(dY),
                                   ptr(g_feats), ptr(g_q), ptr(g_bias) or None, ptr(db), B, Cf, Co // Cf, M, ld, ptr(ws), ws.numel(), s))
        gx = None
        if ctx.needs_input_grad[0]:
            gx = pw_gemm(dY, WcT, torch.ones(Ci, device=dev), torch.zeros(Ci, device=dev), M)[:, :, :M]
        dWq = torch.empty((Co, Ci), device=dev)
        ws2 = torch.empty(int(L.fqss_wgrad_codes_ws_bytes(B, M, Co, Ci)), dtype=torch.uint8, device=dev)
        check(L.fqss_wgrad_codes(ptr(dY), ptr(x_op), B, M, ld, Co, Ci, ptr(qmin), ptr(qmax), ptr(dws), ptr(db), ptr(dWq), ptr(ws2),
                                 ws2.numel(), s))
        gW = torch.empty_like(W, memory_format=torch.contiguous_format)
        gwmin, gwmax = torch.empty_like(wmin), torch.empty_like(wmax)
        check(lib().fqss_fq_weight_bwd(ptr(dWq), ptr(W), ptr(gW), ptr(gwmin), ptr(gwmax), 1, Co, Ci, ptr(wmin), ptr(wmax), 8, s))
        return (gx, None, None, gW, gwmin, gwmax, g_bias, g_q[0:1].reshape(qm_min.shape), g_q[1:2].reshape(qm_max.shape),
                g_feats[:, :, :M], g_q[2:3].reshape(qp_min.shape), g_q[3:4].reshape(qp_max.shape), None)


def mask_head_eligible(conv_layer, q_in, mul_layer, h, feats):
    """True when `mul_layer(conv_layer(h), feats)` (ReLU mask conv + MulQ, both 8-bit, steady state) can run as MaskHead."""
    from .qat import qat_layers as QL
    from .qat.qat_quant import GradientActivationFakeQuantize as AQ
    if not isinstance(conv_layer, QL.Conv1dNlQ) or not isinstance(conv_layer.nl, torch.nn.ReLU) or not isinstance(mul_layer, QL.MulQ):
        return False
    qm, qp = conv_layer.activation_fake_quantize, mul_layer.activation_fake_quantize
    for q in (qm, qp):
        if not isinstance(q, AQ) or q.observing() or q.n_bits != 8:
            return False
    if not code_conv_eligible(conv_layer, q_in, h):
        return False
    if not feats.is_cuda or feats.dtype != torch.float32 or feats.dim() != 3 or feats.shape[0] != h.shape[0] or feats.shape[2] != h.shape[2]:
        return False
    Cf = feats.shape[1]
    return Cf % 32 == 0 and conv_layer.conv1d.out_channels % Cf == 0


def mask_head(conv_layer, q_in, mul_layer, h, feats, want_codes=False):
    """-> (masked features [B, S*F, M] on the grid of the MulQ quantiser, their bf16 integer codes [B, S*F, ld] or None)."""
    conv, wq = conv_layer.conv1d, conv_layer.weight_fake_quantize
    qm, qp = conv_layer.activation_fake_quantize, mul_layer.activation_fake_quantize
    out, codes = MaskHead.apply(h, q_in.min_range, q_in.max_range, conv.weight, wq.min_range, wq.max_range, conv.bias,
                                qm.min_range, qm.max_range, feats, qp.min_range, qp.max_range, bool(want_codes))
    return out, (codes if want_codes else None)


class FusedTCNFunction(Function):
    @staticmethod
    def forward(ctx, x, skip_in, meta, *flat):
        """x: [B,Cio,M] fake-quantised values entering block `start`; skip_in: running skip sum (None when
        start == 0); meta: (quant, dils, q_in, start, total) -- `start`/`total` place the given blocks inside
        the full stack (block 0 starts the skip sum, block total-1 has no residual output); flat: the
        per-block tensors in _BLOCK_SLOTS order (None where a block has no such parameter)."""
        N.require_cuda(x)
        L = _libx()
        quant, dils, q_in, start, total = meta
        nb = len(dils)
        ns = len(_BLOCK_SLOTS)
        dev = x.device
        B, Cio, M = x.shape
        ld = (M + 7) // 8 * 8
        x = _as_pitched(x.detach(), ld)
        s = stream_ptr()
        x_op = torch.empty((B, Cio, ld), dtype=torch.bfloat16, device=dev)
        qi = q_in if quant else (None, None)
        check(L.fqss_tcn_encode(ptr(x), ld, ptr(x_op), ld, B * Cio, M, ptr(qi[0]) or None, ptr(qi[1]) or None, s))
        # pass 1: weight preparation of ALL blocks in two batched launches (weights and ranges are constant during a step)
        tensors, preps, prep_items, wq_items = [], [], [], []
        cur_q = q_in
        for i in range(nb):
            t = dict(zip(_BLOCK_SLOTS, flat[i * ns:(i + 1) * ns]))
            has_res = (start + i) < total - 1
            tensors.append(t)
            preps.append(_prep_items(t, quant, has_res, cur_q, dev, prep_items, wq_items))
            cur_q = (t["qaddmin"], t["qaddmax"])
        _run_batches(prep_items, wq_items)
        # pass 2: the blocks
        states = []
        cur_x, cur_op = x, x_op
        cur_skip = _as_pitched(skip_in.detach(), ld) if skip_in is not None else None
        cur_q = q_in
        for i in range(nb):
            t, P = tensors[i], preps[i]
            first, has_res = (start + i) == 0, (start + i) < total - 1
            Chid = t["W1"].shape[0]
            has_skip = t["Wskip"] is not None
            A = _alloc_acts(B, Cio, Chid, ld, has_res, dev, quant, has_skip)
            blk = TcnBlock()
            _fill_block(blk, t, P, quant, first, has_res, dils[i], B, M, ld, cur_q)
            blk.x_op, blk.x_in, blk.skip_in = ptr(cur_op), ptr(cur_x), ptr(cur_skip) or None
            for k in ("y1", "stats1", "y3", "stats3", "a4_op", "skip_y", "skip_out", "rc1", "rc3"):
                setattr(blk, k, ptr(A[k]) or None)
            if has_res:
                blk.res_y, blk.x_out, blk.x_out_op = ptr(A["res_y"]), ptr(A["x_out"]), ptr(A["x_out_op"])
            if quant:
                blk.code1, blk.code3 = ptr(A["code1"]), ptr(A["code3"])
            check(L.fqss_tcn_block_fwd(C.byref(blk), s))
            st = BlockState()
            st.t, st.dil, st.first, st.has_res, st.prep, st.act, st.blk = t, dils[i], first, has_res, P, A, blk
            A["x_in"], A["x_op"], A["skip_in"] = cur_x, cur_op, cur_skip
            states.append(st)
            if has_res:
                cur_x, cur_op = A["x_out"], A["x_out_op"]
            cur_skip = A["skip_out"]
            cur_q = (t["qaddmin"], t["qaddmax"])
        ctx.states = states
        ctx.meta = (quant, B, Cio, M, ld)
        ctx.nflat = len(flat)
        if KEEP_STATES:
            LAST_STATES[:] = states
        x_last = states[-1].act.get("x_out")
        if x_last is None:
            x_last = torch.zeros((B, Cio, M), device=dev)          # dead output of the last block
            ctx.mark_non_differentiable(x_last)
            return x_last, cur_skip[:, :, :M]
        if cur_skip is None:                                       # skip-less stack: there is no skip sum
            no_skip_sum = torch.zeros(1, device=dev)
            ctx.mark_non_differentiable(no_skip_sum)
            return x_last[:, :, :M], no_skip_sum
        return x_last[:, :, :M], cur_skip[:, :, :M]

    @staticmethod
    @once_differentiable
    def backward(ctx, g_xo, g_skip):
        L = _libx()
        quant, B, Cio, M, ld = ctx.meta
        states = ctx.states
        ctx.states = None
        dev = (g_skip if g_skip is not None else g_xo).device
        nb = len(states)
        has_skip = states[0].t["Wskip"] is not None
        ns = len(_BLOCK_SLOTS)
        Chid = states[0].t["W1"].shape[0]
        bf = torch.bfloat16
        s = stream_ptr()
        g_ss = torch.zeros((B, Cio, ld), device=dev) if has_skip else None
        if g_skip is not None and has_skip:
            g_ss[:, :, :M].copy_(g_skip)
        g_x = torch.zeros((B, Cio, ld), device=dev)
        if g_xo is not None and states[-1].has_res:
            g_x[:, :, :M].copy_(g_xo)
        # g_hid_b (g_y3) exists only for the two-kernel A/B path; the fused gLN2+depthwise kernel keeps it in shared memory
        two_kernel = os.environ.get("FQSS_SPLIT_P2D", "0") not in ("", "0")
        scratch = dict(dY2=torch.empty((B, (2 if has_skip else 1) * Cio, ld), dtype=bf, device=dev), ga=torch.empty((B, Chid, ld), dtype=bf, device=dev),
                       gb=torch.empty((B, Chid, ld), dtype=bf, device=dev) if two_kernel else None,
                       dY1=torch.empty((B, Chid, ld), dtype=bf, device=dev), gxd=torch.empty((B, Cio, ld), device=dev))
        ws = torch.empty(int(L.fqss_tcn_ws_bytes(B, Cio, Chid)), dtype=torch.uint8, device=dev)
        grads = [None] * ctx.nflat
        wq_items, keep = [], []
        for i in range(nb - 1, -1, -1):
            st = states[i]
            t = st.t
            n2 = (Cio if st.has_res else 0) + (Cio if has_skip else 0)
            G = dict(dW1q=torch.empty((Chid, Cio), device=dev), db1=torch.empty(Chid, device=dev),
                     dW2q=torch.empty((n2, Chid), device=dev), db2=torch.empty(n2, device=dev),
                     dwdw=torch.empty((Chid, 3), device=dev), dbdw=torch.empty(Chid, device=dev),
                     g1w=torch.empty(Chid, device=dev), g1b=torch.empty(Chid, device=dev),
                     g2w=torch.empty(Chid, device=dev), g2b=torch.empty(Chid, device=dev),
                     sl1=torch.empty(1, device=dev), sl3=torch.empty(1, device=dev), gq=torch.zeros(16, device=dev))
            g = TcnBlockGrads()
            g.g_x_out, g.g_skip_out, g.g_x_in, g.g_skip_in = ptr(g_x), ptr(g_ss) or None, ptr(g_x), ptr(g_ss) or None
            g.dY2, g.g_hid_a, g.g_hid_b, g.dY1, g.g_xd = ptr(scratch["dY2"]), ptr(scratch["ga"]), ptr(scratch["gb"]) or None, ptr(scratch["dY1"]), ptr(scratch["gxd"])
            g.dW1q, g.db1, g.dW2q, g.db2, g.dwdw, g.dbdw = ptr(G["dW1q"]), ptr(G["db1"]), ptr(G["dW2q"]), ptr(G["db2"]), ptr(G["dwdw"]), ptr(G["dbdw"])
            g.g_gn1_w, g.g_gn1_b, g.g_gn2_w, g.g_gn2_b = ptr(G["g1w"]), ptr(G["g1b"]), ptr(G["g2w"]), ptr(G["g2b"])
            g.g_slope1, g.g_slope3, g.g_q = ptr(G["sl1"]), ptr(G["sl3"]), ptr(G["gq"])
            g.ws, g.ws_bytes = ptr(ws), ws.numel()
            check(L.fqss_tcn_block_bwd(C.byref(st.blk), C.byref(g), s))
            base = i * ns
            out = {}
            # gradients w.r.t. the fake-quantised weights -> raw weights + per-channel ranges
            def wback(gq, wname, lo, hi, shape):
                w = t[wname]
                if quant:      # queued: ONE batched launch for the weight quantisers of all blocks after the loop
                    gw = torch.empty_like(w, memory_format=torch.contiguous_format)
                    gmin, gmax = torch.empty_like(t[lo]), torch.empty_like(t[hi])
                    it = N.WqItem()
                    it.g, it.w, it.out, it.g_rmin, it.g_rmax = ptr(gq), ptr(w), ptr(gw), ptr(gmin), ptr(gmax)
                    it.rmin, it.rmax = ptr(t[lo]), ptr(t[hi])
                    it.outer, it.ch, it.inner, it.n_bits = 1, w.shape[0], w.numel() // w.shape[0], 8
                    wq_items.append(it)
                    keep.append(gq)
                    return gw, gmin, gmax
                return gq.view(shape), None, None
            out["W1"], out["w1min"], out["w1max"] = wback(G["dW1q"], "W1", "w1min", "w1max", t["W1"].shape)
            out["b1"] = G["db1"]
            out["Wdw"], out["wdmin"], out["wdmax"] = wback(G["dwdw"], "Wdw", "wdmin", "wdmax", t["Wdw"].shape)
            out["bdw"] = G["dbdw"]
            off = 0
            if st.has_res:
                out["Wres"], out["wrmin"], out["wrmax"] = wback(G["dW2q"][:Cio], "Wres", "wrmin", "wrmax", t["Wres"].shape)
                out["bres"] = G["db2"][:Cio]
                off = Cio
            if has_skip:
                out["Wskip"], out["wsmin"], out["wsmax"] = wback(G["dW2q"][off:off + Cio], "Wskip", "wsmin", "wsmax", t["Wskip"].shape)
                out["bskip"] = G["db2"][off:off + Cio]
            out["slope1"], out["slope3"] = G["sl1"], G["sl3"]
            out["g1w"], out["g1b"], out["g2w"], out["g2b"] = G["g1w"], G["g1b"], G["g2w"], G["g2b"]
            if quant:
                gq = G["gq"]
                for j, nm in enumerate(("q1", "q2", "q3", "q4", "qres", "qskip", "qadd", "qadds")):
                    if t[nm + "min"] is not None and not (nm in ("qres", "qadd") and not st.has_res) and not (nm == "qadds" and st.first):
                        out[nm + "min"], out[nm + "max"] = gq[2 * j:2 * j + 1], gq[2 * j + 1:2 * j + 2]
            for j, name in enumerate(_BLOCK_SLOTS):
                if t[name] is not None and name in out and out[name] is not None and ctx.needs_input_grad[3 + base + j]:
                    grads[base + j] = out[name].reshape(t[name].shape)
        check(L.fqss_tcn_bwd_join(s))        # the blocks' weight-gradient tails ran on the library's side stream
        _run_batches([], wq_items, wq_bwd=True)
        del keep
        g_skip_in = g_ss[:, :, :M] if (has_skip and not states[0].first and ctx.needs_input_grad[1]) else None
        return (g_x[:, :, :M], g_skip_in, None) + tuple(grads)


def _fused_tcn_infer(x, skip_in, meta, flat):
    """Forward-only run of the quantised stack (no_grad / nothing requires grad): the stores only backward needs (y1, y3,
    res_y, skip_y) are skipped and ONE set of hidden-width buffers serves all blocks; block inputs / outputs ping-pong.
    Same kernels, same codes, same result as the training forward."""
    L = _libx()
    quant, dils, q_in, start, total = meta
    nb, ns = len(dils), len(_BLOCK_SLOTS)
    dev = x.device
    B, Cio, M = x.shape
    ld = (M + 7) // 8 * 8
    x = _as_pitched(x.detach(), ld)
    s = stream_ptr()
    bf = torch.bfloat16
    x_op0 = torch.empty((B, Cio, ld), dtype=bf, device=dev)
    check(L.fqss_tcn_encode(ptr(x), ld, ptr(x_op0), ld, B * Cio, M, ptr(q_in[0]), ptr(q_in[1]), s))
    tensors, preps, prep_items, wq_items = [], [], [], []
    cur_q = q_in
    for i in range(nb):
        t = dict(zip(_BLOCK_SLOTS, flat[i * ns:(i + 1) * ns]))
        tensors.append(t)
        preps.append(_prep_items(t, True, (start + i) < total - 1, cur_q, dev, prep_items, wq_items))
        cur_q = (t["qaddmin"], t["qaddmax"])
    _run_batches(prep_items, wq_items)
    Chid = tensors[0]["W1"].shape[0]
    hid = dict(stats1=torch.empty(2 * B + 1, dtype=torch.float64, device=dev), stats3=torch.empty(2 * B + 1, dtype=torch.float64, device=dev),
               a4_op=torch.empty((B, Chid, ld), dtype=bf, device=dev), rc1=torch.empty(16 + 2 * B, device=dev),
               rc3=torch.empty(16 + 2 * B, device=dev), code1=torch.empty((B, Chid, ld), dtype=torch.uint8, device=dev),
               code3=torch.empty((B, Chid, ld), dtype=torch.uint8, device=dev))
    xs = [torch.empty((B, Cio, ld), device=dev) for _ in range(2)]
    xops = [torch.empty((B, Cio, ld), dtype=bf, device=dev) for _ in range(2)]
    skips = [torch.empty((B, Cio, ld), device=dev) for _ in range(2)]
    cur_x, cur_op = x, x_op0
    cur_skip = _as_pitched(skip_in.detach(), ld) if skip_in is not None else None
    cur_q = q_in
    for i in range(nb):
        t, P = tensors[i], preps[i]
        first, has_res = (start + i) == 0, (start + i) < total - 1
        blk = TcnBlock()
        _fill_block(blk, t, P, True, first, has_res, dils[i], B, M, ld, cur_q)
        blk.x_op, blk.x_in, blk.skip_in = ptr(cur_op), ptr(cur_x), ptr(cur_skip) or None
        for k in ("stats1", "stats3", "a4_op", "rc1", "rc3", "code1", "code3"):
            setattr(blk, k, ptr(hid[k]))
        blk.y1 = blk.y3 = blk.res_y = blk.skip_y = None
        blk.skip_out = ptr(skips[i & 1])
        if has_res:
            blk.x_out, blk.x_out_op = ptr(xs[i & 1]), ptr(xops[i & 1])
        check(L.fqss_tcn_block_fwd(C.byref(blk), s))
        if has_res:
            cur_x, cur_op = xs[i & 1], xops[i & 1]
        cur_skip = skips[i & 1]
        cur_q = (t["qaddmin"], t["qaddmax"])
    has_last_res = (start + nb - 1) < total - 1
    x_last = cur_x[:, :, :M] if has_last_res else torch.zeros((B, Cio, M), device=dev)
    return x_last, cur_skip[:, :, :M]


def fused_tcn(x, blocks, adds, quant, q_in, start=0, total=None, skip_in=None):
    """Run blocks [start, start+len(blocks)) of a stack of `total` blocks.  blocks: ConvBlock modules; adds: the
    AddQ that merges each block's skip into the running sum (entry i belongs to blocks[i]; None for block 0).
    Returns (x_out, skip_sum)."""
    total = len(blocks) if total is None else total
    flat, dils = [], []
    for i, blk in enumerate(blocks):
        t, dil = block_tensors(blk, adds[i] if adds is not None else None, quant)
        dils.append(dil)
        flat.extend(t[k] for k in _BLOCK_SLOTS)
    meta = (quant, tuple(dils), q_in, start, total)
    needs_grad = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in [x, skip_in] + flat)
    if quant and not needs_grad and INFERENCE_MODE:
        return _fused_tcn_infer(x, skip_in, meta, flat)
    return FusedTCNFunction.apply(x, skip_in, meta, *flat)


KEEP_STATES = False         # True: the training forward leaves its per-block BlockState list (prepared weights, saved
LAST_STATES = []            # activations / codes) in LAST_STATES -- read by the kernel-level parity tests, never by the product
INFERENCE_MODE = True       # False: forward-only calls also run the training forward (tests compare the two)


def rows_fit(M, max_dil):
    """The fused row kernels stage one whole (sample, channel) row plus its dilation halo in shared memory
    (csrc/tcn_fwd.cu `validate_block`): about 22 k frames.  Longer inputs (whole utterances in val.py / infer.py) take
    the per-layer path, which has no such bound."""
    ld = (M + 7) // 8 * 8
    return (ld * 2 + 4 * ((max_dil + 3) & ~3) + 1280) * 4 + ld <= 200 * 1024


def fused_eligible_noskip(masker, x):
    """True when the music model's MaskGenerator (convtasnetq_music.py:53-114) can hand its skip-less block stack to the
    fused engine: quantised, observers off, 8-bit, channel counts the GEMM tiles cover, rows that fit shared memory."""
    from .qat import qat_layers as QL
    from .qat.qat_quant import GradientActivationFakeQuantize as AQ, GradientWeightFakeQuantize as WQm
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 3:
        return False
    net = masker.network
    if not isinstance(net[1], QL.Conv1dQ) or isinstance(net[1].activation_fake_quantize, torch.nn.Identity):
        return False
    blocks = [b for rep in net[2] for b in rep]
    b0 = blocks[0]
    if not hasattr(b0, "net") or not hasattr(b0.net[3], "net"):
        return False
    ds = b0.net[3].net
    if not (isinstance(b0.net[0], QL.Conv1dNlQ) and isinstance(b0.net[2], QL.GroupNormQ) and isinstance(ds[0], QL.Conv1dNlQ)
            and isinstance(ds[2], QL.GroupNormQ) and isinstance(ds[3], QL.Conv1dQ) and isinstance(b0.add, QL.AddQ)):
        return False
    Chid, Cio = b0.net[0].conv1d.weight.shape[0], b0.net[0].conv1d.weight.shape[1]
    if Cio % 128 or Chid % 128 or ds[0].conv1d.kernel_size[0] != 3 or ds[0].conv1d.groups != Chid:
        return False
    if not rows_fit(x.shape[-1], max(b.net[3].net[0].conv1d.dilation[0] for b in blocks)):
        return False
    for blk in blocks:
        for m in blk.modules():
            if isinstance(m, AQ) and (m.observing() or m.n_bits != 8):
                return False
            if isinstance(m, WQm) and (m.observer_mode or m.n_bits != 8):
                return False
            if isinstance(m, QL.LayerQ) and type(m).__name__ in ("Conv1dNlQ", "Conv1dQ") and isinstance(m.weight_fake_quantize, torch.nn.Identity):
                return False
            if isinstance(m, QL.LayerQ) and isinstance(m.activation_fake_quantize, torch.nn.Identity):
                return False
            if isinstance(m, torch.nn.PReLU) and m.weight.numel() != 1:
                return False
    q = net[1].activation_fake_quantize
    return not q.observing() and q.n_bits == 8


def fused_eligible(masker, x):
    """True when MaskGenerator can hand its TCN stack to the fused engine (quantised, steady state)."""
    from .qat import qat_layers as QL
    from .qat.qat_quant import GradientActivationFakeQuantize as AQ, GradientWeightFakeQuantize as WQm
    if not x.is_cuda or x.dtype != torch.float32:
        return False
    blocks = list(masker.TCN)
    b0 = blocks[0]
    sb = b0.shared_block
    if not (isinstance(sb[0], QL.Conv1dNlQ) and isinstance(sb[2], QL.GroupNormQ) and isinstance(sb[3], QL.Conv1dNlQ)
            and isinstance(sb[5], QL.GroupNormQ) and isinstance(b0.res_conv, QL.Conv1dQ) and isinstance(b0.add, QL.AddQ)):
        return False
    Chid, Cio = sb[0].conv1d.weight.shape[0], sb[0].conv1d.weight.shape[1]
    if Cio % 128 or Chid % 128 or sb[3].conv1d.kernel_size[0] != 3:
        return False
    if x.dim() != 3 or not rows_fit(x.shape[-1], max(b.shared_block[3].conv1d.dilation[0] for b in blocks)):
        return False
    for m in masker.modules():
        if isinstance(m, AQ) and (m.observing() or m.n_bits != 8):
            return False
        if isinstance(m, WQm) and (m.observer_mode or m.n_bits != 8):
            return False
        if isinstance(m, QL.LayerQ) and type(m).__name__ in ("Conv1dNlQ", "Conv1dQ") and isinstance(m.weight_fake_quantize, torch.nn.Identity):
            return False
        if isinstance(m, QL.LayerQ) and isinstance(m.activation_fake_quantize, torch.nn.Identity):
            return False
        if isinstance(m, torch.nn.PReLU) and m.weight.numel() != 1:
            return False
    return True
