"""Filterbank edges of the quantised separator on the tcgen05 GEMM (north_star item (c)).

Reference: ConvTr1dDecoderQ (qat_layers.py:1305-1361), ResidualErrorBlock (:1105-1220), Conv1dEncoderQ (:993-1046).

Decoder.  A ConvTranspose1d F -> 1 with L taps and stride H is an overlap-add of per-frame tap vectors,

    frames[r, k, m] = sum_o w[o, k] * Y[r, o, m]          y[r, 8m + k] += frames[r, k, m]

i.e. a GEMM over the F filters with the L taps as output channels.  Y is the output of an 8-bit activation quantiser and
w is fake-quantised per tensor, so both operands are INTEGER CODES (exact fp32 accumulation on the tensor core; the result
differs from the reference's fp32 conv only by the final affine  s1 * acc + s0, two roundings), exactly as in the 1x1
convolutions of the ConvBlocks.  The tensor core computes 128 columns (its minimum N here) and the epilogue keeps the L real
ones (fqss_pw_gemm_nstore); fqss_ola_fwd adds the frames up.

Backward.  The framed output gradient G[r, k, m] = g[r, 8m + k] becomes a three-term bf16 operand [hi ; mid ; lo] (together
the fp32 value: the operand is padded to 128 rows anyway, so the third term is free): dgrad = fqss_pw_gemm(G, [code | code |
code | 0]) * dw, wgrad = fqss_wgrad_codes(G, Y codes) folded over (hi, mid, lo) by fqss_dec_wgrad_fold.  The 524 MB feature tensor is read as 2-byte codes by all three GEMMs (the SIMT kernels
of conv_edge.cu read it as fp32 and are FMA-bound at 0.4 of HBM).

RQB.  FQ(Y - Yq) is produced as codes only (fqss_sub_fq_codes: the fp32 residual tensor is never written); its backward is
the library's SUB + FQ backward on the saved (Y, Yq).

Encoder-type convs (Conv1d C -> N, L taps, stride H) whose input lies on an 8-bit grid (the splitter's output, the
quantised decoder output inside the RQB): the input is framed into codes [R][KP][ld] (fqss_frames_encode, KP = C*L padded
to 64) and the conv is one code-operand GEMM; their backward stays on conv_edge.cu (the big operand there is an fp32
gradient).
"""
import ctypes as C

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _native as N
from . import ops
from ._native import PwGrads, check, lib, ptr, stream_ptr, workspace

BF = torch.bfloat16
ENABLED = True          # module switch: False keeps the SIMT filterbank kernels (A/B runs, tests)


def _ld8(M):
    return (M + 7) // 8 * 8


def _pitched(x, ld):
    """fp32 [R,C,M] with row pitch exactly ld (view when possible)."""
    R, Cc, M = x.shape
    if x.stride(2) == 1 and x.stride(1) == ld and x.stride(0) == Cc * ld and x.data_ptr() % 16 == 0:
        return x
    buf = torch.empty((R, Cc, ld), device=x.device, dtype=x.dtype)
    buf[:, :, :M].copy_(x)
    return buf[:, :, :M]


class _DecConsts:
    """GEMM operands of the fake-quantised decoder weight for codes of the quantiser {amin, amax}."""

    def __init__(self, W, wmin, wmax, amin, amax):
        F, one, L = W.shape
        dev = W.device
        self.F, self.L = F, L
        self.Wc = torch.empty((128, F), dtype=BF, device=dev)
        self.WT = torch.empty((F, 128), dtype=BF, device=dev)
        self.s1, self.s0 = torch.empty(128, device=dev), torch.empty(128, device=dev)
        self.dgs = torch.empty(F, device=dev)
        check(lib().fqss_edge_dec_prep(ptr(W.detach().contiguous()), ptr(wmin), ptr(wmax), ptr(amin), ptr(amax), ptr(self.Wc),
                                       ptr(self.WT), ptr(self.s1), ptr(self.s0), ptr(self.dgs), F, L, stream_ptr()))


def _decode_fwd(codes, k, M, ld, H):
    """codes bf16 [R,F,ld] -> y fp32 [R,1,(M-1)H+L]."""
    R = codes.shape[0]
    dev = codes.device
    frames = torch.empty((R, k.L, ld), device=dev)
    check(lib().fqss_pw_gemm_nstore(ptr(codes), ptr(k.Wc), ptr(k.s1), ptr(k.s0), ptr(frames), k.L, R, k.F, 128, M, ld, stream_ptr()))
    T = (M - 1) * H + k.L
    y = ops.alloc_rows((R, 1, T), dev)
    check(lib().fqss_ola_fwd(ptr(frames), ld, ptr(y), ops.ld_of(y), R, 1, k.L, H, M, stream_ptr()))
    return y


_SCRATCH = {}


def _frames_operand(R, ld, dev):
    """bf16 [R,128,ld] operand of the framed gradient, zero-initialised: fqss_frames_split writes the 3L real rows, the padding
    rows up to the tile height must be zero.  A fresh tensor per call (a 65 MB memset at the recipe's size, ~15 us): a cached
    buffer would be baked into captured CUDA graphs and could not be released safely."""
    return torch.zeros((R, 128, ld), dtype=BF, device=dev)


def _ones128(dev):
    key = ("1", dev.index)
    buf = _SCRATCH.get(key)
    if buf is None:
        buf = torch.ones(128, device=dev)
        _SCRATCH[key] = buf
    return buf


def _decode_bwd(g, codes, k, amin, amax, M, ld, H, want_gx, addend=None):
    """g [R,1,T] -> (gx fp32 [R,F,ld] or None (+ addend: a gradient reaching the same tensor on another path), dWfq fp32
    [F,1,L]: gradient w.r.t. the fake-quantised weight)."""
    L_ = lib()
    R = codes.shape[0]
    dev = codes.device
    s = stream_ptr()
    g, _, _, ldg = ops.rows_view(g)
    G = _frames_operand(R, ld, dev)
    rowsum = torch.zeros(128, dtype=torch.float64, device=dev)
    check(L_.fqss_frames_split(ptr(g), ldg, ptr(G), ld, R, M, k.L, H, 0, ptr(rowsum), s))
    gx = None
    if want_gx:
        gx = torch.empty((R, k.F, ld), device=dev)
        add = _pitched(addend, ld) if addend is not None else None
        check(L_.fqss_pw_gemm(ptr(G), ptr(k.WT), ptr(k.dgs), None, ptr(gx), None, ptr(add) or None, R, 128, k.F, M, ld, 0, s))
    elif addend is not None:
        gx = addend
    part = torch.empty((128, k.F), device=dev)
    ones = _ones128(dev)
    ws = torch.empty(int(L_.fqss_wgrad_codes_ws_bytes(R, M, 128, k.F)), dtype=torch.uint8, device=dev)
    check(L_.fqss_wgrad_codes(ptr(G), ptr(codes), R, M, ld, 128, k.F, ptr(amin), ptr(amax), ptr(ones), ptr(rowsum), ptr(part), ptr(ws),
                              ws.numel(), s))
    dW = torch.empty((k.F, 1, k.L), device=dev)
    check(L_.fqss_dec_wgrad_fold(ptr(part), ptr(dW), k.F, k.L, s))
    return gx, dW


class DecodeCodes(Function):
    """ConvTranspose1d(F -> 1) of a tensor on the grid of the 8-bit quantiser {qmin, qmax}.  `x` carries the autograd edge
    (its values are not read when `codes`, the same tensor as bf16 integer codes [R,F,ld], is given).  Returns (y, alias of
    x): the alias is for a second consumer of x (the RQB) -- its gradient arrives here and is added in the dgrad GEMM's
    epilogue, so the sum of the two gradient paths costs no pass of its own."""

    @staticmethod
    def forward(ctx, x, codes, qmin, qmax, w_fq, W, wmin, wmax, stride):
        N.require_cuda(x, codes, qmin, qmax, w_fq, W, wmin, wmax)
        R, F, M = x.shape
        ld = _ld8(M)
        if codes is None:
            xv = _pitched(x.detach(), ld)
            codes = torch.empty((R, F, ld), dtype=BF, device=x.device)
            check(lib().fqss_tcn_encode(ptr(xv), ld, ptr(codes), ld, R * F, M, ptr(qmin), ptr(qmax), stream_ptr()))
        k = _DecConsts(W, wmin, wmax, qmin, qmax)
        ctx.k, ctx.meta = k, (M, ld, stride)
        ctx.save_for_backward(codes, qmin, qmax)
        return _decode_fwd(codes, k, M, ld, stride), x.view_as(x)

    @staticmethod
    @once_differentiable
    def backward(ctx, g, g_alias):
        codes, qmin, qmax = ctx.saved_tensors
        M, ld, H = ctx.meta
        if g is None:
            return g_alias, None, None, None, None, None, None, None, None
        gx, dW = _decode_bwd(g, codes, ctx.k, qmin, qmax, M, ld, H, ctx.needs_input_grad[0], g_alias)
        return (gx[:, :, :M] if gx is not None else None), None, None, None, dW, None, None, None, None


class SubFQDecode(Function):
    """RQB tail (qat_layers.py:1195-1199): decode(FQ_r(Y - Yq)) with the decoder's fake-quantised weight; the residual exists
    only as integer codes."""

    @staticmethod
    def forward(ctx, Y, Yq, rmin, rmax, w_fq, W, wmin, wmax, stride):
        N.require_cuda(Y, Yq, rmin, rmax, w_fq, W, wmin, wmax)
        R, F, M = Y.shape
        ld = _ld8(M)
        Yv, _, _, ldy = ops.rows_view(Y.detach())
        Qv, _, _, ldq = ops.rows_view(Yq.detach())
        codes = torch.empty((R, F, ld), dtype=BF, device=Y.device)
        check(lib().fqss_sub_fq_codes(ptr(Yv), ldy, ptr(Qv), ldq, ptr(codes), ld, R * F, M, ptr(rmin), ptr(rmax), stream_ptr()))
        k = _DecConsts(W, wmin, wmax, rmin, rmax)
        ctx.k, ctx.meta = k, (M, ld, stride, ldy, ldq)
        ctx.save_for_backward(codes, Yv, Qv, rmin, rmax)
        return _decode_fwd(codes, k, M, ld, stride)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        codes, Yv, Qv, rmin, rmax = ctx.saved_tensors
        M, ld, H, ldy, ldq = ctx.meta
        R, F = codes.shape[0], codes.shape[1]
        dev = codes.device
        g1, dW = _decode_bwd(g, codes, ctx.k, rmin, rmax, M, ld, H, True)
        # SUB + FQ backward of the library on the saved operands: g1 -> (gY, gYq, range gradients)
        rows = R * F
        gY = ops.alloc_rows((R, F, M), dev)
        gQ = ops.alloc_rows((R, F, M), dev) if ctx.needs_input_grad[1] else None
        gmin, gmax = torch.empty_like(rmin), torch.empty_like(rmax)
        d = ops._desc(N.PW_SUB, True, 8, Yv, (rows, M, ldy), Qv, (rows, M, ldq), None, 0, None, None, None, None, 0.0, rmin, rmax, 1, 1)
        o = PwGrads()
        o.g, o.ldg = ptr(g1), ld
        o.gx1, o.ldg1 = ptr(gY), ops.ld_of(gY)
        o.gx2, o.ldg2 = (ptr(gQ), ops.ld_of(gQ)) if gQ is not None else (None, 0)
        o.g_rmin, o.g_rmax = ptr(gmin), ptr(gmax)
        ws = workspace(rows, dev)
        check(lib().fqss_pw_bwd(C.byref(d), C.byref(o), ptr(ws), ws.numel(), stream_ptr()))
        return gY, gQ, gmin, gmax, dW, None, None, None, None


class FramedCodeConv(Function):
    """Conv1d(C -> N, L taps, stride H, no padding / bias) of an input on the grid of an 8-bit quantiser {amin, amax}, as one
    code-operand GEMM over the framed input.  Backward: conv_edge.cu (fqss_sconv_bwd) on the saved fp32 input and the
    fake-quantised weight."""

    @staticmethod
    def forward(ctx, x, amin, amax, w_fq, W, wmin, wmax, stride):
        N.require_cuda(x, amin, amax, w_fq, W, wmin, wmax)
        L_ = lib()
        xv, rows, T, ldx = ops.rows_view(x.detach())
        Nf, Cin, Lt = W.shape
        R = rows // Cin
        M = (T - Lt) // stride + 1
        ld = _ld8(M)
        KP = (Cin * Lt + 63) // 64 * 64
        dev = x.device
        s = stream_ptr()
        fr = torch.empty((R, KP, ld), dtype=BF, device=dev)
        check(L_.fqss_frames_encode(ptr(xv), ldx, ptr(fr), ld, R, Cin, M, Lt, stride, KP, ptr(amin), ptr(amax), s))
        Wc = torch.empty((Nf, KP), dtype=BF, device=dev)
        s1, s0 = torch.empty(Nf, device=dev), torch.empty(Nf, device=dev)
        check(L_.fqss_edge_enc_prep(ptr(W.detach().contiguous()), ptr(wmin), ptr(wmax), ptr(amin), ptr(amax), ptr(Wc), ptr(s1), ptr(s0),
                                    Nf, Cin * Lt, KP, s))
        y = torch.empty((R, Nf, ld), device=dev)
        check(L_.fqss_pw_gemm(ptr(fr), ptr(Wc), ptr(s1), ptr(s0), ptr(y), None, None, R, KP, Nf, M, ld, 0, s))
        ctx.save_for_backward(xv, w_fq.detach().contiguous())
        ctx.meta = (R, Cin, Nf, T, Lt, stride, ldx, tuple(x.shape))
        return y[:, :, :M]

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        R, Cin, Nf, T, Lt, stride, ldx, xshape = ctx.meta
        g, _, _, ldg = ops.rows_view(g)
        gx = ops.alloc_rows(xshape, x.device) if ctx.needs_input_grad[0] else None
        gw = torch.empty_like(w)
        ws = workspace(Nf * Cin * Lt // 4 + 1, x.device)
        check(lib().fqss_sconv_bwd(ptr(g), ldg, ptr(x), ldx, ptr(w), ptr(gx) or None, ops.ld_of(gx) if gx is not None else 0,
                                   ptr(gw), R, Cin, Nf, T, Lt, stride, ptr(ws), ws.numel(), stream_ptr()))
        return gx, None, None, gw, None, None, None, None


# ---------------------------------------------------------------------------------------------
# eligibility + composition (called by qat_layers.ConvTr1dDecoderQ / Conv1dEncoderQ / ResidualErrorBlock)
# ---------------------------------------------------------------------------------------------
def _steady_aq(q):
    from .qat.qat_quant import GradientActivationFakeQuantize as AQ
    return isinstance(q, AQ) and not q.observing() and q.n_bits == 8


def _steady_wq(q, numel):
    from .qat.qat_quant import GradientWeightFakeQuantize as WQm
    return isinstance(q, WQm) and not q.observer_mode and q.n_bits == 8 and q.min_range.numel() == numel


def decoder_wants_codes(layer):
    """Static part of decoder_eligible: True when `layer` is a ConvTr1dDecoderQ in the steady state the GEMM path covers (the
    mask head then also emits its output as integer codes)."""
    from .qat.qat_layers import ConvTr1dDecoderQ
    if not ENABLED or not isinstance(layer, ConvTr1dDecoderQ):
        return False
    dec = layer.convTr1d
    F, L = dec.in_channels, dec.kernel_size[0]
    if F % 128 or F > 1024 or L % 16 or 3 * L > 128 or dec.out_channels != 1:
        return False
    if not _steady_wq(layer.weight_fake_quantize, 1) or layer.n_combiner > 2:
        return False
    if layer.n_combiner == 1:
        return True
    rqb = layer.residual_error_block
    if rqb.train_res_dec and not _steady_wq(rqb.weight_fake_quantize_dec, 1):
        return False
    return _steady_aq(rqb.activation_fake_quantize)


def decoder_eligible(layer, x):
    """ConvTr1dDecoderQ.forward(x) can run on the GEMM path: x is a CUDA fp32 [R,F,M] tensor tagged with the steady-state
    8-bit quantiser that produced it, per-tensor 8-bit weight quantiser, tap / filter counts the tensor-core tiles cover."""
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 3 or x.shape[1] != layer.convTr1d.in_channels:
        return False
    return decoder_wants_codes(layer) and _steady_aq(getattr(x, "_fq_src", None))


def decoder_forward(layer, x):
    """ConvTr1dDecoderQ.forward on the GEMM path -> [n_combiner, R, 1, T] (or [R,1,T])."""
    dec = layer.convTr1d
    H = dec.stride[0]
    q_in = x._fq_src
    codes = getattr(x, "_fq_codes", None)
    wq = layer.weight_fake_quantize
    w_dec = wq(dec.weight)
    y_pre, x_r = DecodeCodes.apply(x, codes, q_in.min_range, q_in.max_range, w_dec, dec.weight, wq.min_range, wq.max_range, H)
    y = layer._finish(N.PW_IDENT, y_pre)
    if layer.do_mac_op:
        Ci, Co, kk = dec.weight.shape
        layer.mac_op = x.shape[0] * Co * Ci * ((x.shape[-1] - 1) * H + kk) * (kk // H)
    if layer.n_combiner == 1:
        return y
    rqb = layer.residual_error_block
    Yq = rqb.reencode(y)
    rq = rqb.activation_fake_quantize
    Wr, wqr = rqb.decode_weight(dec, wq)            # the decoder's own filterbank, or the block's (train_res_dec)
    w_res = w_dec if Wr is dec.weight else wqr(Wr)
    y1_pre = SubFQDecode.apply(x_r, Yq, rq.min_range, rq.max_range, w_res, Wr, wqr.min_range, wqr.max_range, H)
    y1 = layer._finish(N.PW_IDENT, y1_pre, quantizer=layer.activation_fake_quantize_residual)
    return torch.stack([y, y1])


def framed_conv_eligible(conv, wq, x, grid):
    """Encoder-type conv on the GEMM path: `grid` = (amin, amax) device tensors of the 8-bit grid x lies on."""
    if not ENABLED or grid is None or not x.is_cuda or x.dtype != torch.float32 or x.dim() != 3:
        return False
    if conv.padding[0] or conv.dilation[0] != 1 or conv.groups != 1 or conv.bias is not None:
        return False
    Nf, Cin, Lt = conv.weight.shape
    if Nf % 128 or Nf > 1024 or Cin * Lt > 1024:
        return False
    return _steady_wq(wq, Nf)


def framed_conv(conv, wq, x, grid):
    w_fq = wq(conv.weight)
    return FramedCodeConv.apply(x, grid[0], grid[1], w_fq, conv.weight, wq.min_range, wq.max_range, conv.stride[0])
