"""Autograd wrappers over the C ABI (include/fqss.h).  Host-side plumbing only: shapes, pitches,
saved tensors.  Every op requires CUDA tensors and raises otherwise (no CPU / eager fallback).

Row-tensor layout: activations [..., C, M] are allocated with a row pitch padded to a multiple of
4 floats so that every row starts 16-byte aligned (M = 3999 in the recipe).  They are ordinary
torch tensors (non-contiguous views), so foreign code can consume them; foreign inputs that do not
have the layout are repacked once at the boundary.
"""
import ctypes as C

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _native as N
from ._native import PwDesc, PwGrads, check, lib, ptr, stream_ptr, workspace

PAD = 8      # row pitch in floats: 32-byte rows keep fp32 (16 B) and bf16 (TMA: 16 B strides) views aligned


def _pitch(m):
    return (m + PAD - 1) // PAD * PAD


def alloc_rows(shape, device, dtype=torch.float32):
    """Uninitialised tensor of `shape` whose last-dim rows sit at a 16-byte aligned pitch."""
    shape = tuple(int(s) for s in shape)
    m = shape[-1]
    rows = 1
    for s in shape[:-1]:
        rows *= s
    base = torch.empty((rows, _pitch(m)), device=device, dtype=dtype)
    return base[:, :m].view(shape)


def _row_layout(t):
    """(ok, ld): whether `t` is a uniformly pitched, 16-byte aligned row tensor, and its pitch."""
    cols = t.shape[-1]
    if t.dim() == 0 or (t.stride(-1) != 1 and cols != 1):
        return False, 0
    dims = [(t.shape[i], t.stride(i)) for i in range(t.dim() - 1) if t.shape[i] > 1]
    if dims:
        ld = dims[-1][1]
        exp = ld
        for size, stride in reversed(dims):
            if stride != exp:
                return False, 0
            exp *= size
    else:
        ld = _pitch(cols)
    return (ld >= cols and ld % PAD == 0 and t.data_ptr() % 16 == 0), ld


def ld_of(t):
    ok, ld = _row_layout(t)
    if not ok:
        raise N.FqssError("internal: tensor is not a pitched row tensor")
    return ld


def rows_view(t):
    """-> (tensor, rows, cols, ld).  Repacks `t` into the padded layout when its strides do not
    describe uniformly pitched rows (or the pitch / base is not 16-byte aligned)."""
    if t.dtype != torch.float32:
        raise N.FqssError("fqss_b200 ops are fp32; got %s" % t.dtype)
    ok, ld = _row_layout(t)
    if not ok:
        new = alloc_rows(t.shape, t.device)
        new.copy_(t)
        t = new
        ok, ld = _row_layout(t)
    cols = t.shape[-1]
    rows = t.numel() // cols if cols else 0
    return t, rows, cols, ld


def _aligned_contig(t):
    """Contiguous AND 16-byte aligned (the flat quantiser kernels use 128-bit accesses): a contiguous view into the middle of
    a buffer is repacked instead of being refused by the library."""
    t = t.contiguous()
    return t if t.data_ptr() % 16 == 0 else t.clone(memory_format=torch.contiguous_format)


def _scalar_like(p):
    return torch.empty_like(p, memory_format=torch.contiguous_format)


# =============================================================================================
# Q2: standalone activation fake-quant (GradientActivationFakeQuantize.forward, quantise branch)
# =============================================================================================
class FakeQuantAct(Function):
    @staticmethod
    def forward(ctx, x, rmin, rmax, n_bits):
        N.require_cuda(x, rmin, rmax)
        xc = _aligned_contig(x)
        y = torch.empty_like(xc)
        check(lib().fqss_fq_act_fwd(ptr(xc), ptr(y), None, xc.numel(), ptr(rmin), ptr(rmax), n_bits, stream_ptr()))
        ctx.save_for_backward(xc, rmin, rmax)
        ctx.n_bits = n_bits
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, rmin, rmax = ctx.saved_tensors
        g = _aligned_contig(g)
        gx = torch.empty_like(x)
        gmin, gmax = _scalar_like(rmin), _scalar_like(rmax)
        ws = workspace(0, x.device)
        check(lib().fqss_fq_act_bwd(ptr(g), ptr(x), ptr(gx), ptr(gmin), ptr(gmax), x.numel(), ptr(rmin), ptr(rmax),
                                    ctx.n_bits, ptr(ws), ws.numel(), stream_ptr()))
        return gx, gmin, gmax, None


def fake_quant_codes(x, rmin, rmax, n_bits=8):
    """(y, uint8 codes) of the activation quantiser -- test / export helper."""
    N.require_cuda(x, rmin, rmax)
    xc = _aligned_contig(x)
    y = torch.empty_like(xc)
    code = torch.empty(xc.shape, dtype=torch.uint8, device=xc.device)
    check(lib().fqss_fq_act_fwd(ptr(xc), ptr(y), ptr(code), xc.numel(), ptr(rmin), ptr(rmax), n_bits, stream_ptr()))
    return y, code


# =============================================================================================
# Q3: weight fake-quant
# =============================================================================================
def _w_geometry(w, axis):
    outer = 1
    for s in w.shape[:axis]:
        outer *= s
    inner = 1
    for s in w.shape[axis + 1:]:
        inner *= s
    return int(outer), int(w.shape[axis]), int(inner)


class FakeQuantWeight(Function):
    @staticmethod
    def forward(ctx, w, rmin, rmax, axis, n_bits):
        N.require_cuda(w, rmin, rmax)
        wc = w.contiguous()
        outer, ch, inner = _w_geometry(wc, axis)
        if rmin.numel() != ch:
            raise N.FqssError("weight quantiser: %d ranges for %d channels" % (rmin.numel(), ch))
        wq = torch.empty_like(wc)
        check(lib().fqss_fq_weight_fwd(ptr(wc), ptr(wq), None, outer, ch, inner, ptr(rmin), ptr(rmax), n_bits, stream_ptr()))
        ctx.save_for_backward(wc, rmin, rmax)
        ctx.geom = (outer, ch, inner, n_bits)
        return wq

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        w, rmin, rmax = ctx.saved_tensors
        outer, ch, inner, n_bits = ctx.geom
        g = g.contiguous()
        gw = torch.empty_like(w)
        gmin, gmax = _scalar_like(rmin), _scalar_like(rmax)
        check(lib().fqss_fq_weight_bwd(ptr(g), ptr(w), ptr(gw), ptr(gmin), ptr(gmax), outer, ch, inner, ptr(rmin), ptr(rmax),
                                       n_bits, stream_ptr()))
        return gw, gmin, gmax, None, None


def weight_codes(w, rmin, rmax, axis=0, n_bits=8):
    wc = w.contiguous()
    outer, ch, inner = _w_geometry(wc, axis)
    code = torch.empty(wc.shape, dtype=torch.int8, device=wc.device)
    check(lib().fqss_fq_weight_fwd(ptr(wc), None, ptr(code), outer, ch, inner, ptr(rmin), ptr(rmax), n_bits, stream_ptr()))
    return code


# =============================================================================================
# X1: export-time quantisers (torch.fake_quantize_per_{tensor,channel}_affine arithmetic; qat_quant.py:15-72)
# =============================================================================================
class FakeQuantAffineTensor(Function):
    @staticmethod
    def forward(ctx, x, scale, zero_point, qmin, qmax):
        N.require_cuda(x)
        xc = x.contiguous()
        y = torch.empty_like(xc)
        mask = torch.empty(xc.shape, dtype=torch.uint8, device=xc.device) if x.requires_grad else None
        rc = lib().fqss_fq_affine_tensor(ptr(xc), ptr(y), ptr(mask) or None, None, xc.numel(), float(scale), int(zero_point),
                                         int(qmin), int(qmax), stream_ptr())
        if rc == -1 and b"zero_point" in lib().fqss_last_error():
            raise RuntimeError(lib().fqss_last_error().decode())        # ATen's own error type and text
        check(rc)
        ctx.mask = mask
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        g = g.contiguous()
        gx = torch.empty_like(g)
        check(lib().fqss_fq_affine_bwd(ptr(g), ptr(ctx.mask), ptr(gx), g.numel(), stream_ptr()))
        return gx, None, None, None, None


class FakeQuantAffineChannel(Function):
    @staticmethod
    def forward(ctx, x, scales, axis, qmin, qmax):
        N.require_cuda(x, scales)
        xc = x.contiguous()
        outer, ch, inner = _w_geometry(xc, axis)
        if scales.numel() != ch:
            raise RuntimeError("Expected `scale` to have the same length as the quantised axis (%d vs %d)" % (scales.numel(), ch))
        y = torch.empty_like(xc)
        mask = torch.empty(xc.shape, dtype=torch.uint8, device=xc.device) if x.requires_grad else None
        check(lib().fqss_fq_affine_channel(ptr(xc), ptr(y), ptr(mask) or None, None, outer, ch, inner, ptr(scales), int(qmin),
                                           int(qmax), stream_ptr()))
        ctx.mask = mask
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        g = g.contiguous()
        gx = torch.empty_like(g)
        check(lib().fqss_fq_affine_bwd(ptr(g), ptr(ctx.mask), ptr(gx), g.numel(), stream_ptr()))
        return gx, None, None, None, None


def affine_codes_tensor(x, scale, zero_point, qmin, qmax):
    """(y, int32 clamped integer codes) of the per-tensor export quantiser."""
    xc = x.contiguous()
    y = torch.empty_like(xc)
    code = torch.empty(xc.shape, dtype=torch.int32, device=xc.device)
    check(lib().fqss_fq_affine_tensor(ptr(xc), ptr(y), None, ptr(code), xc.numel(), float(scale), int(zero_point), int(qmin),
                                      int(qmax), stream_ptr()))
    return y, code


def affine_codes_channel(x, scales, axis, qmin, qmax):
    xc = x.contiguous()
    outer, ch, inner = _w_geometry(xc, axis)
    y = torch.empty_like(xc)
    code = torch.empty(xc.shape, dtype=torch.int32, device=xc.device)
    check(lib().fqss_fq_affine_channel(ptr(xc), ptr(y), None, ptr(code), outer, ch, inner, ptr(scales), int(qmin), int(qmax),
                                       stream_ptr()))
    return y, code


def weight_observe_(w, rmin, rmax, axis):
    N.require_cuda(w, rmin, rmax)
    wc = w.detach().contiguous()
    outer, ch, inner = _w_geometry(wc, axis)
    check(lib().fqss_weight_observe(ptr(wc), outer, ch, inner, ptr(rmin), ptr(rmax), stream_ptr()))


def act_observe_(x, rmin, rmax, alpha):
    N.require_cuda(x, rmin, rmax)
    x, rows, cols, ld = rows_view(x.detach())
    ws = workspace(rows, x.device)
    check(lib().fqss_act_observe(ptr(x), rows, cols, ld, ptr(rmin), ptr(rmax), float(alpha), ptr(ws), ws.numel(), stream_ptr()))


# =============================================================================================
# L1: op -> nonlinearity -> fake-quant in one pass
# =============================================================================================
def _desc(kind, quant, n_bits, x1, r1, x2, r2, y, ry, slope, gamma, beta, stats, eps, rmin, rmax, C_, bcast):
    d = PwDesc()
    d.kind, d.quant, d.n_bits, d.C, d.bcast = kind, int(bool(quant)), int(n_bits), int(C_), int(bcast)
    d.rows, d.cols = r1[0], r1[1]
    d.x1, d.ld1 = ptr(x1), r1[2]
    d.x2, d.ld2 = (ptr(x2), r2[2]) if x2 is not None else (None, 0)
    d.y, d.ldy = (ptr(y), ry) if y is not None else (None, 0)
    d.slope, d.gamma, d.beta, d.stats = ptr(slope) or None, ptr(gamma) or None, ptr(beta) or None, ptr(stats) or None
    d.eps = float(eps)
    d.rmin, d.rmax = ptr(rmin) or None, ptr(rmax) or None
    return d


class PointwiseFQ(Function):
    """y = FQ(op(x1[, x2])) for op in {ident, prelu, relu, add, sub, mul(bcast), gLN}."""

    @staticmethod
    def forward(ctx, kind, x1, x2, slope, gamma, beta, rmin, rmax, quant, n_bits, eps):
        N.require_cuda(x1, x2, slope, gamma, beta, rmin, rmax)
        x1, rows, cols, ld1 = rows_view(x1)
        C_, bcast, r2, stats = 1, 1, None, None
        if x2 is not None:
            if kind == N.PW_MUL and x2.shape != x1.shape:
                # MulQ: mask [B,S,N,M] * feats [B,1,N,M]
                if not (x1.dim() == 4 and x2.dim() == 4 and x2.shape[1] == 1 and x2.shape[0] == x1.shape[0]
                        and x2.shape[2:] == x1.shape[2:]):
                    raise N.FqssError("MulQ: unsupported broadcast %s * %s" % (tuple(x1.shape), tuple(x2.shape)))
                bcast, C_ = x1.shape[1], x1.shape[2]
            elif x2.shape != x1.shape:
                raise N.FqssError("binary op shapes differ: %s vs %s" % (tuple(x1.shape), tuple(x2.shape)))
            elif kind == N.PW_MUL:
                C_ = 1
            x2, rows2, _, ld2 = rows_view(x2)
            r2 = (rows2, cols, ld2)
        if kind == N.PW_GLN:
            C_ = x1.shape[-2]
            B = rows // C_
            stats = torch.empty(2 * B, dtype=torch.float64, device=x1.device)
            check(lib().fqss_gln_stats(ptr(x1), rows, cols, ld1, C_, ptr(stats), stream_ptr()))
        y = alloc_rows(x1.shape, x1.device)
        d = _desc(kind, quant, n_bits, x1, (rows, cols, ld1), x2, r2, y, ld_of(y),
                  slope, gamma, beta, stats, eps, rmin, rmax, C_, bcast)
        check(lib().fqss_pw_fwd(C.byref(d), stream_ptr()))
        ctx.save_for_backward(x1, x2, slope, gamma, beta, rmin, rmax, stats)
        ctx.meta = (kind, quant, n_bits, eps, C_, bcast, rows, cols, ld1, r2)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x1, x2, slope, gamma, beta, rmin, rmax, stats = ctx.saved_tensors
        kind, quant, n_bits, eps, C_, bcast, rows, cols, ld1, r2 = ctx.meta
        g, _, _, ldg = rows_view(g)
        dev = x1.device
        gx1 = alloc_rows(x1.shape, dev)
        gx2 = None
        if x2 is not None and kind in (N.PW_SUB, N.PW_MUL) and ctx.needs_input_grad[2]:
            gx2 = alloc_rows(x2.shape, dev)
        gmin = _scalar_like(rmin) if quant and rmin is not None else None
        gmax = _scalar_like(rmax) if quant and rmax is not None else None
        gslope = _scalar_like(slope) if kind == N.PW_PRELU else None
        ggamma = torch.empty_like(gamma) if kind == N.PW_GLN else None
        gbeta = torch.empty_like(beta) if kind == N.PW_GLN else None
        d = _desc(kind, quant, n_bits, x1, (rows, cols, ld1), x2, r2, None, 0, slope, gamma, beta, stats, eps, rmin, rmax,
                  C_, bcast)
        o = PwGrads()
        o.g, o.ldg = ptr(g), ldg
        o.gx1, o.ldg1 = (ptr(gx1), ld_of(gx1)) if gx1 is not None else (None, 0)
        o.gx2, o.ldg2 = (ptr(gx2), ld_of(gx2)) if gx2 is not None else (None, 0)
        o.g_rmin, o.g_rmax, o.g_slope = ptr(gmin) or None, ptr(gmax) or None, ptr(gslope) or None
        o.g_gamma, o.g_beta = ptr(ggamma) or None, ptr(gbeta) or None
        ws = workspace(rows, dev)
        check(lib().fqss_pw_bwd(C.byref(d), C.byref(o), ptr(ws), ws.numel(), stream_ptr()))
        if kind == N.PW_ADD:
            gx2 = gx1
        return None, gx1, gx2, gslope, ggamma, gbeta, gmin, gmax, None, None, None


class Fanout2(Function):
    """x -> (x, x) for a tensor with two consumers.  Autograd would sum the two incoming gradients with an ATen add
    into a DENSE tensor, which the next stage then has to repack into pitched rows (a second full pass); this node
    does the sum with the library's own add kernel straight into a pitched tensor."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x), x.view_as(x)

    @staticmethod
    @once_differentiable
    def backward(ctx, ga, gb):
        if ga is None or gb is None:
            return ga if gb is None else gb
        N.require_cuda(ga, gb)
        ga, rows, cols, lda = rows_view(ga)
        gb, rows_b, _, ldb = rows_view(gb)
        y = alloc_rows(ga.shape, ga.device)
        d = _desc(N.PW_ADD, False, 8, ga, (rows, cols, lda), gb, (rows_b, cols, ldb), y, ld_of(y),
                  None, None, None, None, 0.0, None, None, 1, 1)
        check(lib().fqss_pw_fwd(C.byref(d), stream_ptr()))
        return y


def fanout2(x):
    """Two aliases of `x` whose gradients are summed by the library (pitched) instead of by autograd (dense)."""
    if not (x.is_cuda and x.requires_grad and torch.is_grad_enabled()):
        return x, x
    a, b = Fanout2.apply(x)
    src = getattr(x, "_fq_src", None)       # the 8-bit quantiser that produced x (qat_layers.LayerQ._finish): both aliases carry it
    if src is not None:
        a._fq_src = b._fq_src = src
    return a, b


def pointwise_fq(kind, x1, x2=None, slope=None, gamma=None, beta=None, rmin=None, rmax=None, quant=False, n_bits=8, eps=0.0):
    return PointwiseFQ.apply(kind, x1, x2, slope, gamma, beta, rmin, rmax, bool(quant), int(n_bits), float(eps))


# =============================================================================================
# convolutions
# =============================================================================================
class Conv1x1(Function):
    @staticmethod
    def forward(ctx, x, w, bias):
        N.require_cuda(x, w, bias)
        x, rows, M, ldx = rows_view(x)
        Co, Ci = w.shape[0], w.shape[1]
        B = rows // Ci
        wc = w.contiguous()
        y = alloc_rows(x.shape[:-2] + (Co, M), x.device)
        check(lib().fqss_conv1x1_fwd(ptr(x), ldx, ptr(wc), ptr(bias), ptr(y), ld_of(y), B, Ci, Co, M, stream_ptr()))
        ctx.save_for_backward(x, wc)
        ctx.meta = (B, Ci, Co, M, ldx, bias is not None)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        B, Ci, Co, M, ldx, has_bias = ctx.meta
        g, _, _, ldg = rows_view(g)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = alloc_rows(x.shape, x.device)
            check(lib().fqss_conv1x1_dgrad(ptr(g), ldg, ptr(w), ptr(gx), ld_of(gx), B, Ci, Co, M, stream_ptr()))
        if ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2]):
            gw = torch.empty_like(w)
            gb = torch.empty(Co, device=x.device) if has_bias else None
            ws = workspace(Co, x.device)
            check(lib().fqss_conv1x1_wgrad(ptr(g), ldg, ptr(x), ldx, ptr(gw), ptr(gb) or None, B, Ci, Co, M, ptr(ws), ws.numel(),
                                           stream_ptr()))
        return gx, gw, gb


class DepthwiseConv(Function):
    @staticmethod
    def forward(ctx, x, w, bias, dil):
        N.require_cuda(x, w, bias)
        x, rows, M, ldx = rows_view(x)
        Cc, K = w.shape[0], w.shape[-1]
        B = rows // Cc
        wc = w.contiguous()
        y = alloc_rows(x.shape, x.device)
        check(lib().fqss_dwconv_fwd(ptr(x), ldx, ptr(wc), ptr(bias), ptr(y), ld_of(y), B, Cc, M, K, dil, stream_ptr()))
        ctx.save_for_backward(x, wc)
        ctx.meta = (B, Cc, M, K, dil, ldx, bias is not None)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        B, Cc, M, K, dil, ldx, has_bias = ctx.meta
        g, _, _, ldg = rows_view(g)
        gx = alloc_rows(x.shape, x.device) if ctx.needs_input_grad[0] else None
        gw = torch.empty_like(w)
        gb = torch.empty(Cc, device=x.device) if has_bias else None
        ws = workspace(Cc * 2, x.device)
        check(lib().fqss_dwconv_bwd(ptr(g), ldg, ptr(x), ldx, ptr(w), ptr(gx) or None, ld_of(gx) if gx is not None else 0,
                                    ptr(gw), ptr(gb) or None, B, Cc, M, K, dil, ptr(ws), ws.numel(), stream_ptr()))
        return gx, gw, gb, None


class StridedConv(Function):
    """Encoder-type conv: no padding, no bias, dilation 1 (qat_layers.py:1030, :1189)."""

    @staticmethod
    def forward(ctx, x, w, stride):
        N.require_cuda(x, w)
        x, rows, T, ldx = rows_view(x)
        Co, Cin, K = w.shape
        B = rows // Cin
        Mo = (T - K) // stride + 1
        wc = w.contiguous()
        y = alloc_rows(x.shape[:-2] + (Co, Mo), x.device)
        check(lib().fqss_sconv_fwd(ptr(x), ldx, ptr(wc), ptr(y), ld_of(y), B, Cin, Co, T, K, stride, stream_ptr()))
        ctx.save_for_backward(x, wc)
        ctx.meta = (B, Cin, Co, T, K, stride, ldx)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        B, Cin, Co, T, K, stride, ldx = ctx.meta
        g, _, _, ldg = rows_view(g)
        gx = alloc_rows(x.shape, x.device) if ctx.needs_input_grad[0] else None
        gw = torch.empty_like(w) if ctx.needs_input_grad[1] else None
        ws = workspace(Co * Cin * K // 4 + 1, x.device)
        check(lib().fqss_sconv_bwd(ptr(g), ldg, ptr(x), ldx, ptr(w), ptr(gx) or None, ld_of(gx) if gx is not None else 0,
                                   ptr(gw) or None, B, Cin, Co, T, K, stride, ptr(ws), ws.numel(), stream_ptr()))
        return gx, gw, None


class TransposedConv1(Function):
    """Decoder-type ConvTranspose1d to one channel, no padding / bias (qat_layers.py:1332, :1194)."""

    @staticmethod
    def forward(ctx, x, w, stride):
        N.require_cuda(x, w)
        x, rows, M, ldx = rows_view(x)
        Ci, one, K = w.shape
        if one != 1:
            raise N.FqssError("transposed conv: only 1 output channel is implemented (decoder)")
        B = rows // Ci
        T = (M - 1) * stride + K
        wc = w.contiguous()
        y = alloc_rows(x.shape[:-2] + (1, T), x.device)
        check(lib().fqss_tconv_fwd(ptr(x), ldx, ptr(wc), ptr(y), ld_of(y), B, Ci, M, K, stride, stream_ptr()))
        ctx.save_for_backward(x, wc)
        ctx.meta = (B, Ci, M, K, stride, ldx)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        B, Ci, M, K, stride, ldx = ctx.meta
        g, _, _, ldg = rows_view(g)
        gx = alloc_rows(x.shape, x.device) if ctx.needs_input_grad[0] else None
        gw = torch.empty_like(w) if ctx.needs_input_grad[1] else None
        ws = workspace(Ci * K // 4 + 1, x.device)
        check(lib().fqss_tconv_bwd(ptr(g), ldg, ptr(x), ldx, ptr(w), ptr(gx) or None, ld_of(gx) if gx is not None else 0,
                                   ptr(gw) or None, B, Ci, M, K, stride, ptr(ws), ws.numel(), stream_ptr()))
        return gx, gw, None


# =============================================================================================
# P1: splitter / reconstructor
# =============================================================================================
_SPLIT_GRID = {}


def _splitter_grid(device):
    g = _SPLIT_GRID.get(device)
    if g is None:
        g = (torch.full((1,), -1.0, device=device), torch.full((1,), 127.0 / 128.0, device=device))
        _SPLIT_GRID[device] = g
    return g


def split_input(x, n_split, n_bits=8, normalize=True):
    """process.preprocess for n_splitter >= 2: [B,C,T] -> [B,n_split*C,T] (no gradient: model input).
    normalize=True: x / max|x| with threshold 1 (the speech recipe, process.py:23-24); False: threshold = max|x|
    (ConvTasNetMusicQ.pre_process, convtasnetq_music.py:233-234)."""
    N.require_cuda(x)
    if x.dim() == 2:
        x = x.unsqueeze(1)
    if x.dim() != 3:
        raise N.FqssError("splitter expects [B,T] or [B,C,T], got %s" % (tuple(x.shape),))
    x, rows, T, ldx = rows_view(x.detach())
    B, Cc = x.shape[0], x.shape[1]
    peak = torch.empty(1, device=x.device)
    ws = workspace(rows, x.device)
    check(lib().fqss_absmax(ptr(x), rows, T, ldx, ptr(peak), ptr(ws), ws.numel(), stream_ptr()))
    from . import parallel          # data-parallel runs with global-batch parity: one peak over ALL shards
    parallel.sync_splitter_peak_(peak)
    y = alloc_rows((B, n_split * Cc, T), x.device)
    if Cc == 1 and normalize:
        check(lib().fqss_split(ptr(x), ldx, ptr(peak), ptr(y), ld_of(y), B, T, n_split, n_bits, stream_ptr()))
    else:
        check(lib().fqss_split_ex(ptr(x), ldx, ptr(peak), ptr(y), ld_of(y), B, Cc, T, n_split, n_bits, int(bool(normalize)),
                                  stream_ptr()))
    if normalize and n_bits == 8:
        # every part is k/128 with k in [-128, 127] (process.py:10-14): an 8-bit grid {min = -1, max = 127/128}; the encoder's
        # tensor-core path (edge_engine.FramedCodeConv) re-reads the parts as integer codes
        y._fq_grid = _splitter_grid(x.device)
    return y


class ChannelLayerNorm(Function):
    """nn.LayerNorm(C) over the CHANNEL axis of an NCL tensor, per frame (ChannelWiseLayerNorm,
    convtasnetq_music.py:32-50): y[b,c,m] = (x - mean_bm) * rstd_bm * gamma_c + beta_c."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        N.require_cuda(x, gamma, beta)
        if x.dim() != 3:
            raise N.FqssError("channel layer norm expects [B,C,M]")
        x, rows, M, ld = rows_view(x)
        B, Cc = x.shape[0], x.shape[1]
        y = alloc_rows(x.shape, x.device)
        mean, rstd = torch.empty((B, M), device=x.device), torch.empty((B, M), device=x.device)
        g_, b_ = gamma.detach().contiguous(), beta.detach().contiguous()
        check(lib().fqss_cln_fwd(ptr(x), ld, ptr(g_), ptr(b_), float(eps), ptr(y), ld_of(y), ptr(mean), ptr(rstd), B, Cc, M,
                                 stream_ptr()))
        ctx.save_for_backward(x, g_, mean, rstd)
        ctx.meta = (B, Cc, M, ld)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, gamma, mean, rstd = ctx.saved_tensors
        B, Cc, M, ld = ctx.meta
        g, _, _, ldg = rows_view(g)
        gx = alloc_rows(x.shape, x.device) if ctx.needs_input_grad[0] else None
        want_aff = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        gg = torch.empty(Cc, device=x.device) if want_aff else None
        gb = torch.empty(Cc, device=x.device) if want_aff else None
        ws = workspace(Cc, x.device)
        check(lib().fqss_cln_bwd(ptr(g), ldg, ptr(x), ld, ptr(gamma), ptr(mean), ptr(rstd), ptr(gx) or None,
                                 ld_of(gx) if gx is not None else 0, ptr(gg) or None, ptr(gb) or None, B, Cc, M, ptr(ws), ws.numel(),
                                 stream_ptr()))
        return gx, gg, gb, None


class OverlapAdd(Function):
    """overlap_and_add (convtasnetq_music.py:10-30) of Linear-decoder outputs kept channels-first:
    y [..., A*L, K] (row (a, j) = sample j of audio channel a, column k = frame) -> [..., A, (K-1)*H + L]."""

    @staticmethod
    def forward(ctx, y, A, L, H):
        N.require_cuda(y)
        y, rows, K, ldy = rows_view(y)
        if y.shape[-2] != A * L:
            raise N.FqssError("overlap_add: expected %d rows per item, got %d" % (A * L, y.shape[-2]))
        R = rows // (A * L)
        T = (K - 1) * H + L
        out = alloc_rows(tuple(y.shape[:-2]) + (A, T), y.device)
        check(lib().fqss_ola_fwd(ptr(y), ldy, ptr(out), ld_of(out), R, A, L, H, K, stream_ptr()))
        ctx.meta = (tuple(y.shape), R, A, L, H, K)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        shape, R, A, L, H, K = ctx.meta
        g, _, _, ldo = rows_view(g)
        gy = alloc_rows(shape, g.device)
        check(lib().fqss_ola_bwd(ptr(g), ldo, ptr(gy), ld_of(gy), R, A, L, H, K, stream_ptr()))
        return gy, None, None, None


class Combine(Function):
    """process.postprocess for n_combiner >= 2: parts [n, ..., T] -> sum_i parts[i]*(0.5/128)^i."""

    @staticmethod
    def forward(ctx, parts, n_bits):
        N.require_cuda(parts)
        n = parts.shape[0]
        p, rows_all, T, ld = rows_view(parts)
        rows = rows_all // n
        y = alloc_rows(parts.shape[1:], parts.device)
        check(lib().fqss_combine(ptr(p), rows * ld, ld, ptr(y), ld_of(y), rows, T, n, n_bits,
                                 stream_ptr()))
        ctx.meta = (n, n_bits)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        n, n_bits = ctx.meta
        base = 0.5 / (2 ** (n_bits - 1))
        # d/d parts[i] = g * base^i : tiny [n,B,S,1,T] tensor, assembled with two strided copies
        out = alloc_rows((n,) + tuple(g.shape), g.device)
        for i in range(n):
            out[i].copy_(g) if i == 0 else torch.mul(g, base ** i, out=out[i])
        return out, None


# =============================================================================================
# S1-S3: KD SI-SDR loss
# =============================================================================================
class KDLoss(Function):
    """returns a 3-vector: [loss, kd_loss (logged), val_loss]; gradient flows from element 0 to est."""

    @staticmethod
    def forward(ctx, est, fest, tgt, kd_lambda):
        N.require_cuda(est, fest, tgt)
        if est.dim() != 3 or est.shape[1] != 2 or est.shape != fest.shape or est.shape != tgt.shape:
            raise N.FqssError("kd_loss expects est/fest/tgt of shape [B,2,T]")
        est_r, _, T, lde = rows_view(est)
        fest_r, _, _, ldf = rows_view(fest.detach())
        tgt_r, _, _, ldt = rows_view(tgt.detach())
        B = est.shape[0]
        out = torch.empty(3, device=est.device)
        gest = alloc_rows(est.shape, est.device) if est.requires_grad else None
        ws = workspace(B * 8, est.device)
        from . import parallel
        if parallel.global_batch_parity():
            # loss of the GLOBAL batch (mysystem.py:145 takes the log of batch means): local means -> one all-reduce of
            # 3 doubles -> loss / gradient from the global means (include/fqss.h, fqss_kd_loss_dp)
            import torch.distributed as dist
            means = torch.empty(3, dtype=torch.float64, device=est.device)
            args = (ptr(est_r), lde, ptr(fest_r), ldf, ptr(tgt_r), ldt, B, T, float(kd_lambda))
            check(lib().fqss_kd_loss_dp(*args, None, None, 0, ptr(ws), ws.numel(), ptr(means), 0, stream_ptr()))
            dist.all_reduce(means, op=dist.ReduceOp.SUM)
            means.mul_(1.0 / dist.get_world_size())
            check(lib().fqss_kd_loss_dp(*args, ptr(out), ptr(gest) or None, ld_of(gest) if gest is not None else 0, ptr(ws),
                                        ws.numel(), ptr(means), 1, stream_ptr()))
            ctx.gest = gest
            return out
        check(lib().fqss_kd_loss(ptr(est_r), lde, ptr(fest_r), ldf, ptr(tgt_r), ldt, B, T, float(kd_lambda), ptr(out),
                                 ptr(gest) or None, ld_of(gest) if gest is not None else 0, ptr(ws), ws.numel(),
                                 stream_ptr()))
        ctx.gest = gest
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        gest = ctx.gest
        ctx.gest = None
        if gest is None:
            return None, None, None, None
        # dL/dest was produced in forward for d(loss)=1; scale by the incoming scalar (1.0 in training)
        return gest * gout[0], None, None, None


def kd_loss(est, fest, tgt, kd_lambda=0.1):
    return KDLoss.apply(est, fest, tgt, kd_lambda)


# =============================================================================================
# music recipe: L1 + SDR-weighted KD loss (train_env/tasnet_musdbhq/musdbhq_train.py:87-109)
# =============================================================================================
class MusicKDLoss(Function):
    """returns a 3-vector [loss, kd term, task term]; gradient flows from element 0 to wavs."""

    @staticmethod
    def forward(ctx, wavs, fwavs, sources, kd_lambda):
        N.require_cuda(wavs, fwavs, sources)
        if wavs.dim() < 2 or wavs.shape != sources.shape or (fwavs is not None and fwavs.shape != wavs.shape):
            raise N.FqssError("music_kd_loss expects wavs / fwavs / sources of one shape [B, ..., T]")
        B, T = wavs.shape[0], wavs.shape[-1]
        w_r, rows, _, ldw = rows_view(wavs)
        s_r, _, _, lds = rows_view(sources.detach())
        f_r, ldf = None, 0
        if fwavs is not None:
            f_r, _, _, ldf = rows_view(fwavs.detach())
        out = torch.empty(3, device=wavs.device)
        g = alloc_rows(wavs.shape, wavs.device) if ctx.needs_input_grad[0] else None
        ws = torch.empty(int(lib().fqss_music_loss_ws_bytes(B)), dtype=torch.uint8, device=wavs.device)
        check(lib().fqss_music_kd_loss(ptr(w_r), ldw, ptr(f_r) or None, ldf, ptr(s_r), lds, B, rows // B, T, float(kd_lambda),
                                       ptr(out), ptr(g) or None, ld_of(g) if g is not None else 0, ptr(ws), ws.numel(),
                                       stream_ptr()))
        ctx.g = g
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        g = ctx.g
        ctx.g = None
        if g is None:
            return None, None, None, None
        return g * gout[0], None, None, None


def music_kd_loss(wavs, fwavs, sources, kd_lambda=0.1):
    return MusicKDLoss.apply(wavs, fwavs, sources, kd_lambda)
