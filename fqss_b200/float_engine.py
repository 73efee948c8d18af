"""Float ConvTasNet inference on the sm_100a kernels: the KD teacher of the FQSS recipe.

`System.common_step` (train_env/asteroid_librimix/mysystem.py:128-131) runs `fmodel(inputs)` under
`no_grad` every training step; `fmodel` is the deep copy of ConvTasNetQ taken BEFORE quantisation
(train_env/train_utils.py:25), i.e. plain nn.Conv1d / GroupNorm / PReLU modules.  This module runs
that forward (convtasnetq.py:182-223, float branch) without any ATen compute kernel:

    encoder (strided conv kernel) -> gLN -> bottleneck 1x1 -> 24 x fused ConvBlock -> PReLU
    -> mask 1x1 with a fused ReLU * features epilogue -> transposed-conv decoder

All 1x1 convolutions run on the tcgen05 GEMM with split-bf16 operands (x = hi + lo, three-term
product, see include/fqss.h `fqss_pw_gemm`): fp32-grade (~2^-16) contractions on the bf16 tensor pipe,
so the teacher matches the reference's fp32 CPU forward far inside the 1e-3 tier.  Forward only (no
activations are saved; ConvBlock scratch is reused across the stack).  Host side = plumbing only.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _native as N
from . import ops
from . import tcn_engine as E
from ._native import check, lib, ptr, stream_ptr


def eligible(model, x):
    """True when `model` is an un-quantised ConvTasNetQ whose forward this engine implements."""
    if not (x.is_cuda and x.dtype == torch.float32):
        return False
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in model.parameters())):
        return False            # training a float model is the reference's ATen path, not this engine
    enc, dec, mk = model.encoder, model.decoder, model.masker
    if type(enc) is not nn.Conv1d or type(dec) is not nn.ConvTranspose1d or type(model.mul).__name__ != "Mul":
        return False
    if enc.bias is not None or dec.bias is not None or dec.out_channels != 1 or enc.padding[0] or dec.padding[0]:
        return False
    bn, mn = mk.bottleneck, mk.mask_net
    if type(bn[0]) is not nn.GroupNorm or type(bn[1]) is not nn.Conv1d or type(mn[0]) is not nn.PReLU \
            or type(mn[1]) is not nn.Conv1d or type(mn[2]) is not nn.ReLU or mn[0].weight.numel() != 1:
        return False
    b0 = mk.TCN[0]
    sb = b0.shared_block
    if type(sb[0]) is not nn.Conv1d or type(b0.res_conv) is not nn.Conv1d or type(sb[3]) is not nn.Conv1d:
        return False
    Chid, Cio = sb[0].weight.shape[0], sb[0].weight.shape[1]
    F_ = enc.out_channels
    M = (x.shape[-1] - enc.kernel_size[0]) // enc.stride[0] + 1
    if M < 1 or not E.rows_fit(M, max(b.shared_block[3].dilation[0] for b in mk.TCN)):
        return False            # whole-utterance inputs beyond the row kernels' shared-memory staging: plain torch modules
    return Cio % 128 == 0 and Chid % 128 == 0 and F_ % 64 == 0 and sb[3].kernel_size[0] == 3 \
        and sb[1].weight.numel() == 1 and (F_ * model.n_srcs) % 128 == 0


def _generation(params):
    """parallel.param_generation() if any of `params` lives in a ParamArena (updated in place by kernels that do not touch
    torch's version counters), else 0: the frozen teacher's prepared weights stay cached across steps."""
    from . import parallel
    return parallel.param_generation() if any(getattr(p, "_fqss_in_arena", False) for p in params) else 0


def _versions(model):
    ps = list(model.parameters())
    return (_generation(ps),) + tuple((p.data_ptr(), p._version) for p in ps)


def _prepare(model, dev):
    """Weight preparation (split-bf16 operands, folded biases); cached until a parameter changes."""
    key = _versions(model)
    cache = getattr(model, "_fqss_float_prep", None)
    if cache is not None and cache[0] == key:
        return cache[1]
    L = E._libx()
    s = stream_ptr()
    mk = model.masker
    P = {"blocks": []}
    f32 = dict(device=dev, dtype=torch.float32)

    def head(conv):
        w = E.split_bf16_weights(conv.weight)
        n = conv.weight.shape[0]
        return w, torch.ones(n, **f32), (conv.bias.detach().clone() if conv.bias is not None else torch.zeros(n, **f32))
    P["bn"] = head(mk.bottleneck[1])
    P["mask"] = head(mk.mask_net[1])
    nblk = len(mk.TCN)
    for i, blk in enumerate(mk.TCN):
        t, dil = E.block_tensors(blk, None, False)
        has_res = i < nblk - 1
        Chid, Cio = t["W1"].shape[0], t["W1"].shape[1]
        n2 = 2 * Cio if has_res else Cio
        bf = torch.bfloat16
        Q = dict(Wc1=torch.empty((Chid, 3 * Cio), dtype=bf, device=dev), s1_1=torch.empty(Chid, **f32),
                 s0_1=torch.empty(Chid, **f32), dws1=torch.empty(Chid, **f32),
                 Wc2=torch.empty((n2, 3 * Chid), dtype=bf, device=dev), s1_2=torch.empty(n2, **f32),
                 s0_2=torch.empty(n2, **f32), dws2=torch.empty(n2, **f32))
        check(L.fqss_tcn_prep(ptr(t["W1"]), None, None, ptr(t["b1"]) or None, None, None, ptr(Q["Wc1"]), None, ptr(Q["s1_1"]),
                              ptr(Q["s0_1"]), ptr(Q["dws1"]), Chid, Cio, Chid, 0, 1, s))
        # res/skip conv with the second gLN folded in (fqss_tcn_prep_fold): s1_2 = u, s0_2 = v
        off = 0
        if has_res:
            check(L.fqss_tcn_prep_fold(ptr(t["Wres"]), ptr(t["bres"]) or None, ptr(t["g2w"]), ptr(t["g2b"]), ptr(Q["Wc2"]),
                                       ptr(Q["s1_2"]), ptr(Q["s0_2"]), Cio, Chid, n2, 0, s))
            off = Cio
        check(L.fqss_tcn_prep_fold(ptr(t["Wskip"]), ptr(t["bskip"]) or None, ptr(t["g2w"]), ptr(t["g2b"]), ptr(Q["Wc2"]),
                                   ptr(Q["s1_2"]), ptr(Q["s0_2"]), Cio, Chid, n2, off, s))
        Q["wdw"] = t["Wdw"].detach().contiguous()
        P["blocks"].append((t, dil, has_res, Q))
    model._fqss_float_prep = (key, P)
    return P


def _tcn_infer(x0, P, B, Cio, M, ld, dev):
    """24 x ConvBlock, forward only.  x0: fp32 [B,Cio,ld] (pitched).  Returns the skip sum fp32 [B,Cio,ld]."""
    L = E._libx()
    s = stream_ptr()
    bf = torch.bfloat16
    Chid = P["blocks"][0][0]["W1"].shape[0]
    y1 = torch.empty((B, Chid, ld), device=dev)
    a4 = torch.empty((B, 2 * Chid, ld), dtype=bf, device=dev)
    st1 = torch.empty(2 * B + 1, dtype=torch.float64, device=dev)
    st3 = torch.empty(2 * B + 1, dtype=torch.float64, device=dev)
    rc1 = torch.empty(16 + 2 * B, device=dev)
    rc3 = torch.empty(16 + 2 * B, device=dev)
    xs = [x0, torch.empty((B, Cio, ld), device=dev), torch.empty((B, Cio, ld), device=dev)]
    xops = [E.split_bf16_acts(x0[:, :, :M], ld), torch.empty((B, 2 * Cio, ld), dtype=bf, device=dev)]
    skips = [torch.empty((B, Cio, ld), device=dev), torch.empty((B, Cio, ld), device=dev)]
    cur_x, cur_op, cur_skip = 0, 0, None
    for i, (t, dil, has_res, Q) in enumerate(P["blocks"]):
        blk = E.TcnBlock()
        blk.B, blk.M, blk.dil, blk.quant, blk.first_block, blk.has_res = B, M, dil, 0, int(i == 0), int(has_res)
        blk.Cio, blk.Chid, blk.split, blk.ld = Cio, Chid, 2, ld
        for k in ("Wc1", "s1_1", "s0_1", "dws1", "Wc2", "s1_2", "s0_2", "dws2", "wdw"):
            setattr(blk, k, ptr(Q[k]))
        blk.bdw = ptr(t["bdw"])
        blk.slope1, blk.slope3 = ptr(t["slope1"]), ptr(t["slope3"])
        blk.gn1_w, blk.gn1_b, blk.gn2_w, blk.gn2_b = ptr(t["g1w"]), ptr(t["g1b"]), ptr(t["g2w"]), ptr(t["g2b"])
        blk.x_op, blk.x_in = ptr(xops[cur_op]), ptr(xs[cur_x])
        blk.skip_in = ptr(skips[cur_skip]) if cur_skip is not None else None
        blk.y1, blk.stats1, blk.y3, blk.stats3, blk.a4_op = ptr(y1), ptr(st1), None, ptr(st3), ptr(a4)
        blk.rc1, blk.rc3 = ptr(rc1), ptr(rc3)
        nxt_skip = 0 if cur_skip is None else 1 - cur_skip
        blk.skip_out = ptr(skips[nxt_skip])
        if has_res:
            nxt_x = 1 if cur_x != 1 else 2
            blk.x_out, blk.x_out_op = ptr(xs[nxt_x]), ptr(xops[1 - cur_op])
        check(L.fqss_tcn_block_fwd(C.byref(blk), s))
        if has_res:
            cur_x, cur_op = nxt_x, 1 - cur_op
        cur_skip = nxt_skip
    return skips[cur_skip]


def forward(model, x):
    """ConvTasNetQ.forward for the float model: x [B,T] or [B,1,T] -> [B, n_srcs, T]."""
    N.require_cuda(x)
    dev = x.device
    with torch.no_grad():
        P = _prepare(model, dev)
        xin = model.pre_process(x)                                          # [B, n_splitter, T]
        B = xin.shape[0]
        enc, dec, mk = model.encoder, model.decoder, model.masker
        feats = ops.StridedConv.apply(xin, enc.weight, enc.stride[0])       # [B,F,M] fp32, pitched rows
        F_, M = feats.shape[1], feats.shape[2]
        ld = ops.ld_of(feats)
        gn = mk.bottleneck[0]
        normed = ops.pointwise_fq(N.PW_GLN, feats, gamma=gn.weight, beta=gn.bias, quant=False, eps=gn.eps)
        w, s1, s0 = P["bn"]
        x0 = E.pw_gemm(E.split_bf16_acts(normed, ld), w, s1, s0, M)         # [B,Cio,ld]
        Cio = x0.shape[1]
        skip = _tcn_infer(x0, P, B, Cio, M, ld, dev)
        act = ops.pointwise_fq(N.PW_PRELU, skip[:, :, :M], slope=mk.mask_net[0].weight, quant=False)
        w, s1, s0 = P["mask"]
        S = model.n_srcs
        masked = E.pw_gemm(E.split_bf16_acts(act, ld), w, s1, s0, M, mul=feats)   # relu(conv)*feats: [B,S*F,ld]
        dec_in = masked.view(B * S, F_, ld)[:, :, :M]
        out = ops.TransposedConv1.apply(dec_in, dec.weight, dec.stride[0])  # [B*S,1,T]
        out = out.reshape((model.n_combiner, B, S, 1, -1)) if model.n_combiner == 1 else None
        if out is None:
            raise N.FqssError("float engine: a float model with an output combiner is not on the FQSS path")
        return model.post_process(out)


# ---------------------------------------------------------------------------------------------
# Skip-less block stack of the music model's float teacher (convtasnetq_music.py:53-199 with quantisation disabled:
# musdbhq_train.py builds the teacher from the same module tree).  Forward only; the layers around the stack stay on the
# per-layer wrappers.
# ---------------------------------------------------------------------------------------------
def _noskip_blocks(masker):
    return [b for rep in masker.network[2] for b in rep]


def noskip_eligible(masker, x):
    """True when `masker` is the music MaskGenerator with every quantiser disabled and nothing needs a gradient."""
    from .qat import qat_layers as QL
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 3):
        return False
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in masker.parameters())):
        return False
    blocks = _noskip_blocks(masker)
    b0 = blocks[0]
    if not hasattr(b0, "net") or not hasattr(b0.net[3], "net"):
        return False
    ds = b0.net[3].net
    if not (isinstance(b0.net[0], QL.Conv1dNlQ) and isinstance(b0.net[2], QL.GroupNormQ) and isinstance(ds[0], QL.Conv1dNlQ)
            and isinstance(ds[2], QL.GroupNormQ) and isinstance(ds[3], QL.Conv1dQ) and isinstance(b0.add, QL.AddQ)):
        return False
    for blk in blocks:
        for m in blk.modules():
            if isinstance(m, QL.LayerQ) and not (isinstance(m.activation_fake_quantize, nn.Identity)
                                                 and isinstance(m.weight_fake_quantize, nn.Identity)):
                return False
            if isinstance(m, nn.PReLU) and m.weight.numel() != 1:
                return False
    Chid, Cio = b0.net[0].conv1d.weight.shape[0], b0.net[0].conv1d.weight.shape[1]
    if Cio % 128 or Chid % 128 or ds[0].conv1d.kernel_size[0] != 3 or ds[0].conv1d.groups != Chid:
        return False
    return E.rows_fit(x.shape[-1], max(b.net[3].net[0].conv1d.dilation[0] for b in blocks))


def _prepare_noskip(masker, dev):
    blocks = _noskip_blocks(masker)
    ps = [p for b in blocks for p in b.parameters()]
    key = (_generation(ps),) + tuple((p.data_ptr(), p._version) for p in ps)
    cache = getattr(masker, "_fqss_float_prep", None)
    if cache is not None and cache[0] == key:
        return cache[1]
    L = E._libx()
    s = stream_ptr()
    f32 = dict(device=dev, dtype=torch.float32)
    bf = torch.bfloat16
    out = []
    for blk in blocks:
        t, dil = E.block_tensors_noskip(blk)
        Chid, Cio = t["W1"].shape[0], t["W1"].shape[1]
        Q = dict(Wc1=torch.empty((Chid, 3 * Cio), dtype=bf, device=dev), s1_1=torch.empty(Chid, **f32),
                 s0_1=torch.empty(Chid, **f32), dws1=torch.empty(Chid, **f32),
                 Wc2=torch.empty((Cio, 3 * Chid), dtype=bf, device=dev), s1_2=torch.empty(Cio, **f32),
                 s0_2=torch.empty(Cio, **f32), dws2=torch.empty(Cio, **f32))
        check(L.fqss_tcn_prep(ptr(t["W1"]), None, None, ptr(t["b1"]) or None, None, None, ptr(Q["Wc1"]), None, ptr(Q["s1_1"]),
                              ptr(Q["s0_1"]), ptr(Q["dws1"]), Chid, Cio, Chid, 0, 1, s))
        # residual 1x1 conv with the block's second gLN folded in (fqss_tcn_prep_fold): s1_2 = u, s0_2 = v
        check(L.fqss_tcn_prep_fold(ptr(t["Wres"]), ptr(t["bres"]) or None, ptr(t["g2w"]), ptr(t["g2b"]), ptr(Q["Wc2"]),
                                   ptr(Q["s1_2"]), ptr(Q["s0_2"]), Cio, Chid, Cio, 0, s))
        Q["wdw"] = t["Wdw"].detach().contiguous()
        out.append((t, dil, Q))
    masker._fqss_float_prep = (key, out)
    return out


def tcn_noskip_infer(masker, h):
    """The float block stack of the music MaskGenerator on the fused engine: h [B,Cio,M] -> [B,Cio,M]."""
    N.require_cuda(h)
    dev = h.device
    L = E._libx()
    s = stream_ptr()
    bf = torch.bfloat16
    with torch.no_grad():
        P = _prepare_noskip(masker, dev)
        B, Cio, M = h.shape
        ld = (M + 7) // 8 * 8
        x0 = E._as_pitched(h.detach(), ld)
        x0 = x0 if x0.shape[-1] == ld else x0.as_strided((B, Cio, ld), (Cio * ld, ld, 1))
        Chid = P[0][0]["W1"].shape[0]
        y1 = torch.empty((B, Chid, ld), device=dev)
        a4 = torch.empty((B, 2 * Chid, ld), dtype=bf, device=dev)
        st1 = torch.empty(2 * B + 1, dtype=torch.float64, device=dev)
        st3 = torch.empty(2 * B + 1, dtype=torch.float64, device=dev)
        rc1 = torch.empty(16 + 2 * B, device=dev)
        rc3 = torch.empty(16 + 2 * B, device=dev)
        xs = [x0, torch.empty((B, Cio, ld), device=dev), torch.empty((B, Cio, ld), device=dev)]
        xops = [E.split_bf16_acts(x0[:, :, :M], ld), torch.empty((B, 2 * Cio, ld), dtype=bf, device=dev)]
        cur_x, cur_op = 0, 0
        for i, (t, dil, Q) in enumerate(P):
            blk = E.TcnBlock()
            blk.B, blk.M, blk.dil, blk.quant, blk.first_block, blk.has_res, blk.no_skip = B, M, dil, 0, int(i == 0), 1, 1
            blk.Cio, blk.Chid, blk.split, blk.ld = Cio, Chid, 2, ld
            for k in ("Wc1", "s1_1", "s0_1", "dws1", "Wc2", "s1_2", "s0_2", "dws2", "wdw"):
                setattr(blk, k, ptr(Q[k]))
            blk.bdw = ptr(t["bdw"]) if t["bdw"] is not None else ptr(E.zero_bias(Chid, dev))
            blk.slope1, blk.slope3 = ptr(t["slope1"]), ptr(t["slope3"])
            blk.gn1_w, blk.gn1_b, blk.gn2_w, blk.gn2_b = ptr(t["g1w"]), ptr(t["g1b"]), ptr(t["g2w"]), ptr(t["g2b"])
            blk.x_op, blk.x_in = ptr(xops[cur_op]), ptr(xs[cur_x])
            blk.y1, blk.stats1, blk.y3, blk.stats3, blk.a4_op = ptr(y1), ptr(st1), None, ptr(st3), ptr(a4)
            blk.rc1, blk.rc3 = ptr(rc1), ptr(rc3)
            nxt_x = 1 if cur_x != 1 else 2
            blk.x_out, blk.x_out_op = ptr(xs[nxt_x]), ptr(xops[1 - cur_op])
            check(L.fqss_tcn_block_fwd(C.byref(blk), s))
            cur_x, cur_op = nxt_x, 1 - cur_op
        return xs[cur_x][:, :, :M]
