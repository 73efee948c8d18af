"""GPU: kernel-level parity of the fused ConvBlock engine at the benchmarked widths (128 / 512 channels, M = 3999).

Two levels, both against the CPU oracle (oracle/fqss_oracle.py, pinned bit-exact to the unmodified reference):

 * EXACT: every quantisation code a fused kernel stores is re-derived on the CPU with the oracle's quantiser
   (qat_quant.py:136-147) from the kernel's OWN stored pre-activation and must be `torch.equal`:
   x_op <- x, code1 <- PReLU(y1), code3 <- PReLU(y3), a4_op <- gLN2(decode(code3)), x_out / x_out_op <- x + FQ(res_y),
   skip_out <- skip_in + FQ(skip_y).  The 1x1 convolutions run on integer codes, so y1 / res_y / skip_y must equal
   fma(sum_k cw*ca, s1, s0) of the exact integer dot product; the depthwise output is held to fp32 round-off.
 * TEACHER FORCED at batch 32 (the benchmarked configuration) and batch 4 (strong scaling at 8 GPUs): one block is fed
   the same on-grid input and the same output gradients as the oracle's block; outputs must sit on the oracle's grid
   (flip rate, +-1 code), input / parameter gradients within the bf16-gradient tier (north_star: 1e-2).

Every test records what it measured through tests/parity_log.py."""
import pytest
import torch
import torch.nn.functional as F

import fqss_oracle as O
from parity_log import reassociated_pointwise_convs, record

pytestmark = pytest.mark.gpu
DEV = "cuda"
HOT_KW = dict(n_spks=2, kernel_size=16, stride=8, n_filters=512, bn_chan=128, hid_chan=512, n_blocks=8, n_repeats=1)
TOTAL = 24          # the blocks under test are placed inside a 24-block stack (block TOTAL-1 alone has no residual output)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


_CACHE = {}


def _hot_model(T=32000):
    """Recipe-width student (8 blocks, dilations 1..128), ranges calibrated by two observer passes on the GPU."""
    if "model" in _CACHE:
        return _CACHE["model"]
    from fqss_b200.qat.models.load_model import enable_observer
    from fqss_b200.testing import model_pair
    model, _ = model_pair(HOT_KW, DEV, seed=0)
    gen = torch.Generator().manual_seed(11)
    mix = (torch.randn(2, 2, T, generator=gen) * 0.05).sum(1, keepdim=True).to(DEV)
    with torch.no_grad():
        model(mix)
        model(mix)
    enable_observer(model, False)
    _CACHE["model"] = model
    return model


def _block_inputs(model, B, T, seed):
    """Realistic on-grid block inputs: run the per-layer path once and capture every block's input and the running skip
    sum that enters its AddQ."""
    from fqss_b200.qat.models.convtasnetq import MaskGenerator
    gen = torch.Generator().manual_seed(seed)
    mix = (torch.randn(B, 2, T, generator=gen) * 0.05).sum(1, keepdim=True).to(DEV)
    xs, totals, hooks = {}, {}, []
    for i, blk in enumerate(model.masker.TCN):
        hooks.append(blk.register_forward_pre_hook(lambda m, inp, i=i: xs.__setitem__(i, inp[0].detach().clone())))
    for i, add in enumerate(model.masker.adds):
        hooks.append(add.register_forward_pre_hook(lambda m, inp, i=i: totals.__setitem__(i + 1, inp[0].detach().clone())))
    MaskGenerator.use_fused = False
    try:
        with torch.no_grad():
            model(mix)
    finally:
        MaskGenerator.use_fused = True
        for h in hooks:
            h.remove()
    return xs, totals


def _q(mod):
    q = mod.activation_fake_quantize
    return q.min_range.detach().cpu(), q.max_range.detach().cpu()


def _decode(code, rmin, rmax):
    return (rmax - rmin) / 255 * code + rmin          # actqf_decode: delta * c, then + min (two roundings)


def _gln_from_rc(a, rc, gamma, beta, B):
    """gln_apply of csrc/tcn_common.cuh with the kernel's own per-sample constants: scale = rstd*gamma,
    shift = (-scale*mu) + beta, n = a*scale + shift -- every op separately rounded, as on the device."""
    mu = rc[16:16 + 2 * B:2].reshape(B, 1, 1)
    rstd = rc[17:17 + 2 * B:2].reshape(B, 1, 1)
    scale = rstd * gamma.reshape(1, -1, 1)
    shift = (-scale) * mu + beta.reshape(1, -1, 1)
    return a * scale + shift


def _run_fused_block(model, i, x, skip_in, keep=True, grad=False):
    from fqss_b200 import tcn_engine as E
    masker = model.masker
    blk = masker.TCN[i]
    qin = masker.bottleneck[1].activation_fake_quantize if i == 0 else masker.TCN[i - 1].add.activation_fake_quantize
    adds = [masker.adds[i - 1] if i > 0 else None]
    E.KEEP_STATES = keep
    try:
        if grad:
            x = x.detach().clone().requires_grad_(True)
            if skip_in is not None:
                skip_in = skip_in.detach().clone().requires_grad_(True)
            xo, ss = E.FusedTCNFunction.apply(x, skip_in, (True, (blk.shared_block[3].conv1d.dilation[0],), (qin.min_range, qin.max_range), i, TOTAL),
                                              *[E.block_tensors(blk, adds[0], True)[0][k] for k in E._BLOCK_SLOTS])
        else:
            xg = x.detach().clone().requires_grad_(True)      # grad-enabled: the training forward (saves everything)
            xo, ss = E.fused_tcn(xg, [blk], adds, True, (qin.min_range, qin.max_range), start=i, total=TOTAL, skip_in=skip_in)
        st = E.LAST_STATES[0] if keep else None
    finally:
        E.KEEP_STATES = False
    return xo, ss, st, qin, x, skip_in


def _exact_chain(model, i, x, skip_in, tag):
    """Re-derive every stored tensor of block i on the CPU from the kernel's own inputs; returns the measured counts."""
    xo, ss, st, qin, _, _ = _run_fused_block(model, i, x, skip_in)
    torch.cuda.synchronize()
    blk = model.masker.TCN[i]
    sb = blk.shared_block
    B, Cio, M = x.shape
    A = {k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in st.act.items()}
    Pp = {k: v.detach().cpu() for k, v in st.prep.items()}
    cut = lambda t: t[:, :, :M]
    out = {}
    # ---- operand of the expand conv: integer codes of x w.r.t. the producing quantiser
    qmin, qmax = qin.min_range.detach().cpu(), qin.max_range.detach().cpu()
    xc = x.detach().cpu()
    x_codes = O.act_codes(xc, qmin, qmax)
    assert torch.equal(cut(A["x_op"]).float(), x_codes), (tag, "x_op")
    # ---- weight codes (qat_quant.py:126-135)
    wq = sb[0].weight_fake_quantize
    w1c = O.weight_codes(sb[0].conv1d.weight.detach().cpu(), wq.min_range.detach().cpu(), wq.max_range.detach().cpu()).reshape(-1, Cio)
    assert torch.equal(Pp["Wc1"].float(), w1c), (tag, "Wc1")
    # ---- y1 = fma(exact integer dot, s1, s0)
    acc = torch.einsum("ok,bkm->bom", w1c.double(), x_codes.double())
    y1_ref = (acc * Pp["s1_1"].double().reshape(1, -1, 1) + Pp["s0_1"].double().reshape(1, -1, 1)).float()
    y1 = cut(A["y1"])
    out["y1_mismatch_frac"] = (y1 != y1_ref).float().mean().item()
    out["y1_max_rel_dev"] = ((y1 - y1_ref).abs() / (y1_ref.abs() + 1e-20)).max().item() if out["y1_mismatch_frac"] else 0.0
    assert out["y1_mismatch_frac"] < 1e-6 and out["y1_max_rel_dev"] < 2.5e-7, (tag, out)
    # how far the reference's fp32 convolution and ours are from exact arithmetic
    w_hat = O.fq_weight(sb[0].conv1d.weight.detach().cpu(), wq.min_range.detach().cpu(), wq.max_range.detach().cpu())
    y_o = F.conv1d(xc, w_hat, sb[0].conv1d.bias.detach().cpu())
    y_64 = F.conv1d(xc.double(), w_hat.double(), sb[0].conv1d.bias.detach().cpu().double())
    out["y1_err_vs_fp64_ours"] = rel(y1, y_64)
    out["y1_err_vs_fp64_oracle"] = rel(y_o, y_64)
    out["y1_rel_vs_oracle"] = rel(y1, y_o)
    assert out["y1_rel_vs_oracle"] < 1e-5, (tag, out)
    # ---- code1 = FQ1 code of PReLU(y1): bit-exact
    q1min, q1max = _q(sb[0])
    slope1 = sb[0].nl.weight.detach().cpu()
    code1 = O.act_codes(F.prelu(y1, slope1), q1min, q1max)
    assert torch.equal(cut(A["code1"]).float(), code1), (tag, "code1", (cut(A["code1"]).float() != code1).float().mean().item())
    # ---- gLN1 statistics (integer code sums on the device) against fp64
    a1 = _decode(code1, q1min, q1max)
    mu64 = a1.double().mean(dim=(1, 2))
    rstd64 = 1.0 / torch.sqrt(a1.double().var(dim=(1, 2), unbiased=False) + 1e-8)
    rc1 = A["rc1"]
    out["gln1_mu_rel"] = ((rc1[16:16 + 2 * B:2].double() - mu64).abs() / (mu64.abs() + 1e-12)).max().item()
    out["gln1_rstd_rel"] = ((rc1[17:17 + 2 * B:2].double() - rstd64).abs() / rstd64).max().item()
    assert out["gln1_mu_rel"] < 1e-5 and out["gln1_rstd_rel"] < 1e-6, (tag, out)
    # ---- a2 = FQ2(gLN1(a1)) with the kernel's constants, depthwise conv -> y3 (fp32 round-off), code3 exact
    n1 = _gln_from_rc(a1, rc1, sb[2].groupnorm.weight.detach().cpu(), sb[2].groupnorm.bias.detach().cpu(), B)
    q2min, q2max = _q(sb[2])
    a2 = O.fq_act(n1, q2min, q2max)
    dwq = sb[3].weight_fake_quantize
    wdw_hat = O.fq_weight(sb[3].conv1d.weight.detach().cpu(), dwq.min_range.detach().cpu(), dwq.max_range.detach().cpu())
    assert torch.equal(Pp["wdw"].reshape(wdw_hat.shape), wdw_hat), (tag, "wdw")
    d = sb[3].conv1d.dilation[0]
    y3_ref = F.conv1d(a2.double(), wdw_hat.double(), sb[3].conv1d.bias.detach().cpu().double(), padding=d, dilation=d, groups=a2.shape[1])
    y3 = cut(A["y3"])
    scale3 = y3_ref.abs().max().item()
    out["y3_max_abs_dev_over_peak"] = ((y3.double() - y3_ref).abs().max().item() / scale3)
    assert out["y3_max_abs_dev_over_peak"] < 1e-6, (tag, out)
    q3min, q3max = _q(sb[3])
    slope3 = sb[3].nl.weight.detach().cpu()
    code3 = O.act_codes(F.prelu(y3, slope3), q3min, q3max)
    assert torch.equal(cut(A["code3"]).float(), code3), (tag, "code3", (cut(A["code3"]).float() != code3).float().mean().item())
    # ---- a4 operand = FQ4 code of gLN2(decode(code3)) with the kernel's constants: bit-exact
    a3 = _decode(code3, q3min, q3max)
    rc3 = A["rc3"]
    mu64 = a3.double().mean(dim=(1, 2))
    rstd64 = 1.0 / torch.sqrt(a3.double().var(dim=(1, 2), unbiased=False) + 1e-8)
    out["gln2_mu_rel"] = ((rc3[16:16 + 2 * B:2].double() - mu64).abs() / (mu64.abs() + 1e-12)).max().item()
    out["gln2_rstd_rel"] = ((rc3[17:17 + 2 * B:2].double() - rstd64).abs() / rstd64).max().item()
    assert out["gln2_mu_rel"] < 1e-5 and out["gln2_rstd_rel"] < 1e-6, (tag, out)
    n3 = _gln_from_rc(a3, rc3, sb[5].groupnorm.weight.detach().cpu(), sb[5].groupnorm.bias.detach().cpu(), B)
    q4min, q4max = _q(sb[5])
    code4 = O.act_codes(n3, q4min, q4max)
    assert torch.equal(cut(A["a4_op"]).float(), code4), (tag, "a4_op", (cut(A["a4_op"]).float() != code4).float().mean().item())
    # ---- res / skip convs on integer codes, then the four 128-wide quantisers: bit-exact
    has_res = st.has_res
    wr, ws = blk.res_conv, blk.skip_conv
    wcs = []
    if has_res:
        wcs.append(O.weight_codes(wr.conv1d.weight.detach().cpu(), wr.weight_fake_quantize.min_range.detach().cpu(),
                                  wr.weight_fake_quantize.max_range.detach().cpu()).reshape(Cio, -1))
    wcs.append(O.weight_codes(ws.conv1d.weight.detach().cpu(), ws.weight_fake_quantize.min_range.detach().cpu(),
                              ws.weight_fake_quantize.max_range.detach().cpu()).reshape(Cio, -1))
    w2c = torch.cat(wcs, 0)
    assert torch.equal(Pp["Wc2"].float(), w2c), (tag, "Wc2")
    acc2 = torch.einsum("ok,bkm->bom", w2c.double(), code4.double())
    y2_ref = (acc2 * Pp["s1_2"].double().reshape(1, -1, 1) + Pp["s0_2"].double().reshape(1, -1, 1)).float()
    off = Cio if has_res else 0
    skip_y = cut(A["skip_y"])
    mism = (skip_y != y2_ref[:, off:]).float().mean().item()
    if has_res:
        res_y = cut(A["res_y"])
        mism = max(mism, (res_y != y2_ref[:, :Cio]).float().mean().item())
    out["resskip_mismatch_frac"] = mism
    assert mism < 1e-6, (tag, out)
    if has_res:
        qrmin, qrmax = _q(blk.res_conv)
        qamin, qamax = _q(blk.add)
        z = xc + O.fq_act(res_y, qrmin, qrmax)
        cz = O.act_codes(z, qamin, qamax)
        assert torch.equal(cut(A["x_out_op"]).float(), cz), (tag, "x_out_op")
        assert torch.equal(cut(A["x_out"]), _decode(cz, qamin, qamax)), (tag, "x_out")
        assert torch.equal(xo.detach().cpu(), _decode(cz, qamin, qamax)), (tag, "returned x_out")
    qsmin, qsmax = _q(blk.skip_conv)
    sk = O.fq_act(skip_y, qsmin, qsmax)
    if i > 0:
        qdmin, qdmax = _q(model.masker.adds[i - 1])
        sk = O.fq_act(skip_in.detach().cpu() + sk, qdmin, qdmax)
    assert torch.equal(cut(A["skip_out"]), sk), (tag, "skip_out")
    assert torch.equal(ss.detach().cpu(), sk), (tag, "returned skip_out")
    out["elements_checked_per_hidden_tensor"] = float(code1.numel())
    record("exact_chain/" + tag, **out)
    return out


@pytest.mark.parametrize("i,B", [(0, 2), (3, 2), (7, 2)])
def test_fused_block_stored_codes_bit_exact(i, B):
    """M = 3999 (pitch 4000, the specialised epilogues), dilations 1 / 8 / 128."""
    model = _hot_model()
    xs, totals = _block_inputs(model, B, 32000, seed=21)
    _exact_chain(model, i, xs[i], totals.get(i), "M3999_B%d_block%d" % (B, i))


def test_fused_block_stored_codes_bit_exact_ragged():
    """Ragged rows (M = 1001: M % 4 = 1, pitch 1008, generic epilogue addressing) and a dilation that is not a power of two
    (generic tap path)."""
    model = _hot_model()
    xs, totals = _block_inputs(model, 3, 8 * 1001 + 8, seed=22)
    assert xs[2].shape[-1] == 1001
    conv = model.masker.TCN[2].shared_block[3].conv1d
    old = conv.dilation, conv.padding
    conv.dilation, conv.padding = (3,), (3,)
    try:
        _exact_chain(model, 2, xs[2], totals.get(2), "M1001_B3_block2_dil3")
    finally:
        conv.dilation, conv.padding = old


def test_fused_block_stored_codes_bit_exact_batch32():
    """The benchmarked configuration: batch 32, M = 3999 (16 384-CTA row grids, integer statistic sums over 2 M frames)."""
    model = _hot_model()
    xs, totals = _block_inputs(model, 32, 32000, seed=23)
    _exact_chain(model, 1, xs[1], totals.get(1), "M3999_B32_block1")


def _oracle_params(model):
    return O.Params({k: v.detach().cpu().clone() for k, v in model.state_dict().items()})


def _teacher_forced(model, i, B, seed, tag, chunk=8):
    cfg = O.SeparatorConfig(n_filters=512, bn_chan=128, hid_chan=512, n_blocks=8, n_repeats=1)
    xs, totals = _block_inputs(model, B, 32000, seed=seed)
    x, skip_in = xs[i], totals.get(i)
    gen = torch.Generator().manual_seed(seed + 1)
    g_out = torch.randn(x.shape, generator=gen) * 1e-3
    g_skip = torch.randn(x.shape, generator=gen) * 1e-3
    model.zero_grad(set_to_none=True)
    xo, ss, _, qin, xg, sg = _run_fused_block(model, i, x, skip_in, keep=False, grad=True)
    torch.autograd.backward([xo, ss], [g_out.to(DEV), g_skip.to(DEV)])
    torch.cuda.synchronize()
    # oracle: same sub-graph, chunked over the batch (the block is per-sample; parameter gradients accumulate)
    st = O.QuantState(observe=False, weights_seen=True)
    pre = "masker.TCN.%d." % i

    def oracle_block():
        P = _oracle_params(model).leafify()
        outs, skips, gxs, gss = [], [], [], []
        for b0 in range(0, B, chunk):
            sl = slice(b0, min(B, b0 + chunk))
            ctx = O._Ctx(P, cfg, st, True, None)
            xin = x[sl].cpu().clone().requires_grad_(True)
            out_o, skip_o = O._tcn_block(ctx, i, xin)
            sin = None
            if i > 0:
                sin = skip_in[sl].cpu().clone().requires_grad_(True)
                skip_o = ctx.aq("masker.adds.%d.activation_fake_quantize" % (i - 1), sin + skip_o)
            torch.autograd.backward([out_o, skip_o], [g_out[sl], g_skip[sl]])
            outs.append(out_o.detach()); skips.append(skip_o.detach()); gxs.append(xin.grad)
            if sin is not None:
                gss.append(sin.grad)
        return P, outs, skips, gxs, gss
    P, outs, skips, gxs, gss = oracle_block()
    # the same oracle with its 1x1 convolutions summed in another channel order: the reference's own reassociation noise
    with reassociated_pointwise_convs():
        P2, outs2, _, gxs2, _ = oracle_block()
    out_o, skip_o, gx_o = torch.cat(outs), torch.cat(skips), torch.cat(gxs)
    m = {}
    for name, ours, ref, qmod in (("out", xo, out_o, model.masker.TCN[i].add), ("skip", ss, skip_o, model.masker.adds[i - 1] if i > 0 else model.masker.TCN[i].skip_conv)):
        qmin, qmax = _q(qmod)
        step = ((qmax - qmin) / 255).item()
        dd = (ours.detach().cpu() - ref).abs() / step
        m[name + "_flip_rate"] = (dd > 0.5).float().mean().item()
        m[name + "_max_code_diff"] = dd.max().item()
    m["gx_rel"] = rel(xg.grad, gx_o)
    if gss:
        m["gskip_in_rel"] = rel(sg.grad, torch.cat(gss))
    m["oracle_self_out_flip_rate"] = ((torch.cat(outs2) - out_o).abs() > 0.5 * ((_q(model.masker.TCN[i].add)[1] - _q(model.masker.TCN[i].add)[0]) / 255).item()).float().mean().item()
    m["oracle_self_gx_rel"] = rel(torch.cat(gxs2), gx_o)
    # per parameter: our deviation from the oracle, and the oracle's deviation from its reassociated self
    excess = ("", 0.0)
    names = [(k, p.grad, P[pre + k].grad, P2[pre + k].grad) for k, p in model.masker.TCN[i].named_parameters()]
    if i > 0:
        ka = "masker.adds.%d.activation_fake_quantize." % (i - 1)
        q = model.masker.adds[i - 1].activation_fake_quantize
        names += [("adds." + nm, pp.grad, P[ka + nm].grad, P2[ka + nm].grad) for nm, pp in (("min_range", q.min_range), ("max_range", q.max_range))]
    # Scalar parameters (activation ranges, PReLU slopes) are single sums over ~10^7 elements of rounding residuals /
    # clipped gradients with heavy cancellation: the relative error of ONE such number is ill-conditioned (the oracle moves
    # some of them by 10 % against itself).  They are judged (a) as a group -- rel-L2 over the vector of the block's scalar
    # gradients -- and (b) each one against max(1e-2 |ref|, 3 x the oracle's own move, 1e-2 x the group's rms).
    sc = [(k, ours, go, go2) for k, ours, go, go2 in names if go is not None and go.numel() == 1]
    v_ours = torch.stack([o.detach().cpu().reshape(()) for _, o, _, _ in sc]).double()
    v_ref = torch.stack([g_.reshape(()) for _, _, g_, _ in sc]).double()
    v_ref2 = torch.stack([g_.reshape(()) for _, _, _, g_ in sc]).double()
    m["scalar_vec_rel"], m["oracle_self_scalar_vec_rel"] = rel(v_ours, v_ref), rel(v_ref2, v_ref)
    rms = float(v_ref.pow(2).mean().sqrt())
    for k, ours, go, go2 in names:
        if go is None:
            assert ours is None or float(ours.abs().max()) == 0.0, (tag, k)
            continue
        assert ours is not None, (tag, k)
        r, r_self = rel(ours, go), rel(go2, go)
        m["grad/" + k] = r
        m["oracle_self/" + k] = r_self
        if go.numel() == 1:
            err = abs(float(ours) - float(go))
            # 3e-2: a PReLU-slope / range gradient behind a gLN is the small remainder of a projection that the gLN backward
            # removes (its output is orthogonal to {1, xhat}); the bf16 rounding of the 512-wide gradient tensors (the
            # documented 1e-2 tier) breaks that orthogonality at 2^-9 and the remainder sees it amplified
            bound = max(3e-2 * abs(float(go)), 3.0 * abs(float(go2) - float(go)), 1e-2 * rms)
            ratio = err / bound
        else:
            # tensors: the bf16-GEMM tier (north_star 1e-2), or 3x what the reference's own reassociation does to this tensor
            # (per-channel weight-range gradients are sums of dWq x rounding residual over the fan-in: 2e-2)
            ratio = r / max(2e-2 if "range" in k else 1e-2, 3.0 * r_self)
        if ratio > excess[1]:
            excess = (k, ratio)
    m["worst_ratio_to_bound"], m["worst_ratio_name"] = excess[1], excess[0]
    m["worst_tensor_grad"] = max(v for k, v in m.items() if k.startswith("grad/") and k.endswith(("conv1d.weight", "conv1d.bias", "groupnorm.weight", "groupnorm.bias")))
    record("teacher_forced/" + tag, **m)
    return m


# north_star tolerances: quantisation codes on the oracle's grid (identical inputs: only fp32 reassociation inside the
# block can move a value across a rounding boundary -> rare +-1 moves), gradients within the bf16-GEMM tier 1e-2
def _assert_teacher_forced(m, tag):
    assert m["out_flip_rate"] <= 1e-3 and m["out_max_code_diff"] <= 1.01, (tag, m)
    assert m["skip_flip_rate"] <= 1e-3 and m["skip_max_code_diff"] <= 1.01, (tag, m)
    assert m["gx_rel"] < 1e-2, (tag, m)
    assert m.get("gskip_in_rel", 0.0) < 1e-3, (tag, m)                # fp32 end to end
    assert m["worst_tensor_grad"] < 1e-2, (tag, m)                    # weights, biases, gLN affine: 1e-2 outright
    # range / slope gradients are sums of rounding residuals and of clipped elements: +-1 code moves (which the oracle
    # produces against ITSELF at the same rate when only its summation order changes) decide them
    assert m["scalar_vec_rel"] <= max(1e-2, 3.0 * m["oracle_self_scalar_vec_rel"]), (tag, m["scalar_vec_rel"], m["oracle_self_scalar_vec_rel"])
    assert m["worst_ratio_to_bound"] <= 1.0, (tag, m["worst_ratio_name"], m)


@pytest.mark.parametrize("i", [0, 7])
def test_fused_block_teacher_forced_batch32(i):
    model = _hot_model()
    tag = "B32_M3999_block%d" % i
    _assert_teacher_forced(_teacher_forced(model, i, 32, seed=31 + i, tag=tag), tag)


def test_fused_block_teacher_forced_batch4():
    model = _hot_model()
    tag = "B4_M3999_block3"
    _assert_teacher_forced(_teacher_forced(model, 3, 4, seed=41, tag=tag, chunk=4), tag)
