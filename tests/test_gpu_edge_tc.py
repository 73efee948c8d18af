"""GPU: the filterbank edges on the tcgen05 GEMM (fqss_b200/edge_engine.py; SURVEY.md 8a row L2, north_star (c)) --
ConvTr1dDecoderQ qat_layers.py:1305-1361, ResidualErrorBlock :1105-1220, Conv1dEncoderQ :993-1046.

Each GEMM-path op against the fp64 definition of the conv it replaces (its operands are integer codes, so the forward differs
from the exact result only by the final affine: 1e-6), its gradients against fp64 autograd (the framed output gradient is a
three-term bf16 operand = the fp32 value: bound 2e-6), the RQB tail against the per-layer composition of the library, and the
whole decoder / encoder layers teacher-forced against the ORACLE on a model wide enough for the tensor-core tiles (128
filters): output codes on the oracle's grid (rare +-1 moves), gradients to the fp32-path tolerance 1e-3."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import fqss_oracle as O
from parity_log import record

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _on_grid(shape, lo, hi, seed):
    """random tensor on the grid of an 8-bit quantiser {lo, hi} + its uint8 codes + the range tensors"""
    from fqss_b200 import ops
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(shape, generator=g) * (hi - lo) * 1.1 + lo - 0.05 * (hi - lo)).to(DEV)
    rmin, rmax = torch.tensor([lo], device=DEV), torch.tensor([hi], device=DEV)
    y, code = ops.fake_quant_codes(x, rmin, rmax)
    return y, code, rmin, rmax


def _dec_weight(Fn, L, seed):
    from fqss_b200 import ops
    g = torch.Generator().manual_seed(seed)
    W = (torch.randn(Fn, 1, L, generator=g) * 0.1).to(DEV)
    wmin, wmax = W.min().reshape(1, 1, 1).clone(), W.max().reshape(1, 1, 1).clone()
    wq = ops.FakeQuantWeight.apply(W, wmin, wmax, 1, 8)
    return W, wmin, wmax, wq


@pytest.mark.parametrize("R,Fn,M,L,H", [(3, 128, 77, 16, 8), (2, 256, 500, 16, 8), (1, 128, 129, 32, 16)])      # 3L <= 128
def test_decode_codes_vs_fp64(R, Fn, M, L, H):
    from fqss_b200 import edge_engine as EE
    x, code, qmin, qmax = _on_grid((R, Fn, M), -0.3, 1.7, 1)
    W, wmin, wmax, wq = _dec_weight(Fn, L, 2)
    xg = x.clone().requires_grad_(True)
    wqg = wq.detach().clone().requires_grad_(True)
    y, _ = EE.DecodeCodes.apply(xg, None, qmin, qmax, wqg, W, wmin, wmax, H)
    xd = x.double().requires_grad_(True)
    wd = wq.detach().double().requires_grad_(True)
    ref = F.conv_transpose1d(xd, wd, None, stride=H)
    assert y.shape == ref.shape
    e_fwd = rel(y, ref)
    g = torch.randn(ref.shape, generator=torch.Generator().manual_seed(3)).to(DEV)
    y.backward(g)
    ref.backward(g.double())
    e_gx, e_gw = rel(xg.grad, xd.grad), rel(wqg.grad, wd.grad)
    record("edge_tc/decode_codes_R%d_F%d_M%d_L%d" % (R, Fn, M, L), fwd=e_fwd, gx=e_gx, gw=e_gw)
    assert e_fwd < 1e-6 and e_gx < 2e-6 and e_gw < 2e-6, (e_fwd, e_gx, e_gw)      # three-term split: fp32-exact operand
    # the same tensor handed over as integer codes (what the mask head emits): identical result
    ld = (M + 7) // 8 * 8
    cb = torch.zeros((R, Fn, ld), dtype=torch.bfloat16, device=DEV)
    cb[:, :, :M] = code.to(torch.bfloat16)
    with torch.no_grad():
        y2, _ = EE.DecodeCodes.apply(x, cb, qmin, qmax, wq, W, wmin, wmax, H)
    assert torch.equal(y2, y.detach())


def test_sub_fq_decode_vs_layer_composition():
    from fqss_b200 import edge_engine as EE, ops
    from fqss_b200 import _native as NN
    R, Fn, M, L, H = 4, 128, 203, 16, 8
    Y, _, _, _ = _on_grid((R, Fn, M), 0.0, 2.0, 5)
    Yq = Y + 0.05 * torch.randn(Y.shape, generator=torch.Generator().manual_seed(6)).to(DEV)
    W, wmin, wmax, wq = _dec_weight(Fn, L, 7)
    rmin, rmax = torch.tensor([-0.08], device=DEV), torch.tensor([0.09], device=DEV)       # clips a few percent
    leaves = [t.detach().clone().requires_grad_(True) for t in (Y, Yq, rmin, rmax, wq)]
    y = EE.SubFQDecode.apply(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], W, wmin, wmax, H)
    ref_l = [t.detach().clone().requires_grad_(True) for t in (Y, Yq, rmin, rmax, wq)]
    Y1 = ops.pointwise_fq(NN.PW_SUB, ref_l[0], ref_l[1], rmin=ref_l[2], rmax=ref_l[3], quant=True)
    ref = ops.TransposedConv1.apply(Y1, ref_l[4], H)
    # the residual codes are bit-exact (same quantiser arithmetic), so the decodes differ only by fp32 summation order
    assert rel(y, ref) < 1e-6
    g = torch.randn(ref.shape, generator=torch.Generator().manual_seed(8)).to(DEV)
    y.backward(g)
    ref.backward(g)
    meas = {n: rel(a.grad, b.grad) for n, a, b in zip(("gY", "gYq", "gmin", "gmax", "gw"), leaves, ref_l)}
    record("edge_tc/sub_fq_decode", **meas)
    assert all(v < 2e-4 for v in meas.values()), meas


@pytest.mark.parametrize("Cin", [1, 2])
def test_framed_code_conv_vs_fp64(Cin):
    from fqss_b200 import edge_engine as EE, ops
    R, Nf, L, H, T = 3, 128, 16, 8, 16 + 8 * 150
    if Cin == 2:          # the splitter's output and its grid
        g = torch.Generator().manual_seed(9)
        x = ops.split_input((torch.randn(R, 1, T, generator=g) * 0.1).to(DEV), 2)
        amin, amax = x._fq_grid
    else:
        x, _, amin, amax = _on_grid((R, 1, T), -0.4, 0.5, 10)
    g = torch.Generator().manual_seed(11)
    W = (torch.randn(Nf, Cin, L, generator=g) * 0.1).to(DEV)
    wmin = W.amin(dim=(1, 2), keepdim=True).clone()
    wmax = W.amax(dim=(1, 2), keepdim=True).clone()
    wq = ops.FakeQuantWeight.apply(W, wmin, wmax, 0, 8)
    xg = x.detach().clone().requires_grad_(Cin == 1)      # the input gradient exists for one input channel (RQB re-encoder)
    wqg = wq.detach().clone().requires_grad_(True)
    y = EE.FramedCodeConv.apply(xg, amin, amax, wqg, W, wmin, wmax, H)
    xd, wd = x.detach().double().requires_grad_(True), wq.detach().double().requires_grad_(True)
    ref = F.conv1d(xd, wd, None, stride=H)
    assert y.shape == ref.shape
    gr = torch.randn(ref.shape, generator=torch.Generator().manual_seed(12)).to(DEV)
    y.backward(gr)
    ref.backward(gr.double())
    meas = dict(fwd=rel(y, ref), gx=rel(xg.grad, xd.grad) if Cin == 1 else 0.0, gw=rel(wqg.grad, wd.grad))
    record("edge_tc/framed_conv_C%d" % Cin, **meas)
    assert meas["fwd"] < 1e-6 and meas["gx"] < 1e-5 and meas["gw"] < 1e-5, meas


# ---------------------------------------------------------------------------------------------
# whole layers, teacher-forced against the oracle (a model with 128 filters: eligible for the GEMM path)
# ---------------------------------------------------------------------------------------------
def _oracle_setup():
    from fqss_b200.testing import FUSED_SMALL_KW, model_pair, oracle_params, _oracle_cfg
    from fqss_b200.qat.models.load_model import enable_observer
    cfg = _oracle_cfg(FUSED_SMALL_KW)
    model, fmodel = model_pair(FUSED_SMALL_KW, DEV, seed=0)
    g = torch.Generator().manual_seed(1)
    src = torch.randn(3, 2, 2400, generator=g) * 0.05
    mix = src.sum(1, keepdim=True)
    P, fP = oracle_params(model), oracle_params(fmodel)
    st = O.calibrate(P, mix, cfg, passes=2)
    model.load_state_dict({k: v for k, v in P.items()}, strict=True)          # the oracle's calibrated ranges
    enable_observer(model, False)
    for m in model.modules():
        if hasattr(m, "observer_mode"):
            m.observer_mode = False
    P.leafify()
    taps = {}
    est = O.separator_forward(P, mix, cfg, st, quant=True, tap=taps)
    for k in ("split", "encoder", "masked", "decoder"):
        if taps[k].requires_grad:
            taps[k].retain_grad()
    with torch.no_grad():
        fest = O.separator_forward(fP, mix, cfg, quant=False)
    loss, _ = O.fqss_kd_loss(est, fest, src, 0.1)
    loss.backward()
    return model, P, taps, cfg


def _flips(t, ref, step):
    d = (t.detach().cpu() - ref.detach()).abs()
    return (d > 0.5 * step).float().mean().item(), d.max().item() / step


def _param_grads(mod, prefix, P):
    out = {}
    for k, p in mod.named_parameters():
        go = P[prefix + k].grad
        if go is None or p.grad is None:
            continue
        out["grad/" + k] = (rel(p.grad, go), (p.grad.cpu() - go).abs().max().item())
    return out


def test_decoder_layer_on_gemm_path_vs_oracle():
    from fqss_b200 import edge_engine as EE
    model, P, taps, cfg = _oracle_setup()
    Y = taps["masked"].detach()
    Yd = Y.reshape(Y.shape[0] * cfg.n_src, cfg.n_filters, -1).to(DEV).requires_grad_(True)
    Yd._fq_src = model.mul.activation_fake_quantize          # the producer tag the model's forward attaches
    assert EE.decoder_eligible(model.decoder, Yd)
    model.zero_grad(set_to_none=True)
    out = model.decoder(Yd)
    ref = taps["decoder"]
    meas = {}
    for i, qname in enumerate(("decoder.activation_fake_quantize", "decoder.activation_fake_quantize_residual")):
        step = (P[qname + ".max_range"] - P[qname + ".min_range"]).item() / 255
        meas["y%d_flip_rate" % i], meas["y%d_max_code_diff" % i] = _flips(out[i], ref[i], step)
        assert meas["y%d_flip_rate" % i] <= 2e-3 and meas["y%d_max_code_diff" % i] <= 1.01, (i, meas)
    out.backward(ref.grad.to(DEV))
    meas["g_masked_rel"] = rel(Yd.grad, taps["masked"].grad.reshape(Yd.shape))
    pg = _param_grads(model.decoder, "decoder.", P)
    meas.update({k: v[0] for k, v in pg.items()})
    record("edge_tc/decoder_layer_vs_oracle", **meas)
    assert meas["g_masked_rel"] < 1e-3, meas
    bad = [(k, v) for k, v in pg.items() if not (v[0] < 1e-3 or v[1] < 1e-6)]
    assert not bad, bad
    # A/B: the SIMT composition on the same input
    EE.ENABLED = False
    try:
        Y2 = Yd.detach().clone().requires_grad_(True)
        Y2._fq_src = Yd._fq_src
        out2 = model.decoder(Y2)
    finally:
        EE.ENABLED = True
    step = (P["decoder.activation_fake_quantize.max_range"] - P["decoder.activation_fake_quantize.min_range"]).item() / 255
    fr, worst = _flips(out[0], out2[0].cpu(), step)
    assert fr <= 2e-3 and worst <= 1.01, (fr, worst)


def test_encoder_layer_on_gemm_path_vs_oracle():
    from fqss_b200 import edge_engine as EE, ops
    model, P, taps, cfg = _oracle_setup()
    x = taps["split"].detach().to(DEV)
    x._fq_grid = ops._splitter_grid(x.device)
    enc = model.encoder
    assert EE.framed_conv_eligible(enc.conv1d, enc.weight_fake_quantize, x, x._fq_grid)
    model.zero_grad(set_to_none=True)
    feats = enc(x)
    q = "encoder.activation_fake_quantize"
    step = (P[q + ".max_range"] - P[q + ".min_range"]).item() / 255
    frac, worst = _flips(feats, taps["encoder"], step)
    feats.backward(taps["encoder"].grad.to(DEV))
    pg = _param_grads(enc, "encoder.", P)
    record("edge_tc/encoder_layer_vs_oracle", flip_rate=frac, max_code_diff=worst, **{k: v[0] for k, v in pg.items()})
    assert frac <= 1e-3 and worst <= 1.01, (frac, worst)
    bad = [(k, v) for k, v in pg.items() if not (v[0] < 1e-3 or v[1] < 1e-6)]
    assert not bad, bad


def test_model_forward_takes_the_gemm_edges():
    """The full model's forward reaches the tensor-core edge ops (counted by the profiler's kernel classes)."""
    from fqss_b200 import edge_engine as EE
    from fqss_b200.testing import FUSED_SMALL_KW, model_pair
    from fqss_b200.qat.models.load_model import enable_observer
    model, _ = model_pair(FUSED_SMALL_KW, DEV, seed=0)
    g = torch.Generator().manual_seed(1)
    mix = (torch.randn(2, 1, 2400, generator=g) * 0.05).to(DEV)
    with torch.no_grad():
        model(mix); model(mix)
    enable_observer(model, False)
    calls = []
    orig = EE.DecodeCodes.forward
    origf = EE.FramedCodeConv.forward

    def spy(ctx, *a):
        calls.append("decode:" + ("codes" if a[1] is not None else "values"))
        return orig(ctx, *a)

    def spyf(ctx, *a):
        calls.append("framed")
        return origf(ctx, *a)
    EE.DecodeCodes.forward = staticmethod(spy)
    EE.FramedCodeConv.forward = staticmethod(spyf)
    try:
        est = model(mix)
        est.square().mean().backward()
    finally:
        EE.DecodeCodes.forward = staticmethod(orig)
        EE.FramedCodeConv.forward = staticmethod(origf)
    assert "decode:codes" in calls and calls.count("framed") == 2, calls
    assert all(p.grad is None or torch.isfinite(p.grad).all() for p in model.parameters())
