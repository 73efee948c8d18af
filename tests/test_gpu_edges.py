"""GPU: parity of the filterbank edges (Conv1dEncoderQ qat_layers.py:993-1046, ConvTr1dDecoderQ
:1305-1361, ResidualErrorBlock :1105-1220) with the ORACLE's exact inputs and output gradients (teacher-forced, the
same contract as the ConvBlock tests): every quantiser bit-exact on the kernel's own pre-activation, every conv against
its fp64 definition, the composition on the oracle's quantisation grid (rare +-1 code moves where an fp32 reassociation
of a 1 024-term sum crosses a rounding boundary), input / parameter gradients to the fp32-path tolerance 1e-3."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import fqss_oracle as O
from parity_log import record

pytestmark = pytest.mark.gpu
DEV = "cuda"


def T(a):
    return torch.from_numpy(np.asarray(a))


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _setup(golden):
    from fqss_b200.testing import small_model_pair
    g = golden("model_small.npz")
    model, fmodel = small_model_pair(DEV, seed=0)
    calib = {k[6:]: T(g[k]) for k in g.files if k.startswith("calib/")}
    model.load_state_dict(calib, strict=True)
    for m in model.modules():
        if hasattr(m, "observer_mode"):
            m.observer_mode = False
    cfg = O.SeparatorConfig(n_filters=64, bn_chan=32, hid_chan=64, n_blocks=3, n_repeats=2)
    P = O.Params({k: v.clone() for k, v in calib.items()})
    fP = O.Params({k[8:]: T(g[k]) for k in g.files if k.startswith("teacher/")})
    st = O.QuantState(observe=False, weights_seen=True)
    P.leafify()
    taps = {}
    est = O.separator_forward(P, T(g["mix"]), cfg, st, quant=True, tap=taps)
    for k in ("split", "encoder", "masked", "decoder"):
        if taps[k].requires_grad:
            taps[k].retain_grad()
    with torch.no_grad():
        fest = O.separator_forward(fP, T(g["mix"]), cfg, quant=False)
    loss, _ = O.fqss_kd_loss(est, fest, T(g["src"]), 0.1)
    loss.backward()
    return g, model, P, taps, cfg


def _flips(t, ref, step):
    d = (t.detach().cpu() - ref.detach()).abs()
    return (d > 0.5 * step).float().mean().item(), d.max().item() / step


def _check_param_grads(mod, prefix, P, tag, tol=1e-3):
    bad, meas = [], {}
    for k, p in mod.named_parameters():
        go = P[prefix + k].grad
        if go is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        meas["grad/" + k] = rel(p.grad, go)
        if not (meas["grad/" + k] < tol or (p.grad.cpu() - go).abs().max() < 1e-6):
            bad.append((k, meas["grad/" + k], (p.grad.cpu() - go).abs().max().item()))
    record(tag, **meas)
    assert not bad, bad


def test_decoder_rqb_teacher_forced(golden):
    g, model, P, taps, cfg = _setup(golden)
    Y = taps["masked"].detach()
    B = Y.shape[0]
    Yd = Y.reshape(B * cfg.n_src, cfg.n_filters, -1).to(DEV).requires_grad_(True)
    model.zero_grad(set_to_none=True)
    out = model.decoder(Yd)                                            # [2, B*S, 1, T]
    ref = taps["decoder"]
    meas = {}
    for i, qname in enumerate(("decoder.activation_fake_quantize", "decoder.activation_fake_quantize_residual")):
        step = (P[qname + ".max_range"] - P[qname + ".min_range"]).item() / 255
        meas["y%d_flip_rate" % i], meas["y%d_max_code_diff" % i] = _flips(out[i], ref[i], step)
        assert meas["y%d_flip_rate" % i] <= 1e-3 and meas["y%d_max_code_diff" % i] <= 1.01, (i, meas)
    out.backward(ref.grad.to(DEV))
    meas["g_masked_rel"] = rel(Yd.grad, taps["masked"].grad.reshape(Yd.shape))
    record("edges_teacher_forced/decoder_rqb", **meas)
    assert meas["g_masked_rel"] < 1e-3, meas
    _check_param_grads(model.decoder, "decoder.", P, "edges_teacher_forced/decoder_rqb")


def test_decoder_rqb_kernels_exact_on_own_inputs(golden):
    """Kernel by kernel through the RQB (qat_layers.py:1188-1202, 1330-1354): transposed conv vs its fp64 definition, out-FQ
    bit-exact on the kernel's own output, re-encoder conv vs fp64, FQ(Y - Yq) bit-exact, second decode, residual FQ."""
    from fqss_b200 import ops
    from fqss_b200 import _native as NN
    g, model, P, taps, cfg = _setup(golden)
    dec = model.decoder
    Y = taps["masked"].detach()
    Yd = Y.reshape(-1, cfg.n_filters, Y.shape[-1]).to(DEV)
    stride = cfg.stride
    with torch.no_grad():
        wd = dec.weight_fake_quantize(dec.convTr1d.weight)
        assert torch.equal(wd.cpu(), O.fq_weight(P["decoder.convTr1d.weight"].detach(), P["decoder.weight_fake_quantize.min_range"].detach(),
                                                P["decoder.weight_fake_quantize.max_range"].detach()))
        y0p = ops.TransposedConv1.apply(Yd, wd, stride)
        assert rel(y0p, F.conv_transpose1d(Yd.double(), wd.double(), None, stride=stride)) < 1e-6
        q = dec.activation_fake_quantize
        y0 = dec._finish(NN.PW_IDENT, y0p)
        assert torch.equal(y0.cpu(), O.fq_act(y0p.cpu(), q.min_range.detach().cpu(), q.max_range.detach().cpu()))
        rqb = dec.residual_error_block
        we = rqb.weight_fake_quantize(rqb.residual_encoder.weight)
        Yq = ops.StridedConv.apply(y0, we, stride)
        assert rel(Yq, F.conv1d(y0.double(), we.double(), None, stride=stride)) < 1e-6
        q = rqb.activation_fake_quantize
        Y1 = rqb._finish(NN.PW_SUB, Yd, Yq)
        assert torch.equal(Y1.cpu(), O.fq_act(Yd.cpu() - Yq.cpu(), q.min_range.detach().cpu(), q.max_range.detach().cpu()))
        y1p = ops.TransposedConv1.apply(Y1, wd, stride)
        assert rel(y1p, F.conv_transpose1d(Y1.double(), wd.double(), None, stride=stride)) < 1e-6
        q = dec.activation_fake_quantize_residual
        y1 = dec._finish(NN.PW_IDENT, y1p, quantizer=q)
        assert torch.equal(y1.cpu(), O.fq_act(y1p.cpu(), q.min_range.detach().cpu(), q.max_range.detach().cpu()))
        full = dec(Yd)
        assert torch.equal(full[0], y0) and torch.equal(full[1], y1)


def test_encoder_teacher_forced(golden):
    g, model, P, taps, cfg = _setup(golden)
    x = taps["split"].detach().to(DEV)
    model.zero_grad(set_to_none=True)
    feats = model.encoder(x)
    q = "encoder.activation_fake_quantize"
    step = (P[q + ".max_range"] - P[q + ".min_range"]).item() / 255
    frac, worst = _flips(feats, taps["encoder"], step)
    record("edges_teacher_forced/encoder", flip_rate=frac, max_code_diff=worst)
    assert frac <= 1e-3 and worst <= 1.01, (frac, worst)
    feats.backward(taps["encoder"].grad.to(DEV))
    _check_param_grads(model.encoder, "encoder.", P, "edges_teacher_forced/encoder")
    # the conv itself against its fp64 definition, the quantiser bit-exact on the kernel's own output
    from fqss_b200 import ops
    from fqss_b200 import _native as NN
    with torch.no_grad():
        w = model.encoder.weight_fake_quantize(model.encoder.conv1d.weight)
        yp = ops.StridedConv.apply(x, w, cfg.stride)
        assert rel(yp, F.conv1d(x.double(), w.double(), None, stride=cfg.stride)) < 1e-6
        qa = model.encoder.activation_fake_quantize
        assert torch.equal(model.encoder._finish(NN.PW_IDENT, yp).cpu(),
                           O.fq_act(yp.cpu(), qa.min_range.detach().cpu(), qa.max_range.detach().cpu()))
