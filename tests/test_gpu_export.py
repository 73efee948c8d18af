"""GPU: export path.  (1) the export-time quantiser kernels (csrc/export.cu) against golden vectors of
the unmodified reference classes, bit-exact, incl. the reference's error behaviour; (2) `replace_*_quantizer` surgery;
(3) int8 checkpoint: codes bit-exact against the oracle, and a model rebuilt from the integers alone reproduces the
fake-quantised model's output bit-identically on the fused tcgen05 engine (integer-code operands end to end)."""
import numpy as np
import pytest
import torch

import fqss_oracle as O
import fqss_oracle_export as OE

pytestmark = pytest.mark.gpu
DEV = "cuda"


def T(a):
    return torch.from_numpy(np.asarray(a))


def _act_q(lo, hi):
    from fqss_b200.qat.qat_quant import GradientActivationFakeQuantize
    q = GradientActivationFakeQuantize(True).to(DEV)
    with torch.no_grad():
        q.min_range.fill_(float(lo))
        q.max_range.fill_(float(hi))
    q.enable_observer(False)
    return q


def test_export_activation_quantiser_bit_exact(golden):
    from fqss_b200 import ops
    from fqss_b200.qat.qat_quant import TorchActivationFakeQuantize
    g = golden("export.npz")
    for ci in range(6):
        lo, hi = g[f"act{ci}_range"]
        t = TorchActivationFakeQuantize(_act_q(lo, hi))
        assert t.scale == float(g[f"act{ci}_scale"]) and t.zero_point == int(g[f"act{ci}_zp"])
        x = T(g[f"act{ci}_x"]).to(DEV).requires_grad_(True)
        y = t(x)
        assert torch.equal(y.detach().cpu(), T(g[f"act{ci}_y"])), ci
        y.backward(T(g[f"act{ci}_go"]).to(DEV))
        assert torch.equal(x.grad.cpu(), T(g[f"act{ci}_gx"])), ci
        y2, code = ops.affine_codes_tensor(x.detach(), t.scale, t.zero_point, 0, 255)
        _, q, _ = OE.act_export_forward(g[f"act{ci}_x"], lo, hi)
        assert torch.equal(code.cpu().long(), T(np.clip(q, 0, 255))) and torch.equal(y2, y.detach())


def test_export_weight_quantiser_bit_exact(golden):
    from fqss_b200.qat.qat_quant import GradientWeightFakeQuantize, TorchWeightFakeQuantize
    g = golden("export.npz")
    for ci in range(4):
        axis = int(g[f"w{ci}_axis"])
        w = T(g[f"w{ci}_w"]).to(DEV).requires_grad_(True)
        q = GradientWeightFakeQuantize(True, tuple(w.shape), n_bits=8, ch_out_idx=axis).to(DEV)
        with torch.no_grad():
            q.min_range.copy_(T(g[f"w{ci}_min"]))
            q.max_range.copy_(T(g[f"w{ci}_max"]))
        t = TorchWeightFakeQuantize(q)
        assert torch.equal(t.scales.cpu(), T(g[f"w{ci}_scales"])) and t.axis == axis
        y = t(w)
        assert torch.equal(y.detach().cpu(), T(g[f"w{ci}_y"])), ci
        y.backward(T(g[f"w{ci}_go"]).to(DEV))
        assert torch.equal(w.grad.cpu(), T(g[f"w{ci}_gw"])), ci


def test_export_negative_range_raises_like_reference(golden):
    from fqss_b200.qat.qat_quant import TorchActivationFakeQuantize
    g = golden("export.npz")
    t = TorchActivationFakeQuantize(_act_q(-2.0, -0.5))
    assert t.zero_point == int(g["neg_range_zp"])
    with pytest.raises(RuntimeError, match="zero_point"):
        t(torch.zeros(8, device=DEV))


def test_replace_quantizer_surgery_and_dynamic():
    from fqss_b200.qat import qat_utils as U
    from fqss_b200.qat.qat_quant import TorchActivationFakeQuantize, TorchDymActivationFakeQuantize, TorchWeightFakeQuantize
    from fqss_b200.testing import small_model_pair
    model, _ = small_model_pair(DEV, seed=0)
    x = torch.randn(2, 1, 1600, device=DEV) * 0.1
    with torch.no_grad():
        model(x)
        model(x)
    from fqss_b200.qat.models.load_model import enable_observer
    enable_observer(model, False)
    blk = model.masker.TCN[0]
    U.replace_weight_quantizer(model, "masker.TCN.0.res_conv.weight_fake_quantize", blk.res_conv.weight_fake_quantize)
    U.replace_activation_quantizer(model, "masker.TCN.0.res_conv.activation_fake_quantize", blk.res_conv.activation_fake_quantize)
    assert isinstance(model.masker.TCN[0].res_conv.weight_fake_quantize, TorchWeightFakeQuantize)
    assert isinstance(model.masker.TCN[0].res_conv.activation_fake_quantize, TorchActivationFakeQuantize)

    class _Dyn:
        n_bits, factor = 8, 0.9
    d = TorchDymActivationFakeQuantize(_Dyn())
    v = torch.randn(4096, device=DEV)
    lo, hi = np.float32(0.9 * float(v.min())), np.float32(0.9 * float(v.max()))
    want, _, _ = OE.act_export_forward(v.cpu().numpy(), lo, hi)
    assert torch.equal(d(v).cpu(), T(want))


@pytest.mark.parametrize("kw_name", ["SMALL_KW", "FUSED_SMALL_KW"])
def test_int8_checkpoint_round_trip_bit_identical(kw_name):
    from fqss_b200 import export as X
    from fqss_b200 import testing as TS
    from fqss_b200.qat.models.load_model import enable_observer
    kw = getattr(TS, kw_name)
    model, _ = TS.model_pair(kw, DEV, seed=0)
    torch.manual_seed(3)
    x = torch.randn(3, 1, 4000, device=DEV) * 0.1
    with torch.no_grad():
        model(x)
        model(x)
    enable_observer(model, False)
    model.eval()
    with torch.no_grad():
        want = model(x)
    ckpt = X.export_int8(model)
    # (a) every weight code equals the oracle's quantiser on the same weight / ranges; de-quantised weight == FQ(weight)
    sd = model.state_dict()
    for pname, e in ckpt["weights"].items():
        assert e["code"].dtype == torch.int8
        w = sd[pname].cpu()
        lo, hi = e["min_range"], e["max_range"]
        delta = 2 * torch.maximum(lo.abs(), hi.abs()) / 255
        assert torch.equal(e["code"].float(), torch.clip(torch.round(w / delta), -128, 127)), pname
        assert torch.equal(delta * e["code"].float(), O.fq_weight(w, lo, hi)), pname
    qb, fb = X.checkpoint_bytes(ckpt)
    assert qb < 0.5 * fb
    # (b) a fresh model built from the integers alone gives the same output, bit for bit
    model2, _ = TS.model_pair(kw, DEV, seed=123)          # different init: everything must come from the checkpoint
    X.load_int8(model2, ckpt)
    model2.eval()
    with torch.no_grad():
        got = model2(x)
    assert torch.equal(got, want)
    # (c) re-exporting the reloaded model reproduces the same integers (idempotence)
    ckpt2 = X.export_int8(model2)
    for pname, e in ckpt["weights"].items():
        assert torch.equal(e["code"], ckpt2["weights"][pname]["code"]), pname
    # (d) activation entries carry the reference's export convention
    for name, e in ckpt["activations"].items():
        s, zp = OE.act_export_params(float(e["min_range"]), float(e["max_range"]))
        assert e["scale"] == s and e["zero_point"] == zp, name
