"""GPU: DPTNetQ (BASELINE configs[2]) on the sm_100a quantiser kernels against golden vectors produced
by the UNMODIFIED reference (tests/golden/make_golden_dptnet.py): calibration (two observer passes), forward with the
reference's calibrated ranges, and every parameter gradient.

Tolerances.  The quantisers are bit-exact on identical inputs (tests/test_gpu_ops.py); what differs from the reference's CPU
run is the fp32 summation order of the dense ops between them (cuBLAS / cuDNN vs ATen CPU), i.e. ~1e-6 relative on a
quantiser's input, which moves a code by one step wherever the input sits on a rounding boundary.  The bounds below are the
same kind as for the ConvTasNet model tests: calibrated ranges to 1e-4 of the range, output to 2e-2 relative (one step of
the 8-bit output quantiser is 4e-3 of its range, i.e. a few percent of this low-level signal: the output is judged in
units of that step -- rate of samples that moved and the largest move -- against the reference's OWN sensitivity to
rounding-level perturbations, its float64 evaluation of the same model, stored with the fixture: 13.8 % of the output
samples move, by up to 2 steps), loss to 2e-3, gradients by cosine against the reference's."""
import numpy as np
import pytest
import torch

from parity_log import record
from test_dptnet_cpu import build

pytestmark = pytest.mark.gpu
DEV = "cuda"


def T(a):
    return torch.from_numpy(np.asarray(a))


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _load(model, g, prefix):
    model.load_state_dict({str(k): T(g[prefix + str(k)]) for k in g["keys"]}, strict=True)


def test_dptnet_calibration_matches_reference(golden):
    from fqss_b200.qat.models.load_model import enable_observer
    g = golden("dptnet_small.npz")
    model = build().to(DEV)
    _load(model, g, "init/")
    mix = T(g["mix"]).to(DEV)
    model.train()
    with torch.no_grad():
        model(mix); model(mix)
    enable_observer(model, False)
    sd = model.state_dict()
    worst, where = 0.0, None
    for k in g["keys"]:
        k = str(k)
        if not k.endswith(("min_range", "max_range")):
            assert torch.equal(sd[k].cpu(), T(g["calib/" + k])), k          # weights untouched by calibration
            continue
        other = k[:-9] + ("max_range" if k.endswith("min_range") else "min_range")
        span = np.abs(g["calib/" + k] - g["calib/" + other]).max() + 1e-12
        d = float(np.abs(sd[k].cpu().numpy() - g["calib/" + k]).max() / span)
        if d > worst:
            worst, where = d, k
    record("dptnet/calibration", worst_range_dev=worst)
    assert worst < 1e-4, (worst, where)


def test_dptnet_forward_backward_vs_reference(golden):
    from fqss_b200.qat.models.load_model import enable_observer
    g = golden("dptnet_small.npz")
    model = build().to(DEV)
    _load(model, g, "calib/")
    enable_observer(model, False)
    for m in model.modules():
        if hasattr(m, "observer_mode"):
            m.observer_mode = False
    mix, src = T(g["mix"]).to(DEV), T(g["src"]).to(DEV)
    model.train()
    est = model(mix)
    assert est.shape == tuple(g["est"].shape)
    loss = ((est - src[..., :est.shape[-1]]) ** 2).mean()
    loss.backward()
    meas = dict(est_rel=rel(est, T(g["est"])), loss_rel=abs(loss.item() - float(g["loss"])) / float(g["loss"]))
    q = "decoder.basis_signals.activation_fake_quantize."
    step = float(g["calib/" + q + "max_range"][0] - g["calib/" + q + "min_range"][0]) / 255
    d = (est.detach().cpu() - T(g["est"])).abs() / step
    meas["out_flip_rate"], meas["out_max_steps"] = (d > 0.5).float().mean().item(), d.max().item()
    num = n1 = n2 = 0.0
    missing = []
    for k, p in model.named_parameters():
        key = "grad/" + k
        if key not in g.files:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        if p.grad is None:
            missing.append(k)
            continue
        a, b = p.grad.detach().double().cpu(), T(g[key]).double()
        num += (a * b).sum().item(); n1 += a.pow(2).sum().item(); n2 += b.pow(2).sum().item()
    assert not missing, missing[:5]
    meas["grad_cos"] = num / ((n1 * n2) ** 0.5 + 1e-30)
    meas["grad_norm_ratio"] = (n1 / (n2 + 1e-30)) ** 0.5
    record("dptnet/forward_backward", **meas)
    # yardstick: the reference's OWN response to rounding-level perturbations (same model evaluated in float64, stored with
    # the fixture): the CUDA path must stay within it
    ref_flip, ref_max, ref_rel = float(g["self_flip_rate"]), float(g["self_max_steps"]), float(g["self_est_rel"])
    meas.update(ref_self_flip_rate=ref_flip, ref_self_max_steps=ref_max, ref_self_est_rel=ref_rel)
    record("dptnet/forward_backward", **meas)
    assert meas["out_flip_rate"] <= 1.25 * ref_flip and meas["out_max_steps"] <= ref_max + 1.01, meas
    assert meas["est_rel"] <= 1.25 * ref_rel and meas["loss_rel"] < 2e-3, meas
    assert meas["grad_cos"] > 0.98 and 0.9 < meas["grad_norm_ratio"] < 1.1, meas
