"""GPU: SepformerQ (BASELINE configs[3]) on the sm_100a quantiser kernels against golden vectors of the
UNMODIFIED reference (tests/golden/make_golden_sepformer.py); same contract and yardstick as tests/test_gpu_dptnet.py (the
reference's own float64 evaluation bounds the output deviation; large gradient tensors are compared through the stored
fingerprints, relative to |g|_1-scaled tolerances)."""
import numpy as np
import pytest
import torch

from parity_log import record
from test_sepformer_cpu import build, fp

pytestmark = pytest.mark.gpu
DEV = "cuda"


def T(a):
    return torch.from_numpy(np.asarray(a))


def _calibrated(g):
    from fqss_b200.qat.models.load_model import enable_observer
    model = build().to(DEV)
    sd = model.state_dict()
    for k in g["keys"]:
        k = str(k)
        if k.endswith(("min_range", "max_range")):
            sd[k] = T(g["calib/" + k]).to(DEV)
    model.load_state_dict(sd, strict=True)
    enable_observer(model, False)
    for m in model.modules():
        if hasattr(m, "observer_mode"):
            m.observer_mode = False
    return model


def test_sepformer_calibration_matches_reference(golden):
    from fqss_b200.qat.models.load_model import enable_observer
    g = golden("sepformer_small.npz")
    model = build().to(DEV)
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
            continue
        other = k[:-9] + ("max_range" if k.endswith("min_range") else "min_range")
        span = np.abs(g["calib/" + k] - g["calib/" + other]).max() + 1e-12
        d = float(np.abs(sd[k].cpu().numpy() - g["calib/" + k]).max() / span)
        if d > worst:
            worst, where = d, k
    record("sepformer/calibration", worst_range_dev=worst)
    assert worst < 1e-4, (worst, where)


def test_sepformer_forward_backward_vs_reference(golden):
    g = golden("sepformer_small.npz")
    model = _calibrated(g)
    mix, src = T(g["mix"]).to(DEV), T(g["src"]).to(DEV)
    model.train()
    est = model(mix)
    assert est.shape == tuple(g["est"].shape)
    loss = ((est - src[..., :est.shape[-1]]) ** 2).mean()
    loss.backward()
    ref = T(g["est"])
    meas = dict(est_rel=((est.detach().cpu() - ref).norm() / ref.norm()).item(),
                loss_rel=abs(loss.item() - float(g["loss"])) / float(g["loss"]))
    q = "decoder.activation_fake_quantize."
    step = float(g["calib/" + q + "max_range"][0] - g["calib/" + q + "min_range"][0]) / 255
    d = (est.detach().cpu() - ref).abs() / step
    meas["out_flip_rate"], meas["out_max_steps"] = (d > 0.5).float().mean().item(), d.max().item()
    # gradients: small tensors element-wise (pooled cosine), large ones through their fingerprints
    num = n1 = n2 = 0.0
    fp_dev = 0.0
    for k, p in model.named_parameters():
        if "grad/" + k in g.files:
            a, b = p.grad.detach().double().cpu(), T(g["grad/" + k]).double()
            num += (a * b).sum().item(); n1 += a.pow(2).sum().item(); n2 += b.pow(2).sum().item()
        elif "gradfp/" + k in g.files:
            mine, theirs = fp(p.grad), g["gradfp/" + k]
            # |sum| and |<g, r>| deviations relative to |g|_1 (<= 1 for identical gradients up to noise ~ 1/sqrt(n))
            fp_dev = max(fp_dev, abs(mine[0] - theirs[0]) / (theirs[1] + 1e-30), abs(mine[1] - theirs[1]) / (theirs[1] + 1e-30),
                         abs(mine[2] - theirs[2]) / (theirs[1] + 1e-30))
        else:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
    meas["grad_cos_small"] = num / ((n1 * n2) ** 0.5 + 1e-30)
    meas["grad_norm_ratio_small"] = (n1 / (n2 + 1e-30)) ** 0.5
    meas["grad_fp_dev_large"] = fp_dev
    ref_flip, ref_max, ref_rel = float(g["self_flip_rate"]), float(g["self_max_steps"]), float(g["self_est_rel"])
    meas.update(ref_self_flip_rate=ref_flip, ref_self_max_steps=ref_max, ref_self_est_rel=ref_rel)
    record("sepformer/forward_backward", **meas)
    assert meas["out_flip_rate"] <= 1.25 * ref_flip and meas["out_max_steps"] <= ref_max + 1.01, meas
    assert meas["est_rel"] <= 1.25 * ref_rel and meas["loss_rel"] < 2e-3, meas
    assert meas["grad_cos_small"] > 0.98 and 0.9 < meas["grad_norm_ratio_small"] < 1.1 and fp_dev < 5e-2, meas
