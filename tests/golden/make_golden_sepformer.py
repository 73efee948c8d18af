"""Golden fixture of SepformerQ (BASELINE configs[3]) from the UNMODIFIED reference on CPU.

    python tests/golden/make_golden_sepformer.py        ->  tests/golden/sepformer_small.npz

Recipe as make_golden_dptnet.py.  The model has 16 transformer layers with 1024-wide FFNs even in its smallest
configuration, so large tensors are stored as fingerprints instead of values: `fp(t)` = [sum, sum |t|, <t, r>] in float64
with r a seeded N(0,1) vector of t's size (seed = t.numel()).  The mirror is built from the same seed on the test side, so
matching initial fingerprints pin its module tree, key order and initialisation; gradients of large tensors are compared
through the same fingerprints, small ones (all quantiser ranges, norms, biases) element-wise.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import as R  # noqa: E402

KW = dict(n_spks=2, kernel_size=16, stride=8, n_filters=32, n_repeats=1, n_heads=4, chunk_size=10)
QCFG = dict(qat=True, gradient_based=True, weight_quant=True, weight_n_bits=8, act_quant=True, act_n_bits=8,
            in_quant=False, in_act_n_bits=8, out_quant=True, out_act_n_bits=8, n_splitter=2, n_combiner=2, observer=True)
SMALL = 4096


def fp(t):
    t = t.detach().double().flatten().cpu()
    r = torch.randn(t.numel(), generator=torch.Generator().manual_seed(t.numel()), dtype=torch.float64)
    return np.array([t.sum().item(), t.abs().sum().item(), (t * r).sum().item()])


def main():
    R.install()
    from quantization.qat.models.sepformerq import SepformerQ
    from quantization.qat.models import load_model as LM
    torch.manual_seed(0)
    model = LM.quantize_model(SepformerQ(**KW), dict(QCFG))
    d = {"keys": np.array(list(model.state_dict().keys()))}
    for k, v in model.state_dict().items():
        d["initfp/" + k] = fp(v)
    g = torch.Generator().manual_seed(1)
    src = torch.randn(2, 2, 808, generator=g) * 0.05
    mix = src.sum(1, keepdim=True)
    model.train()
    with torch.no_grad():
        model(mix); model(mix)
    LM.enable_observer(model, False)
    for k, v in model.state_dict().items():
        if k.endswith(("min_range", "max_range")):
            d["calib/" + k] = v.detach().numpy().copy()
    est = model(mix)
    loss = ((est - src[..., :est.shape[-1]]) ** 2).mean()
    loss.backward()
    d.update(mix=mix.numpy(), src=src.numpy(), est=est.detach().numpy(), loss=np.float32(loss.item()))
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        if p.numel() <= SMALL:
            d["grad/" + k] = p.grad.numpy().copy()
        else:
            d["gradfp/" + k] = fp(p.grad)
    import copy
    m64 = copy.deepcopy(model).double()
    with torch.no_grad():
        est64 = m64(mix.double()).float()
    q = "decoder.activation_fake_quantize."
    sd = model.state_dict()
    step = float(sd[q + "max_range"] - sd[q + "min_range"]) / 255
    dev = (est64 - est.detach()).abs() / step
    d.update(self_flip_rate=np.float32((dev > 0.5).float().mean().item()), self_max_steps=np.float32(dev.max().item()),
             self_est_rel=np.float32(((est64 - est.detach()).norm() / est.detach().norm()).item()))
    out = os.path.join(HERE, "sepformer_small.npz")
    np.savez_compressed(out, **d)
    print("wrote", out, "%.1f KB" % (os.path.getsize(out) / 1024), "keys", len(d["keys"]), "est", tuple(est.shape), "loss", loss.item())
    print("fp64 self-sensitivity: flip rate %.4f, max %.2f steps, rel %.3e" % (d["self_flip_rate"], d["self_max_steps"], d["self_est_rel"]))


if __name__ == "__main__":
    main()
