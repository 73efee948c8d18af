"""Golden fixture of the next scope row: a reduced DPTNetQ run by the UNMODIFIED reference on CPU.

    python tests/golden/make_golden_dptnet.py        ->  tests/golden/dptnet_small.npz

Recipe (train_env/train_utils.py:8-27 + the quantisation block of configs/dptnet_2spks_8k.yaml): seed, create, quantise
(W8A8, splitter / combiner 2 / 2), two observer passes, observers off, forward, backward of mean((est - src)^2).
Stored: the state dict right after quantisation (`init/...`, in state-dict order: pins the mirror's module tree, key order
and seeded initialisation), after calibration (`calib/...`), the input, the estimate and every parameter gradient.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import as R  # noqa: E402

KW = dict(n_spks=2, kernel_size=2, enc_dim=64, feature_dim=32, hidden_dim=32, layer=1, segment_size=20)
QCFG = dict(qat=True, gradient_based=True, weight_quant=True, weight_n_bits=8, act_quant=True, act_n_bits=8,
            in_quant=False, in_act_n_bits=8, out_quant=True, out_act_n_bits=8, n_splitter=2, n_combiner=2, observer=True)


def main():
    R.install()
    from quantization.qat.models.dptnetq import DPTNetQ
    from quantization.qat.models import load_model as LM
    torch.manual_seed(0)
    model = LM.quantize_model(DPTNetQ(**KW), dict(QCFG))
    d = {"keys": np.array(list(model.state_dict().keys()))}
    for k, v in model.state_dict().items():
        d["init/" + k] = v.detach().numpy().copy()
    g = torch.Generator().manual_seed(1)
    src = torch.randn(2, 2, 401, generator=g) * 0.05
    mix = src.sum(1, keepdim=True)
    model.train()
    with torch.no_grad():
        model(mix); model(mix)
    LM.enable_observer(model, False)
    for k, v in model.state_dict().items():
        d["calib/" + k] = v.detach().numpy().copy()
    est = model(mix)
    loss = ((est - src[..., :est.shape[-1]]) ** 2).mean()
    loss.backward()
    d.update(mix=mix.numpy(), src=src.numpy(), est=est.detach().numpy(), loss=np.float32(loss.item()))
    # yardstick for the GPU test: the reference's own sensitivity to rounding-level perturbations -- the same model and
    # ranges evaluated in float64 (every dense op and every quantiser decision made in double instead of float)
    import copy
    m64 = copy.deepcopy(model).double()
    with torch.no_grad():
        est64 = m64(mix.double()).float()
    q = "decoder.basis_signals.activation_fake_quantize."
    sd = model.state_dict()
    step = float(sd[q + "max_range"] - sd[q + "min_range"]) / 255
    dev = (est64 - est.detach()).abs() / step
    d.update(self_flip_rate=np.float32((dev > 0.5).float().mean().item()), self_max_steps=np.float32(dev.max().item()),
             self_est_rel=np.float32(((est64 - est.detach()).norm() / est.detach().norm()).item()))
    print("fp64 self-sensitivity: flip rate %.4f, max %.2f steps, rel %.3e" % (d["self_flip_rate"], d["self_max_steps"], d["self_est_rel"]))
    for k, p in model.named_parameters():
        if p.grad is not None:
            d["grad/" + k] = p.grad.numpy().copy()
    out = os.path.join(HERE, "dptnet_small.npz")
    np.savez_compressed(out, **d)
    print("wrote", out, "%.1f KB" % (os.path.getsize(out) / 1024), "est", tuple(est.shape), "loss", loss.item(),
          "grads", sum(1 for k in d if k.startswith("grad/")))


if __name__ == "__main__":
    main()
