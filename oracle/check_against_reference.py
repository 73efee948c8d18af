"""Pin the oracle against the UNMODIFIED reference, run on CPU where /root/reference exists.

    python oracle/check_against_reference.py [--full]

Builds the real ConvTasNetQ (small config by default, the full cfg-1 model with --full), runs the
reference recipe (2 observer passes, observers off, student fwd, float-teacher fwd, FQSS KD loss,
backward) and the oracle on the same state_dict and inputs, and compares output, loss, ranges and
every gradient.  Exit code 0 only if everything matches to fp32 round-off (the two programs run
the same ATen CPU kernels in the same order, so the expectation is bit-identity).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "tests", "golden"))
sys.path.insert(0, HERE)

import _ref_import as R  # noqa: E402
import fqss_oracle as O  # noqa: E402

QCFG = dict(qat=True, gradient_based=True, weight_quant=True, weight_n_bits=8, act_quant=True, act_n_bits=8,
            in_quant=False, in_act_n_bits=8, out_quant=True, out_act_n_bits=8, n_splitter=2, n_combiner=2,
            observer=True)
SMALL = dict(n_spks=2, kernel_size=16, stride=8, n_filters=64, bn_chan=32, hid_chan=64, n_blocks=3, n_repeats=2)
FULL = dict(n_spks=2, kernel_size=16, stride=8)


def ref_pit_loss_factory():
    """asteroid is absent: the recipe's PITLossWrapper is restated in the oracle (S3)."""
    from train_env.asteroid_librimix.wsdr import PairwiseWSDR
    neg = PairwiseWSDR("sisdr", take_log=True)
    kd = PairwiseWSDR("sisdr", take_log=False)      # = pairwise_wsisdr (wsdr.py:100)

    def loss_func(e, t):
        return O.pit_min_mean(-neg(e, t))[0]

    def kd_func(e, t, weights=None):
        return O.pit_min_mean(kd(e, t, weights))[0]

    return loss_func, kd_func


def reference_common_step(model, fmodel, inputs, targets, kd_lambda, loss_func, kd_func):
    """mysystem.py:124-146 with self.* replaced by arguments (Lightning is absent)."""
    est = model(inputs)
    with torch.no_grad():
        fest = fmodel(inputs).detach()
        a, b = [], []
        for i in range(len(fest)):
            a.append(loss_func(fest[i:i + 1], targets[i:i + 1]).detach())
            b.append(loss_func(est[i:i + 1], targets[i:i + 1]).detach())
        w = 10 ** ((torch.stack(a) - torch.stack(b)) / 10)
    kd = -kd_func(est, fest, weights=w)
    task = -kd_func(est, targets)
    loss = -10 * torch.log10((1 - kd_lambda) * task + kd_lambda * kd + 1e-8)
    return loss, est, fest


def main(full=False, B=2, T=4000, seed=0):
    kw = FULL if full else SMALL
    if full:
        T = 32000
    model, fmodel, LM = R.build_reference_model(kw, QCFG, seed)
    cfg = O.SeparatorConfig(n_src=2, kernel_size=16, stride=8, n_filters=kw.get("n_filters", 512),
                            bn_chan=kw.get("bn_chan", 128), hid_chan=kw.get("hid_chan", 512),
                            n_blocks=kw.get("n_blocks", 8), n_repeats=kw.get("n_repeats", 3))
    g = torch.Generator().manual_seed(seed + 1)
    src = torch.randn(B, 2, T, generator=g) * 0.05
    mix = src.sum(1, keepdim=True)

    P = O.Params({k: v.clone() for k, v in model.state_dict().items()})
    fP = O.Params({k: v.clone() for k, v in fmodel.state_dict().items()})

    # reference: two observer passes then observers off
    model.train()
    with torch.no_grad():
        model(mix); model(mix)
    LM.enable_observer(model, False)
    st = O.calibrate(P, mix, cfg, passes=2)
    worst = 0.0
    for k, v in model.state_dict().items():
        d = (v - P[k]).abs().max().item()
        worst = max(worst, d)
    print("ranges/params after calibration: max |ref-oracle| = %.3e" % worst)

    loss_func, kd_func = ref_pit_loss_factory()
    loss_r, est_r, fest_r = reference_common_step(model, fmodel, mix, src, 0.1, loss_func, kd_func)
    loss_r.backward()

    P.leafify()
    est_o = O.separator_forward(P, mix, cfg, st, quant=True)
    with torch.no_grad():
        fest_o = O.separator_forward(fP, mix, cfg, quant=False)
    loss_o, _ = O.fqss_kd_loss(est_o, fest_o, src, 0.1)
    loss_o.backward()

    ok = worst == 0.0
    print("est   max|d| %.3e" % (est_r - est_o).abs().max().item())
    print("fest  max|d| %.3e" % (fest_r - fest_o).abs().max().item())
    print("loss  ref %.6f oracle %.6f" % (loss_r.item(), loss_o.item()))
    ok &= torch.equal(est_r, est_o) and torch.equal(fest_r, fest_o)
    ok &= abs(loss_r.item() - loss_o.item()) <= 1e-5 * abs(loss_r.item())
    gmax, n_none = 0.0, 0
    for k, p in model.named_parameters():
        go = P[k].grad
        if p.grad is None:
            n_none += 1
            ok &= go is None or float(go.abs().max()) == 0.0
            continue
        den = p.grad.abs().max().item() + 1e-30
        gmax = max(gmax, (p.grad - go).abs().max().item() / den)
    print("grads: worst max-normalised diff %.3e ; params without grad in reference: %d" % (gmax, n_none))
    ok &= gmax <= 1e-4
    print("ORACLE == REFERENCE" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main(full="--full" in sys.argv))
