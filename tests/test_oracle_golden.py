"""CPU: pin the oracle (oracle/fqss_oracle.py) to golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py).  Same ATen CPU kernels in the same order => bit-exact."""
import itertools

import numpy as np
import torch

import fqss_oracle as O


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_act_fake_quant_values_codes_grads(golden):
    g = golden("q_ops.npz")
    for ci in range(6):
        x = T(g[f"act{ci}_x"]).requires_grad_(True)
        lo, hi = g[f"act{ci}_range"]
        rmin = torch.tensor([lo], requires_grad=True)
        rmax = torch.tensor([hi], requires_grad=True)
        y = O.fq_act(x, rmin, rmax)
        assert torch.equal(y, T(g[f"act{ci}_y"]))
        assert torch.equal(O.act_codes(x, rmin, rmax).to(torch.uint8), T(g[f"act{ci}_code"]))
        y.backward(T(g[f"act{ci}_go"]))
        assert torch.equal(x.grad, T(g[f"act{ci}_gx"]))
        assert torch.equal(rmin.grad, T(g[f"act{ci}_gmin"]))
        assert torch.equal(rmax.grad, T(g[f"act{ci}_gmax"]))


def test_weight_fake_quant_values_codes_grads(golden):
    g = golden("q_ops.npz")
    for ci in range(4):
        w = T(g[f"w{ci}_w"]).requires_grad_(True)
        rmin = T(g[f"w{ci}_min"]).requires_grad_(True)
        rmax = T(g[f"w{ci}_max"]).requires_grad_(True)
        y = O.fq_weight(w, rmin, rmax)
        assert torch.equal(y, T(g[f"w{ci}_y"]))
        assert torch.equal(O.weight_codes(w, rmin, rmax).to(torch.int8), T(g[f"w{ci}_code"]))
        y.backward(T(g[f"w{ci}_go"]))
        assert torch.equal(w.grad, T(g[f"w{ci}_gw"]))
        assert torch.equal(rmin.grad, T(g[f"w{ci}_gmin"]))
        assert torch.equal(rmax.grad, T(g[f"w{ci}_gmax"]))


def test_observer_ema(golden):
    g = golden("q_ops.npz")
    rmin, rmax = torch.tensor([-0.5]), torch.tensor([0.5])
    for x, (emin, emax) in zip(T(g["obs_x"]), g["obs_trace"]):
        rmin, rmax = O.observe_act(rmin, rmax, x)
        assert rmin.item() == emin and rmax.item() == emax


def test_splitter_reconstructor(golden):
    g = golden("split.npz")
    x = T(g["x"])
    assert torch.equal(O.split_input(x.clone(), 2), T(g["y2"]))
    assert torch.equal(O.split_input(x.clone(), 3), T(g["y3"]))
    assert torch.equal(O.combine_output(T(g["dec"]).clone(), 2), T(g["z"]))
    # codes are integers in [-128,127] over 1/128; msb + (lsb+1)*(0.5/128) re-assembles x/peak to 16 bits
    y2 = T(g["y2"])
    assert torch.equal(y2 * 128, torch.round(y2 * 128)) and y2.min() >= -1 and y2.max() <= 127 / 128
    rec = y2[:, 0] + (y2[:, 1] + 1.0) * (0.5 / 128)
    err = (rec - x[:, 0] / x.abs().max())[y2[:, 0] < 127 / 128]       # (the +peak sample is clipped)
    assert err.abs().max() < 1.0 / 128 / 256 + 1e-6


def test_pairwise_and_kd_loss(golden):
    g = golden("loss.npz")
    tgt, fest, w = T(g["tgt"]), T(g["fest"]), T(g["w"])
    est = T(g["est"]).requires_grad_(True)
    assert torch.equal(-O.pairwise_sisdr_ratio(est, tgt, w), T(g["pw_lin"]))
    assert torch.equal(10 * torch.log10(O.pairwise_sisdr_ratio(est, tgt) + 1e-8), T(g["pw_log"]))
    loss, _ = O.fqss_kd_loss(est, fest, tgt, 0.1)
    loss.backward()
    assert loss.item() == float(g["loss"])
    assert torch.equal(est.grad, T(g["gest"]))


def test_pit_against_bruteforce():
    """asteroid 0.6 PIT (absent third-party) restated: pin against an explicit per-sample loop."""
    gen = torch.Generator().manual_seed(3)
    for S in (2, 3):
        pw = torch.randn(5, S, S, generator=gen)
        mean, per_b = O.pit_min_mean(pw)
        for b in range(5):
            best = min(sum(pw[b, p[j], j] for j in range(S)) / S for p in itertools.permutations(range(S)))
            assert abs(per_b[b] - best) < 1e-6
        assert abs(mean - per_b.mean()) < 1e-7


def _small_cfg():
    return O.SeparatorConfig(n_filters=64, bn_chan=32, hid_chan=64, n_blocks=3, n_repeats=2)


def test_small_model_calibration_forward_backward(golden):
    g = golden("model_small.npz")
    cfg = _small_cfg()
    P = O.Params({k[5:]: T(g[k]).clone() for k in g.files if k.startswith("init/")})
    fP = O.Params({k[8:]: T(g[k]).clone() for k in g.files if k.startswith("teacher/")})
    mix, src = T(g["mix"]), T(g["src"])
    st = O.calibrate(P, mix, cfg, passes=2)
    for k in P:
        assert torch.equal(P[k], T(g["calib/" + k])), k
    P.leafify()
    taps = {}
    est = O.separator_forward(P, mix, cfg, st, quant=True, tap=taps)
    with torch.no_grad():
        fest = O.separator_forward(fP, mix, cfg, quant=False)
    assert torch.equal(est, T(g["est"])) and torch.equal(fest, T(g["fest"]))
    for k in g.files:
        if k.startswith("tap/"):
            assert torch.equal(taps[k[4:]], T(g[k])), k
    loss, _ = O.fqss_kd_loss(est, fest, src, 0.1)
    loss.backward()
    assert loss.item() == float(g["loss"])
    n = 0
    for k in g.files:
        if k.startswith("grad/"):
            assert torch.equal(P[k[5:]].grad, T(g[k])), k
            n += 1
    assert n > 200
    # block 5 (last) res_conv / add never reach the output (dead in the reference too)
    assert P["masker.TCN.5.res_conv.conv1d.weight"].grad is None


def test_ddp_mean_semantics():
    a = {"w": torch.ones(3), "dead": None}
    b = {"w": 3 * torch.ones(3), "dead": None}
    m = O.ddp_mean_grads([a, b])
    assert torch.equal(m["w"], 2 * torch.ones(3)) and m["dead"] is None
