"""Generate the committed golden fixtures from the UNMODIFIED reference (CPU, build container).

    python tests/golden/make_golden.py

Outputs (all small, fp32, np.savez_compressed):
  tests/golden/q_ops.npz       Q1-Q3: act / weight fake-quant forward values, integer codes and
                               autograd gradients on seeded + adversarial inputs (qat_quant.py:88-147)
  tests/golden/split.npz       P1: splitter / reconstructor (process.py:10-52)
  tests/golden/loss.npz        S1/S2: PairwiseWSDR matrices (wsdr.py:46-95) + common_step loss/grad
  tests/golden/infer.npz       process.model_infer (process.py:154-194): whole-signal and chunked overlap-add inference
                               of a deterministic toy separator
  tests/golden/export.npz      X1: the export-time quantisers TorchActivationFakeQuantize / TorchWeightFakeQuantize
                               (qat_quant.py:15-53): scale, zero-point, outputs and gradients on seeded + boundary inputs
  tests/golden/music_loss.npz  the KD training loss of the music recipe (musdbhq_train.py:87-109) computed with the
                               reference's own calc_nsdr (process.py:70-75), center_trim and nn.L1Loss: loss, terms, gradient
  tests/golden/model_small.npz M1-M3/L1/L2: a reduced ConvTasNetQ (64 filters, 2x3 blocks): state_dict
                               before/after 2 observer passes, input, per-layer taps, output, teacher
                               output, FQSS loss and every parameter gradient

The reference cannot travel to the GPU box; these vectors can.  tests/test_oracle_golden.py pins
the oracle to them on CPU; the GPU tests pin the CUDA path to them and to the oracle.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import _ref_import as R  # noqa: E402

R.install()
from quantization.qat import qat_quant as RQ  # noqa: E402
import process as RP  # noqa: E402
from train_env.asteroid_librimix.wsdr import PairwiseWSDR  # noqa: E402
import check_against_reference as CK  # noqa: E402  (reference common_step restated without Lightning)


def _np(t):
    return t.detach().cpu().numpy()


def adversarial_act_input(rmin, rmax, n=4096, seed=0):
    """Seeded values plus points sitting exactly on rounding / clipping boundaries."""
    g = torch.Generator().manual_seed(seed)
    step = (rmax - rmin) / 255.0
    k = torch.arange(-3, 260, dtype=torch.float32)
    on_half = rmin + (k + 0.5) * step
    on_half_lo = torch.nextafter(on_half, torch.tensor(-1e30))
    on_half_hi = torch.nextafter(on_half, torch.tensor(1e30))
    on_int = rmin + k * step
    rnd = rmin + (rmax - rmin) * (torch.rand(n, generator=g) * 1.3 - 0.15)
    special = torch.tensor([0.0, -0.0, 1e-40, -1e-40, 1e3, -1e3, float(rmin), float(rmax)])
    return torch.cat([on_half, on_half_lo, on_half_hi, on_int, rnd, special])


def gen_q_ops(out):
    d = {}
    cases = [(-0.5, 0.5), (-1.7320508, 3.1415927), (0.0, 6.0), (0.25, 0.75), (-2.0, -0.5), (-3e-3, 2e-3)]
    for ci, (lo, hi) in enumerate(cases):
        rmin = torch.tensor([lo], requires_grad=True)
        rmax = torch.tensor([hi], requires_grad=True)
        x = adversarial_act_input(rmin.detach(), rmax.detach(), seed=ci).requires_grad_(True)
        y = RQ.linear_quantize(x, rmin, rmax, 8, True, False, False)
        g = torch.Generator().manual_seed(100 + ci)
        go = torch.randn(y.shape, generator=g)
        y.backward(go)
        step = (rmax - rmin) / 255
        code = torch.clip(torch.round((x - rmin) / step), 0, 255)
        d.update({f"act{ci}_x": _np(x), f"act{ci}_range": np.array([lo, hi], np.float32), f"act{ci}_y": _np(y),
                  f"act{ci}_code": _np(code).astype(np.uint8), f"act{ci}_go": _np(go), f"act{ci}_gx": _np(x.grad),
                  f"act{ci}_gmin": _np(rmin.grad), f"act{ci}_gmax": _np(rmax.grad)})
    # weights: [Co,Ci,k] per-out-channel (axis 0) and ConvTranspose style (axis 1, single channel)
    wcases = [((16, 8, 1), 0), ((8, 1, 3), 0), ((12, 2, 16), 0), ((24, 1, 16), 1)]
    for ci, (shape, axis) in enumerate(wcases):
        g = torch.Generator().manual_seed(200 + ci)
        w = (torch.randn(shape, generator=g) * 0.2).requires_grad_(True)
        q = RQ.GradientWeightFakeQuantize(True, shape, n_bits=8, ch_out_idx=axis)
        q(w)                       # observer call: captures amin/amax, returns w
        with torch.no_grad():      # perturb so that clipping and |min|>|max| / ties all occur
            q.max_range.mul_(0.8)
            if q.max_range.numel() > 2:
                q.min_range.view(-1)[0] = -q.max_range.view(-1)[0]          # exact tie
                q.min_range.view(-1)[1] = -2.0 * q.max_range.view(-1)[1]    # |min| dominates
        y = q(w)
        go = torch.randn(y.shape, generator=g)
        y.backward(go)
        bound = torch.maximum(q.min_range.abs(), q.max_range.abs())
        code = torch.clip(torch.round(w / (2 * bound / 255)), -128, 127)
        d.update({f"w{ci}_w": _np(w), f"w{ci}_axis": np.array(axis), f"w{ci}_min": _np(q.min_range),
                  f"w{ci}_max": _np(q.max_range), f"w{ci}_y": _np(y), f"w{ci}_code": _np(code).astype(np.int8),
                  f"w{ci}_go": _np(go), f"w{ci}_gw": _np(w.grad), f"w{ci}_gmin": _np(q.min_range.grad),
                  f"w{ci}_gmax": _np(q.max_range.grad)})
    # observer EMA (qat_quant.py:228-233)
    q = RQ.GradientActivationFakeQuantize(True)
    g = torch.Generator().manual_seed(300)
    xs = [torch.randn(3, 5, 7, generator=g) * (i + 1) for i in range(3)]
    tr = []
    for x in xs:
        y = q(x)
        assert torch.equal(y, x)
        tr.append([q.min_range.item(), q.max_range.item()])
    d["obs_x"] = np.stack([_np(x) for x in xs])
    d["obs_trace"] = np.array(tr, np.float32)
    np.savez_compressed(out, **d)
    print("wrote", out, len(d), "arrays")


def gen_split(out):
    g = torch.Generator().manual_seed(7)
    x = torch.randn(3, 1, 1000, generator=g) * 0.3
    x[0, 0, 0] = x.abs().max() * 1.5            # a clear peak
    y = RP.preprocess(x.clone(), n_splitter=2)
    y3 = RP.preprocess(x.clone(), n_splitter=3)
    dec = torch.randn(2, 3, 2, 1, 1000, generator=g)
    z = RP.postprocess(dec.clone(), n_combiner=2)
    np.savez_compressed(out, x=_np(x), y2=_np(y), y3=_np(y3), dec=_np(dec), z=_np(z))
    print("wrote", out)


def gen_loss(out):
    g = torch.Generator().manual_seed(11)
    B, S, T = 3, 2, 1600
    tgt = torch.randn(B, S, T, generator=g) * 0.05
    est = (tgt[:, [1, 0]] + 0.02 * torch.randn(B, S, T, generator=g)).requires_grad_(True)   # permuted
    fest = tgt[:, [1, 0]] + 0.01 * torch.randn(B, S, T, generator=g)
    w = torch.rand(B, generator=g) + 0.5
    pw_lin = PairwiseWSDR("sisdr", take_log=False)(est, tgt, w)
    pw_log = PairwiseWSDR("sisdr", take_log=True)(est, tgt)
    loss_func, kd_func = CK.ref_pit_loss_factory()

    class _M(torch.nn.Module):          # stand-ins so common_step's model(inputs) returns est / fest
        def __init__(s, t):
            super().__init__()
            s.t = t

        def forward(s, _):
            return s.t
    loss, _, _ = CK.reference_common_step(_M(est), _M(fest), None, tgt, 0.1, loss_func, kd_func)
    loss.backward()
    np.savez_compressed(out, tgt=_np(tgt), est=_np(est), fest=_np(fest), w=_np(w), pw_lin=_np(pw_lin),
                        pw_log=_np(pw_log), loss=_np(loss), gest=_np(est.grad))
    print("wrote", out, "loss", float(loss))


def gen_export(out):
    """Export-time quantisers of the unmodified reference (torch.fake_quantize_* on CPU)."""
    d = {}
    cases = [(-0.5, 0.5), (-1.7320508, 3.1415927), (0.0, 6.0), (0.25, 0.75), (-3e-3, 2e-3), (-7.3, 0.01)]
    for ci, (lo, hi) in enumerate(cases):
        q = RQ.GradientActivationFakeQuantize(True)
        with torch.no_grad():
            q.min_range.fill_(lo)
            q.max_range.fill_(hi)
            t = RQ.TorchActivationFakeQuantize(q)
        x = adversarial_act_input(torch.tensor([lo]), torch.tensor([hi]), seed=40 + ci)
        g = torch.Generator().manual_seed(400 + ci)
        # scale-grid boundaries of the EXPORT quantiser as well (its grid differs from the training one)
        k = torch.arange(-130, 130, dtype=torch.float32)
        x = torch.cat([x, (k + 0.5) * t.scale, torch.nextafter((k + 0.5) * t.scale, torch.tensor(1e30))]).requires_grad_(True)
        y = t(x)
        go = torch.randn(y.shape, generator=g)
        y.backward(go)
        d.update({f"act{ci}_range": np.array([lo, hi], np.float32), f"act{ci}_scale": np.array(t.scale, np.float64),
                  f"act{ci}_zp": np.array(t.zero_point), f"act{ci}_x": _np(x), f"act{ci}_y": _np(y), f"act{ci}_go": _np(go),
                  f"act{ci}_gx": _np(x.grad)})
    for ci, (shape, axis) in enumerate([((16, 8, 1), 0), ((8, 1, 3), 0), ((12, 2, 16), 0), ((24, 1, 16), 1)]):
        g = torch.Generator().manual_seed(500 + ci)
        w = (torch.randn(shape, generator=g) * 0.2).requires_grad_(True)
        q = RQ.GradientWeightFakeQuantize(True, shape, n_bits=8, ch_out_idx=axis)
        q(w)
        with torch.no_grad():
            q.max_range.mul_(0.8)
            t = RQ.TorchWeightFakeQuantize(q)
        y = t(w)
        go = torch.randn(y.shape, generator=g)
        y.backward(go)
        d.update({f"w{ci}_w": _np(w), f"w{ci}_axis": np.array(axis), f"w{ci}_min": _np(q.min_range), f"w{ci}_max": _np(q.max_range),
                  f"w{ci}_scales": _np(t.scales), f"w{ci}_y": _np(y), f"w{ci}_go": _np(go), f"w{ci}_gw": _np(w.grad)})
    # error behaviour: a range that excludes zero on the negative side cannot be exported (zero_point lands beyond quant_max)
    q = RQ.GradientActivationFakeQuantize(True)
    with torch.no_grad():
        q.min_range.fill_(-2.0)
        q.max_range.fill_(-0.5)
        t = RQ.TorchActivationFakeQuantize(q)
    try:
        t(torch.zeros(4))
        msg = ""
    except RuntimeError as e:
        msg = str(e)
    d["neg_range_error"] = np.array(msg)
    d["neg_range_zp"] = np.array(t.zero_point)
    np.savez_compressed(out, **d)
    print("wrote", out, len(d), "arrays; negative-range error:", msg)


def gen_music_loss(out):
    """The loss lines of the reference's training loop (musdbhq_train.py:83-109) on seeded tensors.  The loop is not a
    function in the reference, so its few lines are composed here from the reference's OWN helpers: process.calc_nsdr,
    musdbhq_utils.center_trim and nn.L1Loss (musdbhq_train.py:249)."""
    sys.path.insert(0, os.path.join(R.REFERENCE_ROOT, "train_env", "tasnet_musdbhq"))
    from musdbhq_utils import center_trim
    g = torch.Generator().manual_seed(21)
    B, S, C, Tw, Ts = 3, 4, 2, 1210, 1237
    sources_full = torch.randn(B, S, C, Ts, generator=g) * 0.1
    trimmed = center_trim(sources_full, Tw)
    wavs = (trimmed + 0.03 * torch.randn(B, S, C, Tw, generator=g)).clone().requires_grad_(True)
    fwavs = trimmed + 0.01 * torch.randn(B, S, C, Tw, generator=g) * torch.tensor([0.5, 1.0, 3.0]).view(B, 1, 1, 1)
    wavs.data[0, 0, 0, :5] = trimmed[0, 0, 0, :5]                   # exact zeros of w - s: sign(0) = 0
    loss_fn = torch.nn.L1Loss()
    kd_lambda = 0.1
    sources = center_trim(sources_full, wavs)
    sdrs, sdrqs = [], []
    with torch.no_grad():
        for i in range(len(fwavs)):
            sdrs.append(RP.calc_nsdr(fwavs[i:i + 1], sources[i:i + 1]))
            sdrqs.append(RP.calc_nsdr(wavs[i:i + 1], sources[i:i + 1]))
        w = 10 ** ((torch.Tensor(sdrs) - torch.Tensor(sdrqs)) / 10)
    kd_loss = torch.mean(w * torch.stack([loss_fn(wavs[i:i + 1], fwavs[i:i + 1]) for i in range(len(fwavs))], dim=0))
    task_loss = loss_fn(wavs, sources)
    loss = (1 - kd_lambda) * task_loss + kd_lambda * kd_loss
    loss.backward()
    g_kd = wavs.grad.clone()
    wavs.grad = None
    loss0 = loss_fn(wavs, sources)
    loss0.backward()
    np.savez_compressed(out, sources_full=_np(sources_full), wavs=_np(wavs), fwavs=_np(fwavs), kd_lambda=np.float32(kd_lambda),
                        w=_np(w), loss=_np(loss), kd=_np(kd_loss), task=_np(task_loss), g=_np(g_kd), loss0=_np(loss0),
                        g0=_np(wavs.grad))
    print("wrote", out, "loss", float(loss), "weights", w.tolist())


def gen_model_small(out):
    import fqss_oracle as O
    model, fmodel, LM = R.build_reference_model(CK.SMALL, CK.QCFG, seed=0)
    B, T = 2, 2400
    g = torch.Generator().manual_seed(1)
    src = torch.randn(B, 2, T, generator=g) * 0.05
    mix = src.sum(1, keepdim=True)
    d = {"mix": _np(mix), "src": _np(src)}
    for k, v in model.state_dict().items():
        d["init/" + k] = _np(v)
    for k, v in fmodel.state_dict().items():
        d["teacher/" + k] = _np(v)
    model.train()
    with torch.no_grad():
        model(mix); model(mix)
    LM.enable_observer(model, False)
    for k, v in model.state_dict().items():
        d["calib/" + k] = _np(v)
    # taps via forward hooks on the reference modules
    taps = {}

    def hook(name):
        def f(m, i, o):
            taps[name] = o.detach().clone()
        return f
    hs = [model.encoder.register_forward_hook(hook("encoder")),
          model.masker.bottleneck.register_forward_hook(hook("masker.bottleneck")),
          model.masker.register_forward_hook(hook("mask")),
          model.mul.register_forward_hook(hook("masked")),
          model.decoder.register_forward_hook(hook("decoder"))]
    for i, blk in enumerate(model.masker.TCN):
        hs.append(blk.register_forward_hook(lambda m, inp, o, i=i: taps.__setitem__("masker.TCN.%d.out" % i, o[0].detach().clone())))
        hs.append(blk.shared_block.register_forward_hook(hook("masker.TCN.%d.hidden" % i)))
    loss_func, kd_func = CK.ref_pit_loss_factory()
    loss, est, fest = CK.reference_common_step(model, fmodel, mix, src, 0.1, loss_func, kd_func)
    loss.backward()
    for h in hs:
        h.remove()
    for k, v in taps.items():
        d["tap/" + k] = _np(v)
    d["est"], d["fest"], d["loss"] = _np(est), _np(fest), _np(loss)
    for k, p in model.named_parameters():
        if p.grad is not None:
            d["grad/" + k] = _np(p.grad)
    np.savez_compressed(out, **d)
    print("wrote", out, "loss", float(loss), "arrays", len(d), "bytes", os.path.getsize(out))


class _ToySeparator(torch.nn.Module):
    """Deterministic stand-in for a separator (model_infer only needs `model(x) -> [B, S, T]`): non-linear, with a
    per-call peak normalisation like the FQSS splitter, so that chunking / padding / batching mistakes show up."""
    n_srcs = 2

    def forward(self, x):                      # x: [1, C, T] -> [1, 2, T - 3] (shorter than the input, like the codec)
        x = x[:, 0, :]
        x = x / x.abs().max().clamp_min(1e-8)
        a = torch.tanh(2.0 * x)[:, :-3]
        b = (x * x.abs())[:, 3:] * 0.5
        return torch.stack([a, b], dim=1)


def gen_infer(out):
    g = torch.Generator().manual_seed(5)
    mix = torch.randn(1, 5000, generator=g) * 0.3
    model = _ToySeparator()
    full = RP.model_infer(model, mix[0:1], device="cpu")
    ola = RP.model_infer(model, mix, segment=1600, overlap=0.25, device="cpu")
    ola2 = RP.model_infer(model, mix, segment=999, overlap=0.5, device="cpu")
    np.savez_compressed(out, mix=_np(mix), full=_np(full), ola=_np(ola), ola2=_np(ola2))
    print("wrote", out, "bytes", os.path.getsize(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "music_loss":
        gen_music_loss(os.path.join(HERE, "music_loss.npz"))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "export":
        gen_export(os.path.join(HERE, "export.npz"))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "infer":
        gen_infer(os.path.join(HERE, "infer.npz"))
        sys.exit(0)
    torch.set_num_threads(8)
    gen_q_ops(os.path.join(HERE, "q_ops.npz"))
    gen_split(os.path.join(HERE, "split.npz"))
    gen_loss(os.path.join(HERE, "loss.npz"))
    gen_model_small(os.path.join(HERE, "model_small.npz"))
    gen_infer(os.path.join(HERE, "infer.npz"))
    gen_export(os.path.join(HERE, "export.npz"))
    gen_music_loss(os.path.join(HERE, "music_loss.npz"))
