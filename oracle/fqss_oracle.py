"""CPU ORACLE for the FQSS ConvTasNet QAT hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, from the reference's algorithm, the arithmetic of the fake-quantized
ConvTasNet separator (forward + autograd backward), the FQSS input splitter / output
reconstructor, the observer calibration, and the knowledge-distillation SI-SDR loss.  It is a
plain PyTorch-CPU, *functional* program over a flat parameter dict that uses the reference's
state_dict key names; it shares no code with the reference and never touches CUDA.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it, and only as the checker / reported CPU baseline.  The product
(`fqss_b200`) must never import it.

Parity pin: the reference ships no tests or golden vectors.  The oracle is
pinned against outputs of the UNMODIFIED reference run in the build container
(`tests/golden/make_golden.py` -> `tests/golden/*.npz`, checked by `tests/test_oracle_golden.py`)
and, where /root/reference is present, directly by `oracle/check_against_reference.py`.
The PIT search of asteroid 0.6 (third-party, absent) is restated from its published algorithm
and pinned against a brute-force permutation loop: "parity unpinned" for that one function.

Reference citations are file:line under the reference tree.
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch
import torch.nn.functional as F

EPS_GLN = 1e-8          # convtasnetq.py:8 (EPS used for every GroupNorm)
EPS_SDR = 1e-8          # wsdr.py:48 / mysystem.py:14


# --------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------
@dataclass
class SeparatorConfig:
    """Shape hyper-parameters of ConvTasNetQ (convtasnetq.py:124-137) + quantisation block."""
    n_src: int = 2
    kernel_size: int = 16
    stride: int = 8
    n_filters: int = 512
    bn_chan: int = 128
    hid_chan: int = 512
    n_blocks: int = 8
    n_repeats: int = 3
    mask_kernel: int = 3
    n_splitter: int = 2
    n_combiner: int = 2
    act_bits: int = 8
    weight_bits: int = 8
    out_bits: int = 8
    alpha: float = 0.9              # qat_quant.py:217 observer EMA
    max_observations: int = 50      # qat_quant.py:215

    @property
    def n_tcn(self) -> int:
        return self.n_blocks * self.n_repeats

    def dilation(self, i: int) -> int:
        return 2 ** (i % self.n_blocks)     # convtasnetq.py:71-80


# --------------------------------------------------------------------------------------
# Q1/Q2/Q3: quantisers (qat_quant.py:88-147)
# --------------------------------------------------------------------------------------
def _ste_rint(v: torch.Tensor) -> torch.Tensor:
    # value rint(v) (half-to-even), gradient identity: qat_quant.py:88-89
    return v + (torch.round(v) - v).detach()


def act_codes(x, rmin, rmax, n_bits=8):
    """Integer codes of the activation quantiser, clip(rint((x-min)/delta), 0, 2^b-1)."""
    levels = 2 ** n_bits - 1
    step = (rmax - rmin) / levels
    return torch.clip(torch.round((x - rmin) / step), 0, levels)


def fq_act(x, rmin, rmax, n_bits=8):
    """Asymmetric per-tensor fake-quant, qat_quant.py:136-147 (sym=False, scale_grad=False)."""
    levels = 2 ** n_bits - 1
    step = (rmax - rmin) / levels
    q = _ste_rint((x - rmin) / step)
    return step * torch.clip(q, 0, levels) + rmin


def weight_codes(w, rmin, rmax, n_bits=8):
    levels = 2 ** n_bits - 1
    bound = torch.maximum(rmin.abs(), rmax.abs())
    step = 2 * bound / levels
    return torch.clip(torch.round(w / step), -(2 ** (n_bits - 1)), 2 ** (n_bits - 1) - 1)


def fq_weight(w, rmin, rmax, n_bits=8):
    """Symmetric signed per-channel fake-quant, qat_quant.py:126-135 (sym=True, sign=True)."""
    levels = 2 ** n_bits - 1
    lo, hi = -(2 ** (n_bits - 1)), 2 ** (n_bits - 1) - 1
    bound = torch.maximum(rmin.abs(), rmax.abs())
    step = 2 * bound / levels
    q = _ste_rint(w / step)
    return step * torch.clip(q, lo, hi)


def observe_act(rmin, rmax, x, alpha=0.9):
    """EMA range update, qat_quant.py:228-232.  Returns the new (min,max) tensors."""
    return alpha * rmin + (1 - alpha) * x.min(), alpha * rmax + (1 - alpha) * x.max()


def observe_weight(w, ch_axis):
    """First-call range capture, qat_quant.py:373-375."""
    dims = [d for d in range(w.dim()) if d != ch_axis]
    return torch.amin(w, dim=dims, keepdim=True), torch.amax(w, dim=dims, keepdim=True)


# --------------------------------------------------------------------------------------
# P1: FQSS splitter / reconstructor (process.py:10-52)
# --------------------------------------------------------------------------------------
def floor_quant(x, threshold=1.0, n_bits=8):
    step = threshold / (2 ** (n_bits - 1))
    return torch.clip(torch.floor(x / step), -(2 ** (n_bits - 1)), 2 ** (n_bits - 1) - 1) * step


def split_input(x, n_splitter, n_bits=8):
    if x.dim() == 2:
        x = x.unsqueeze(1)
    if n_splitter <= 1:
        return x
    peak = max(abs(x.min()), abs(x.max()))      # ONE scalar over the whole batch (process.py:23)
    x = x / peak
    step = 1.0 / (2 ** (n_bits - 1))
    parts = []
    for _ in range(n_splitter):
        q = floor_quant(x, 1.0, n_bits)
        parts.append(q)
        x = 2 * (x - q) * 1.0 / step - 1.0
    return torch.cat(parts, dim=1)


def combine_output(out, n_combiner, n_bits=8):
    # out: [n_combiner, B, S, 1, T]  (process.py:39-52)
    if n_combiner == 1:
        y = out.squeeze(0)
    else:
        step = 1.0 / (2 ** (n_bits - 1))
        y = out[0]
        for i in range(1, n_combiner):
            y = y + out[i] * (0.5 * step) ** i
    if y.dim() <= 4 and y.shape[-2] == 1:
        y = y.squeeze(-2)
    return y


# --------------------------------------------------------------------------------------
# parameter dict helpers
# --------------------------------------------------------------------------------------
class Params(dict):
    """Flat {state_dict key: tensor}.  `leafify` turns entries into autograd leaves."""

    def leafify(self, requires_grad=True):
        for k in list(self.keys()):
            self[k] = self[k].detach().clone().requires_grad_(requires_grad)
        return self

    def grads(self) -> Dict[str, Optional[torch.Tensor]]:
        return {k: v.grad for k, v in self.items()}


@dataclass
class QuantState:
    """Non-state_dict quantiser state (qat_quant.py:214-220): observer switch and call counters."""
    observe: bool = False
    weights_seen: bool = False
    n_iter: Dict[str, int] = field(default_factory=dict)


class _Ctx:
    def __init__(self, P: Params, cfg: SeparatorConfig, st: QuantState, quant: bool, tap):
        self.P, self.cfg, self.st, self.quant, self.tap = P, cfg, st, quant, tap

    def rec(self, name, t):
        if self.tap is not None:
            self.tap[name] = t
        return t

    # activation quantiser named `<prefix>.min_range/.max_range`
    def aq(self, prefix, x, n_bits=None):
        if not self.quant:
            return x
        n_bits = n_bits or self.cfg.act_bits
        kmin, kmax = prefix + ".min_range", prefix + ".max_range"
        it = self.st.n_iter.get(prefix, 0)
        if self.st.observe and it < self.cfg.max_observations:
            self.st.n_iter[prefix] = it + 1
            with torch.no_grad():
                nmin, nmax = observe_act(self.P[kmin], self.P[kmax], x, self.cfg.alpha)
                self.P[kmin].data = nmin.reshape(1)
                self.P[kmax].data = nmax.reshape(1)
            return x
        if self.tap is not None:
            self.tap[prefix + "::pre"] = x
        return fq_act(x, self.P[kmin], self.P[kmax], n_bits)

    def wq(self, prefix, w, ch_axis=0):
        if not self.quant:
            return w
        kmin, kmax = prefix + ".min_range", prefix + ".max_range"
        if not self.st.weights_seen_for(prefix):
            with torch.no_grad():
                nmin, nmax = observe_weight(w, ch_axis)
                self.P[kmin].data = nmin
                self.P[kmax].data = nmax
            return w
        return fq_weight(w, self.P[kmin], self.P[kmax], self.cfg.weight_bits)


def _weights_seen_for(self, prefix):
    # GradientWeightFakeQuantize.observer_mode starts True and flips after the first call
    # (qat_quant.py:372-377); enable_observer(False) also clears it (load_model.py:16-19).
    seen = self.n_iter.get("W:" + prefix, 0) > 0 or self.weights_seen
    self.n_iter["W:" + prefix] = 1
    return seen


QuantState.weights_seen_for = _weights_seen_for


# --------------------------------------------------------------------------------------
# L1/L2/M1/M2/M3: the separator (convtasnetq.py:182-223 after quantize_model :243-288)
# --------------------------------------------------------------------------------------
def _gln(x, gamma, beta):
    return F.group_norm(x, 1, gamma, beta, EPS_GLN)


def _tcn_block(c: _Ctx, i: int, x):
    P, cfg = c.P, c.cfg
    p = "masker.TCN.%d." % i
    sb = p + "shared_block."
    d = cfg.dilation(i)
    # 1x1 expand + PReLU + FQ            (qat_layers.py:188-212)
    w = c.wq(sb + "0.weight_fake_quantize", P[sb + "0.conv1d.weight"])
    y = F.conv1d(x, w, P.get(sb + "0.conv1d.bias"))
    y = c.aq(sb + "0.activation_fake_quantize", F.prelu(y, P[sb + "0.nl.weight"]))
    # gLN + FQ                           (qat_layers.py:438-449)
    y = c.aq(sb + "2.activation_fake_quantize",
             _gln(y, P[sb + "2.groupnorm.weight"], P[sb + "2.groupnorm.bias"]))
    # depthwise dilated conv + PReLU + FQ
    w = c.wq(sb + "3.weight_fake_quantize", P[sb + "3.conv1d.weight"])
    y = F.conv1d(y, w, P.get(sb + "3.conv1d.bias"), padding=d,
                 dilation=d, groups=y.shape[1])
    y = c.aq(sb + "3.activation_fake_quantize", F.prelu(y, P[sb + "3.nl.weight"]))
    y = c.aq(sb + "5.activation_fake_quantize",
             _gln(y, P[sb + "5.groupnorm.weight"], P[sb + "5.groupnorm.bias"]))
    c.rec(p + "hidden", y)
    # residual / skip 1x1 + FQ           (qat_layers.py:124-146)
    w = c.wq(p + "res_conv.weight_fake_quantize", P[p + "res_conv.conv1d.weight"])
    res = c.aq(p + "res_conv.activation_fake_quantize", F.conv1d(y, w, P.get(p + "res_conv.conv1d.bias")))
    w = c.wq(p + "skip_conv.weight_fake_quantize", P[p + "skip_conv.conv1d.weight"])
    skip = c.aq(p + "skip_conv.activation_fake_quantize", F.conv1d(y, w, P.get(p + "skip_conv.conv1d.bias")))
    # AddQ                               (qat_layers.py:62-71)
    out = c.aq(p + "add.activation_fake_quantize", x + res)
    c.rec(p + "in", x)
    c.rec(p + "skip", skip)
    return out, skip


def _mask_generator(c: _Ctx, feats):
    P, cfg = c.P, c.cfg
    B = feats.shape[0]
    y = c.aq("masker.bottleneck.0.activation_fake_quantize",
             _gln(feats, P["masker.bottleneck.0.groupnorm.weight"], P["masker.bottleneck.0.groupnorm.bias"]))
    w = c.wq("masker.bottleneck.1.weight_fake_quantize", P["masker.bottleneck.1.conv1d.weight"])
    y = c.aq("masker.bottleneck.1.activation_fake_quantize",
             F.conv1d(y, w, P.get("masker.bottleneck.1.conv1d.bias")))
    c.rec("masker.bottleneck", y)
    y, acc = _tcn_block(c, 0, y)                       # convtasnetq.py:106-111
    c.rec("masker.TCN.0.out", y)
    for i in range(1, cfg.n_tcn):
        y, skip = _tcn_block(c, i, y)
        acc = c.aq("masker.adds.%d.activation_fake_quantize" % (i - 1), acc + skip)
        c.rec("masker.TCN.%d.out" % i, y)
    c.rec("masker.skip_sum", acc)
    y = c.aq("masker.mask_net.0.activation_fake_quantize", F.prelu(acc, P["masker.mask_net.0.nl.weight"]))
    w = c.wq("masker.mask_net.1.weight_fake_quantize", P["masker.mask_net.1.conv1d.weight"])
    y = F.conv1d(y, w, P.get("masker.mask_net.1.conv1d.bias"))
    y = c.aq("masker.mask_net.1.activation_fake_quantize", F.relu(y))
    return y.reshape(B, cfg.n_src, cfg.n_filters, -1)


def separator_forward(P: Params, x, cfg: SeparatorConfig, st: Optional[QuantState] = None,
                      quant: bool = True, tap: Optional[dict] = None):
    """ConvTasNetQ.forward.  quant=False is the float teacher (deep-copied before
    set_splitter_combiner, train_utils.py:25-26): 1-channel encoder, no splitter, no RQB."""
    st = st or QuantState()
    c = _Ctx(P, cfg, st, quant, tap)
    n_split = cfg.n_splitter if quant else 1
    n_comb = cfg.n_combiner if quant else 1
    xin = split_input(x, n_split)                                   # [B, n_split, T]
    c.rec("split", xin)
    B = xin.shape[0]
    # encoder (qat_layers.py:993-1038): in_quantizer is Identity for in_quant=False
    w = c.wq("encoder.weight_fake_quantize", P["encoder.conv1d.weight"] if quant else P["encoder.weight"])
    feats = c.aq("encoder.activation_fake_quantize", F.conv1d(xin, w, None, stride=cfg.stride))
    c.rec("encoder", feats)
    mask = _mask_generator(c, feats) if quant else _float_masker(P, cfg, feats)
    c.rec("mask", mask)
    masked = c.aq("mul.activation_fake_quantize", mask * feats.unsqueeze(1))   # qat_layers.py:86-96
    c.rec("masked", masked)
    Y = masked.reshape(B * cfg.n_src, cfg.n_filters, -1)
    # decoder (+ RQB)  qat_layers.py:1330-1354, :1188-1202
    if not quant:
        y0 = F.conv_transpose1d(Y, P["decoder.weight"], None, stride=cfg.stride)
        return combine_output(y0.reshape(1, B, cfg.n_src, 1, -1), 1)
    wd = c.wq("decoder.weight_fake_quantize", P["decoder.convTr1d.weight"], ch_axis=1)
    y0 = c.aq("decoder.activation_fake_quantize", F.conv_transpose1d(Y, wd, None, stride=cfg.stride), cfg.out_bits)
    outs = [y0]
    cur_in, cur_out = Y, y0
    for _ in range(1, n_comb):
        we = c.wq("decoder.residual_error_block.weight_fake_quantize",
                  P["decoder.residual_error_block.residual_encoder.weight"])
        Yq = F.conv1d(cur_out, we, None, stride=cfg.stride)
        Y1 = c.aq("decoder.residual_error_block.activation_fake_quantize", cur_in - Yq)
        y1 = F.conv_transpose1d(Y1, wd, None, stride=cfg.stride)
        cur_in = y1                                                  # qat_layers.py:1348-1351 reuses x
        cur_out = c.aq("decoder.activation_fake_quantize_residual", y1, cfg.out_bits)
        outs.append(cur_out)
    stacked = torch.stack(outs) if n_comb > 1 else y0.unsqueeze(0)
    c.rec("decoder", stacked)
    return combine_output(stacked.reshape(n_comb, B, cfg.n_src, 1, -1), n_comb)


def _float_masker(P, cfg, feats):
    """Un-quantised MaskGenerator (convtasnetq.py:101-115) with the float state_dict keys."""
    B = feats.shape[0]
    y = _gln(feats, P["masker.bottleneck.0.weight"], P["masker.bottleneck.0.bias"])
    y = F.conv1d(y, P["masker.bottleneck.1.weight"], P["masker.bottleneck.1.bias"])
    acc = None
    for i in range(cfg.n_tcn):
        p = "masker.TCN.%d." % i
        sb = p + "shared_block."
        d = cfg.dilation(i)
        h = F.prelu(F.conv1d(y, P[sb + "0.weight"], P[sb + "0.bias"]), P[sb + "1.weight"])
        h = _gln(h, P[sb + "2.weight"], P[sb + "2.bias"])
        h = F.prelu(F.conv1d(h, P[sb + "3.weight"], P[sb + "3.bias"], padding=d, dilation=d,
                             groups=h.shape[1]), P[sb + "4.weight"])
        h = _gln(h, P[sb + "5.weight"], P[sb + "5.bias"])
        res = F.conv1d(h, P[p + "res_conv.weight"], P[p + "res_conv.bias"])
        skip = F.conv1d(h, P[p + "skip_conv.weight"], P[p + "skip_conv.bias"])
        y = y + res
        acc = skip if acc is None else acc + skip
    h = F.prelu(acc, P["masker.mask_net.0.weight"])
    h = F.relu(F.conv1d(h, P["masker.mask_net.1.weight"], P["masker.mask_net.1.bias"]))
    return h.reshape(B, cfg.n_src, cfg.n_filters, -1)


def calibrate(P: Params, x, cfg: SeparatorConfig, passes: int = 2) -> QuantState:
    """`passes` observer forwards (load_model.enable_observer(model, True)), then observers off."""
    st = QuantState(observe=True)
    with torch.no_grad():
        for _ in range(passes):
            separator_forward(P, x, cfg, st, quant=True)
    st.observe = False
    st.weights_seen = True
    return st


# --------------------------------------------------------------------------------------
# S1/S2/S3: FQSS KD SI-SDR loss (mysystem.py:124-151, wsdr.py:46-95, asteroid 0.6 PIT)
# --------------------------------------------------------------------------------------
def pairwise_sisdr_ratio(est, tgt, weights=None):
    """rho[b, i, j] = linear SI-SDR of estimate i against target j (wsdr.py:63-92)."""
    tgt = tgt - tgt.mean(dim=2, keepdim=True)
    est = est - est.mean(dim=2, keepdim=True)
    t = tgt.unsqueeze(1)            # [B,1,S,T]
    e = est.unsqueeze(2)            # [B,S,1,T]
    dot = (e * t).sum(dim=3, keepdim=True)
    energy = (t ** 2).sum(dim=3, keepdim=True) + EPS_SDR
    proj = dot * t / energy
    noise = e - proj
    rho = (proj ** 2).sum(dim=3) / ((noise ** 2).sum(dim=3) + EPS_SDR)
    if weights is not None:
        rho = rho * weights[:, None, None]
    return rho


def pit_min_mean(pw):
    """asteroid 0.6 PITLossWrapper(pit_from='pw_mtx'), factorial search (n_src <= 3):
    loss[b] = min over permutations p of mean_j pw[b, p(j), j]; returns (mean_b loss, loss[b])."""
    S = pw.shape[-1]
    cands = []
    for perm in itertools.permutations(range(S)):
        cands.append(sum(pw[:, perm[j], j] for j in range(S)) / S)
    per_b = torch.stack(cands, dim=1).min(dim=1).values
    return per_b.mean(), per_b


def neg_sisdr_db_pit(est, tgt):
    """loss_func of the recipe = PITLossWrapper(pairwise_neg_sisdr) (asteroid_librimix_trainer.py:105)."""
    pw = -10 * torch.log10(pairwise_sisdr_ratio(est, tgt) + EPS_SDR)
    return pit_min_mean(pw)


def fqss_kd_loss(est, fest, tgt, kd_lambda=0.1):
    """System.common_step(train=True), mysystem.py:124-146.  Returns (loss, kd_loss_logged)."""
    with torch.no_grad():
        _, sdr_f = neg_sisdr_db_pit(fest.detach(), tgt)
        _, sdr_q = neg_sisdr_db_pit(est.detach(), tgt)
        w = 10 ** ((sdr_f - sdr_q) / 10)
    kd = -pit_min_mean(-pairwise_sisdr_ratio(est, fest.detach(), w))[0]
    task = -pit_min_mean(-pairwise_sisdr_ratio(est, tgt))[0]
    loss = -10 * torch.log10((1 - kd_lambda) * task + kd_lambda * kd + EPS_SDR)
    return loss, -10 * torch.log10(kd + EPS_SDR)


# --------------------------------------------------------------------------------------
# D1: data-parallel semantics (Lightning DDP): mean over ranks of per-rank gradients
# --------------------------------------------------------------------------------------
def ddp_mean_grads(per_rank_grads):
    keys = per_rank_grads[0].keys()
    n = len(per_rank_grads)
    out = {}
    for k in keys:
        gs = [g[k] if g[k] is not None else None for g in per_rank_grads]
        if all(g is None for g in gs):
            out[k] = None
        else:
            ref = next(g for g in gs if g is not None)
            out[k] = sum(g if g is not None else torch.zeros_like(ref) for g in gs) / n
    return out


def qat_step(P: Params, fP: Params, mix, tgt, cfg: SeparatorConfig, st: QuantState, kd_lambda=0.1):
    """One full QAT step on CPU: student fwd, teacher fwd, KD loss, backward.  Returns (loss, est)."""
    est = separator_forward(P, mix, cfg, st, quant=True)
    with torch.no_grad():
        fest = separator_forward(fP, mix, cfg, quant=False)
    loss, _ = fqss_kd_loss(est, fest, tgt, kd_lambda)
    loss.backward()
    return loss.detach(), est.detach()
