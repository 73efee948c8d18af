"""Integer export of a trained (or calibrated) fake-quantised model, and its reload.

The reference stops at fake quantisation: the deployable artefact is implied by its quantisers -- int8 weight codes
with one fp32 step per output channel (qat_quant.py:126-135), uint8 activation codes with one (min, max) pair per
tensor (:136-147) -- and by the export hooks `replace_weight_quantizer` / `replace_activation_quantizer`
(qat_utils.py:334-345), which re-express the ranges as torch (scale, zero-point) pairs.  This module materialises it:

    ckpt = export_int8(model)            # every weight as int8 codes + per-channel ranges, every activation quantiser as
                                         # (min, max, scale, zero_point), the remaining float parameters as they are
    load_int8(model2, ckpt)              # rebuilds the fake-quantised model from the integers alone
    model2.eval()(x)                     # -> bit-identical to model.eval()(x)

Weights are reconstructed as  w = delta * code  with  delta = 2 * max(|min|, |max|) / 255  evaluated exactly like the
training quantiser; fake quantisation is idempotent on such a weight (its code is again `code`, its value again `w`), so the
reloaded model runs the same integer-code tensor-core GEMMs (csrc/gemm_tc.cu: uint8 x int8 codes, exact fp32 accumulation)
on the same operands.  Everything is computed by the library's quantiser kernels (`ops.weight_codes`); torch only moves
bytes here.
"""
import torch
import torch.nn as nn

from . import ops
from .qat.qat_quant import GradientActivationFakeQuantize, GradientWeightFakeQuantize

FORMAT = "fqss-int8-v1"


def _owner_weights(model):
    """(weight-quantiser name, quantiser, parameter name, parameter) for every quantised weight of the model."""
    out = []
    for name, mod in model.named_modules():
        wq = getattr(mod, "weight_fake_quantize", None)
        if not isinstance(wq, GradientWeightFakeQuantize):
            continue
        target = None
        for attr in ("conv1d", "convTr1d", "linear", "residual_encoder"):
            sub = getattr(mod, attr, None)
            if sub is not None and hasattr(sub, "weight"):
                target = (attr, sub)
                break
        if target is None:
            raise NotImplementedError("export: %s has a weight quantiser but no known weight owner" % name)
        pname = (name + "." if name else "") + target[0] + ".weight"
        out.append(((name + "." if name else "") + "weight_fake_quantize", wq, pname, target[1].weight))
    return out


def export_int8(model):
    """-> dict: format tag, int8 weight codes + ranges, activation quantiser parameters, float remainder of the state dict."""
    if any(isinstance(m, GradientActivationFakeQuantize) and m.observing() for m in model.modules()):
        raise RuntimeError("export_int8: the model is still calibrating (observer mode); call enable_observer(model, False) first")
    sd = model.state_dict()
    weights, acts, taken = {}, {}, set()
    with torch.no_grad():
        for qname, wq, pname, w in _owner_weights(model):
            if wq.observer_mode:
                raise RuntimeError("export_int8: weight quantiser %s has not seen its weight yet" % qname)
            code = ops.weight_codes(w.detach(), wq.min_range.detach(), wq.max_range.detach(), wq.axis, wq.n_bits)
            weights[pname] = {"code": code.cpu(), "axis": wq.axis, "n_bits": wq.n_bits, "quantizer": qname,
                              "min_range": wq.min_range.detach().cpu().clone(), "max_range": wq.max_range.detach().cpu().clone()}
            taken.update((pname, qname + ".min_range", qname + ".max_range"))
        for name, mod in model.named_modules():
            if isinstance(mod, GradientActivationFakeQuantize):
                lo, hi = mod.min_range.detach().cpu().clone(), mod.max_range.detach().cpu().clone()
                scale = (hi - lo) / (2 ** mod.n_bits - 1)
                entry = {"min_range": lo, "max_range": hi, "n_bits": mod.n_bits, "scale": float(scale)}
                zp = int(torch.round(lo / float(scale)))
                entry["zero_point"] = -zp if float(lo) < 0 else zp          # the reference's convention (qat_quant.py:44-45)
                acts[name] = entry
                taken.update((name + ".min_range", name + ".max_range"))
    rest = {k: v.detach().cpu().clone() for k, v in sd.items() if k not in taken}
    return {"format": FORMAT, "weights": weights, "activations": acts, "float": rest}


def int8_state_dict(ckpt, device="cpu"):
    """The full fake-quant state dict implied by an int8 checkpoint (weights de-quantised as delta * code)."""
    if ckpt.get("format") != FORMAT:
        raise ValueError("not an %s checkpoint" % FORMAT)
    sd = {k: v.to(device) for k, v in ckpt["float"].items()}
    for pname, e in ckpt["weights"].items():
        lo, hi = e["min_range"].cpu(), e["max_range"].cpu()
        levels = float(2 ** e["n_bits"] - 1)
        # host arithmetic: the IEEE quotient the training quantiser uses (torch's CUDA tensor / scalar multiplies by 1/scalar)
        delta = (2.0 * torch.maximum(lo.abs(), hi.abs())) / levels
        sd[pname] = (delta * e["code"].cpu().to(torch.float32)).to(device)
        sd[e["quantizer"] + ".min_range"] = lo.to(device)
        sd[e["quantizer"] + ".max_range"] = hi.to(device)
    for name, e in ckpt["activations"].items():
        sd[name + ".min_range"] = e["min_range"].to(device)
        sd[name + ".max_range"] = e["max_range"].to(device)
    return sd


def load_int8(model, ckpt):
    """Load an int8 checkpoint into a quantised model of the same architecture (strict) and end calibration."""
    dev = next(model.parameters()).device
    model.load_state_dict(int8_state_dict(ckpt, dev), strict=True)
    for m in model.modules():
        if isinstance(m, (GradientActivationFakeQuantize, GradientWeightFakeQuantize)):
            m.enable_observer(False)
    return model


def checkpoint_bytes(ckpt):
    """(bytes of the int8 checkpoint, bytes of the fp32 state dict it replaces)."""
    q = sum(e["code"].numel() + 4 * (e["min_range"].numel() + e["max_range"].numel()) for e in ckpt["weights"].values())
    q += sum(8 for _ in ckpt["activations"]) + sum(4 * v.numel() for v in ckpt["float"].values())
    f = sum(4 * (e["code"].numel() + e["min_range"].numel() + e["max_range"].numel()) for e in ckpt["weights"].values())
    f += sum(8 for _ in ckpt["activations"]) + sum(4 * v.numel() for v in ckpt["float"].values())
    return q, f
