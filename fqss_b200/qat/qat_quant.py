"""B200 mirror of `quantization/qat/qat_quant.py` (hot-path subset).

Same public names, constructor signatures, Parameter names/shapes and observer semantics as the
reference; the arithmetic runs in libfqss_sm100.so (no torch elementwise chain, no host syncs):
  round_ste / linear_quantize ............ qat_quant.py:88-89, :124-147
  GradientActivationFakeQuantize ......... qat_quant.py:206-242
  GradientWeightFakeQuantize ............. qat_quant.py:350-381
  get_activation_quantizer / get_weight_quantizer ... :384-396
  TorchWeightFakeQuantize / TorchActivationFakeQuantize / TorchDymActivationFakeQuantize (export) ... :15-72
Not provided (never reached by the ConvTasNet recipe): mu-law quantiser (`nl=True`), MSE-histogram
observer, the dynamic training quantiser, scale_grad=True.
"""
import torch
import torch.nn as nn

from .. import ops


def _default_device():
    return torch.device("cuda" if torch.cuda.is_available() else "cpu")


def round_ste(x):
    """rint with identity gradient.  Kept for API parity; the kernels fuse it."""
    return x + (torch.round(x) - x).detach()


def linear_quantize(x, min_range, max_range, n_bits, sign=True, sym=False, scale_grad=False):
    """Uniform (sym=False) / symmetric per-channel (sym=True) fake-quant with STE gradients."""
    if scale_grad:
        raise NotImplementedError("scale_grad=True is not on the FQSS ConvTasNet path")
    if not sym:
        if min_range.numel() != 1:
            raise NotImplementedError("activation quantiser is per-tensor")
        return ops.FakeQuantAct.apply(x, min_range, max_range, int(n_bits))
    if not sign:
        raise NotImplementedError("unsigned symmetric weights are not on the FQSS ConvTasNet path")
    axis = 0
    for d in range(min_range.dim()):
        if min_range.shape[d] > 1:
            axis = d
            break
    else:
        # per-tensor range: any size-1 axis of x works as the "channel" axis (decoder: [Ci,1,K] -> axis 1)
        ones = [d for d in range(x.dim()) if x.shape[d] == 1]
        if not ones:
            raise NotImplementedError("per-tensor symmetric range needs a size-1 channel axis")
        axis = ones[0]
    return ops.FakeQuantWeight.apply(x, min_range, max_range, axis, int(n_bits))


class GradientActivationFakeQuantize(nn.Module):
    """Per-tensor asymmetric activation quantiser with learnable range and an EMA observer."""

    def __init__(self, gradient_based, n_bits=8, sym=False, scale_grad=False):
        super().__init__()
        if sym or scale_grad:
            raise NotImplementedError("sym / scale_grad activation quantisers are not on the ConvTasNet path")
        self.n_bits = n_bits
        self.sym = sym
        self.min_range = nn.Parameter(torch.tensor([-0.5]), requires_grad=gradient_based)
        self.max_range = nn.Parameter(torch.tensor([0.5]), requires_grad=gradient_based)
        self.max_observations = 50
        self.observer_mode = True
        self.alpha = 0.9
        self.n_iter = 0
        self.sign = True
        self.scale_grad = scale_grad

    def enable_observer(self, observer_mode):
        self.observer_mode = observer_mode

    # -- protocol used by the fused LayerQ wrappers -------------------------------------------
    def observing(self):
        """True while this call must record ranges and pass data through (qat_quant.py:228)."""
        return self.observer_mode and self.n_iter < self.max_observations

    def observe_(self, x):
        """min/max EMA update on the device (no .item(), no assert): qat_quant.py:229-232."""
        self.n_iter += 1
        ops.act_observe_(x, self.min_range.data, self.max_range.data, self.alpha)

    def forward(self, x):
        if self.observing():
            self.observe_(x)
            return x
        return ops.FakeQuantAct.apply(x, self.min_range, self.max_range, self.n_bits)


def ranges_ok(model):
    """0-dim bool tensor ON THE DEVICE: every activation quantiser of `model` has max_range > min_range and finite ranges,
    every weight quantiser a non-zero finite range.  The reference asserts `max_range >= min_range` on the host in every
    forward call (qat_quant.py:238, one sync per quantiser per step); the kernels never sync, so the check is a separate,
    cheap call for the training loop to make once per epoch / validation (`assert bool(ranges_ok(model))`)."""
    spans = []
    for m in model.modules():
        if isinstance(m, GradientActivationFakeQuantize):
            spans.append((m.max_range.detach() - m.min_range.detach()).reshape(-1))
        elif isinstance(m, GradientWeightFakeQuantize) and not m.observer_mode:
            spans.append(torch.maximum(m.min_range.detach().abs(), m.max_range.detach().abs()).reshape(-1))
    if not spans:
        return torch.ones((), dtype=torch.bool)
    v = torch.cat(spans)
    return torch.isfinite(v).all() & (v > 0).all()


class GradientWeightFakeQuantize(nn.Module):
    """Per-output-channel symmetric signed weight quantiser; first call captures amin/amax."""

    def __init__(self, gradient_based, weight_shape, n_bits=8, sym=True, ch_out_idx=0, scale_grad=False):
        super().__init__()
        if not sym or scale_grad:
            raise NotImplementedError("only the symmetric weight quantiser is on the ConvTasNet path")
        self.n_bits = n_bits
        self.sym = sym
        self.axis = ch_out_idx
        self.x_dims = [d for d in range(len(weight_shape)) if d != ch_out_idx]
        shape = [1] * len(weight_shape)
        shape[ch_out_idx] = weight_shape[ch_out_idx]
        dev = _default_device()
        self.min_range = nn.Parameter(-0.5 * torch.ones(shape, device=dev), requires_grad=gradient_based)
        self.max_range = nn.Parameter(0.5 * torch.ones(shape, device=dev), requires_grad=gradient_based)
        self.observer_mode = True
        self.sign = True
        self.scale_grad = scale_grad

    def enable_observer(self, observer_mode):
        self.observer_mode = observer_mode

    def forward(self, w):
        if self.observer_mode:
            ops.weight_observe_(w, self.min_range.data, self.max_range.data, self.axis)
            self.observer_mode = False
            return w
        return ops.FakeQuantWeight.apply(w, self.min_range, self.max_range, self.axis, self.n_bits)


# ---------------------------------------------------------------------------------------------
# export-time quantisers (qat_quant.py:15-72): learned ranges -> (scale, zero-point), evaluated with the arithmetic of
# torch.fake_quantize_per_{tensor,channel}_affine (csrc/export.cu)
# ---------------------------------------------------------------------------------------------
class TorchWeightFakeQuantize(nn.Module):
    def __init__(self, quantizer):
        super().__init__()
        with torch.no_grad():
            bound = torch.maximum(torch.abs(quantizer.min_range), torch.abs(quantizer.max_range))
            scales = bound / (2 ** (quantizer.n_bits - int(quantizer.sign)))
        self.scales = scales.flatten()
        self.zero_points = torch.zeros_like(self.scales)
        self.axis = quantizer.axis
        self.sign = quantizer.sign
        self.n_bits = quantizer.n_bits

    def forward(self, x):
        qmin = -2 ** (self.n_bits - 1) if self.sign else 0
        qmax = 2 ** (self.n_bits - 1) - 1 if self.sign else 2 ** self.n_bits - 1
        return ops.FakeQuantAffineChannel.apply(x, self.scales, self.axis, qmin, qmax)


class TorchActivationFakeQuantize(nn.Module):
    def __init__(self, quantizer):
        super().__init__()
        # host arithmetic, as in the reference (an export-time read): torch's CUDA `tensor / python_scalar` multiplies by the
        # reciprocal and would differ from the IEEE quotient by an ulp
        lo, hi = quantizer.min_range.detach().cpu(), quantizer.max_range.detach().cpu()
        scale = (hi - lo) / (2 ** quantizer.n_bits - 1)
        self.scale = float(scale)                       # export-time host read, as in the reference (:43)
        zp = int(torch.round(lo / self.scale))
        self.zero_point = -zp if float(lo) < 0 else zp  # the reference keeps a positive minimum's zero-point as is (:45)
        self.n_bits = quantizer.n_bits

    def forward(self, x):
        return ops.FakeQuantAffineTensor.apply(x, self.scale, self.zero_point, 0, 2 ** self.n_bits - 1)


class TorchDymActivationFakeQuantize(nn.Module):
    def __init__(self, quantizer):
        super().__init__()
        self.n_bits = quantizer.n_bits
        self.factor = quantizer.factor

    def forward(self, x):
        lo_t = torch.zeros(1, device=x.device)
        hi_t = torch.zeros(1, device=x.device)
        ops.act_observe_(x.detach(), lo_t, hi_t, 0.0)   # alpha = 0: ranges <- 0*0 + 1*min(x), max(x) (device reduction)
        lo, hi = self.factor * float(lo_t), self.factor * float(hi_t)
        import numpy as np
        f32 = np.float32
        scale32 = f32(f32(f32(hi) - f32(lo)) / f32(2 ** self.n_bits - 1))
        scale = float(scale32)
        zp = int(np.rint(f32(f32(lo) / scale32)))
        zp = -zp if lo < 0 else zp
        return ops.FakeQuantAffineTensor.apply(x, scale, zp, 0, 2 ** self.n_bits - 1)


def get_activation_quantizer(gradient_based=True, nl=False, n_bits=8):
    if nl:
        raise NotImplementedError("mu-law (inout_nl_quant) quantiser is not on the FQSS ConvTasNet path")
    return GradientActivationFakeQuantize(gradient_based, n_bits=n_bits)


def get_weight_quantizer(gradient_based=True, weight_shape=(1, 1, 1), n_bits=8, ch_out_idx=0):
    return GradientWeightFakeQuantize(gradient_based, weight_shape, n_bits=n_bits, ch_out_idx=ch_out_idx)


# reference names outside the ConvTasNet hot path (imported by the reference's other model files): importable placeholders
# that raise NotImplementedError when used -- see fqss_b200/shim.py
from ..shim import module_getattr as _module_getattr  # noqa: E402
__getattr__ = _module_getattr("quantization.qat.qat_quant")
