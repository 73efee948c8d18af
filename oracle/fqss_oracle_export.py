"""CPU oracle of the reference's export-time quantisers (TEST INFRASTRUCTURE ONLY -- imported by tests/, never by fqss_b200/).

Restates `TorchWeightFakeQuantize` (quantization/qat/qat_quant.py:15-35), `TorchActivationFakeQuantize` (:38-53) and
`TorchDymActivationFakeQuantize` (:56-72): the reference converts its learned ranges into (scale, zero-point) pairs and
evaluates `torch.fake_quantize_per_{tensor,channel}_affine`.  The arithmetic of those ATen ops is restated in numpy
(fp32, one rounding per operation):

    inv = 1.0f / scale ;  q = nearbyint(x * inv) + zero_point ;  y = (clamp(q, qmin, qmax) - zero_point) * scale

Pinned against the unmodified reference classes on CPU by tests/golden/make_golden.py (export.npz) and
tests/test_export_cpu.py -- bit-identical.
"""
import numpy as np

f32 = np.float32


def fake_quantize_affine(x, scale, zero_point, qmin, qmax):
    """x: fp32 array; scale / zero_point: scalars or arrays broadcastable against x.  -> (y fp32, q int64, mask bool)."""
    x = np.asarray(x, f32)
    scale = np.asarray(scale, f32)
    inv = (f32(1.0) / scale).astype(f32)
    q = np.rint((x * inv).astype(f32)).astype(np.float64) + np.asarray(zero_point, np.float64)
    q = q.astype(np.int64)
    qc = np.clip(q, qmin, qmax)
    y = ((qc - np.asarray(zero_point, np.float64)).astype(f32) * scale).astype(f32)
    return y, q, (q >= qmin) & (q <= qmax)


def act_export_params(min_range, max_range, n_bits=8):
    """(scale as fp32, zero_point as python int) exactly as TorchActivationFakeQuantize.__init__ derives them (:42-46)."""
    mn, mx = f32(min_range), f32(max_range)
    scale32 = f32(f32(mx - mn) / f32(2 ** n_bits - 1))
    scale = float(scale32)                                   # python float of the fp32 quotient
    zp = int(np.rint(f32(mn / scale32)))                     # torch.round(min_range / self.scale): fp32 tensor / python float
    zp = -zp if mn < 0 else zp
    return scale, zp


def act_export_forward(x, min_range, max_range, n_bits=8):
    scale, zp = act_export_params(min_range, max_range, n_bits)
    return fake_quantize_affine(x, f32(scale), zp, 0, 2 ** n_bits - 1)


def weight_export_params(min_range, max_range, n_bits=8, sign=True):
    """Per-channel scales of TorchWeightFakeQuantize (:19-24): max(|min|,|max|) / 2^(n_bits - sign); zero-points 0."""
    a = np.maximum(np.abs(np.asarray(min_range, f32)), np.abs(np.asarray(max_range, f32)))
    return (a / f32(2 ** (n_bits - int(sign)))).astype(f32).reshape(-1)


def weight_export_forward(w, min_range, max_range, axis, n_bits=8, sign=True):
    scales = weight_export_params(min_range, max_range, n_bits, sign)
    shape = [1] * np.ndim(w)
    shape[axis] = -1
    qmin, qmax = (-2 ** (n_bits - 1), 2 ** (n_bits - 1) - 1) if sign else (0, 2 ** n_bits - 1)
    return fake_quantize_affine(w, scales.reshape(shape), 0, qmin, qmax)
