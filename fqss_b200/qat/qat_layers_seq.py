"""Quantised wrappers of the sequence-model layers (DPTNetQ / SepformerQ): LinearQ, LinearNlQ,
LSTMQ, MultiheadAttentionQ, Conv2dQ, Conv2dNlQ, ConstQ (reference: qat_layers.py:100-122, 156-186, 521-613, 865-990), with the
reference's attribute names (= state-dict keys) and constructor signatures.

Scope of the native code on this row: every QUANTISER -- per-channel weight fake-quant, per-tensor activation fake-quant
with learnable ranges, observers, their straight-through backward and range gradients -- runs on the sm_100a kernels of
libfqss_sm100 (csrc/fq_ops.cu: bit-exact against the reference's arithmetic, no host syncs), and so do the LSTM RECURRENCE
(csrc/lstm.cu: the 8-bit codes of the fake-quantised recurrent weights stay in registers for the whole sequence, forward
and backward) and the ATTENTION CORE softmax(q k^T) v (csrc/attention.cu: one pass per direction, the score tensor is never
materialised).  The remaining dense float math between two quantisers (F.linear incl. the LSTM's batched input projection
and the attention's in / out projections, LayerNorm) is delegated to torch, as in the reference itself (which calls _VF.lstm / torch.bmm):
these models are the "next" rows of the scope table, not the hot path.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import _VF

from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import _native as N
from .._native import check, lib, ptr, stream_ptr
from .qat_layers import LayerQ
from .qat_quant import GradientWeightFakeQuantize, get_activation_quantizer, get_weight_quantizer

import os as _os

NATIVE_LSTM = _os.environ.get("FQSS_NATIVE_LSTM", "1") not in ("", "0")      # False: the recurrence of LSTMQ stays on torch's LSTM (A/B runs, tests)


NATIVE_ATTENTION = _os.environ.get("FQSS_NATIVE_ATTENTION", "1") not in ("", "0")      # False: torch.bmm / softmax / bmm


class SmallHeadAttention(Function):
    """o = softmax(q k^T) v per (batch, head) on csrc/attention.cu (scores never materialised); q [BH,Lq,hd], k / v [BH,Lk,hd]."""

    @staticmethod
    def forward(ctx, q, k, v):
        N.require_cuda(q, k, v)
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        BH, Lq, hd = q.shape
        Lk = k.shape[1]
        o = torch.empty_like(q)
        lse = torch.empty((BH, Lq), device=q.device)
        check(lib().fqss_attn_fwd(ptr(q), ptr(k), ptr(v), ptr(o), ptr(lse), BH, Lq, Lk, hd, stream_ptr()))
        ctx.save_for_backward(q, k, v, o, lse)
        return o

    @staticmethod
    @once_differentiable
    def backward(ctx, do):
        q, k, v, o, lse = ctx.saved_tensors
        BH, Lq, hd = q.shape
        Lk = k.shape[1]
        do = do.contiguous()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        ws = torch.empty((BH, Lq), device=q.device)
        check(lib().fqss_attn_bwd(ptr(q), ptr(k), ptr(v), ptr(o), ptr(do), ptr(lse), ptr(dq), ptr(dk), ptr(dv), ptr(ws), BH, Lq, Lk, hd,
                                  stream_ptr()))
        return dq, dk, dv


def attention_eligible(q, k, observing):
    """The fused kernel covers fp32 CUDA tensors, head sizes 8 / 16 / 32 and heads whose K, V fit shared memory; while the
    (result-discarding) score / softmax quantisers of the reference still observe, the scores must exist: library path."""
    if not NATIVE_ATTENTION or observing or not q.is_cuda or q.dtype != torch.float32 or q.shape[0] >= 65536:
        return False
    hd = q.shape[-1]
    return hd in (8, 16, 32) and int(lib().fqss_attn_smem_bytes(q.shape[1], k.shape[1], hd)) <= 200 * 1024


class LSTMRecurrence(Function):
    """h_t of a one-layer (bi)directional LSTM from its input projections gx [D,T,N,4H], on the register-resident integer-code
    kernels of csrc/lstm.cu.  whh_fq [D,4H,H] (the fake-quantised recurrent weights) carries the autograd edge; the kernels
    re-derive the same codes / steps from the raw weights and their quantiser ranges."""

    @staticmethod
    def forward(ctx, gx, whh_fq, W0, W1, mn0, mn1, mx0, mx1):
        N.require_cuda(gx, whh_fq, W0, W1, mn0, mn1, mx0, mx1)
        D, T, nb, G = gx.shape
        H = G // 4
        gx = gx.contiguous()
        dev = gx.device
        out = torch.empty((T, nb, D * H), device=dev)
        gates = torch.empty((D, T, nb, G), device=dev)
        cseq = torch.empty((D, T, nb, H), device=dev)
        W0c, W1c = W0.detach().contiguous(), (W1.detach().contiguous() if W1 is not None else None)
        check(lib().fqss_lstm_rec_fwd(ptr(gx), ptr(W0c), ptr(W1c) or None, ptr(mn0), ptr(mn1) or None, ptr(mx0), ptr(mx1) or None,
                                      ptr(out), ptr(gates), ptr(cseq), T, nb, H, D, stream_ptr()))
        ctx.save_for_backward(out, gates, cseq, W0c, W1c, mn0, mn1, mx0, mx1)
        ctx.dims = (D, T, nb, H)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        out, gates, cseq, W0c, W1c, mn0, mn1, mx0, mx1 = ctx.saved_tensors
        D, T, nb, H = ctx.dims
        dout = dout.contiguous()
        dG = torch.empty((D, T, nb, 4 * H), device=dout.device)
        check(lib().fqss_lstm_rec_bwd(ptr(dout), ptr(gates), ptr(cseq), ptr(W0c), ptr(W1c) or None, ptr(mn0), ptr(mn1) or None, ptr(mx0),
                                      ptr(mx1) or None, ptr(dG), T, nb, H, D, stream_ptr()))
        # dW_hh[d] = sum_t dG[d, t]^T h_{t-1}: one batched GEMM per direction on the previous hidden states
        zero = out.new_zeros(1, nb, H)
        dW = []
        for d in range(D):
            hd = out[:, :, d * H:(d + 1) * H]
            hprev = torch.cat([zero, hd[:-1]], 0) if d == 0 else torch.cat([hd[1:], zero], 0)
            dW.append(dG[d].reshape(T * nb, 4 * H).t() @ hprev.reshape(T * nb, H))
        return dG, torch.stack(dW), None, None, None, None, None, None


class Const(nn.Module):
    """Marker module for a constant tensor entering the graph (qat_layers.py:40-46)."""

    def __init__(self, shape=None):
        super().__init__()
        self.shape = shape

    def forward(self, x):
        return x


class Div(nn.Module):
    def forward(self, x1, x2):
        return torch.div(x1, x2)


class ConstQ(LayerQ):
    def __init__(self, const, gradient_based=True, act_quant=True, act_n_bits=8):
        super().__init__(gradient_based=gradient_based, act_quant=act_quant, act_n_bits=act_n_bits)
        self.const = const

    def forward(self, x):
        return self.activation_fake_quantize(x)


class DivQ(LayerQ):
    def __init__(self, div, gradient_based=True, act_quant=True, act_n_bits=8):
        super().__init__(gradient_based=gradient_based, act_quant=act_quant, act_n_bits=act_n_bits)
        if not isinstance(div, Div):
            raise Exception("DivQ wraps Div, got %s" % type(div))
        self.div = div

    def forward(self, x1, x2):
        return self.activation_fake_quantize(torch.div(x1, x2))


def _weighted(layer_cls_name, mod, want, **kw):
    if not isinstance(mod, want):
        raise Exception("%s wraps %s, got %s" % (layer_cls_name, want.__name__, type(mod)))


class LinearQ(LayerQ):
    def __init__(self, linear, gradient_based=True, weight_quant=True, act_quant=True, act_n_bits=8, weight_n_bits=8):
        super().__init__(gradient_based=gradient_based, weight_quant=weight_quant, act_quant=act_quant,
                         weight_shape=linear.weight.shape, act_n_bits=act_n_bits, weight_n_bits=weight_n_bits)
        _weighted("LinearQ", linear, nn.Linear)
        self.linear = linear

    def _linear(self, x):
        if self.do_mac_op:
            self.mac_op = x.numel() // x.shape[-1] * self.linear.weight.numel()
        return F.linear(x, self.weight_fake_quantize(self.linear.weight), self.linear.bias)

    def forward(self, x):
        return self.activation_fake_quantize(self._linear(x))


class LinearNlQ(LinearQ):
    def __init__(self, linear, nl, gradient_based=True, weight_quant=True, act_quant=True, act_n_bits=8, weight_n_bits=8):
        super().__init__(linear, gradient_based, weight_quant, act_quant, act_n_bits, weight_n_bits)
        self.nl = nl

    def forward(self, x):
        return self.activation_fake_quantize(self.nl(self._linear(x)))


class Conv2dQ(LayerQ):
    def __init__(self, conv2d, gradient_based=True, weight_quant=True, act_quant=True, act_n_bits=8, weight_n_bits=8):
        super().__init__(gradient_based=gradient_based, weight_quant=weight_quant, act_quant=act_quant,
                         weight_shape=conv2d.weight.shape, act_n_bits=act_n_bits, weight_n_bits=weight_n_bits)
        _weighted("Conv2dQ", conv2d, nn.Conv2d)
        self.conv2d = conv2d

    def _conv(self, x):
        c = self.conv2d
        return F.conv2d(x, self.weight_fake_quantize(c.weight), c.bias, c.stride, c.padding, c.dilation, c.groups)

    def forward(self, x):
        return self.activation_fake_quantize(self._conv(x))


class Conv2dNlQ(Conv2dQ):
    def __init__(self, conv2d, nl, gradient_based=True, weight_quant=True, act_quant=True, act_n_bits=8, weight_n_bits=8):
        super().__init__(conv2d, gradient_based, weight_quant, act_quant, act_n_bits, weight_n_bits)
        self.nl = nl

    def forward(self, x):
        return self.activation_fake_quantize(self.nl(self._conv(x)))


class LSTMQ(LayerQ):
    """nn.LSTM with one per-output-row weight quantiser per weight matrix (`weight_quantizers_dict`, keyed by the LSTM's own
    flat weight names) and one activation quantiser on the output sequence; zero initial state (qat_layers.py:571-613)."""

    def __init__(self, lstm, gradient_based=True, weight_quant=True, act_quant=True, act_n_bits=8, weight_n_bits=8):
        super().__init__(gradient_based=gradient_based, act_quant=act_quant, act_n_bits=act_n_bits)
        _weighted("LSTMQ", lstm, nn.LSTM)
        self.lstm = lstm
        self.num_directions = 2 if lstm.bidirectional else 1
        self.real_hidden_size = lstm.proj_size if lstm.proj_size > 0 else lstm.hidden_size
        self.weight_quantizers_dict = nn.ModuleDict()
        for name, w in zip(lstm._flat_weights_names, lstm._flat_weights):
            if name.startswith("weight"):
                self.weight_quantizers_dict[name] = (get_weight_quantizer(gradient_based, w.shape, n_bits=weight_n_bits)
                                                     if weight_quant else nn.Identity())

    def _native_ok(self, x):
        lstm = self.lstm
        if not NATIVE_LSTM or not x.is_cuda or x.dtype != torch.float32 or x.dim() != 3:
            return False
        if lstm.num_layers != 1 or lstm.proj_size != 0 or not lstm.bias or lstm.hidden_size not in (32, 64, 128):
            return False
        for name in lstm._flat_weights_names:
            if name.startswith("weight_hh"):
                q = self.weight_quantizers_dict[name]
                if not isinstance(q, GradientWeightFakeQuantize) or q.observer_mode or q.n_bits != 8 or q.axis != 0:
                    return False
        return True

    def _forward_native(self, x):
        """Input projections of all steps as one GEMM per direction (torch), the recurrence on csrc/lstm.cu."""
        lstm = self.lstm
        if lstm.batch_first:
            x = x.transpose(0, 1)
        gx, whh, raw, mn, mx = [], [], [], [], []
        for sfx in ("", "_reverse")[: self.num_directions]:
            q_ih, q_hh = self.weight_quantizers_dict["weight_ih_l0" + sfx], self.weight_quantizers_dict["weight_hh_l0" + sfx]
            w_ih, w_hh = getattr(lstm, "weight_ih_l0" + sfx), getattr(lstm, "weight_hh_l0" + sfx)
            bias = getattr(lstm, "bias_ih_l0" + sfx) + getattr(lstm, "bias_hh_l0" + sfx)
            gx.append(F.linear(x, q_ih(w_ih), bias))
            whh.append(q_hh(w_hh))
            raw.append(w_hh)
            mn.append(q_hh.min_range)
            mx.append(q_hh.max_range)
        two = self.num_directions == 2
        y = LSTMRecurrence.apply(torch.stack(gx), torch.stack(whh), raw[0], raw[1] if two else None, mn[0], mn[1] if two else None,
                                 mx[0], mx[1] if two else None)
        return y.transpose(0, 1) if lstm.batch_first else y

    def forward(self, x):
        if self._native_ok(x):
            return [self.activation_fake_quantize(self._forward_native(x))]
        lstm = self.lstm
        flat = []
        for name in lstm._flat_weights_names:
            w = getattr(lstm, name)
            flat.append(self.weight_quantizers_dict[name](w) if name.startswith("weight") else w)
        nb = x.size(0) if lstm.batch_first else x.size(1)
        h0 = x.new_zeros(lstm.num_layers * self.num_directions, nb, self.real_hidden_size)
        c0 = x.new_zeros(lstm.num_layers * self.num_directions, nb, lstm.hidden_size)
        y = _VF.lstm(x, (h0, c0), flat, lstm.bias, lstm.num_layers, lstm.dropout, lstm.training, lstm.bidirectional,
                     lstm.batch_first)
        if self.do_mac_op:
            B, Li, Ci = x.shape
            for name in lstm._flat_weights_names:
                if name.startswith("weight"):
                    Fw, Cw = getattr(lstm, name).shape
                    self.mac_op += (B * Li if Cw == Ci else B * self.real_hidden_size) * Fw * Cw + 3 * B * self.real_hidden_size
        return [self.activation_fake_quantize(y[0])]


class MultiheadAttentionQ(LayerQ):
    """Self/cross attention with quantised projections (qat_layers.py:865-990).  Arithmetic as the reference's `forward`:
    the packed in-projection is applied to query, key and value separately (each followed by its own quantiser, then the
    matching third is kept); q / sqrt(head_dim) is quantised; the two statements `attn - fq(attn)` of the reference discard
    their result, i.e. neither the scores nor the softmax output are quantised (their quantisers still observe / are
    called, so they keep receiving calibration updates and zero gradients as in the reference); the head outputs and the
    out-projection are quantised."""

    def __init__(self, mha, gradient_based=True, weight_quant=True, act_quant=True, act_n_bits=8, weight_n_bits=8):
        super().__init__(gradient_based=gradient_based, act_quant=act_quant, act_n_bits=act_n_bits)
        _weighted("MultiheadAttentionQ", mha, nn.MultiheadAttention)
        self.mha = mha
        self.do = mha.out_proj.weight.shape[0]
        self.head_dim = mha.embed_dim // mha.num_heads

        def aq():
            return get_activation_quantizer(gradient_based, n_bits=act_n_bits) if act_quant else nn.Identity()
        self.activation_fake_quantize_q = aq()
        self.activation_fake_quantize_k = aq()
        self.activation_fake_quantize_v = aq()
        self.activation_fake_quantize_div = aq()
        self.activation_fake_quantize_attn = aq()
        self.activation_fake_quantize_softmax = aq()
        self.activation_fake_quantize_head = aq()
        self.weight_fake_quantize_in = (get_weight_quantizer(gradient_based, mha.in_proj_weight.shape, n_bits=weight_n_bits)
                                        if weight_quant else nn.Identity())
        self.weight_fake_quantize_out = (get_weight_quantizer(gradient_based, mha.out_proj.weight.shape, n_bits=weight_n_bits)
                                         if weight_quant else nn.Identity())

    @staticmethod
    def _observe_only(q, x):
        """The reference evaluates the quantiser and drops the result: the only lasting effect is the range update while it
        observes (in steady state the call changes nothing, so it is skipped)."""
        if hasattr(q, "observing") and q.observing():
            q.observe_(x.detach())

    def forward(self, query, key, value, attn_mask=None, key_padding_mask=None, need_weights=False, is_causal=False):
        mha = self.mha
        Wi = self.weight_fake_quantize_in(mha.in_proj_weight)
        Wo = self.weight_fake_quantize_out(mha.out_proj.weight)
        if mha.batch_first:
            query, key, value = query.transpose(1, 0), key.transpose(1, 0), value.transpose(1, 0)
        Lq, nb, _ = query.shape
        Lk, Lv = key.shape[0], value.shape[0]
        E, H, hd = mha.embed_dim, mha.num_heads, self.head_dim
        same = key is query and value is query
        Pq = F.linear(query, Wi, mha.in_proj_bias)
        Pk = Pq if same else F.linear(key, Wi, mha.in_proj_bias)
        Pv = Pq if same else F.linear(value, Wi, mha.in_proj_bias)
        Q = self.activation_fake_quantize_q(Pq)[..., :E]
        K = self.activation_fake_quantize_k(Pk)[..., E:2 * E]
        V = self.activation_fake_quantize_v(Pv)[..., 2 * E:]
        q = Q.reshape(Lq, nb * H, hd).permute(1, 0, 2)
        k = K.reshape(Lk, nb * H, hd).permute(1, 0, 2)
        v = V.reshape(Lv, nb * H, hd).permute(1, 0, 2)
        q = self.activation_fake_quantize_div(q / math.sqrt(hd))
        observing = any(hasattr(m, "observing") and m.observing()
                        for m in (self.activation_fake_quantize_attn, self.activation_fake_quantize_softmax))
        if attention_eligible(q, k, observing):
            ctxv = SmallHeadAttention.apply(q, k, v)                            # csrc/attention.cu: one pass, no score tensor
        else:
            attn = torch.bmm(q, k.transpose(-2, -1))
            self._observe_only(self.activation_fake_quantize_attn, attn)        # qat_layers.py:934: `attn - fq(attn)`, result discarded
            attn = torch.softmax(attn, dim=-1)
            self._observe_only(self.activation_fake_quantize_softmax, attn)     # qat_layers.py:936: likewise
            ctxv = torch.bmm(attn, v)
        heads = self.activation_fake_quantize_head(ctxv)
        flat = heads.transpose(1, 0).reshape(Lq * nb, E)
        y = F.linear(flat, Wo, mha.out_proj.bias).reshape(Lq, nb, self.do)
        if self.do_mac_op:
            self.mac_op = (Lq + (0 if same else Lk + Lv)) * nb * Wi.numel() + flat.shape[0] * Wo.numel() \
                + nb * H * Lq * Lk * hd * 2
        if mha.batch_first:
            y = y.transpose(1, 0)
        return self.activation_fake_quantize(y),
