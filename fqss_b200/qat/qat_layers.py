"""B200 mirror of `quantization/qat/qat_layers.py` (the nine wrappers ConvTasNetQ uses; SURVEY.md
8a rows L1/L2).  Class names, constructor signatures and attribute names (= state_dict keys) match
the reference; each forward is a short sequence of libfqss_sm100 kernels:

    conv kernel (1x1 SGEMM / depthwise / strided / transposed)  ->  one fused pass that applies the
    nonlinearity and the activation fake-quant (or, while observing, the nonlinearity + range EMA).

Reference: LayerQ :49-59, AddQ :62, SubQ :74, MulQ :86, Conv1dQ :124, Conv1dNlQ :188, GroupNormQ :438,
NlQ :511, Conv1dEncoderQ :993, ResidualErrorBlock :1105, ConvTr1dDecoderQ :1305.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _native as N
from .. import ops
from .qat_quant import GradientActivationFakeQuantize, get_activation_quantizer, get_weight_quantizer


class Add(nn.Module):
    def forward(self, x1, x2):
        return torch.add(x1, x2)


class Sub(nn.Module):
    def forward(self, x1, x2):
        return torch.sub(x1, x2)


class Mul(nn.Module):
    def forward(self, x1, x2):
        return torch.mul(x1, x2)


class LayerQ(nn.Module):
    """Base: owns `activation_fake_quantize` / `weight_fake_quantize` (Identity when disabled)."""

    def __init__(self, gradient_based=True, weight_quant=False, act_quant=False, act_nl_quantizer=False,
                 weight_shape=(1, 1, 1), ch_out_idx=0, act_n_bits=8, weight_n_bits=8, do_mac_op=False):
        super().__init__()
        self.weight_quant = weight_quant
        self.act_quant = act_quant
        self.gradient_based = gradient_based
        self.activation_fake_quantize = (get_activation_quantizer(gradient_based, n_bits=act_n_bits, nl=act_nl_quantizer)
                                         if act_quant else nn.Identity())
        self.weight_fake_quantize = (get_weight_quantizer(gradient_based, weight_shape, ch_out_idx=ch_out_idx,
                                                          n_bits=weight_n_bits) if weight_quant else nn.Identity())
        self.do_mac_op = do_mac_op
        self.mac_op = 0

    # one fused pass: y = FQ(op(x1[,x2])); while the quantiser observes: y = op(...), ranges <- EMA
    def _finish(self, kind, x1, x2=None, slope=None, gamma=None, beta=None, eps=0.0, quantizer=None):
        q = self.activation_fake_quantize if quantizer is None else quantizer
        if isinstance(q, GradientActivationFakeQuantize):
            if q.observing():
                y = ops.pointwise_fq(kind, x1, x2, slope, gamma, beta, quant=False, eps=eps)
                q.observe_(y)
                return y
            y = ops.pointwise_fq(kind, x1, x2, slope, gamma, beta, q.min_range, q.max_range, True, q.n_bits, eps)
            y._fq_src = q         # tag: y lies on q's grid -- a quantised 1x1 conv consuming THIS tensor object runs on the
            return y              # tensor cores with integer-code operands (_conv1d_q); views / copies drop the tag
        if isinstance(q, nn.Identity):
            return ops.pointwise_fq(kind, x1, x2, slope, gamma, beta, quant=False, eps=eps)
        raise NotImplementedError("unsupported activation quantiser %s" % type(q).__name__)


_TORCH_NL = (nn.Tanh, nn.Sigmoid, nn.GELU, nn.LeakyReLU)      # nonlinearities without a fused kernel (DPTNetQ / SepformerQ gates)


def _nl_kind(nl):
    """(kernel kind, slope tensor) for the nonlinearities the recipe uses."""
    if nl is None or isinstance(nl, nn.Identity):
        return N.PW_IDENT, None
    if isinstance(nl, nn.PReLU):
        if nl.weight.numel() != 1:
            raise NotImplementedError("per-channel PReLU is not on the ConvTasNet path")
        return N.PW_PRELU, nl.weight
    if isinstance(nl, nn.ReLU):
        return N.PW_RELU, None
    raise NotImplementedError("nonlinearity %s has no sm_100a kernel here" % type(nl).__name__)


def _conv1d(x, w, conv):
    """Dispatch an nn.Conv1d geometry to its kernel; anything else is out of scope (no fallback)."""
    k, s, p, d, g = conv.kernel_size[0], conv.stride[0], conv.padding[0], conv.dilation[0], conv.groups
    if isinstance(p, str):
        raise NotImplementedError("string padding")
    if k == 1 and s == 1 and p == 0 and g == 1:
        return ops.Conv1x1.apply(x, w, conv.bias)
    if g == conv.in_channels == conv.out_channels and s == 1 and (k % 2) == 1 and p == d * (k - 1) // 2:
        return ops.DepthwiseConv.apply(x, w, conv.bias, d)
    if g == 1 and p == 0 and d == 1 and conv.bias is None:
        return ops.StridedConv.apply(x, w, s)
    raise NotImplementedError("Conv1d geometry k=%d s=%d p=%d d=%d groups=%d bias=%s is not on the FQSS ConvTasNet path"
                              % (k, s, p, d, g, conv.bias is not None))


TENSOR_CORE_CONV1X1 = True      # False: every 1x1 conv of the per-layer wrappers stays on the fp32 SIMT kernels (tests / A/B)


def _conv1d_q(layer, x):
    """conv part of Conv1dQ / Conv1dNlQ.forward.  A 1x1 conv whose input carries the tag of the 8-bit quantiser that produced
    it, with an 8-bit per-channel weight quantiser in steady state, is the code-operand tcgen05 GEMM of the fused engine
    (tcn_engine.CodeConv1x1: exact integer accumulation, quantised weights never materialised); an un-quantised 1x1 conv under
    no_grad (the float teacher of a recipe without a dedicated engine) takes the split-bf16 tcgen05 GEMM (fp32-grade, 2^-16);
    everything else the per-layer SIMT kernels."""
    from .. import tcn_engine as E
    conv = layer.conv1d
    if TENSOR_CORE_CONV1X1 and conv.kernel_size[0] == 1 and x.is_cuda and x.dim() == 3:
        q_in = getattr(x, "_fq_src", None)
        if q_in is not None and E.code_conv_eligible(layer, q_in, x):
            return E.code_conv(layer, q_in, x)
        if isinstance(layer.weight_fake_quantize, nn.Identity) and E.float_conv_eligible(conv, x):
            return E.float_conv(conv, x)
    return _conv1d(x, layer.weight_fake_quantize(conv.weight), conv)


def _linear_1x1(x, weight, wq):
    """Per-frame Linear layer of the music model's decoder / RQB (qat_layers.py:1110-1121, 1256-1302) as a channels-first 1x1
    conv: on the code-operand tcgen05 GEMM when `x` carries the tag of the steady 8-bit quantiser that produced it and the
    weight quantiser is an 8-bit per-row one in steady state (40-row / 40-column shapes are zero-padded to the tiles,
    tcn_engine.CodeConv1x1); the fp32 SIMT kernel on the materialised fake-quantised weight otherwise."""
    from .. import tcn_engine as E
    src = getattr(x, "_fq_src", None)
    if TENSOR_CORE_CONV1X1 and src is not None and E.code_linear_eligible(weight, wq, src, x):
        return E.CodeConv1x1.apply(x, src.min_range, src.max_range, weight, wq.min_range, wq.max_range, None)
    if TENSOR_CORE_CONV1X1 and isinstance(wq, nn.Identity) and E.float_linear_eligible(weight, x):
        return E.float_linear(weight, x)          # un-quantised teacher under no_grad: split-bf16 GEMM (fp32-grade)
    return ops.Conv1x1.apply(x, wq(weight).unsqueeze(-1), None)


def _conv_out_len(conv, L):
    return math.floor((L + 2 * conv.padding[0] - conv.dilation[0] * (conv.kernel_size[0] - 1) - 1) / conv.stride[0] + 1)


class AddQ(LayerQ):
    def __init__(self, add, gradient_based=True, act_quant=True, act_n_bits=8):
        super().__init__(gradient_based=gradient_based, act_quant=act_quant, act_n_bits=act_n_bits)
        if not isinstance(add, Add):
            raise Exception("AddQ wraps Add, got %s" % type(add))
        self.add = add

    def forward(self, x1, x2):
        if x1.shape != x2.shape:         # e.g. SepformerQ's positional table [1, T, F] added to [B, T, F]
            x1, x2 = torch.broadcast_tensors(x1, x2)
        return self._finish(N.PW_ADD, x1, x2)


class SubQ(LayerQ):
    def __init__(self, sub, gradient_based=True, act_quant=True, act_n_bits=8):
        super().__init__(gradient_based=gradient_based, act_quant=act_quant, act_n_bits=act_n_bits)
        if not isinstance(sub, Sub):
            raise Exception("SubQ wraps Sub, got %s" % type(sub))
        self.sub = sub

    def forward(self, x1, x2):
        return self._finish(N.PW_SUB, x1, x2)


class MulQ(LayerQ):
    def __init__(self, mul, gradient_based=True, act_quant=True, act_n_bits=8):
        super().__init__(gradient_based=gradient_based, act_quant=act_quant, act_n_bits=act_n_bits)
        if not isinstance(mul, Mul):
            raise Exception("MulQ wraps Mul, got %s" % type(mul))
        self.mul = mul

    def forward(self, x1, x2):
        if torch.is_tensor(x2) and x1.dim() == 4 and x2.dim() == 4 and x1.shape[1] == 1 and x2.shape[1] > 1:
            x1, x2 = x2, x1          # DPTNetQ multiplies (features[:, None], masks): the kernel broadcasts its SECOND operand
        if self.do_mac_op:
            self.mac_op = x1.numel()
        return self._finish(N.PW_MUL, x1, x2)


class Conv1dQ(LayerQ):
    def __init__(self, conv1d, gradient_based=True, weight_quant=True, act_quant=True, act_n_bits=8, weight_n_bits=8):
        super().__init__(gradient_based=gradient_based, weight_quant=weight_quant, act_quant=act_quant,
                         weight_shape=conv1d.weight.shape, act_n_bits=act_n_bits, weight_n_bits=weight_n_bits)
        if not isinstance(conv1d, nn.Conv1d):
            raise Exception("Conv1dQ wraps Conv1d, got %s" % type(conv1d))
        self.conv1d = conv1d

    def forward(self, x):
        y = _conv1d_q(self, x)
        self.calc_mac_op(x.shape)
        return self._finish(N.PW_IDENT, y)

    def calc_mac_op(self, x_shape):
        if self.do_mac_op:
            Co, Ci, k = self.conv1d.weight.shape
            self.mac_op = x_shape[0] * Co * _conv_out_len(self.conv1d, x_shape[-1]) * Ci * k


class Conv1dNlQ(LayerQ):
    def __init__(self, conv1d, nl, gradient_based=True, weight_quant=True, act_quant=True, act_n_bits=8, weight_n_bits=8):
        super().__init__(gradient_based=gradient_based, weight_quant=weight_quant, act_quant=act_quant,
                         weight_shape=conv1d.weight.shape, act_n_bits=act_n_bits, weight_n_bits=weight_n_bits)
        if not isinstance(conv1d, nn.Conv1d):
            raise Exception("Conv1dNlQ wraps Conv1d, got %s" % type(conv1d))
        self.conv1d = conv1d
        self.nl = nl

    def forward(self, x):
        y = _conv1d_q(self, x)
        self.calc_mac_op(x.shape)
        if isinstance(self.nl, _TORCH_NL):      # gates of the sequence models (tanh / sigmoid): torch evaluates the nl, the
            return self._finish(N.PW_IDENT, self.nl(y))     # quantiser runs on the library's kernel
        kind, slope = _nl_kind(self.nl)
        return self._finish(kind, y, slope=slope)

    calc_mac_op = Conv1dQ.calc_mac_op


class GroupNormQ(LayerQ):
    def __init__(self, groupnorm, gradient_based=True, act_quant=True, act_n_bits=8):
        super().__init__(gradient_based=gradient_based, act_quant=act_quant, act_n_bits=act_n_bits)
        if not isinstance(groupnorm, nn.GroupNorm):
            raise Exception("GroupNormQ wraps GroupNorm, got %s" % type(groupnorm))
        self.groupnorm = groupnorm

    def forward(self, x):
        gn = self.groupnorm
        if gn.num_groups != 1 or not gn.affine:
            raise NotImplementedError("only gLN = GroupNorm(1, C, affine=True) is on the ConvTasNet path")
        if self.do_mac_op:
            self.mac_op = 2 * x.numel()
        if x.dim() == 4:        # [B, C, K, S] chunked features of the dual-path models: per-sample statistics over (C, K, S)
            B, Cc, K, S = x.shape
            y = self._finish(N.PW_GLN, x.reshape(B, Cc, K * S), gamma=gn.weight, beta=gn.bias, eps=gn.eps)
            return y.reshape(B, Cc, K, S)
        return self._finish(N.PW_GLN, x, gamma=gn.weight, beta=gn.bias, eps=gn.eps)


class LayerNormQ(LayerQ):
    """nn.LayerNorm + FQ (qat_layers.py:455-468).  The one use on the scoped paths is ConvTasNetMusicQ's channel-wise cLN
    (convtasnetq_music.py:32-50): the module is handed `x.transpose(1, 2)` of an NCL tensor and normalises its last axis =
    the channels of every frame.  The kernel works on the NCL tensor itself (threads along frames), so the transposes stay
    views."""

    def __init__(self, layernorm, gradient_based=True, act_quant=True, act_n_bits=8):
        super().__init__(gradient_based=gradient_based, act_quant=act_quant, act_n_bits=act_n_bits)
        if not isinstance(layernorm, nn.LayerNorm):
            raise Exception("LayerNormQ wraps LayerNorm, got %s" % type(layernorm))
        self.layernorm = layernorm

    def forward(self, x):
        ln = self.layernorm
        if x.stride(-1) == 1 and x.shape[-1] > 1:
            # feature-last tensors of the sequence models (DPTNetQ / SepformerQ: [seq, batch, d]): torch normalises, the
            # quantiser runs on the library's kernel (see qat_layers_seq.py for the scope of that row)
            if self.do_mac_op:
                self.mac_op = 2 * x.numel()
            return self.activation_fake_quantize(F.layer_norm(x, ln.normalized_shape, ln.weight, ln.bias, ln.eps))
        if len(ln.normalized_shape) != 1 or not ln.elementwise_affine or ln.bias is None or x.dim() != 3 \
                or x.shape[-1] != ln.normalized_shape[0]:
            raise NotImplementedError("LayerNormQ: only LayerNorm(C) over the last axis of [B, M, C] (channel-wise cLN) has a kernel")
        if self.do_mac_op:
            self.mac_op = 2 * x.numel()
        y = ops.ChannelLayerNorm.apply(x.transpose(1, 2), ln.weight, ln.bias, ln.eps)      # [B, C, M]
        return self._finish(N.PW_IDENT, y).transpose(1, 2)


class NlQ(LayerQ):
    def __init__(self, nl, gradient_based=True, act_quant=True, act_n_bits=8):
        super().__init__(gradient_based=gradient_based, act_quant=act_quant, act_n_bits=act_n_bits)
        self.nl = nl

    def forward(self, x):
        if isinstance(self.nl, _TORCH_NL):
            return self._finish(N.PW_IDENT, self.nl(x))
        kind, slope = _nl_kind(self.nl)
        return self._finish(kind, x, slope=slope)


class Conv1dEncoderQ(LayerQ):
    """Analysis filterbank; with n_splitter >= 2 the conv is rebuilt with n_splitter input channels:
    channel 0 copies the float filters, the others are drawn around their mean (qat_layers.py:1009-1026)."""

    def __init__(self, encoder, n_splitter=1, gradient_based=True, weight_quant=True, act_quant=True, in_quant=False,
                 inout_nl_quant=False, act_n_bits=8, weight_n_bits=8, in_act_n_bits=8):
        super().__init__(gradient_based=gradient_based, weight_quant=weight_quant, act_quant=act_quant,
                         weight_shape=encoder[0].weight.shape, act_n_bits=act_n_bits, weight_n_bits=weight_n_bits)
        if not isinstance(encoder[0], nn.Conv1d):
            raise Exception("Conv1dEncoderQ wraps Conv1d, got %s" % type(encoder[0]))
        self.in_quantizer = (get_activation_quantizer(self.gradient_based, nl=inout_nl_quant, n_bits=in_act_n_bits)
                             if in_quant else nn.Identity())
        self.conv1d = encoder[0]
        self.nl = nn.Identity() if len(encoder) == 1 else encoder[1]
        if n_splitter >= 2:
            old = self.conv1d
            cin = old.in_channels
            grown = nn.Conv1d(n_splitter * cin, old.out_channels, old.kernel_size, stride=old.stride,
                              padding=old.padding, bias=old.bias is not None).to(old.weight.device)
            with torch.no_grad():
                w0 = old.weight.detach()
                w = w0.repeat(1, n_splitter, 1)
                for rep in range(1, n_splitter):
                    for c in range(cin):
                        col = w0[:, c, :]
                        # same RNG draw order as the reference: one randn_like per (rep, channel)
                        w[:, rep * cin + c, :] = torch.mean(col) + torch.randn_like(col) * (torch.std(col) ** rep)
                grown.weight.copy_(w)
                if old.bias is not None:
                    grown.bias.copy_(old.bias)
            self.conv1d = grown

    def forward(self, x):
        from .. import edge_engine as EE
        grid = getattr(x, "_fq_grid", None)          # the splitter's 8-bit grid (ops.split_input)
        if not isinstance(self.in_quantizer, nn.Identity):
            x = self.in_quantizer(x)
            q = self.in_quantizer
            grid = (q.min_range, q.max_range) if EE._steady_aq(q) else None
        if EE.framed_conv_eligible(self.conv1d, self.weight_fake_quantize, x, grid):
            # input on an 8-bit grid, 8-bit weights: one integer-code tcgen05 GEMM over the framed input
            y = EE.framed_conv(self.conv1d, self.weight_fake_quantize, x, grid)
        else:
            y = _conv1d(x, self.weight_fake_quantize(self.conv1d.weight), self.conv1d)
        if self.do_mac_op:
            Co, Ci, k = self.conv1d.weight.shape
            self.mac_op = x.shape[0] * Ci * Co * _conv_out_len(self.conv1d, x.shape[-1]) * k
        kind, slope = _nl_kind(self.nl)
        return self._finish(kind, y, slope=slope)


class ResidualErrorBlock(LayerQ):
    """RQB: re-encode the quantised output, quantise the feature-domain error, decode it with the
    SAME quantised decoder weight (qat_layers.py:1188-1202, ConvTranspose1d branch only)."""

    def __init__(self, decoder, gradient_based, weight_quant, act_quant, act_nl_quantizer=False, act_n_bits=8,
                 weight_n_bits=8, train_res_dec=False):
        super().__init__(gradient_based=gradient_based, act_quant=act_quant, act_nl_quantizer=act_nl_quantizer,
                         act_n_bits=act_n_bits)
        if decoder.bias is not None or (train_res_dec and type(decoder) is not nn.ConvTranspose1d):
            raise NotImplementedError("RQB is implemented for bias-free decoders (train_res_dec: ConvTranspose1d only)")
        self.train_res_dec = train_res_dec
        self.decoder_bias = None
        if type(decoder) is nn.Linear:       # ConvTasNetMusicQ: per-frame Linear decoder (qat_layers.py:1110-1121)
            self.decoder_type = nn.Linear
            self.residual_encoder = nn.Linear(decoder.out_features, decoder.in_features, bias=False).to(decoder.weight.device)
            self.weight_fake_quantize = (get_weight_quantizer(gradient_based, self.residual_encoder.weight.shape,
                                                              n_bits=weight_n_bits) if weight_quant else nn.Identity())
            return
        if type(decoder) is not nn.ConvTranspose1d:
            raise NotImplementedError("RQB is implemented for ConvTranspose1d and Linear decoders")
        if decoder.padding[0] or decoder.output_padding[0] or decoder.dilation[0] != 1:
            raise NotImplementedError("decoder geometry not on the ConvTasNet path")
        self.decoder_type = nn.ConvTranspose1d
        self.residual_encoder = nn.Conv1d(decoder.out_channels, decoder.in_channels, decoder.kernel_size,
                                          stride=decoder.stride, bias=False).to(decoder.weight.device)
        self.decoder_kernel, self.decoder_stride = decoder.kernel_size, decoder.stride
        self.decoder_padding, self.decoder_output_padding = decoder.padding, decoder.output_padding
        self.decoder_dilation, self.decoder_groups = decoder.dilation, decoder.groups
        self.decoder_in_channels, self.decoder_out_channels = decoder.in_channels, decoder.out_channels
        if train_res_dec:       # SepformerQ: the residual is decoded by its own trainable filterbank (qat_layers.py:1138-1147)
            self.residual_decoder = nn.ConvTranspose1d(decoder.in_channels, decoder.out_channels, decoder.kernel_size,
                                                       stride=decoder.stride, bias=False).to(decoder.weight.device)
            self.weight_fake_quantize_dec = (get_weight_quantizer(gradient_based, self.residual_decoder.weight.shape, ch_out_idx=1,
                                                                  n_bits=weight_n_bits) if weight_quant else nn.Identity())
        self.weight_fake_quantize = (get_weight_quantizer(gradient_based, self.residual_encoder.weight.shape,
                                                          n_bits=weight_n_bits) if weight_quant else nn.Identity())

    def decode_weight(self, dec, wq):
        """(raw weight, its quantiser) of the filterbank that decodes the residual: the block's own when train_res_dec, else
        the decoder's (`dec`, `wq`)."""
        if self.train_res_dec:
            return self.residual_decoder.weight, self.weight_fake_quantize_dec
        return dec.weight, wq

    def reencode(self, y_q):
        """Yq = residual_encoder(y_q) with the fake-quantised re-encoder weight (qat_layers.py:1189); y_q lies on the grid of
        the decoder's output quantiser, so the conv runs as an integer-code GEMM when that quantiser is in steady state."""
        from .. import edge_engine as EE
        conv, wq = self.residual_encoder, self.weight_fake_quantize
        src = getattr(y_q, "_fq_src", None)
        grid = (src.min_range, src.max_range) if EE._steady_aq(src) else None
        if EE.framed_conv_eligible(conv, wq, y_q, grid):
            return EE.framed_conv(conv, wq, y_q, grid)
        return ops.StridedConv.apply(y_q, wq(conv.weight), self.decoder_stride[0])

    def forward(self, Y, y_q, w_decoder):
        if self.decoder_type is nn.Linear:
            # channels-first: Y [R, N, K] features, y_q [R, F, K] quantised decoder output; the Linear layers are 1x1 convs
            Yq = _linear_1x1(y_q, self.residual_encoder.weight, self.weight_fake_quantize)
            Y1 = self._finish(N.PW_SUB, Y, Yq)
            if isinstance(w_decoder, tuple):      # (raw weight, its quantiser): LinearDecoderQ hands both over
                return _linear_1x1(Y1, w_decoder[0], w_decoder[1])
            return ops.Conv1x1.apply(Y1, w_decoder.unsqueeze(-1), None)
        Yq = self.reencode(y_q)
        Y1 = self._finish(N.PW_SUB, Y, Yq)
        if self.train_res_dec:
            w_decoder = self.weight_fake_quantize_dec(self.residual_decoder.weight)
        return ops.TransposedConv1.apply(Y1, w_decoder, self.decoder_stride[0])


class ConvTr1dDecoderQ(LayerQ):
    def __init__(self, decoder, n_combiner=1, gradient_based=True, weight_quant=True, weight_n_bits=8, act_quant=True,
                 act_n_bits=8, inout_nl_quant=False, out_quant=True, out_act_n_bits=8, train_res_dec=False):
        super().__init__(gradient_based=gradient_based, weight_quant=weight_quant, act_quant=out_quant,
                         act_nl_quantizer=inout_nl_quant, weight_shape=decoder[0].weight.shape, ch_out_idx=1,
                         act_n_bits=out_act_n_bits, weight_n_bits=weight_n_bits)
        dec = decoder[0]
        if not isinstance(dec, nn.ConvTranspose1d):
            raise Exception("ConvTr1dDecoderQ wraps ConvTranspose1d, got %s" % type(dec))
        if dec.out_channels != 1 or dec.bias is not None or dec.padding[0] or dec.output_padding[0] or dec.groups != 1:
            raise NotImplementedError("decoder geometry not on the ConvTasNet path")
        self.n_combiner = n_combiner
        self.convTr1d = dec
        if self.n_combiner >= 2:
            self.residual_error_block = ResidualErrorBlock(dec, gradient_based, weight_quant=weight_quant,
                                                           act_quant=act_quant, weight_n_bits=weight_n_bits,
                                                           act_n_bits=act_n_bits, train_res_dec=train_res_dec)
            self.activation_fake_quantize_residual = (get_activation_quantizer(gradient_based, n_bits=out_act_n_bits)
                                                      if out_quant else nn.Identity())

    def forward(self, x):
        from .. import edge_engine as EE
        if EE.decoder_eligible(self, x):       # features on an 8-bit grid, 8-bit weights: integer-code tcgen05 GEMMs + overlap-add
            return EE.decoder_forward(self, x)
        stride = self.convTr1d.stride[0]
        w_dec = self.weight_fake_quantize(self.convTr1d.weight)
        x_dec = x
        if self.n_combiner >= 2:      # x also feeds the residual block: sum the two gradients in the library
            x_dec, x = ops.fanout2(x)
        y = self._finish(N.PW_IDENT, ops.TransposedConv1.apply(x_dec, w_dec, stride))
        if self.do_mac_op:
            Ci, Co, k = self.convTr1d.weight.shape
            self.mac_op = x.shape[0] * Co * Ci * ((x.shape[-1] - 1) * stride + k) * (k // stride)
        if self.n_combiner == 1:
            return y
        outs = [y]
        for _ in range(1, self.n_combiner):
            x = self.residual_error_block(x, y, w_dec)
            y = self._finish(N.PW_IDENT, x, quantizer=self.activation_fake_quantize_residual)
            outs.append(y)
        return torch.stack(outs)


class LinearDecoderQ(LayerQ):
    """Per-frame nn.Linear decoder + out-FQ (+ RQB) of ConvTasNetMusicQ (qat_layers.py:1256-1302).  `forward` keeps the
    reference's contract ([..., K, N] -> [n_combiner, ..., K, F]); the arithmetic runs channels-first (`forward_ncl`:
    [R, N, K] -> [n_combiner, R, F, K]), where the Linear layers are 1x1 convolutions and every transpose is a view."""

    def __init__(self, decoder, n_combiner=1, gradient_based=True, weight_quant=True, weight_n_bits=8, act_quant=True,
                 inout_nl_quant=False, act_n_bits=8, out_quant=True, out_act_n_bits=8, train_res_dec=False):
        super().__init__(gradient_based=gradient_based, weight_quant=weight_quant, act_quant=out_quant,
                         act_nl_quantizer=inout_nl_quant, weight_shape=decoder[0].weight.shape,
                         act_n_bits=out_act_n_bits, weight_n_bits=weight_n_bits)
        if not isinstance(decoder[0], nn.Linear):
            raise Exception("LinearDecoderQ wraps Linear, got %s" % type(decoder[0]))
        if decoder[0].bias is not None:
            raise NotImplementedError("decoder bias is not on the ConvTasNetMusic path")
        self.linear = decoder[0]
        self.n_combiner = n_combiner
        if self.n_combiner >= 2:
            self.residual_error_block = ResidualErrorBlock(self.linear, gradient_based, weight_quant, act_quant,
                                                           act_n_bits=act_n_bits, weight_n_bits=weight_n_bits,
                                                           train_res_dec=train_res_dec)
            self.activation_fake_quantize_residual = (get_activation_quantizer(gradient_based, n_bits=out_act_n_bits)
                                                      if out_quant else nn.Identity())

    def forward_ncl(self, x):
        from .. import tcn_engine as E
        src = getattr(x, "_fq_src", None)
        tc = TENSOR_CORE_CONV1X1 and src is not None and E.code_linear_eligible(self.linear.weight, self.weight_fake_quantize, src, x)
        # per-layer route: the weight quantiser is applied ONCE per forward and its output shared with the RQB, as in the
        # reference (its observer counts calls); tensor-core route: raw weight + quantiser, the codes are derived in the GEMM prep
        w_dec = (self.linear.weight, self.weight_fake_quantize) if tc else self.weight_fake_quantize(self.linear.weight)
        x_dec = x
        if self.n_combiner >= 2:      # x also feeds the residual block: sum the two gradients in the library
            x_dec, x = ops.fanout2(x)
        if tc or isinstance(self.weight_fake_quantize, nn.Identity):
            y_lin = _linear_1x1(x_dec, self.linear.weight, self.weight_fake_quantize)
        else:
            y_lin = ops.Conv1x1.apply(x_dec, w_dec.unsqueeze(-1), None)
        y = self._finish(N.PW_IDENT, y_lin)
        if self.do_mac_op:
            self.mac_op = x.numel() * self.linear.weight.shape[0]
        if self.n_combiner == 1:
            return y.unsqueeze(0)
        outs = [y]
        for _ in range(1, self.n_combiner):
            x = self.residual_error_block(x, y, w_dec)
            y = self._finish(N.PW_IDENT, x, quantizer=self.activation_fake_quantize_residual)
            outs.append(y)
        return torch.stack(outs)

    def forward(self, x):
        lead = x.shape[:-2]
        K, Nf = x.shape[-2], x.shape[-1]
        out = self.forward_ncl(x.transpose(-1, -2).reshape(-1, Nf, K))           # [n, R, F, K]
        out = out.reshape((out.shape[0],) + tuple(lead) + (out.shape[-2], K)).transpose(-1, -2)
        return out if self.n_combiner >= 2 else out[0]


# sequence-model layers (DPTNetQ / SepformerQ) live in qat_layers_seq.py; the reference keeps them in this module, so they are
# resolved lazily here (qat_layers_seq imports LayerQ from this module).  Any other reference name outside the scoped paths
# (imported by the reference's remaining model files) is an importable placeholder that raises NotImplementedError when used --
# see fqss_b200/shim.py
from ..shim import module_getattr as _module_getattr  # noqa: E402

_SEQ_NAMES = ("Const", "Div", "ConstQ", "DivQ", "LinearQ", "LinearNlQ", "Conv2dQ", "Conv2dNlQ", "LSTMQ", "MultiheadAttentionQ")
_placeholder = _module_getattr("quantization.qat.qat_layers")


def __getattr__(name):
    if name in _SEQ_NAMES:
        from . import qat_layers_seq as QS
        return getattr(QS, name)
    return _placeholder(name)
