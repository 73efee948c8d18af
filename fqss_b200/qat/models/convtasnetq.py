"""B200 mirror of `quantization/qat/models/convtasnetq.py`.

Same topology, constructor signatures, attribute names and state_dict keys as the reference
(ConvBlock :11-42, MaskGenerator :45-115, ConvTasNetQ :118-288).  Before `quantize_model` the
graph is the float model (used as the KD teacher); afterwards every child is a wrapper from
`fqss_b200.qat.qat_layers` running sm_100a kernels.  When the quantised model is in steady state
(observers off) `MaskGenerator.forward` hands the 24 TCN blocks to the fused engine
(`fqss_b200.tcn_engine`) instead of walking the per-layer wrappers.
"""
from typing import Optional, Tuple

import torch
import torch.nn as nn

from ... import ops
from ...process import postprocess, preprocess
from ..qat_layers import Add, Mul
from ..qat_utils import quantize_modules, replace_decoderq, replace_encoderq

EPS = 1e-8


class ConvBlock(nn.Module):
    """1x1 expand -> PReLU -> gLN -> depthwise dilated conv -> PReLU -> gLN -> {res, skip} 1x1."""

    def __init__(self, io_channels: int, hidden_channels: int, kernel_size: int, padding: int, dilation: int = 1):
        super().__init__()
        self.shared_block = nn.Sequential(
            nn.Conv1d(io_channels, hidden_channels, 1),
            nn.PReLU(),
            nn.GroupNorm(1, hidden_channels, eps=EPS),
            nn.Conv1d(hidden_channels, hidden_channels, kernel_size, padding=padding, dilation=dilation,
                      groups=hidden_channels),
            nn.PReLU(),
            nn.GroupNorm(1, hidden_channels, eps=EPS),
        )
        self.res_conv = nn.Conv1d(hidden_channels, io_channels, 1)
        self.skip_conv = nn.Conv1d(hidden_channels, io_channels, 1)
        self.add = Add()

    def forward(self, x: torch.Tensor) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
        h = self.shared_block(x)
        res = self.res_conv(h)
        skip = self.skip_conv(h)
        return self.add(x, res), skip


class MaskGenerator(nn.Module):
    """TCN separation module: bottleneck, num_stacks x num_layers ConvBlocks, mask head."""

    def __init__(self, input_dim: int, n_srcs: int, kernel_size: int, num_feats: int, num_hidden: int, num_layers: int,
                 num_stacks: int, msk_activate: str):
        super().__init__()
        self.input_dim = input_dim
        self.n_srcs = n_srcs
        self.bottleneck = nn.Sequential(nn.GroupNorm(1, input_dim, eps=EPS), nn.Conv1d(input_dim, num_feats, 1))
        self.receptive_field = 0
        self.TCN = nn.ModuleList()
        for s in range(num_stacks):
            for layer in range(num_layers):
                d = 2 ** layer
                self.TCN.append(ConvBlock(num_feats, num_hidden, kernel_size, dilation=d, padding=d))
                self.receptive_field += kernel_size if (s == 0 and layer == 0) else (kernel_size - 1) * d
        self.adds = nn.ModuleList([Add() for _ in range(len(self.TCN) - 1)])
        if msk_activate == "sigmoid":
            act = nn.Sigmoid()
        elif msk_activate == "relu":
            act = nn.ReLU()
        else:
            raise ValueError(f"Unsupported activation {msk_activate}")
        self.mask_net = nn.Sequential(nn.PReLU(), nn.Conv1d(num_feats, input_dim * n_srcs, 1), act)

    use_fused = True      # class-level switch: False forces the per-layer wrappers (tests / debugging)

    def _skip_total(self, x, E):
        """bottleneck + TCN stack -> running skip sum (the input of the mask head)."""
        feats = self._conv_after(self.bottleneck[0], self.bottleneck[1], x, E)
        if self.use_fused and E.fused_eligible(self, feats):
            q = self.bottleneck[1].activation_fake_quantize
            adds = [None] + list(self.adds)
            _, total = E.fused_tcn(feats, list(self.TCN), adds, True, (q.min_range, q.max_range))
        else:
            feats, total = self.TCN[0](feats)
            for i, block in enumerate(self.TCN[1:]):
                feats, skip = block(feats)
                total = self.adds[i](total, skip)
        return total

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        batch = x.shape[0]
        from ... import tcn_engine as E
        total = self._skip_total(x, E)
        out = self._conv_after(self.mask_net[0], self.mask_net[1], total, E)
        out = self.mask_net[2](out)          # Identity after quantize_model (the ReLU moved into mask_net[1])
        return out.reshape(batch, self.n_srcs, self.input_dim, -1)

    def masked_features(self, x: torch.Tensor, feats: torch.Tensor, mul_layer, want_codes=False) -> Optional[torch.Tensor]:
        """`mul_layer(self(x), feats.unsqueeze(1))` with the mask head (mask conv + ReLU + FQ, x features + FQ) fused into
        the mask GEMM's epilogue; None when the layers are not in the quantised steady state the fused kernel covers
        (the caller then composes the modules as the reference does, convtasnetq.py:202-203)."""
        from ... import tcn_engine as E
        if not self.use_fused or not isinstance(self.mask_net[2], nn.Identity):
            return None
        first, conv_layer = self.mask_net[0], self.mask_net[1]
        q_in = getattr(first, "activation_fake_quantize", None)
        # cheap structural checks first: nothing is computed unless the fused path will be taken
        if q_in is None or not E.mask_head_eligible(conv_layer, q_in, mul_layer, _Meta(x, self.bottleneck[1].conv1d.out_channels), feats):
            return None
        total = self._skip_total(x, E)
        h = first(total)
        out, codes = E.mask_head(conv_layer, q_in, mul_layer, h, feats, want_codes)
        conv_layer.calc_mac_op(h.shape)
        if mul_layer.do_mac_op:
            mul_layer.mac_op = out.numel()
        res = out.reshape(x.shape[0], self.n_srcs, self.input_dim, -1)
        res._fq_src = mul_layer.activation_fake_quantize     # the masked features lie on the MulQ quantiser's grid ...
        res._fq_codes = codes                                # ... and exist as integer codes too ([B, S*F, ld] bf16, or None)
        return res

    def _conv_after(self, first, conv_layer, x, E):
        """conv_layer(first(x)); when `first` ends in an 8-bit activation quantiser and `conv_layer` is a quantised 1x1
        conv in steady state, the conv runs on the tensor cores with integer-code operands."""
        h = first(x)
        q_in = getattr(first, "activation_fake_quantize", None)
        if self.use_fused and q_in is not None and E.code_conv_eligible(conv_layer, q_in, h):
            from ..qat_layers import _nl_kind
            y = E.code_conv(conv_layer, q_in, h)
            conv_layer.calc_mac_op(h.shape)
            kind, slope = _nl_kind(getattr(conv_layer, "nl", None))
            return conv_layer._finish(kind, y, slope=slope)
        return conv_layer(h)


class _Meta:
    """Shape / placement stand-in for a tensor that has not been computed yet (eligibility checks only)."""

    def __init__(self, like, channels):
        self.is_cuda, self.dtype, self.shape = like.is_cuda, like.dtype, (like.shape[0], channels, like.shape[2])

    def dim(self):
        return 3


class ConvTasNetQ(nn.Module):
    """Conv-TasNet (non-causal) with the FQSS splitter / RQB hooks; `quantize_model` rewrites it."""

    def __init__(self, n_spks: int = 1, kernel_size: int = 32, stride: int = 16, n_filters: int = 512,
                 mask_kernel_size: int = 3, bn_chan: int = 128, hid_chan: int = 512, n_blocks: int = 8, n_repeats: int = 3,
                 mask_act: str = "relu"):
        super().__init__()
        self.n_srcs = n_spks
        self.enc_num_feats = n_filters
        self.set_splitter_combiner(1, 1)
        self.encoder = nn.Conv1d(1, n_filters, kernel_size, stride=stride, padding=0, bias=False)
        self.masker = MaskGenerator(input_dim=n_filters, n_srcs=n_spks, kernel_size=mask_kernel_size, num_feats=bn_chan,
                                    num_hidden=hid_chan, num_layers=n_blocks, num_stacks=n_repeats, msk_activate=mask_act)
        self.decoder = nn.ConvTranspose1d(n_filters, 1, kernel_size, stride=stride, padding=0, bias=False)
        self.mul = Mul()

    def pre_process(self, x):
        return preprocess(x, n_splitter=self.n_splitter)

    def post_process(self, x):
        return postprocess(x, n_combiner=self.n_combiner)

    use_float_engine = True   # class-level switch: False keeps the un-quantised model on plain torch modules

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.use_float_engine and x.is_cuda and type(self.encoder) is nn.Conv1d:
            from ... import float_engine as FE          # float (teacher) inference on the sm_100a kernels
            if FE.eligible(self, x):
                return FE.forward(self, x)
        x = self.pre_process(x)                                    # [B, n_splitter, T]
        batch = x.shape[0]
        feats = self.encoder(x)                                    # [B, F, M]
        f_mask, f_mul = ops.fanout2(feats)                         # two consumers: gradients summed by the library
        from ... import edge_engine as EE
        masked = self.masker.masked_features(f_mask, f_mul, self.mul,   # fused mask head; None: not applicable
                                             want_codes=EE.decoder_wants_codes(self.decoder))
        if masked is None:
            masked = self.mul(self.masker(f_mask), f_mul.unsqueeze(1))  # [B, S, F, M]
        dec_in = masked.reshape(batch * self.n_srcs, self.enc_num_feats, -1)
        src, codes = getattr(masked, "_fq_src", None), getattr(masked, "_fq_codes", None)
        if src is not None:          # a reshape drops the producer tag; the decoder's tensor-core path needs it
            dec_in._fq_src = src
            if codes is not None:
                dec_in._fq_codes = codes.view(batch * self.n_srcs, self.enc_num_feats, codes.shape[-1])
        dec = self.decoder(dec_in)                                 # [n_combiner, B*S, 1, T] (or [B*S,1,T])
        dec = dec.reshape((self.n_combiner, batch, self.n_srcs, 1, -1))
        return self.post_process(dec)                              # [B, S, T]

    def load_pretrain(self, weights_path):
        """Positional key matching for checkpoints with foreign key names (convtasnetq.py:225-237)."""
        own = self.state_dict()
        src = torch.load(weights_path)
        src = src.get("state_dict", src)
        src = {k: v for k, v in src.items() if not k.startswith("fmodel.")}
        assert len(own) == len(src), \
            "Error: mismatch models weights. Please check if the model configurations match to model weights!"
        self.load_state_dict({mine: src[theirs] for mine, theirs in zip(own.keys(), src.keys())}, strict=True)

    def set_splitter_combiner(self, n_splitter, n_combiner):
        self.n_splitter = n_splitter
        self.n_combiner = n_combiner

    def quantize_model(self, gradient_based=True, weight_quant=True, weight_n_bits=8, act_quant=True, act_n_bits=8,
                       inout_nl_quant=False, in_quant=False, in_act_n_bits=8, out_quant=True, out_act_n_bits=8):
        p = dict(gradient_based=gradient_based, act_quant=act_quant, weight_quant=weight_quant,
                 weight_n_bits=weight_n_bits, act_n_bits=act_n_bits)
        edge = dict(p, inout_nl_quant=inout_nl_quant)
        # snapshot first: the surgery mutates the tree while the reference walks named_modules() lazily,
        # visiting ConvTasNetQ, then the MaskGenerator, then each ConvBlock -- keep that RNG/creation order.
        for _, m in list(self.named_modules()):
            if type(m) is ConvTasNetQ:
                replace_encoderq(m, ["encoder"], dict(edge, n_splitter=self.n_splitter, in_quant=in_quant,
                                                      in_act_n_bits=in_act_n_bits))
                replace_decoderq(m, ["decoder"], dict(edge, n_combiner=self.n_combiner, act_n_bits=out_act_n_bits,
                                                      out_quant=out_quant, out_act_n_bits=out_act_n_bits))
                quantize_modules(m, ["mul"], p)
            elif type(m) is ConvBlock:
                for group in (["0", "1"], ["2"], ["3", "4"], ["5"]):
                    quantize_modules(m.shared_block, group, p)
                for name in ("res_conv", "skip_conv", "add"):
                    quantize_modules(m, [name], p)
            elif type(m) is MaskGenerator:
                quantize_modules(m.bottleneck, ["0"], p)
                quantize_modules(m.bottleneck, ["1"], p)
                quantize_modules(m.mask_net, ["0"], p)
                quantize_modules(m.mask_net, ["1", "2"], p)
                for i in range(len(m.adds)):
                    quantize_modules(m.adds, [str(i)], p)
