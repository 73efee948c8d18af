"""B200 mirror of `quantization/qat/models/sepformerq.py` (BASELINE configs[3]): SepFormer -- a
dual-path transformer separator between a ConvTasNet-style filterbank pair -- with the FQSS hooks (splitter / combiner,
`quantize_model`, `load_pretrain`) and the reference's module tree, so its checkpoints load with strict=True.

Structure (reference lines in brackets): Conv1d encoder + ReLU [:365-370]; mask generator [:178-339]: global LayerNorm,
1x1 conv, 50 %-overlapping chunks, `n_repeats` dual-path blocks [:126-175] each made of an intra-chunk and an inter-chunk
transformer block [:98-123] (sinusoidal positions added once, 8 pre-norm layers of self attention + 2-layer FFN [:50-95], final
LayerNorm) followed by a GroupNorm and a skip `Add`; PReLU, 1x1 Conv2d to the speakers, overlap-add, tanh x sigmoid gate,
1x1 conv + ReLU; masks x features; ConvTranspose1d decoder whose RQB decodes the residual with its own filterbank
(`train_res_dec`, [:492-501]).

What runs where: as DPTNetQ -- quantisers, observers and their backward on libfqss_sm100; encoder / decoder / RQB, the
GroupNorms, gates, `Add` / `Mul` on the per-layer kernels of the ConvTasNet path (the decoder on the integer-code tcgen05
GEMMs when the filter count allows); attention / FFN float math on torch (qat_layers_seq.py)."""
import math
import os

import torch
import torch.nn as nn

from ...process import postprocess, preprocess
from ..qat_layers import Add, Const, Mul
from ..qat_utils import quantize_modules, replace_decoderq, replace_encoderq

EPS_T = 1e-6
EPS = 1e-8


class PositionalEncoding(nn.Module):
    """Absolute sinusoidal positions: pe[p, 2i] = sin(p / 10000^(2i/d)), pe[p, 2i+1] = cos(.)  (a buffer, no parameters)."""

    def __init__(self, input_size, max_len=2500, device="cpu"):
        super().__init__()
        self.max_len = max_len
        pos = torch.arange(0, max_len).unsqueeze(1).float()
        freq = torch.exp(torch.arange(0, input_size, 2).float() * -(math.log(10000.0) / input_size))
        pe = torch.zeros(max_len, input_size, requires_grad=False, device=device)
        pe[:, 0::2] = torch.sin(pos * freq)
        pe[:, 1::2] = torch.cos(pos * freq)
        self.register_buffer("pe", pe.unsqueeze(0))
        self.const = Const()

    def forward(self, x):                                  # x: [batch, time, features]
        return self.const(self.pe[:, : x.size(1)].clone().detach())


class TransformerLayer(nn.Module):
    """Pre-norm transformer layer; the residual sums are plain (un-quantised) adds."""

    def __init__(self, n_filters, n_ffn, n_heads, dropout=0.0):
        super().__init__()
        self.mha = nn.MultiheadAttention(n_filters, n_heads, dropout=dropout, batch_first=False)
        self.ffn = nn.Sequential(nn.Linear(n_filters, n_ffn), nn.ReLU(), nn.Dropout(dropout), nn.Linear(n_ffn, n_filters))
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(n_filters, eps=EPS_T)
        self.norm2 = nn.LayerNorm(n_filters, eps=EPS_T)

    def forward(self, x):                                  # [batch, time, features]
        q = self.norm1(x).permute(1, 0, 2)
        x = x + self.dropout1(self.mha(q, q, q)[0].permute(1, 0, 2))
        h = self.ffn(self.norm2(x).permute(1, 0, 2)).permute(1, 0, 2)
        return x + self.dropout2(h)


class TransformerBlock(nn.Module):
    def __init__(self, n_filters, n_heads, n_ffn, num_layers=8, dropout=0.0, device="cpu"):
        super().__init__()
        self.layers = nn.ModuleList()
        for _ in range(num_layers):
            self.layers.append(TransformerLayer(n_filters, n_heads=n_heads, n_ffn=n_ffn, dropout=dropout))
        self.norm = nn.LayerNorm(n_filters, eps=EPS_T)
        self.pos = PositionalEncoding(n_filters, device=device)
        self.pos_add = Add()

    def forward(self, x):
        h = self.pos_add(x, self.pos(x))
        for layer in self.layers:
            h = layer(h)
        return self.norm(h)


class DualPathBlock(nn.Module):
    def __init__(self, n_filters, n_heads, n_ffn, dropout=0.0, device="cpu"):
        super().__init__()
        self.intra_transformer_block = TransformerBlock(n_filters=n_filters, n_heads=n_heads, n_ffn=n_ffn, dropout=dropout, device=device)
        self.inter_transformer_block = TransformerBlock(n_filters=n_filters, n_heads=n_heads, n_ffn=n_ffn, dropout=dropout, device=device)
        self.intra_norm = nn.GroupNorm(num_groups=1, num_channels=n_filters, eps=EPS)
        self.inter_norm = nn.GroupNorm(num_groups=1, num_channels=n_filters, eps=EPS)
        self.intra_add = Add()
        self.inter_add = Add()

    def forward(self, x):
        B, Fn, K, S = x.shape
        a = self.intra_transformer_block(x.permute(0, 3, 2, 1).contiguous().reshape(B * S, K, Fn))
        a = self.intra_norm(a.reshape(B, S, K, Fn).permute(0, 3, 2, 1).contiguous())
        a = self.intra_add(a, x)
        e = self.inter_transformer_block(a.permute(0, 2, 3, 1).contiguous().reshape(B * K, S, Fn))
        e = self.inter_norm(e.reshape(B, K, S, Fn).permute(0, 3, 1, 2).contiguous())
        return self.inter_add(e, a)


class MaskGenerator(nn.Module):
    def __init__(self, n_srcs: int, n_filters: int, n_repeats: int = 2, n_heads: int = 8, chunk_size: int = 250,
                 n_ffn: int = 1024, dropout: float = 0.0, device: str = "cpu"):
        super().__init__()
        self.n_srcs = n_srcs
        self.chunk_size = chunk_size
        self.norm = nn.GroupNorm(num_groups=1, num_channels=n_filters, eps=EPS)
        self.conv1d = nn.Conv1d(n_filters, n_filters, 1, bias=False)
        self.layers = nn.ModuleList([])
        for _ in range(n_repeats):
            self.layers.append(DualPathBlock(n_filters, n_heads=n_heads, n_ffn=n_ffn, dropout=dropout, device=device))
        self.conv2d = nn.Conv2d(n_filters, n_srcs * n_filters, kernel_size=1, bias=True)
        self.end_conv = nn.Sequential(nn.Conv1d(n_filters, n_filters, 1, bias=False), nn.ReLU())
        self.prelu = nn.PReLU()
        self.net_out = nn.Sequential(nn.Conv1d(n_filters, n_filters, 1, bias=True), nn.Tanh())
        self.net_gate = nn.Sequential(nn.Conv1d(n_filters, n_filters, 1, bias=True), nn.Sigmoid())
        self.mul = Mul()

    def padding(self, x, K):
        """Zero-pad [B, N, L] on the right to a whole number of half-overlapping chunks, plus K/2 on both sides."""
        P = K // 2
        gap = K - (P + x.shape[2] % K) % K
        return nn.functional.pad(x, (P, gap + P)), gap

    def segmentation(self, x, K):
        """[B, N, L] -> ([B, N, K, S], gap): chunk s covers padded frames [s K/2, s K/2 + K)."""
        x, gap = self.padding(x, K)
        return x.unfold(2, K, K // 2).transpose(2, 3).contiguous(), gap

    def over_add(self, x, gap):
        """Inverse layout of `segmentation`: even and odd chunks each tile the padded axis; their (plain) sum, trimmed."""
        B, N, K, S = x.shape
        P = K // 2
        c = x.transpose(2, 3)
        out = c[:, :, 0::2].reshape(B, N, -1)[:, :, P:] + c[:, :, 1::2].reshape(B, N, -1)[:, :, :-P]
        return out[:, :, :-gap] if gap > 0 else out

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, Fn, _ = x.shape
        seg, gap = self.segmentation(self.conv1d(self.norm(x)), self.chunk_size)
        for layer in self.layers:
            seg = layer(seg)
        y = self.conv2d(self.prelu(seg))
        y = y.reshape(B * self.n_srcs, -1, self.chunk_size, y.shape[-1])
        y = self.over_add(y, gap)
        out = self.end_conv(self.mul(self.net_out(y), self.net_gate(y)))          # [B*S, F, L]
        return out.reshape(B, self.n_srcs, Fn, out.shape[-1])


class SepformerQ(nn.Module):
    def __init__(self, n_spks: int = 1, kernel_size: int = 16, stride: int = 8, n_filters: int = 256, n_repeats: int = 2,
                 n_heads: int = 8, chunk_size: int = 250, device: str = "cpu"):
        super().__init__()
        self.n_srcs = n_spks
        self.enc_num_feats = n_filters
        self.set_splitter_combiner(1, 1)
        self.encoder = nn.Sequential(nn.Conv1d(1, n_filters, kernel_size=kernel_size, stride=stride, padding=0, bias=False), nn.ReLU())
        self.masker = MaskGenerator(n_spks, n_filters, n_repeats=n_repeats, n_heads=n_heads, chunk_size=chunk_size, device=device)
        self.decoder = nn.ConvTranspose1d(n_filters, 1, kernel_size=kernel_size, stride=stride, padding=0, bias=False)
        self.mul = Mul()

    def pre_process(self, x):
        return preprocess(x, n_splitter=self.n_splitter)

    def post_process(self, x):
        return postprocess(x, n_combiner=self.n_combiner)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.pre_process(x)
        B = x.shape[0]
        feats = self.encoder(x)                                                   # [B, F, M]
        masked = self.mul(self.masker(feats), feats.unsqueeze(1))                 # [B, S, F, M]
        dec_in = masked.reshape(B * self.n_srcs, self.enc_num_feats, -1)
        src = getattr(masked, "_fq_src", None)
        if src is not None:       # the reshape drops the producer tag the decoder's integer-code path looks for
            dec_in._fq_src = src
        out = self.decoder(dec_in).reshape((self.n_combiner, B, self.n_srcs, 1, -1))
        return self.post_process(out)

    def load_pretrain(self, weights_path):
        """A single checkpoint file (positional key matching, fmodel.* dropped) or a speechbrain directory with encoder.ckpt /
        masknet.ckpt / decoder.ckpt (sepformerq.py:439-464)."""
        own = self.state_dict()
        if os.path.isfile(weights_path):
            src = torch.load(weights_path)
            src = src.get("state_dict", src)
            src = {k: v for k, v in src.items() if not k.startswith("fmodel.")}
            assert len(own) == len(src), \
                "Error: mismatch models weights. Please check if the model configurations match to model weights!"
            own = {mine: src[theirs] for mine, theirs in zip(own.keys(), src.keys())}
        else:
            own["encoder.0.weight"] = torch.load(os.path.join(weights_path, "encoder.ckpt"))["conv1d.weight"]
            mask = torch.load(os.path.join(weights_path, "masknet.ckpt"))
            for mine, theirs in zip(self.masker.state_dict().keys(), mask):
                own["masker." + mine] = mask.get(theirs)
            own["decoder.weight"] = torch.load(os.path.join(weights_path, "decoder.ckpt"))["weight"]
        self.load_state_dict(own, strict=True)

    def set_splitter_combiner(self, n_splitter, n_combiner):
        self.n_splitter = n_splitter
        self.n_combiner = n_combiner

    def quantize_model(self, gradient_based=True, weight_quant=True, weight_n_bits=8, act_quant=True, act_n_bits=8,
                       inout_nl_quant=False, in_quant=False, in_act_n_bits=8, out_quant=True, out_act_n_bits=8):
        p = dict(gradient_based=gradient_based, act_quant=act_quant, weight_quant=weight_quant,
                 weight_n_bits=weight_n_bits, act_n_bits=act_n_bits)
        edge = dict(p, inout_nl_quant=inout_nl_quant)
        for _, m in list(self.named_modules()):          # snapshot: children are replaced while we walk
            if type(m) is SepformerQ:
                replace_encoderq(m.encoder, ["0", "1"], dict(edge, n_splitter=self.n_splitter, in_quant=in_quant,
                                                             in_act_n_bits=in_act_n_bits))
                replace_decoderq(m, ["decoder"], dict(edge, n_combiner=self.n_combiner, out_quant=out_quant,
                                                      out_act_n_bits=out_act_n_bits, train_res_dec=True))
                quantize_modules(m, ["mul"], p)
            elif type(m) is TransformerBlock:
                quantize_modules(m, ["norm"], p)
                quantize_modules(m, ["pos_add"], p)
                quantize_modules(m.pos, ["const"], p)
            elif type(m) is TransformerLayer:
                for name in ("norm1", "norm2", "mha"):
                    quantize_modules(m, [name], p)
                for name in ("0", "1", "3"):
                    quantize_modules(m.ffn, [name], p)
            elif type(m) is DualPathBlock:
                for name in ("inter_norm", "intra_norm", "inter_add", "intra_add"):
                    quantize_modules(m, [name], p)
            elif type(m) is MaskGenerator:
                quantize_modules(m.net_out, ["0", "1"], p)
                quantize_modules(m.net_gate, ["0", "1"], p)
                quantize_modules(m, ["norm"], p)
                quantize_modules(m, ["conv1d"], p)
                quantize_modules(m, ["conv2d"], p)
                quantize_modules(m.end_conv, ["0", "1"], p)
                quantize_modules(m, ["prelu"], p)
                quantize_modules(m, ["mul"], p)
