"""B200 mirror of `quantization/qat/models/dptnetq.py` (BASELINE configs[2]): the dual-path
transformer separator DPTNet with the FQSS hooks -- splitter / combiner, `quantize_model`, `load_pretrain` -- and the
reference's module tree (so a reference checkpoint loads with strict=True and vice versa).

Architecture (reference lines in brackets): 2-tap / stride-1 Conv1d encoder + ReLU [:104-125], global LayerNorm,
1x1 bottleneck, 50 %-overlapping chunks of `segment_size` frames [:214-259], `layer` pairs of (intra-chunk, inter-chunk)
improved-transformer layers -- 4-head self attention, add & norm, bidirectional LSTM -> ReLU -> Linear in place of the FFN,
add & norm [:57-97, :153-204] --, PReLU + 1x1 Conv2d to `n_spks` feature maps, overlap-add of the chunks, tanh x sigmoid
gate, 1x1 mask conv + ReLU, mask x encoder output, Linear decoder + overlap-add [:127-138].

What runs where: the quantisers (weights per output row, activations per tensor), observers and their backward are
libfqss_sm100 kernels; encoder, bottleneck norm, gates, mask conv, `Mul`/`Add` and the Linear decoder + RQB go through the
same per-layer wrappers as ConvTasNetQ; the float math of attention / LSTM / Linear is torch (qat_layers_seq.py).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ...process import postprocess, preprocess
from ..qat_layers import Add, Mul
from ..qat_utils import quantize_modules, replace_decoderq, replace_encoderq


def overlap_and_add(signal, frame_step):
    """[..., frames, frame_length] -> [..., (frames - 1) * frame_step + frame_length]: frame f added at offset f * frame_step
    (dptnetq.py:17-55), as one `fold`."""
    lead = signal.shape[:-2]
    frames, flen = signal.shape[-2:]
    T = (frames - 1) * frame_step + flen
    cols = signal.reshape(-1, frames, flen).transpose(1, 2)                 # [N, flen, frames]
    out = F.fold(cols, output_size=(1, T), kernel_size=(1, flen), stride=(1, frame_step))
    return out.reshape(*lead, T)


class TransformerEncoderLayer(nn.Module):
    """Improved transformer layer of DPTNet: self attention, add & norm, BiLSTM -> ReLU -> Linear, add & norm."""

    def __init__(self, d_model, nhead, hidden_size, dim_feedforward, dropout, activation="relu"):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.lstm = nn.LSTM(d_model, hidden_size, 1, bidirectional=True)
        self.dropout = nn.Dropout(dropout)
        self.linear = nn.Linear(hidden_size * 2, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.add_norm1 = Add()
        self.add_norm2 = Add()
        if activation not in ("relu", "gelu"):
            raise RuntimeError("activation should be relu/gelu, not {}".format(activation))
        self.activation = F.relu if activation == "relu" else F.gelu

    def forward(self, src):
        att = self.self_attn(src, src, src)[0]
        src = self.norm1(self.add_norm1(src, self.dropout1(att)))
        rec = self.linear(self.dropout(self.activation(self.lstm(src)[0])))
        return self.norm2(self.add_norm2(src, self.dropout2(rec)))


class Encoder(nn.Module):
    def __init__(self, W=2, N=64):
        super().__init__()
        self.W, self.N = W, N
        self.conv1d_U = nn.Conv1d(1, N, kernel_size=W, stride=W // 2, bias=False)
        self.relu = nn.ReLU()

    def forward(self, mixture):
        return self.relu(self.conv1d_U(mixture))                             # [B, N, L]


class Decoder(nn.Module):
    def __init__(self, E, W):
        super().__init__()
        self.E, self.W = E, W
        self.basis_signals = nn.Linear(E, W, bias=False)

    def forward(self, mixture_w):
        return overlap_and_add(self.basis_signals(mixture_w), self.W // 2)   # [..., L, E] -> [..., T]


class SingleTransformer(nn.Module):
    """One transformer layer applied to [batch, seq, dim] (the layer itself is sequence-first)."""

    def __init__(self, input_size, hidden_size, dropout):
        super().__init__()
        self.transformer = TransformerEncoderLayer(d_model=input_size, nhead=4, hidden_size=hidden_size,
                                                   dim_feedforward=hidden_size * 2, dropout=dropout)

    def forward(self, input):
        return self.transformer(input.permute(1, 0, 2).contiguous()).permute(1, 0, 2).contiguous()


class DPT(nn.Module):
    """Dual-path stack: per layer one transformer along the chunk axis (intra), one across chunks (inter)."""

    def __init__(self, input_size, hidden_size, output_size, num_layers=1, dropout=0):
        super().__init__()
        self.input_size, self.output_size, self.hidden_size = input_size, output_size, hidden_size
        self.row_transformer = nn.ModuleList([])
        self.col_transformer = nn.ModuleList([])
        for _ in range(num_layers):
            self.row_transformer.append(SingleTransformer(input_size, hidden_size, dropout))
            self.col_transformer.append(SingleTransformer(input_size, hidden_size, dropout))
        self.output = nn.Sequential(nn.PReLU(), nn.Conv2d(input_size, output_size, 1))

    def forward(self, input):
        B, Nf, K, S = input.shape                                            # features, chunk length, chunks
        x = input
        for intra, inter in zip(self.row_transformer, self.col_transformer):
            r = intra(x.permute(0, 3, 2, 1).contiguous().view(B * S, K, Nf))
            x = r.view(B, S, K, Nf).permute(0, 3, 2, 1).contiguous()
            c = inter(x.permute(0, 2, 3, 1).contiguous().view(B * K, S, Nf))
            x = c.view(B, K, S, Nf).permute(0, 3, 1, 2).contiguous()
        return self.output(x)


class DPT_base(nn.Module):
    def __init__(self, input_dim, feature_dim, hidden_dim, num_spk=2, layer=6, segment_size=250):
        super().__init__()
        self.input_dim, self.feature_dim, self.hidden_dim = input_dim, feature_dim, hidden_dim
        self.layer, self.segment_size, self.num_spk = layer, segment_size, num_spk
        self.eps = 1e-8
        self.BN = nn.Conv1d(self.input_dim, self.feature_dim, 1, bias=False)
        self.DPT = DPT(self.feature_dim, self.hidden_dim, self.feature_dim * self.num_spk, num_layers=layer)
        self.add = Add()

    def pad_segment(self, input, segment_size):
        """Right-pad to a whole number of half-overlapping chunks, then half a chunk of zeros on both sides."""
        hop = segment_size // 2
        rest = segment_size - (hop + input.shape[2] % segment_size) % segment_size
        return F.pad(input, (hop, rest + hop)), rest

    def split_feature(self, input, segment_size):
        """[B, N, T] -> ([B, N, segment_size, chunks], rest): chunk s starts at s * segment_size / 2."""
        padded, rest = self.pad_segment(input, segment_size)
        return padded.unfold(2, segment_size, segment_size // 2).transpose(2, 3).contiguous(), rest

    def merge_feature(self, input, rest):
        """Overlap-add of the chunks through `self.add` (the AddQ of the quantised model): even chunks and odd chunks each
        tile the padded axis, shifted by half a chunk."""
        B, Nf, K, S = input.shape
        hop = K // 2
        chunks = input.transpose(2, 3)                                        # [B, N, S, K]
        even = chunks[:, :, 0::2].reshape(B, Nf, -1)[:, :, hop:]
        odd = chunks[:, :, 1::2].reshape(B, Nf, -1)[:, :, :-hop]
        out = self.add(even, odd)
        if rest > 0:
            out = out[:, :, :-rest]
        return out.contiguous()

    def forward(self, input):
        pass


class BF_module(DPT_base):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.output = nn.Sequential(nn.Conv1d(self.feature_dim, self.feature_dim, 1), nn.Tanh())
        self.output_gate = nn.Sequential(nn.Conv1d(self.feature_dim, self.feature_dim, 1), nn.Sigmoid())
        self.mul = Mul()

    def forward(self, input):
        B = input.shape[0]
        feats = self.BN(input)
        segs, rest = self.split_feature(feats, self.segment_size)
        y = self.DPT(segs).view(B * self.num_spk, self.feature_dim, self.segment_size, -1)
        y = self.merge_feature(y, rest)                                       # [B*S, N, T]
        gated = self.mul(self.output(y), self.output_gate(y))
        return gated.transpose(1, 2).contiguous().view(B, self.num_spk, -1, self.feature_dim)


class DPTNetQ(nn.Module):
    def __init__(self, n_spks=2, kernel_size=2, enc_dim=256, feature_dim=64, hidden_dim=128, layer=6, segment_size=250):
        super().__init__()
        self.set_splitter_combiner(1, 1)
        self.window = kernel_size
        self.enc_dim, self.feature_dim, self.hidden_dim, self.segment_size = enc_dim, feature_dim, hidden_dim, segment_size
        self.layer = layer
        self.n_srcs = n_spks
        self.eps = 1e-8
        self.encoder = Encoder(kernel_size, enc_dim)
        self.enc_LN = nn.GroupNorm(1, self.enc_dim, eps=self.eps)
        self.separator = BF_module(self.enc_dim, self.feature_dim, self.hidden_dim, self.n_srcs, self.layer, self.segment_size)
        self.mask_conv1x1 = nn.Sequential(nn.Conv1d(self.feature_dim, self.enc_dim, 1, bias=False), nn.ReLU())
        self.decoder = Decoder(enc_dim, kernel_size)
        self.mul = Mul()

    def pre_process(self, x):
        return preprocess(x, n_splitter=self.n_splitter)

    def post_process(self, x):
        return postprocess(x, n_combiner=self.n_combiner)

    def forward(self, x):
        x = self.pre_process(x)                                               # [B, n_splitter, T]
        B = x.shape[0]
        mixture_w = self.encoder(x)                                           # [B, E, L]
        score = self.separator(self.enc_LN(mixture_w))                        # [B, S, L, N]
        score = score.view(B * self.n_srcs, -1, self.feature_dim).transpose(1, 2).contiguous()
        mask = self.mask_conv1x1(score).view(B, self.n_srcs, self.enc_dim, -1)
        source_w = self.mul(mixture_w.unsqueeze(1), mask).transpose(2, 3)     # [B, S, L, E]
        est = self.decoder(source_w)
        return self.post_process(est.reshape((self.n_combiner, B, self.n_srcs, 1, -1)))

    def load_pretrain(self, weights_path):
        """Positional key matching for checkpoints with foreign key names (dptnetq.py:408-421)."""
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
        for _, m in list(self.named_modules()):          # snapshot: the surgery below replaces children while we walk
            if type(m) is DPTNetQ:
                replace_encoderq(m.encoder, ["conv1d_U", "relu"], dict(edge, n_splitter=self.n_splitter, in_quant=in_quant,
                                                                       in_act_n_bits=in_act_n_bits))
                replace_decoderq(m.decoder, ["basis_signals"], dict(edge, n_combiner=self.n_combiner, act_n_bits=out_act_n_bits,
                                                                    out_quant=out_quant, out_act_n_bits=out_act_n_bits))
                quantize_modules(m, ["enc_LN"], p)
                quantize_modules(m.mask_conv1x1, ["0", "1"], p)
                quantize_modules(m, ["mul"], p)
            elif type(m) is TransformerEncoderLayer:
                for name in ("lstm", "linear", "norm1", "norm2", "add_norm1", "add_norm2", "self_attn"):
                    quantize_modules(m, [name], p)
            elif type(m) is DPT:
                quantize_modules(m.output, ["0"], p)
                quantize_modules(m.output, ["1"], p)
            elif type(m) is BF_module:
                quantize_modules(m.output, ["0", "1"], p)
                quantize_modules(m.output_gate, ["0", "1"], p)
                quantize_modules(m, ["mul"], p)
                quantize_modules(m, ["add"], p)
                quantize_modules(m, ["BN"], p)
