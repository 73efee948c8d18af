"""B200 mirror of `quantization/qat/models/convtasnetq_music.py`: the Demucs-v2-style
Conv-TasNet for 4-stem stereo music -- same module tree, constructor signatures, attribute names and state_dict keys as
the reference (ChannelWiseLayerNorm :32-50, MaskGenerator :53-114, ConvBlock :117-140, DepthwiseSeparableConv :143-176,
ConvTasNetMusicQ :179-333).  After `quantize_model` every child is a wrapper from `fqss_b200.qat.qat_layers` running
sm_100a kernels through the C ABI; the pieces this model adds to the speech path are the un-normalised multi-channel
splitter (`fqss_split_ex`), the channel-wise LayerNorm (`fqss_cln_fwd/bwd`), the per-frame Linear decoder as 1x1
convolutions and the overlap-add (`fqss_ola_fwd/bwd`).

Everything stays channels-first ([B, C, K]); the reference's transposes around LayerNorm / Linear are views here.
In the quantised steady state the TCN blocks run on the fused tcgen05 engine of the speech model in its skip-less mode
(`fqss_tcn_block.no_skip`: these blocks have a residual path only); observer calibration, the float teacher and shapes the
engine does not cover stay on the per-layer wrappers."""
import torch
import torch.nn as nn

from ... import ops
from ...process import postprocess, preprocess
from ..qat_layers import Add, Mul
from ..qat_utils import quantize_modules, replace_decoderq, replace_encoderq

EPS = 1e-8


def overlap_and_add(signal, frame_step):
    """[..., frames, frame_length] -> [..., frame_step * (frames - 1) + frame_length] (convtasnetq_music.py:10-30), for callers
    that hold the reference's frame-major layout; the model itself keeps frames in the last axis and calls ops.OverlapAdd."""
    frames, length = signal.shape[-2], signal.shape[-1]
    y = signal.transpose(-1, -2).reshape(signal.shape[:-2] + (1 * length, frames))
    return ops.OverlapAdd.apply(y, 1, length, frame_step).squeeze(-2)


class ChannelWiseLayerNorm(nn.Module):
    """LayerNorm over the channel axis of [B, C, K]."""

    def __init__(self, N, eps=EPS):
        super().__init__()
        self.norm = nn.LayerNorm(N, eps=eps)

    def forward(self, x):
        return torch.transpose(self.norm(torch.transpose(x, 1, 2)), 1, 2)


class DepthwiseSeparableConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation):
        super().__init__()
        self.net = nn.Sequential(
            nn.Conv1d(in_channels, in_channels, kernel_size, stride=stride, padding=padding, dilation=dilation,
                      groups=in_channels, bias=False),
            nn.PReLU(),
            nn.GroupNorm(num_groups=1, num_channels=in_channels, eps=EPS),
            nn.Conv1d(in_channels, out_channels, 1, bias=False))

    def forward(self, x):
        return self.net(x)


class ConvBlock(nn.Module):
    """1x1 expand -> PReLU -> gLN -> depthwise dilated conv -> PReLU -> gLN -> 1x1, plus the residual."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation):
        super().__init__()
        self.net = nn.Sequential(
            nn.Conv1d(in_channels, out_channels, 1, bias=False),
            nn.PReLU(),
            nn.GroupNorm(num_groups=1, num_channels=out_channels, eps=EPS),
            DepthwiseSeparableConv(out_channels, in_channels, kernel_size, stride, padding, dilation))
        self.add = Add()

    def forward(self, x):
        xa, xb = ops.fanout2(x)                     # two consumers: gradients summed by the library
        return self.add(self.net(xa), xb)


class MaskGenerator(nn.Module):
    def __init__(self, N, B, H, P, X, R, C, mask_act="relu"):
        super().__init__()
        self.C = C
        repeats = []
        for _ in range(R):
            blocks = []
            for x in range(X):
                d = 2 ** x
                blocks.append(ConvBlock(B, H, P, stride=1, padding=(P - 1) * d // 2, dilation=d))
            repeats.append(nn.Sequential(*blocks))
        if mask_act == "sigmoid":
            act = nn.Sigmoid()
        elif mask_act == "relu":
            act = nn.ReLU()
        else:
            raise ValueError(f"Unsupported activation {mask_act}")
        self.network = nn.Sequential(ChannelWiseLayerNorm(N, eps=EPS), nn.Conv1d(N, B, 1, bias=False),
                                     nn.Sequential(*repeats), nn.Conv1d(B, C * N, 1, bias=False), act)

    use_fused = True      # class-level switch: False forces the per-layer wrappers (tests / debugging)

    def forward(self, mixture_w):
        M, N, K = mixture_w.size()
        from ... import tcn_engine as E
        net = self.network
        if self.use_fused and E.fused_eligible_noskip(self, mixture_w):
            # quantised steady state: the 40 blocks run as ONE autograd node on the fused engine (no_skip blocks:
            # expand GEMM, depthwise row kernel, hidden quantiser, residual GEMM; fused backward), everything around them
            # on the per-layer wrappers as before
            h = net[1](net[0](mixture_w))
            q = net[1].activation_fake_quantize
            blocks = [b for rep in net[2] for b in rep]
            h, _ = E.fused_tcn(h, blocks, None, True, (q.min_range, q.max_range), start=0, total=len(blocks) + 1)
            h._fq_src = blocks[-1].add.activation_fake_quantize      # on the last AddQ's grid: the mask conv runs on code operands
            return net[4](net[3](h)).reshape(M, self.C, N, K)
        from ... import float_engine as FE
        if self.use_fused and FE.noskip_eligible(self, mixture_w):
            # the recipe's float teacher (same module tree, quantisation disabled) under no_grad: the block stack on the
            # fused engine's float mode (split-bf16 GEMMs, second gLN folded into the residual conv, 3 kernels per block)
            h = FE.tcn_noskip_infer(self, net[1](net[0](mixture_w)))
            return net[4](net[3](h)).reshape(M, self.C, N, K)
        return net(mixture_w).reshape(M, self.C, N, K)


class ConvTasNetMusicQ(nn.Module):
    def __init__(self, sources=("drums", "bass", "other", "vocals"), audio_channels=2, n_filters=256, kernel=20, stride=10,
                 bn_chan=256, hid_chan=512, conv_kernel=3, n_blocks=10, n_repeats=4, mask_act="relu"):
        super().__init__()
        self.sources = list(sources)
        self.n_srcs = len(self.sources)
        self.set_splitter_combiner(1, 1)
        self.stride = stride
        self.kernel = kernel
        self.audio_channels = audio_channels
        self.encoder = nn.Sequential(nn.Conv1d(audio_channels, n_filters, kernel, stride=stride, padding=0, bias=False),
                                     nn.ReLU())
        self.separator = MaskGenerator(n_filters, bn_chan, hid_chan, conv_kernel, n_blocks, n_repeats, self.n_srcs, mask_act)
        self.decoder = nn.Linear(n_filters, audio_channels * kernel, bias=False)
        self.mul = Mul()

    def pre_process(self, x):
        return preprocess(x, n_splitter=self.n_splitter, normalize=False)

    def post_process(self, x):
        return postprocess(x, n_combiner=self.n_combiner)

    def forward(self, x):
        x = self.pre_process(x)                                      # [B, audio_channels * n_splitter, T]
        B = x.shape[0]
        feats = self.encoder(x)                                      # [B, N, K]
        f_mask, f_mul = ops.fanout2(feats)
        masked = self.mul(self.separator(f_mask), f_mul.unsqueeze(1))   # [B, S, N, K]
        Nf, K = masked.shape[-2], masked.shape[-1]
        dec = self.decoder
        if hasattr(dec, "forward_ncl"):                              # LinearDecoderQ: channels-first, [n, B*S, A*L, K]
            dec_in = masked.reshape(B * self.n_srcs, Nf, K)
            src = getattr(masked, "_fq_src", None)
            if src is not None:
                dec_in._fq_src = src                                 # still on the MulQ quantiser's grid: code-operand GEMMs
            out = dec.forward_ncl(dec_in)
        else:                                                        # float nn.Linear: the reference's frame-major call
            out = dec(torch.transpose(masked, 2, 3)).transpose(2, 3).reshape(1, B * self.n_srcs, -1, K)
        out = ops.OverlapAdd.apply(out, self.audio_channels, out.shape[-2] // self.audio_channels, self.stride)
        out = out.reshape((self.n_combiner, B, self.n_srcs, self.audio_channels, -1))
        return self.post_process(out)                                # [B, S, audio_channels, T']

    def load_pretrain(self, weights_path):
        """Positional key matching for checkpoints with foreign key names (convtasnetq_music.py:277-291)."""
        own = self.state_dict()
        src = torch.load(weights_path)
        src = src.get("state_dict", src)
        src = {k: v for k, v in src.items() if not k.startswith("fmodel.")}
        assert len(own) == len(src), \
            "Error: mismatch models weights. Please check if the model configurations match to model weights!"
        new = {}
        for mine, theirs in zip(own.keys(), src.keys()):
            v = src[theirs]
            new[mine] = v.reshape(-1) if ("beta" in theirs or "gamma" in theirs) else v
        self.load_state_dict(new, strict=True)

    def set_splitter_combiner(self, n_splitter, n_combiner):
        self.n_splitter = n_splitter
        self.n_combiner = n_combiner

    def quantize_model(self, gradient_based=True, weight_quant=True, weight_n_bits=8, act_quant=True, act_n_bits=8,
                       inout_nl_quant=False, in_quant=False, in_act_n_bits=8, out_quant=True, out_act_n_bits=8):
        p = dict(gradient_based=gradient_based, act_quant=act_quant, weight_quant=weight_quant,
                 weight_n_bits=weight_n_bits, act_n_bits=act_n_bits)
        edge = dict(p, inout_nl_quant=inout_nl_quant)
        # snapshot first: the surgery mutates the tree while the reference walks named_modules() lazily -- keep its
        # creation order (the grown encoder draws from the global RNG)
        for _, m in list(self.named_modules()):
            if type(m) is ConvTasNetMusicQ:
                replace_encoderq(m.encoder, ["0", "1"], dict(edge, n_splitter=self.n_splitter, in_quant=in_quant,
                                                             in_act_n_bits=in_act_n_bits))
                replace_decoderq(m, ["decoder"], dict(edge, n_combiner=self.n_combiner, act_n_bits=out_act_n_bits,
                                                      out_quant=out_quant, out_act_n_bits=out_act_n_bits, train_res_dec=False))
                quantize_modules(m, ["mul"], p)
            elif type(m) is ConvBlock:
                quantize_modules(m.net, ["0", "1"], p)
                quantize_modules(m.net, ["2"], p)
                quantize_modules(m, ["add"], p)
            elif type(m) is DepthwiseSeparableConv:
                quantize_modules(m.net, ["0", "1"], p)
                quantize_modules(m.net, ["2"], p)
                quantize_modules(m.net, ["3"], p)
            elif type(m) is MaskGenerator:
                quantize_modules(m.network[0], ["norm"], p)
                quantize_modules(m.network, ["1"], p)
                quantize_modules(m.network, ["3", "4"], p)
