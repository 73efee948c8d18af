"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the next row of the scope table:
the fake-quantised ConvTasNetMusicQ forward (quantization/qat/models/convtasnetq_music.py:53-333 after
`quantize_model` :292-333), written functionally over a flat {state_dict key: tensor} dict like fqss_oracle.py and on
top of its quantiser primitives.  Nothing under fqss_b200/ imports it; there is no CUDA path for this model yet -- the
oracle and its pin against the unmodified reference (oracle/check_music_against_reference.py,
tests/golden/music_small.npz) are the first gate of that row.

Differences from the speech model that the CUDA engine will have to cover:
  * splitter WITHOUT normalisation (`preprocess(..., normalize=False)`, :234): threshold = max|x| of the batch
  * encoder Conv1d(audio_channels * n_splitter -> N, kernel, stride) + ReLU + FQ          (qat_layers.py:993-1046)
  * cLN = nn.LayerNorm over the CHANNEL axis per frame (transpose, LayerNorm(N), transpose) + FQ     (:33-51, :455-468)
  * blocks without a skip path: 1x1 + PReLU + FQ, gLN + FQ, depthwise + PReLU + FQ, gLN + FQ, 1x1 + FQ, residual
    add + FQ; no biases anywhere                                                                      (:117-176)
  * mask 1x1 (no bias) + ReLU + FQ, MulQ
  * decoder nn.Linear(N -> audio_channels * kernel) per frame (per-output-feature weight ranges) + FQ_out, the RQB
    with a Linear re-encoder, then overlap_and_add with hop `stride`                      (qat_layers.py:1256-1302, :8-30)
"""
import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F

from fqss_oracle import EPS_GLN, Params, QuantState, _Ctx, combine_output, floor_quant  # noqa: F401


@dataclass
class MusicConfig:
    n_src: int = 4
    audio_channels: int = 2
    n_filters: int = 256
    kernel: int = 20
    stride: int = 10
    bn_chan: int = 256
    hid_chan: int = 512
    conv_kernel: int = 3
    n_blocks: int = 10
    n_repeats: int = 4
    n_splitter: int = 2
    n_combiner: int = 2
    act_bits: int = 8
    weight_bits: int = 8
    out_bits: int = 8
    max_observations: int = 50
    alpha: float = 0.9


def split_input_unnormalised(x, n_splitter, n_bits=8):
    """process.preprocess(normalize=False), process.py:16-36: the threshold is max|x| over the whole batch."""
    if x.dim() == 2:
        x = x.unsqueeze(1)
    if n_splitter <= 1:
        return x
    threshold = max(abs(x.min()), abs(x.max()))
    delta = threshold / (2 ** (n_bits - 1))
    ys = []
    for _ in range(n_splitter):
        xq = floor_quant(x, threshold=threshold, n_bits=n_bits)
        ys.append(xq)
        x = 2 * (x - xq) * threshold / delta - threshold
    return torch.cat(ys, dim=1)


def overlap_and_add(signal, frame_step):
    """convtasnetq_music.py:10-30 (index_add of gcd-sized sub-frames)."""
    outer = signal.size()[:-2]
    frames, frame_length = signal.size()[-2:]
    sub = math.gcd(frame_length, frame_step)
    sub_step = frame_step // sub
    sub_per_frame = frame_length // sub
    out_size = frame_step * (frames - 1) + frame_length
    out_sub = out_size // sub
    sig = signal.reshape(*outer, -1, sub)
    idx = torch.arange(0, out_sub).unfold(0, sub_per_frame, sub_step).long().contiguous().reshape(-1)
    res = signal.new_zeros(*outer, out_sub, sub)
    res.index_add_(-2, idx, sig)
    return res.reshape(*outer, -1)


def _block(c: _Ctx, prefix: str, x, dilation: int, pad: int):
    """ConvBlock (:117-140) with DepthwiseSeparableConv (:143-176) after quantize_model."""
    P = c.P
    n = prefix + "net."
    w = c.wq(n + "0.weight_fake_quantize", P[n + "0.conv1d.weight"])
    y = c.aq(n + "0.activation_fake_quantize", F.prelu(F.conv1d(x, w, None), P[n + "0.nl.weight"]))
    y = c.aq(n + "2.activation_fake_quantize", F.group_norm(y, 1, P[n + "2.groupnorm.weight"], P[n + "2.groupnorm.bias"], EPS_GLN))
    d = n + "3.net."
    w = c.wq(d + "0.weight_fake_quantize", P[d + "0.conv1d.weight"])
    y = F.conv1d(y, w, None, padding=pad, dilation=dilation, groups=y.shape[1])
    y = c.aq(d + "0.activation_fake_quantize", F.prelu(y, P[d + "0.nl.weight"]))
    y = c.aq(d + "2.activation_fake_quantize", F.group_norm(y, 1, P[d + "2.groupnorm.weight"], P[d + "2.groupnorm.bias"], EPS_GLN))
    w = c.wq(d + "3.weight_fake_quantize", P[d + "3.conv1d.weight"])
    y = c.aq(d + "3.activation_fake_quantize", F.conv1d(y, w, None))
    return c.aq(prefix + "add.activation_fake_quantize", y + x)          # Add(out, residual), AddQ


def music_forward(P: Params, x, cfg: MusicConfig, st: QuantState = None, tap: dict = None):
    """ConvTasNetMusicQ.forward (:236-275), quantised: x [B, audio_channels, T] -> [B, n_src, audio_channels, T']."""
    st = st or QuantState()
    c = _Ctx(P, cfg, st, True, tap)
    xin = split_input_unnormalised(x, cfg.n_splitter)
    B = xin.shape[0]
    w = c.wq("encoder.0.weight_fake_quantize", P["encoder.0.conv1d.weight"])
    feats = c.aq("encoder.0.activation_fake_quantize", F.relu(F.conv1d(xin, w, None, stride=cfg.stride)))
    c.rec("encoder", feats)
    s = "separator.network."
    y = F.layer_norm(feats.transpose(1, 2), (cfg.n_filters,), P[s + "0.norm.layernorm.weight"], P[s + "0.norm.layernorm.bias"], EPS_GLN)
    y = c.aq(s + "0.norm.activation_fake_quantize", y).transpose(1, 2)
    w = c.wq(s + "1.weight_fake_quantize", P[s + "1.conv1d.weight"])
    y = c.aq(s + "1.activation_fake_quantize", F.conv1d(y, w, None))
    c.rec("bottleneck", y)
    for r in range(cfg.n_repeats):
        for b in range(cfg.n_blocks):
            dil = 2 ** b
            y = _block(c, "%s2.%d.%d." % (s, r, b), y, dil, (cfg.conv_kernel - 1) * dil // 2)
    c.rec("tcn", y)
    w = c.wq(s + "3.weight_fake_quantize", P[s + "3.conv1d.weight"])
    mask = c.aq(s + "3.activation_fake_quantize", F.relu(F.conv1d(y, w, None))).reshape(B, cfg.n_src, cfg.n_filters, -1)
    masked = c.aq("mul.activation_fake_quantize", mask * feats.unsqueeze(1))
    c.rec("masked", masked)
    X = masked.transpose(2, 3)                                            # [B, S, K, N]
    wd = c.wq("decoder.weight_fake_quantize", P["decoder.linear.weight"])
    y0 = c.aq("decoder.activation_fake_quantize", F.linear(X, wd, None), cfg.out_bits)
    outs = [y0]
    cur_in, cur_out = X, y0
    for _ in range(1, cfg.n_combiner):                                    # RQB, qat_layers.py:1178-1187, 1286-1293
        we = c.wq("decoder.residual_error_block.weight_fake_quantize", P["decoder.residual_error_block.residual_encoder.weight"])
        Xq = F.linear(cur_out, we, None)
        X1 = c.aq("decoder.residual_error_block.activation_fake_quantize", cur_in - Xq)
        y1 = F.linear(X1, wd, None)
        cur_in = y1
        cur_out = c.aq("decoder.activation_fake_quantize_residual", y1, cfg.out_bits)
        outs.append(cur_out)
    dec = torch.stack(outs) if cfg.n_combiner > 1 else y0
    K = dec.shape[-2]
    dec = dec.reshape((cfg.n_combiner, B, cfg.n_src, K, cfg.audio_channels, -1)).transpose(3, 4)
    out = overlap_and_add(dec, cfg.stride)                                # [n_combiner, B, S, audio_channels, T']
    c.rec("decoder", out)
    return combine_output(out, cfg.n_combiner)


def calibrate_music(P: Params, x, cfg: MusicConfig, passes: int = 2) -> QuantState:
    st = QuantState(observe=True)
    with torch.no_grad():
        for _ in range(passes):
            music_forward(P, x, cfg, st)
    st.observe = False
    st.weights_seen = True
    return st


# --------------------------------------------------------------------------------------
# training loss of the music recipe (train_env/tasnet_musdbhq/musdbhq_train.py:76-109)
# --------------------------------------------------------------------------------------
def center_trim(tensor, reference):
    """musdbhq_utils.py:16-29."""
    ref = reference.size(-1) if hasattr(reference, "size") else int(reference)
    delta = tensor.size(-1) - ref
    if delta < 0:
        raise ValueError("tensor must be larger than reference")
    return tensor[..., delta // 2:-(delta - delta // 2)] if delta else tensor


def new_sdr_db(ref, sig, eps=1e-7):
    """process.py:70-75 (MDX new-SDR), a python float: 10 log10 of an fp32 ratio."""
    import numpy as np
    r = (torch.sum(torch.square(ref)) + eps) / (torch.sum(torch.square(ref - sig)) + eps)
    return 10 * np.log10(r.item())


def music_kd_loss(wavs, fwavs, sources, kd_lambda=0.1):
    """musdbhq_train.py:87-109 with loss_fn = nn.L1Loss() (:249): per-item new-SDR weights (note the argument order of the
    call sites: the ESTIMATE is passed as `ref`), weighted per-item L1 distillation term, batch L1 task term.
    -> (loss, kd term, task term); kd_lambda == 0 or fwavs None: loss = task."""
    task = F.l1_loss(wavs, sources)
    if not (kd_lambda > 0) or fwavs is None:
        return task, torch.zeros(()), task
    n = wavs.shape[0]
    with torch.no_grad():
        sdr = torch.Tensor([new_sdr_db(fwavs[i:i + 1], sources[i:i + 1]) for i in range(n)])
        sdrq = torch.Tensor([new_sdr_db(wavs[i:i + 1], sources[i:i + 1]) for i in range(n)])
        w = 10 ** ((sdr - sdrq) / 10)
    kd = torch.mean(w * torch.stack([F.l1_loss(wavs[i:i + 1], fwavs[i:i + 1]) for i in range(n)], dim=0))
    return (1 - kd_lambda) * task + kd_lambda * kd, kd, task
