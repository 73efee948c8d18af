"""B200 mirror of the hot-path part of the reference's `process.py` (:10-52): the FQSS input
splitter (`preprocess`) and output reconstructor (`postprocess`), plus the evaluation entry `model_infer`
(:154-194, chunked overlap-add inference used by val.py).  The metric helpers (:54-152, torchmetrics) are out of
scope."""
import torch
import torch.nn.functional as F

from . import ops


def quantize(x, threshold=1.0, n_bits=8, sign=True):
    """Floor quantiser of the splitter (process.py:10-14).  Tiny helper kept for API parity."""
    delta = threshold / (2 ** (n_bits - int(sign)))
    lo = -2 ** (n_bits - int(sign)) if sign else 0
    hi = 2 ** (n_bits - int(sign)) - 1
    return torch.clip(torch.floor(x / delta), lo, hi) * delta


def preprocess(x, n_splitter=1, n_bits=8, sign=True, normalize=True):
    if x.dim() == 2:
        x = x.unsqueeze(1)
    if n_splitter > 1:
        if not sign:
            raise NotImplementedError("splitter kernel implements the recipes' signed mode")
        return ops.split_input(x, n_splitter, n_bits, normalize=normalize)
    return x


def postprocess(x, n_combiner=1, n_bits=8, sign=True):
    # x: [n_combiner, batch, sources, audio_channels, T]
    if n_combiner == 1:
        y = x.squeeze(0)
    else:
        if not sign:
            raise NotImplementedError("reconstructor kernel implements the recipe's signed mode")
        y = ops.Combine.apply(x, n_bits)
    if y.dim() <= 4 and y.shape[-2] == 1:
        y = y.squeeze(-2)
    return y


def _si_snr(est, ref, eps=0.0):
    """Scale-invariant SNR in dB of 1-D signals (what torchmetrics' ScaleInvariantSignalNoiseRatio computes:
    zero-mean both, project, 10 log10 of the energy ratio); used only to order sources in `swap_channel_order`."""
    est = est - est.mean(dim=-1, keepdim=True)
    ref = ref - ref.mean(dim=-1, keepdim=True)
    eps = torch.finfo(est.dtype).eps if eps == 0.0 else eps
    alpha = ((est * ref).sum(-1, keepdim=True) + eps) / ((ref * ref).sum(-1, keepdim=True) + eps)
    tgt = alpha * ref
    noise = tgt - est
    return 10.0 * torch.log10(((tgt * tgt).sum(-1) + eps) / ((noise * noise).sum(-1) + eps))


def swap_channel_order(sep_tensor, clean_tensor):
    """Order (and sign) the separated sources like the clean references, by maximal SI-SNR (process.py:106-124)."""
    n_src = clean_tensor.shape[0]
    if n_src == 1:
        return sep_tensor
    new = sep_tensor.clone()
    for src in range(n_src):
        sep_ch = sep_tensor[src:src + 1, :]
        scores = torch.stack([_si_snr(sep_ch.reshape(-1), clean_tensor[i].reshape(-1)) for i in range(n_src)])
        best = int(torch.argmax(scores))          # first maximum, like the reference's strict `>` scan
        new[best, ...] = sep_ch if src == best else -sep_ch
    return new


def model_infer(model, mix, n_srcs=1, segment=None, overlap=0.25, device="cuda", target=None):
    """process.model_infer (process.py:154-194): batch-1 inference of `mix` [channels, length]; with `segment`,
    chunked inference with triangular-window overlap-add.  Same chunking, padding, window and normalisation as the
    reference, chunk by chunk (the splitter normalises by ONE peak per call, so chunks are NOT batched: batching them
    would change the quantised input).  Chunks, window and accumulators stay on `device`; one copy back at the end."""
    if not segment:
        x = mix.unsqueeze(0)
        with torch.no_grad():
            out = model(x.to(device)).detach()
        out = out[0]
        out = F.pad(out, (0, x.size(-1) - out.size(-1)))
        return out.cpu()
    channels, length = mix.shape
    num_srcs = model.n_srcs if hasattr(model, "n_srcs") else n_srcs
    out_shape = (num_srcs, channels, length) if channels > 1 else (num_srcs, length)
    mix_d = mix.to(device)
    tgt_d = target.to(device) if target is not None else None
    out = torch.zeros(*out_shape, device=device)
    sum_weight = torch.zeros(length, device=device)
    stride = int((1 - overlap) * segment)
    weight = torch.cat([torch.arange(1, segment // 2 + 1), torch.arange(segment - segment // 2, 0, -1)])
    assert len(weight) == segment
    weight = (weight / weight.max()).to(device)
    for start in range(0, length, stride):
        stop = min(start + segment, length)
        chunk = mix_d[..., start:stop]
        n = chunk.size(-1)
        chunk = F.pad(chunk, (0, segment - n))
        with torch.no_grad():
            co = model(chunk.unsqueeze(0)).detach()[0]
        co = F.pad(co, (0, segment - co.size(-1)))[..., :n]
        if tgt_d is not None and num_srcs > 1:
            co = swap_channel_order(co, tgt_d[..., start:start + n])
        out[..., start:stop] += weight[:n] * co
        sum_weight[start:stop] += weight[:n]
    assert float(sum_weight.min()) > 0
    out /= sum_weight
    return out.cpu()
