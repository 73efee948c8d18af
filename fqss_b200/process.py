"""B200 mirror of the hot-path part of the reference's `process.py` (:10-52): the FQSS input
splitter (`preprocess`) and output reconstructor (`postprocess`).  Metrics / OLA inference
(:54-194) are out of scope."""
import torch

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
        if not (sign and normalize):
            raise NotImplementedError("splitter kernel implements the recipe's signed, normalised mode")
        return ops.split_input(x, n_splitter, n_bits)
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
