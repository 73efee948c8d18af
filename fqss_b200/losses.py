"""FQSS training-step arithmetic: the body of `System.common_step`
(train_env/asteroid_librimix/mysystem.py:124-151) with the python loop of 2*B PIT calls, the
[B,S,S,T] temporaries of `PairwiseWSDR` (wsdr.py:46-95) and asteroid's PITLossWrapper replaced by
one fused CUDA reduction (csrc/loss.cu)."""
import os

import torch

from . import ops


def fqss_kd_loss(est, fest, targets, kd_lambda=0.1):
    """-> (loss, kd_loss_logged, val_loss) as 0-dim tensors; gradient flows from `loss` into `est`."""
    out = ops.kd_loss(est, fest, targets, kd_lambda)
    return out[0], out[1].detach(), out[2].detach()


_TEACHER_STREAMS = {}


def _teacher_stream(device):
    key = device.index if device.index is not None else torch.cuda.current_device()
    s = _TEACHER_STREAMS.get(key)
    if s is None:
        s = _TEACHER_STREAMS[key] = torch.cuda.Stream(device=device)
    return s


def teacher_overlap_default():
    """The float teacher's forward does not depend on the student's: by default it runs on a side stream next to the student
    forward (fork / join through events, so the pair is captured into the step's CUDA graph as two branches).  The row kernels
    of one model (issue-bound, ~50 % of HBM) and the TMA-fed GEMMs of the other fill each other's idle resources.
    FQSS_TEACHER_OVERLAP=0 serialises them (per-kernel profiling, A/B runs)."""
    if _OVERRIDE[0] is not None:
        return _OVERRIDE[0]
    return os.environ.get("FQSS_TEACHER_OVERLAP", "1") not in ("", "0")


_OVERRIDE = [None]


def set_teacher_overlap(flag):
    """Force (True / False) or release (None) the side-stream teacher; returns the previous setting."""
    prev = _OVERRIDE[0]
    _OVERRIDE[0] = flag
    return prev


def fqss_training_step(model, fmodel, inputs, targets, kd_lambda=0.1, overlap_teacher=None):
    """common_step(train=True): student forward, float-teacher forward (no grad), KD SI-SDR loss."""
    if overlap_teacher is None:
        overlap_teacher = teacher_overlap_default()
    if overlap_teacher and inputs.is_cuda:
        cur = torch.cuda.current_stream(inputs.device)
        side = _teacher_stream(inputs.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():
            fest = fmodel(inputs)
        inputs.record_stream(side)
        est = model(inputs)
        cur.wait_stream(side)
        fest.record_stream(cur)
    else:
        est = model(inputs)
        with torch.no_grad():
            fest = fmodel(inputs)
    loss, kd_logged, _ = fqss_kd_loss(est, fest, targets, kd_lambda)
    return loss, kd_logged, est


def fqss_validation_loss(model, inputs, targets):
    """common_step(train=False): PIT negative SI-SDR (dB), mysystem.py:148-151."""
    with torch.no_grad():
        est = model(inputs)
        return ops.kd_loss(est, est, targets, 0.0)[2]


def center_trim(tensor, reference):
    """Centre-trim the last axis of `tensor` to the length of `reference` (tensor or int); an odd surplus loses its extra
    sample on the right (train_env/tasnet_musdbhq/musdbhq_utils.py:16-29).  A view."""
    ref = reference.size(-1) if hasattr(reference, "size") else int(reference)
    delta = tensor.size(-1) - ref
    if delta < 0:
        raise ValueError("tensor must be larger than reference. Delta is %d." % delta)
    return tensor[..., delta // 2:tensor.size(-1) - (delta - delta // 2)] if delta else tensor


def music_training_step(model, fmodel, mix, sources, kd_lambda=0.1):
    """One training step's forward of the music recipe (musdbhq_train.py:76-109): student, float teacher (no grad), centre
    trim of the sources, L1 task loss + new-SDR-weighted L1 distillation loss as one fused reduction.
    -> (loss, kd term, task term, wavs)."""
    wavs = model(mix)
    src = center_trim(sources, wavs)
    fwavs = None
    if kd_lambda > 0 and fmodel is not None:
        with torch.no_grad():
            fwavs = fmodel(mix).detach()
    out = ops.music_kd_loss(wavs, fwavs, src, kd_lambda)
    return out[0], out[1].detach(), out[2].detach(), wavs
