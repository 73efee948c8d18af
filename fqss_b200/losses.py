"""FQSS training-step arithmetic: the body of `System.common_step`
(train_env/asteroid_librimix/mysystem.py:124-151) with the python loop of 2*B PIT calls, the
[B,S,S,T] temporaries of `PairwiseWSDR` (wsdr.py:46-95) and asteroid's PITLossWrapper replaced by
one fused CUDA reduction (csrc/loss.cu)."""
import torch

from . import ops


def fqss_kd_loss(est, fest, targets, kd_lambda=0.1):
    """-> (loss, kd_loss_logged, val_loss) as 0-dim tensors; gradient flows from `loss` into `est`."""
    out = ops.kd_loss(est, fest, targets, kd_lambda)
    return out[0], out[1].detach(), out[2].detach()


def fqss_training_step(model, fmodel, inputs, targets, kd_lambda=0.1):
    """common_step(train=True): student forward, float-teacher forward (no grad), KD SI-SDR loss."""
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
