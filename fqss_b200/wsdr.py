"""Drop-in for `train_env/asteroid_librimix/wsdr.py` (PairwiseWSDR, pairwise_wsisdr; :46-95) and for the
permutation-invariant wrapper the recipe puts around it (asteroid 0.6 `PITLossWrapper(pit_from="pw_mtx")`, not in the
reference tree; call sites mysystem.py:83,135-144).

The fused training loss is `fqss_b200.losses.fqss_kd_loss` (one library call for the whole of `common_step`).  This
module gives the same arithmetic its reference-shaped API for code that calls the pieces directly: the two passes over
the signals (means, centred inner products) and the gradient pass are the loss kernels of csrc/loss.cu
(`fqss_loss_stats`, `fqss_loss_grad_apply`); what happens in between is O(batch) algebra on [B,2,2] tensors.
Two sources, zero-mean SI-SDR -- the configuration of the recipe; anything else raises.
"""
import itertools

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.loss import _Loss

from . import _native as N
from ._native import check, lib, ptr, stream_ptr
from .ops import alloc_rows, ld_of, rows_view


class _PairwiseSISDRRatio(Function):
    """rho[b,i,j] = |alpha t_j|^2 / (|e_i - alpha t_j|^2 + eps), alpha = <e_i,t_j> / (|t_j|^2 + eps), zero-mean signals."""

    @staticmethod
    def forward(ctx, est, tgt, eps):
        N.require_cuda(est, tgt)
        if est.shape != tgt.shape or est.dim() != 3:
            raise TypeError("Inputs must be of shape [batch, n_src, time], got %s and %s instead" % (tgt.size(), est.size()))
        B, S, T = est.shape
        if S != 2:
            raise NotImplementedError("the loss kernels implement the recipe's two-speaker case (n_src == 2)")
        e, _, _, lde = rows_view(est.detach())
        t, _, _, ldt = rows_view(tgt.detach())
        st = torch.empty((B, 24), dtype=torch.float64, device=est.device)
        check(lib().fqss_loss_stats(ptr(e), lde, ptr(e), lde, ptr(t), ldt, B, T, ptr(st), stream_ptr()))
        ee, tt = st[:, 6:8], st[:, 10:12]                              # [B,2]
        d = st[:, 12:16].reshape(B, 2, 2)                              # <e_i, t_j>
        E = tt[:, None, :] + eps
        alpha = d / E
        P = alpha * alpha * tt[:, None, :]
        Nn = ee[:, :, None] - 2.0 * alpha * d + P + eps
        rho = P / Nn
        ctx.save_for_backward(e, t, st, alpha, P, Nn, E)
        ctx.meta = (B, T, lde, ldt)
        return rho.float()

    @staticmethod
    @once_differentiable
    def backward(ctx, G):
        e, t, st, alpha, P, Nn, E = ctx.saved_tensors
        B, T, lde, ldt = ctx.meta
        G = G.double()
        tt = st[:, 10:12]
        ce = -2.0 * P / (Nn * Nn)                                                          # d rho / d e_i  =  ce e_i + ct t_j
        ct = 2.0 * alpha * tt[:, None, :] / (E * Nn) + 2.0 * alpha * (2.0 - tt[:, None, :] / E) * P / (Nn * Nn)
        coef = torch.zeros((B, 16), dtype=torch.float64, device=G.device)
        coef[:, 0:4] = (G * ct).reshape(B, 4)                                              # a_ij on t_j
        coef[:, 8:10] = (G * ce).sum(dim=2)                                                # c_i on e_i
        coef[:, 10:12] = st[:, 0:2] / T
        coef[:, 12:14] = st[:, 0:2] / T                                                    # "f" = e here (b_ij = 0)
        coef[:, 14:16] = st[:, 4:6] / T
        coef = coef.float().contiguous()
        gest = alloc_rows(e.shape, e.device)
        check(lib().fqss_loss_grad_apply(ptr(e), lde, ptr(e), lde, ptr(t), ldt, B, T, ptr(coef), ptr(gest), ld_of(gest),
                                         stream_ptr()))
        return gest, None, None


class PairwiseWSDR(_Loss):
    """wsdr.py:46-95: pairwise (weighted) SI-SDR matrix [batch, n_src(est), n_src(tgt)].  `take_log=True` returns
    `10 log10(rho + EPS)`, `take_log=False` returns `-rho` -- exactly the reference's (asymmetric) sign convention."""

    def __init__(self, sdr_type, zero_mean=True, take_log=True, EPS=1e-8):
        super().__init__()
        assert sdr_type in ["snr", "sisdr", "sdsdr"]
        if sdr_type != "sisdr" or not zero_mean:
            raise NotImplementedError("only the zero-mean 'sisdr' variant is on the FQSS ConvTasNet path")
        self.sdr_type, self.zero_mean, self.take_log, self.EPS = sdr_type, zero_mean, take_log, EPS

    def forward(self, est_targets, targets, weights=None):
        if targets.size() != est_targets.size() or targets.ndim != 3:
            raise TypeError(f"Inputs must be of shape [batch, n_src, time], got {targets.size()} and {est_targets.size()} instead")
        if targets.requires_grad:
            raise NotImplementedError("gradient flows into the estimates only (the recipe detaches both kinds of target)")
        rho = _PairwiseSISDRRatio.apply(est_targets, targets, float(self.EPS))
        if weights is not None:
            rho = rho * weights[:, None, None]
        if self.take_log:
            return 10 * torch.log10(rho + self.EPS)
        return -rho


pairwise_wsisdr = PairwiseWSDR("sisdr", take_log=False)          # kd_func of the recipe (mysystem.py:83)


class PairwiseNegSISDR(_Loss):
    """asteroid 0.6 `pairwise_neg_sisdr` (loss_func of the recipe, asteroid_librimix_trainer.py:105): -10 log10(rho + EPS)."""

    def __init__(self, EPS=1e-8):
        super().__init__()
        self.inner = PairwiseWSDR("sisdr", take_log=True, EPS=EPS)

    def forward(self, est_targets, targets):
        return -self.inner(est_targets, targets)


pairwise_neg_sisdr = PairwiseNegSISDR()


class PITLossWrapper(torch.nn.Module):
    """asteroid 0.6 `PITLossWrapper(loss_func, pit_from="pw_mtx")`, factorial search (n_src <= 3):
    loss = mean_b min_perm mean_j pw[b, perm(j), j] with pw = loss_func(est, tgt) of shape [B, n_src, n_src]."""

    def __init__(self, loss_func, pit_from="pw_mtx"):
        super().__init__()
        if pit_from != "pw_mtx":
            raise NotImplementedError("the recipe uses pit_from='pw_mtx'")
        self.loss_func = loss_func

    @staticmethod
    def find_best_perm(pw):
        S = pw.shape[-1]
        perms = list(itertools.permutations(range(S)))
        cands = torch.stack([sum(pw[:, p[j], j] for j in range(S)) / S for p in perms], dim=1)      # [B, S!]
        min_loss, idx = cands.min(dim=1)
        return min_loss, torch.tensor(perms, device=pw.device)[idx]

    def forward(self, est_targets, targets, return_est=False, **kwargs):
        pw = self.loss_func(est_targets, targets, **kwargs)
        min_loss, perm = self.find_best_perm(pw)
        mean_loss = min_loss.mean()
        if not return_est:
            return mean_loss
        reordered = torch.stack([torch.index_select(e, 0, p) for e, p in zip(est_targets, perm)])
        return mean_loss, reordered
