"""Data-parallel QAT plumbing: one process per GPU, batch sharded
contiguously across ranks, ONE collective per step -- an all-reduce (sum, then 1/n) of a flat fp32
gradient arena holding all 948 parameter tensors (5 133 123 floats = 20.5 MB) -- followed by the
global-norm clip (gradient_clip_val=5.0) and Adam on the arena, both as single sm_100a kernels.

Reference: pl.Trainer(strategy="ddp", gradient_clip_val=5.0) + make_optimizer(adam, lr 1e-3)
(train_env/asteroid_librimix/asteroid_librimix_trainer.py:94,125-135).  Parameters that receive no
gradient (block 23's dead res_conv/add) stay zero in the arena, which is what DDP's
find_unused_parameters gives the reference.
"""
import torch
import torch.distributed as dist

from . import _native as N
from ._native import check, lib, ptr, stream_ptr, workspace


# ---------------------------------------------------------------------------------------------
# Batch-coupled quirks of the reference under data parallelism.  DEFAULT = the reference's own DDP
# behaviour (every rank normalises / calibrates on its local shard).  With `set_global_batch_parity(True)` a sharded
# run reproduces the single-process run on the global batch instead:
#   (1) splitter: `x / max|x|` uses ONE scalar over the whole batch (process.py:23-24)  -> all-reduce(MAX) of 1 float
#   (2) observers: EMA of the batch min / max (qat_quant.py:230-232); DDP never re-syncs the range parameters, so the
#       reference's ranks silently diverge                                              -> all-reduce(MIN) of
#       [min_range ; -max_range] after every calibration forward (the EMA is monotone in the batch statistic, so
#       min over ranks of the updated value == the update with the global statistic, bit for bit)
#   (3) loss: `-10 log10(mean_b(...))` (mysystem.py:145) is not a mean of per-sample terms, so DDP's mean of per-rank
#       gradients is not the gradient of the global-batch loss                         -> all-reduce(SUM) of the three
#       local means {kd, task, val} (3 doubles) between the statistics pass and the gradient pass of the loss kernel
#       (ops.KDLoss -> fqss_kd_loss_dp); per-sample weights keep 1/(2 B_local), so the arena's mean over ranks is exact.
# With all three on, an n-rank step equals the one-process step on the concatenated batch up to fp32 summation order
# (`bench.py --verify-dp`, run by default at world size > 1, reports the measured difference).
# ---------------------------------------------------------------------------------------------
_GLOBAL_PARITY = False


# Parameters updated by the arena kernels change IN PLACE without touching torch's version counters; caches of prepared
# weights (float_engine / tcn_engine: split-bf16 operands of un-quantised convs) add this counter to their key.  Bumped by
# every optimizer step, eager or replayed (graph.GraphedStep calls it per replay).
_PARAM_GENERATION = 0


def param_generation():
    return _PARAM_GENERATION


def bump_param_generation():
    global _PARAM_GENERATION
    _PARAM_GENERATION += 1


def set_global_batch_parity(on=True):
    global _GLOBAL_PARITY
    _GLOBAL_PARITY = bool(on)


def global_batch_parity():
    return _GLOBAL_PARITY and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def sync_splitter_peak_(peak, group=None):
    """(1): in-place all-reduce(MAX) of the splitter's `max|x|` scalar when global-batch parity is on."""
    if global_batch_parity():
        dist.all_reduce(peak, op=dist.ReduceOp.MAX, group=group)
    return peak


def sync_observer_ranges_(model, group=None):
    """(2): make the activation-quantiser ranges of all ranks equal to what one process would have learned from the
    global batch.  Call after every calibration forward (observer mode).  One collective over 2 x (#quantisers)
    floats.  Returns the number of quantisers synchronised."""
    from .qat.qat_quant import GradientActivationFakeQuantize as AQ
    qs = [m for m in model.modules() if isinstance(m, AQ)]
    if not qs or not global_batch_parity():
        return 0
    flat = torch.cat([q.min_range.data.reshape(-1) for q in qs] + [-q.max_range.data.reshape(-1) for q in qs])
    dist.all_reduce(flat, op=dist.ReduceOp.MIN, group=group)
    n = len(qs)
    for i, q in enumerate(qs):
        q.min_range.data.copy_(flat[i:i + 1].reshape(q.min_range.shape))
        q.max_range.data.copy_((-flat[n + i:n + i + 1]).reshape(q.max_range.shape))
    return n


def shard_bounds(global_batch, rank, world):
    """Contiguous shard [lo, hi) of the global batch owned by `rank`."""
    if global_batch % world != 0:
        raise ValueError("global batch %d is not divisible by world size %d" % (global_batch, world))
    per = global_batch // world
    return rank * per, (rank + 1) * per


class ParamArena:
    """Flat fp32 storage for parameters, gradients and Adam moments.  `p.data` of every parameter is
    re-pointed at a view into the arena so optimiser and collective work on one buffer."""

    def __init__(self, params):
        self.params = [p for p in params]
        for p in self.params:
            p._fqss_in_arena = True       # caches of prepared weights key on param_generation() for these (float_engine._generation)
        self.numel = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        pad = (-self.numel) % 4
        self.flat = torch.zeros(self.numel + pad, device=dev)
        self.grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.sumsq = torch.zeros(1, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.step_count = 0
        self.views, self.grad_views = [], []
        off = 0
        for p in self.params:
            n = p.numel()
            v = self.flat[off:off + n].view(p.shape)
            v.copy_(p.data)
            p.data = v
            self.views.append(v)
            self.grad_views.append(self.grad[off:off + n].view(p.shape))
            off += n
        self._keep = []
        self._gather_items = (N.GatherItem * len(self.params))()
        off = 0
        for i, p in enumerate(self.params):
            self._gather_items[i].offset, self._gather_items[i].numel = off, p.numel()
            off += p.numel()

    def gather_grads(self):
        """Collect p.grad of every parameter into the gradient arena (zeros where a parameter got none): one
        fqss_arena_gather call (3 launches for the 948 tensors) instead of a chunked multi-tensor ATen copy."""
        if not self.flat.is_cuda:
            raise N.FqssError("ParamArena.gather_grads needs CUDA (no CPU fallback)")
        items = self._gather_items
        for i, p in enumerate(self.params):
            g = p.grad
            if g is None:
                items[i].src = None
            else:
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.float().contiguous()
                    self._keep.append(g)
                items[i].src = g.data_ptr()
        check(lib().fqss_arena_gather(items, len(self.params), ptr(self.grad), stream_ptr()))
        self._keep.clear()          # safe: same stream, the copies are already enqueued

    def allreduce_mean(self, group=None):
        """The one exchange step of the path.  Returns the 1/world factor still to be applied (folded
        into the clip kernel so the arena is touched once)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)
            return 1.0 / dist.get_world_size(group)
        return 1.0

    def set_lr(self, lr):
        """Learning rate kept on the device: `clip_and_step` then reads it at run time, so a scheduler (the recipe's
        ReduceLROnPlateau `half_lr`, asteroid_librimix_trainer.py:96-97) takes effect under a replayed CUDA graph too."""
        if getattr(self, "lr_dev", None) is None:
            self.lr_dev = torch.empty(1, device=self.flat.device)
        self.lr_dev.fill_(float(lr))

    def clip_and_step(self, pre_scale=1.0, max_norm=5.0, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if not self.flat.is_cuda:
            raise N.FqssError("ParamArena.clip_and_step needs CUDA (no CPU fallback)")
        n = self.flat.numel()
        ws = workspace(0, self.flat.device)
        s = stream_ptr()
        check(lib().fqss_arena_sumsq(ptr(self.grad), n, ptr(self.sumsq), ptr(ws), ws.numel(), s))
        check(lib().fqss_arena_scale_clip(ptr(self.grad), n, ptr(self.sumsq), float(pre_scale), float(max_norm), s))
        self.step_count += 1          # host mirror; the kernels read the device counter (CUDA-graph replays stay correct)
        bump_param_generation()       # prepared-weight caches keyed on torch's version counters would otherwise go stale
        check(lib().fqss_arena_adam_dev(ptr(self.flat), ptr(self.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq), n, float(lr),
                                        float(betas[0]), float(betas[1]), float(eps), ptr(self.step_dev),
                                        ptr(getattr(self, "lr_dev", None)) or None, s))

    def zero_grad(self):
        for p in self.params:
            p.grad = None


def reference_ddp_step_cpu(per_rank_grads):
    """Semantics of the exchange on plain tensors (used by the gloo test): mean over ranks."""
    n = len(per_rank_grads)
    return [sum(gs) / n for gs in zip(*per_rank_grads)]
