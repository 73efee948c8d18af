"""CPU (gloo, world size 2): the data-parallel exchange of the QAT path -- contiguous batch sharding, ONE
all-reduce(sum) x 1/n of the flat gradient arena, zeros for parameters that got no gradient (SURVEY.md 8e, D1;
reference: pl.Trainer(strategy="ddp"), train_env/asteroid_librimix/asteroid_librimix_trainer.py:125-135).
The arena plumbing is device-agnostic; only clip/Adam are CUDA kernels (covered by the -m gpu tests)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _toy(seed=0):
    torch.manual_seed(seed)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    dead = torch.nn.Parameter(torch.ones(4))           # never used in forward: like block 23's res_conv / add
    return net, dead


def _gather_on_host(arena):
    """Host stand-in for fqss_arena_gather (the product gathers on the GPU only): p.grad -> the arena's gradient views."""
    for p, gv in zip(arena.params, arena.grad_views):
        if p.grad is None:
            gv.zero_()
        else:
            gv.copy_(p.grad)


def _worker(rank, world, port, gb, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fqss_b200.parallel import ParamArena, shard_bounds
    net, dead = _toy()
    params = list(net.parameters()) + [dead]
    arena = ParamArena(params)
    g = torch.Generator().manual_seed(7)
    x, y = torch.randn(gb, 6, generator=g), torch.randn(gb, 3, generator=g)
    lo, hi = shard_bounds(gb, rank, world)
    loss = ((net(x[lo:hi]) - y[lo:hi]) ** 2).mean()     # per-rank loss on the local shard (reference DDP semantics)
    loss.backward()
    _gather_on_host(arena)
    scale = arena.allreduce_mean()
    torch.save({"grad": arena.grad.clone() * scale, "scale": scale, "lo": lo, "hi": hi,
                "flat_is_param": all(p.data_ptr() == v.data_ptr() for p, v in zip(params, arena.views))},
               os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds():
    from fqss_b200.parallel import shard_bounds
    assert [shard_bounds(32, r, 8) for r in range(8)] == [(4 * r, 4 * r + 4) for r in range(8)]
    with pytest.raises(ValueError):
        shard_bounds(30, 0, 8)


def test_arena_allreduce_gloo_world2(tmp_path):
    world, gb = 2, 8
    port = _free_port()
    mp.spawn(_worker, args=(world, port, gb, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(str(tmp_path), "rank%d.pt" % r)) for r in range(world)]
    assert outs[0]["scale"] == 0.5 and (outs[0]["lo"], outs[0]["hi"], outs[1]["lo"], outs[1]["hi"]) == (0, 4, 4, 8)
    assert outs[0]["flat_is_param"] and outs[1]["flat_is_param"]
    assert torch.equal(outs[0]["grad"], outs[1]["grad"])           # every rank holds the same averaged arena
    # expectation: mean over ranks of the per-rank gradients (what DDP computes), dead parameter -> zeros
    from fqss_b200.parallel import reference_ddp_step_cpu
    g = torch.Generator().manual_seed(7)
    x, y = torch.randn(gb, 6, generator=g), torch.randn(gb, 3, generator=g)
    per_rank = []
    for r in range(world):
        net, dead = _toy()
        loss = ((net(x[4 * r:4 * r + 4]) - y[4 * r:4 * r + 4]) ** 2).mean()
        loss.backward()
        per_rank.append([p.grad for p in net.parameters()] + [torch.zeros_like(dead)])
    want = torch.cat([t.reshape(-1) for t in reference_ddp_step_cpu(per_rank)])
    got = outs[0]["grad"][:want.numel()]
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-7)
    assert float(got[-4:].abs().max()) == 0.0
    # equal shard sizes + mean loss: the DDP average equals the gradient of the global-batch mean loss
    net, dead = _toy()
    ((net(x) - y) ** 2).mean().backward()
    full = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert torch.allclose(got[:full.numel()], full, rtol=1e-5, atol=1e-6)


def _parity_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fqss_b200 import parallel
    from fqss_b200.qat.qat_quant import GradientActivationFakeQuantize as AQ
    g = torch.Generator().manual_seed(11)
    x = torch.randn(8, 1, 64, generator=g)                  # global batch; rank r owns rows [4r, 4r+4)
    lo, hi = parallel.shard_bounds(8, rank, world)
    xs = x[lo:hi]
    net = torch.nn.Sequential(AQ(True), torch.nn.Identity(), AQ(True))
    res = {}
    for mode in (False, True):
        parallel.set_global_batch_parity(mode)
        for q in net:
            if isinstance(q, AQ):
                q.min_range.data.fill_(-0.5)
                q.max_range.data.fill_(0.5)
        peak = xs.abs().max().reshape(1).clone()
        parallel.sync_splitter_peak_(peak)
        for step in range(3):                               # three calibration passes: EMA of the batch statistics
            for k, q in enumerate(m for m in net if isinstance(m, AQ)):
                t = xs * (k + 1 + step)
                q.min_range.data.mul_(0.9).add_(0.1 * t.min())      # qat_quant.py:230-231 on the local shard
                q.max_range.data.mul_(0.9).add_(0.1 * t.max())
            n = parallel.sync_observer_ranges_(net)
        res[mode] = dict(peak=peak.clone(), n=n, mins=[q.min_range.data.clone() for q in net if isinstance(q, AQ)],
                         maxs=[q.max_range.data.clone() for q in net if isinstance(q, AQ)])
    parallel.set_global_batch_parity(False)
    torch.save(res, os.path.join(out_dir, "parity%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_global_batch_parity_sync_gloo_world2(tmp_path):
    """with global-batch parity on, the splitter peak and the observer ranges of a
    2-rank run equal those of one process on the global batch (bit for bit); off = the reference's per-rank values."""
    world = 2
    port = _free_port()
    mp.spawn(_parity_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(str(tmp_path), "parity%d.pt" % r)) for r in range(world)]
    g = torch.Generator().manual_seed(11)
    x = torch.randn(8, 1, 64, generator=g)
    # single-process reference on the global batch
    mins = [torch.tensor([-0.5]), torch.tensor([-0.5])]
    maxs = [torch.tensor([0.5]), torch.tensor([0.5])]
    for step in range(3):
        for k in range(2):
            t = x * (k + 1 + step)
            mins[k] = mins[k] * 0.9 + 0.1 * t.min()
            maxs[k] = maxs[k] * 0.9 + 0.1 * t.max()
    for r in range(world):
        on, off = outs[r][True], outs[r][False]
        assert on["n"] == 2 and off["n"] == 0
        assert torch.equal(on["peak"], x.abs().max().reshape(1))
        assert torch.equal(off["peak"], x[4 * r:4 * r + 4].abs().max().reshape(1))
        for k in range(2):
            assert torch.equal(on["mins"][k], mins[k]) and torch.equal(on["maxs"][k], maxs[k]), (r, k)
    # without the sync the two ranks have diverged (what the reference's DDP run does)
    assert not torch.equal(outs[0][False]["mins"][0], outs[1][False]["mins"][0])


def test_param_generation_keys_prepared_weight_caches():
    """Arena kernels update parameters in place without touching torch's version counters: caches of prepared weights add
    parallel.param_generation() to their key for parameters that live in an arena, and only for those (the frozen teacher's
    prepared operands must stay cached across steps)."""
    import torch
    from fqss_b200 import float_engine as FE
    from fqss_b200 import parallel
    a, b = torch.nn.Linear(4, 4), torch.nn.Linear(4, 4)
    for p in a.parameters():
        p._fqss_in_arena = True          # what ParamArena.__init__ does (it needs CUDA tensors)
    k_a, k_b = FE._versions(a), FE._versions(b)
    parallel.bump_param_generation()
    assert FE._versions(a) != k_a and FE._versions(b) == k_b
