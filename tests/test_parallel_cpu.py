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
    arena.gather_grads()
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
