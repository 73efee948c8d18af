"""GPU: kernel-level parity of the C-ABI ops against the golden vectors of the UNMODIFIED reference
and against the CPU oracle.  Integer codes / quantised values are bit-exact; reductions (range
gradients, weight gradients) agree to fp32 summation round-off."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import fqss_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def T(a):
    return torch.from_numpy(np.asarray(a))


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def test_act_fake_quant_bit_exact(golden):
    from fqss_b200 import ops
    g = golden("q_ops.npz")
    for ci in range(6):
        x = T(g[f"act{ci}_x"]).to(DEV).requires_grad_(True)
        lo, hi = g[f"act{ci}_range"]
        rmin = torch.tensor([lo], device=DEV, requires_grad=True)
        rmax = torch.tensor([hi], device=DEV, requires_grad=True)
        y = ops.FakeQuantAct.apply(x, rmin, rmax, 8)
        _, code = ops.fake_quant_codes(x.detach(), rmin.detach(), rmax.detach())
        assert torch.equal(code.cpu(), T(g[f"act{ci}_code"])), "codes differ (case %d)" % ci
        assert torch.equal(y.detach().cpu(), T(g[f"act{ci}_y"]))
        y.backward(T(g[f"act{ci}_go"]).to(DEV))
        assert torch.equal(x.grad.cpu(), T(g[f"act{ci}_gx"]))
        for got, want in ((rmin.grad, g[f"act{ci}_gmin"]), (rmax.grad, g[f"act{ci}_gmax"])):
            assert abs(got.item() - float(want[0])) <= 2e-5 * max(1.0, abs(float(want[0]))), (ci, got.item(), want)


def test_weight_fake_quant_bit_exact(golden):
    from fqss_b200 import ops
    g = golden("q_ops.npz")
    for ci in range(4):
        axis = int(g[f"w{ci}_axis"])
        w = T(g[f"w{ci}_w"]).to(DEV).requires_grad_(True)
        rmin = T(g[f"w{ci}_min"]).to(DEV).requires_grad_(True)
        rmax = T(g[f"w{ci}_max"]).to(DEV).requires_grad_(True)
        y = ops.FakeQuantWeight.apply(w, rmin, rmax, axis, 8)
        assert torch.equal(ops.weight_codes(w.detach(), rmin.detach(), rmax.detach(), axis).cpu(), T(g[f"w{ci}_code"]))
        assert torch.equal(y.detach().cpu(), T(g[f"w{ci}_y"]))
        y.backward(T(g[f"w{ci}_go"]).to(DEV))
        assert torch.equal(w.grad.cpu(), T(g[f"w{ci}_gw"]))
        assert torch.allclose(rmin.grad.cpu(), T(g[f"w{ci}_gmin"]), rtol=1e-4, atol=1e-6)
        assert torch.allclose(rmax.grad.cpu(), T(g[f"w{ci}_gmax"]), rtol=1e-4, atol=1e-6)


def test_observers(golden):
    from fqss_b200.qat.qat_quant import GradientActivationFakeQuantize, GradientWeightFakeQuantize
    g = golden("q_ops.npz")
    q = GradientActivationFakeQuantize(True).to(DEV)
    for x, (emin, emax) in zip(T(g["obs_x"]), g["obs_trace"]):
        xd = x.to(DEV)
        assert q(xd) is xd                               # observer passes data through (qat_quant.py:233)
        assert q.min_range.item() == emin and q.max_range.item() == emax
    w = T(g["w0_w"]).to(DEV)
    wq = GradientWeightFakeQuantize(True, w.shape).to(DEV)
    assert wq(w) is w and wq.observer_mode is False
    lo, hi = O.observe_weight(w.cpu(), 0)
    assert torch.equal(wq.min_range.detach().cpu(), lo) and torch.equal(wq.max_range.detach().cpu(), hi)


def test_splitter_reconstructor_bit_exact(golden):
    from fqss_b200.process import postprocess, preprocess
    g = golden("split.npz")
    x = T(g["x"]).to(DEV)
    assert torch.equal(preprocess(x, n_splitter=2).cpu(), T(g["y2"]))
    assert torch.equal(preprocess(x, n_splitter=3).cpu(), T(g["y3"]))
    dec = T(g["dec"]).to(DEV).requires_grad_(True)
    z = postprocess(dec, n_combiner=2)
    assert torch.equal(z.detach().cpu(), T(g["z"]))
    z.sum().backward()
    assert torch.equal(dec.grad[0].squeeze(-2).cpu(), torch.ones_like(z).cpu())
    assert torch.allclose(dec.grad[1].squeeze(-2).cpu(), torch.full_like(z, 0.5 / 128).cpu())
    # ragged / odd lengths and a 2-D input
    x2 = torch.randn(3, 1001, device=DEV)
    assert torch.equal(preprocess(x2, n_splitter=2).cpu(), O.split_input(x2.cpu(), 2))


def test_kd_loss_golden(golden):
    from fqss_b200.losses import fqss_kd_loss
    g = golden("loss.npz")
    est = T(g["est"]).to(DEV).requires_grad_(True)
    loss, kd, val = fqss_kd_loss(est, T(g["fest"]).to(DEV), T(g["tgt"]).to(DEV), 0.1)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 2e-4
    assert rel(est.grad, T(g["gest"])) < 1e-4
    _, per_b = O.neg_sisdr_db_pit(T(g["est"]), T(g["tgt"]))
    assert abs(val.item() - per_b.mean().item()) < 2e-4


@pytest.mark.parametrize("B,T_", [(1, 257), (5, 4001)])
def test_kd_loss_vs_oracle_random(B, T_):
    from fqss_b200.losses import fqss_kd_loss
    gen = torch.Generator().manual_seed(B)
    tgt = torch.randn(B, 2, T_, generator=gen) * 0.1 + 0.02
    est = (tgt[:, [1, 0]] * 0.9 + 0.03 * torch.randn(B, 2, T_, generator=gen)).requires_grad_(True)
    fest = tgt[:, [1, 0]] + 0.01 * torch.randn(B, 2, T_, generator=gen)
    lo, kdo = O.fqss_kd_loss(est, fest, tgt, 0.1)
    lo.backward()
    e2 = est.detach().to(DEV).requires_grad_(True)
    l, kd, _ = fqss_kd_loss(e2, fest.to(DEV), tgt.to(DEV), 0.1)
    l.backward()
    assert abs(l.item() - lo.item()) < 2e-4 and abs(kd.item() - kdo.item()) < 2e-4
    assert rel(e2.grad, est.grad) < 2e-4


def _pw_case(kind, shape, seed=0, quant=True):
    """run one pointwise+FQ layer on GPU and through the oracle's torch ops on CPU"""
    from fqss_b200 import _native as N
    from fqss_b200 import ops
    gen = torch.Generator().manual_seed(seed)
    x1 = torch.randn(shape, generator=gen)
    x2 = None
    slope = torch.tensor([0.25])
    C = shape[-2]
    gamma, beta = torch.rand(C, generator=gen) + 0.5, torch.randn(C, generator=gen) * 0.1
    if kind in (N.PW_ADD, N.PW_SUB):
        x2 = torch.randn(shape, generator=gen)
    if kind == N.PW_MUL:
        x2 = torch.rand((shape[0], 1) + tuple(shape[2:]), generator=gen)
    rmin, rmax = torch.tensor([-1.3]), torch.tensor([1.9])
    go = torch.randn(shape, generator=gen)

    def leaf(t):
        return None if t is None else t.clone().requires_grad_(True)
    c = [leaf(t) for t in (x1, x2, slope, gamma, beta, rmin, rmax)]
    z = {N.PW_IDENT: lambda: c[0], N.PW_PRELU: lambda: F.prelu(c[0], c[2]), N.PW_RELU: lambda: F.relu(c[0]),
         N.PW_ADD: lambda: c[0] + c[1], N.PW_SUB: lambda: c[0] - c[1], N.PW_MUL: lambda: c[0] * c[1],
         N.PW_GLN: lambda: F.group_norm(c[0], 1, c[3], c[4], 1e-8)}[kind]()
    yo = O.fq_act(z, c[5], c[6]) if quant else z
    yo.backward(go)
    d = [None if t is None else t.clone().to(DEV).requires_grad_(True) for t in (x1, x2, slope, gamma, beta, rmin, rmax)]
    y = ops.pointwise_fq(kind, d[0], d[1], d[2] if kind == N.PW_PRELU else None, d[3] if kind == N.PW_GLN else None,
                         d[4] if kind == N.PW_GLN else None, d[5], d[6], quant, 8, 1e-8)
    y.backward(go.to(DEV))
    return z, yo, c, y, d


@pytest.mark.parametrize("shape", [(2, 8, 999), (1, 3, 4100), (3, 16, 64)])
def test_pointwise_fq_layers_vs_oracle(shape):
    from fqss_b200 import _native as N
    for kind in (N.PW_IDENT, N.PW_PRELU, N.PW_RELU, N.PW_ADD, N.PW_SUB, N.PW_GLN):
        for quant in (True, False):
            z, yo, c, y, d = _pw_case(kind, shape, seed=kind, quant=quant)
            if kind == N.PW_GLN:
                # statistics are reduced in a different order: allow rare +-1 code flips
                step = (1.9 + 1.3) / 255
                diff = (y.detach().cpu() - yo.detach()).abs()
                assert diff.max() <= step * 1.001 + 1e-5 and (diff > 1e-5).float().mean() < 2e-3
            else:
                assert torch.equal(y.detach().cpu(), yo.detach()), (kind, quant)
            assert rel(d[0].grad, c[0].grad) < (2e-3 if kind == N.PW_GLN else 1e-6), (kind, quant)
            if kind in (N.PW_ADD, N.PW_SUB):
                assert rel(d[1].grad, c[1].grad) < 1e-6
            if kind == N.PW_PRELU:
                assert rel(d[2].grad, c[2].grad) < 1e-4
            if kind == N.PW_GLN:
                assert rel(d[3].grad, c[3].grad) < 2e-3 and rel(d[4].grad, c[4].grad) < 2e-3
            if quant:
                # the oracle's own range-gradient sums are fp32 (ours fp64): agreement to fp32 summation error
                tol = 5e-3 if kind == N.PW_GLN else 1e-3
                assert rel(d[5].grad, c[5].grad) < tol and rel(d[6].grad, c[6].grad) < tol, (kind,)


def test_mulq_broadcast_vs_oracle():
    from fqss_b200 import _native as N
    z, yo, c, y, d = _pw_case(N.PW_MUL, (2, 2, 8, 999), seed=5)
    assert torch.equal(y.detach().cpu(), yo.detach())
    assert rel(d[0].grad, c[0].grad) < 1e-6 and rel(d[1].grad, c[1].grad) < 1e-5
    assert rel(d[5].grad, c[5].grad) < 1e-4 and rel(d[6].grad, c[6].grad) < 1e-4


def _grads(fn, tensors, go):
    leaves = [t.clone().requires_grad_(True) if t is not None else None for t in tensors]
    y = fn(*leaves)
    y.backward(go)
    return y.detach(), [None if l is None else l.grad for l in leaves]


@pytest.mark.parametrize("B,Ci,Co,M", [(2, 32, 64, 999), (1, 128, 512, 300), (3, 24, 8, 131)])
def test_conv1x1_vs_torch(B, Ci, Co, M):
    from fqss_b200 import ops
    gen = torch.Generator().manual_seed(M)
    x, w, b = torch.randn(B, Ci, M, generator=gen), torch.randn(Co, Ci, 1, generator=gen) * 0.1, torch.randn(Co, generator=gen)
    go = torch.randn(B, Co, M, generator=gen)
    yo, gso = _grads(lambda x, w, b: F.conv1d(x, w, b), (x, w, b), go)
    y, gs = _grads(lambda x, w, b: ops.Conv1x1.apply(x, w, b), (x.to(DEV), w.to(DEV), b.to(DEV)), go.to(DEV))
    assert rel(y, yo) < 1e-5
    for a, bb in zip(gs, gso):
        assert rel(a, bb) < 1e-5


@pytest.mark.parametrize("dil,M", [(1, 1000), (4, 1000), (128, 1000), (2, 1003), (128, 999), (512, 1999), (8, 6)])
def test_depthwise_vs_torch(dil, M):
    """Per-layer depthwise conv (vectorised K = 3 kernels on aligned rows; ragged row ends, dilations that are / are not a
    multiple of the vector width, halos longer than the row)."""
    from fqss_b200 import ops
    B, C = 2, 16
    gen = torch.Generator().manual_seed(dil)
    x, w, b = torch.randn(B, C, M, generator=gen), torch.randn(C, 1, 3, generator=gen), torch.randn(C, generator=gen)
    go = torch.randn(B, C, M, generator=gen)
    yo, gso = _grads(lambda x, w, b: F.conv1d(x, w, b, padding=dil, dilation=dil, groups=C), (x, w, b), go)
    y, gs = _grads(lambda x, w, b: ops.DepthwiseConv.apply(x, w, b, dil), (x.to(DEV), w.to(DEV), b.to(DEV)), go.to(DEV))
    assert rel(y, yo) < 1e-6
    for a, bb in zip(gs, gso):
        assert rel(a, bb) < 1e-5


@pytest.mark.parametrize("Cin,B,Co,T_,K,s", [(1, 3, 48, 1208, 16, 8), (2, 3, 48, 1208, 16, 8), (2, 2, 512, 32000, 16, 8),
                                             (1, 2, 512, 32000, 16, 8), (1, 2, 40, 1003, 16, 8), (1, 2, 24, 700, 10, 5),
                                             (4, 2, 44, 30007, 20, 10)])      # the music encoder's geometry: tiled generic wgrad
def test_strided_and_transposed_conv_vs_torch(Cin, B, Co, T_, K, s):
    """Encoder / RQB re-encoder / decoder filterbank convs (tiled kernels for kernel 16 / hop 8, generic otherwise):
    forward, input gradient and weight gradient against torch fp32; ragged tails (T not a multiple of the hop)."""
    from fqss_b200 import ops
    gen = torch.Generator().manual_seed(Cin)
    x, w = torch.randn(B, Cin, T_, generator=gen), torch.randn(Co, Cin, K, generator=gen) * 0.1
    Mo = (T_ - K) // s + 1
    go = torch.randn(B, Co, Mo, generator=gen)
    yo, gso = _grads(lambda x, w: F.conv1d(x, w, None, stride=s), (x, w), go)
    xin = x.to(DEV)
    if Cin == 1:
        y, gs = _grads(lambda x, w: ops.StridedConv.apply(x, w, s), (xin, w.to(DEV)), go.to(DEV))
        assert rel(gs[0], gso[0]) < 1e-5
    else:   # main encoder: the input needs no gradient
        wl = w.to(DEV).requires_grad_(True)
        yy = ops.StridedConv.apply(xin, wl, s)
        yy.backward(go.to(DEV))
        y, gs = yy.detach(), [None, wl.grad]
    assert rel(y, yo) < 1e-5 and rel(gs[1], gso[1]) < 1e-5
    # decoder
    X, wd = torch.randn(B, Co, Mo, generator=gen), torch.randn(Co, 1, K, generator=gen) * 0.1
    god = torch.randn(B, 1, (Mo - 1) * s + K, generator=gen)
    yo, gso = _grads(lambda x, w: F.conv_transpose1d(x, w, None, stride=s), (X, wd), god)
    y, gs = _grads(lambda x, w: ops.TransposedConv1.apply(x, w, s), (X.to(DEV), wd.to(DEV)), god.to(DEV))
    assert rel(y, yo) < 1e-5 and rel(gs[0], gso[0]) < 1e-5 and rel(gs[1], gso[1]) < 1e-5


def test_fq_act_full_size_properties():
    """BASELINE-size tensor (B=32 x 512 x 3999 would be 262 MB; use B=8): idempotence, code range,
    monotonicity -- size-independent properties of the quantiser."""
    from fqss_b200 import ops
    x = torch.randn(8 * 512 * 3999, device=DEV)
    rmin, rmax = torch.tensor([-2.0], device=DEV), torch.tensor([3.0], device=DEV)
    y, code = ops.fake_quant_codes(x, rmin, rmax)
    y2, code2 = ops.fake_quant_codes(y, rmin, rmax)
    assert torch.equal(code, code2) and torch.equal(y, y2)                   # FQ(FQ(x)) == FQ(x)
    assert int(code.max()) == 255 and int(code.min()) == 0
    xs, idx = torch.sort(x[: 1 << 20])
    assert bool((code[: 1 << 20][idx].to(torch.int16).diff() >= 0).all())    # monotone in x
    step = 5.0 / 255
    inside = (x > -2.0) & (x < 3.0)
    assert float((y - x)[inside].abs().max()) <= step / 2 * 1.0001


@pytest.mark.parametrize("B,K,N,M", [(2, 128, 512, 3999), (1, 512, 256, 1000), (3, 64, 128, 100), (2, 128, 1024, 257)])
def test_tcgen05_pw_gemm_exact_integer_codes(B, K, N, M):
    """The tensor-core 1x1 conv on integer fake-quant codes: every product / partial sum is an integer
    < 2^24, so the fp32 accumulator is exact and must equal an int64 reference bit for bit."""
    from fqss_b200 import tcn_engine as E
    gen = torch.Generator().manual_seed(K + N)
    ld = (M + 7) // 8 * 8
    act = torch.randint(0, 256, (B, K, ld), generator=gen)
    w = torch.randint(-128, 128, (N, K), generator=gen)
    ref = torch.einsum("ok,bkm->bom", w.double(), act[:, :, :M].double())
    s1 = torch.ones(N, device=DEV)
    s0 = torch.zeros(N, device=DEV)
    out = E.pw_gemm(act.to(DEV).to(torch.bfloat16), w.to(DEV).to(torch.bfloat16), s1, s0, M)
    assert torch.equal(out[:, :, :M].double().cpu(), ref)
    # affine + bf16 output + addend variants on real-valued operands
    actf = (torch.randn(B, K, ld, generator=gen)).to(torch.bfloat16)
    wf = (torch.randn(N, K, generator=gen) * 0.1).to(torch.bfloat16)
    s1 = torch.rand(N, generator=gen) + 0.5
    s0 = torch.randn(N, generator=gen)
    add = torch.randn(B, N, ld, generator=gen)
    ref = torch.einsum("ok,bkm->bom", wf.double(), actf[:, :, :M].double()) * s1.double()[None, :, None] + s0.double()[None, :, None]
    o1 = E.pw_gemm(actf.to(DEV), wf.to(DEV), s1.to(DEV), s0.to(DEV), M)
    assert rel(o1[:, :, :M], ref) < 1e-5
    o2 = E.pw_gemm(actf.to(DEV), wf.to(DEV), s1.to(DEV), s0.to(DEV), M, addend=add.to(DEV))
    assert rel(o2[:, :, :M], ref + add[:, :, :M].double()) < 1e-5
    o3 = E.pw_gemm(actf.to(DEV), wf.to(DEV), s1.to(DEV), s0.to(DEV), M, out_dtype=torch.bfloat16)
    assert rel(o3[:, :, :M].float(), ref) < 5e-3


def test_param_arena_gather_clip_adam_vs_torch():
    """D1 tail (asteroid_librimix_trainer.py:94,132): gather of ragged gradient tensors into the flat arena (incl. a
    parameter without gradient), global-norm clip 5.0 and Adam(lr 1e-3), against torch's own clip + Adam."""
    from fqss_b200.parallel import ParamArena
    torch.manual_seed(0)
    shapes = [(1,), (512, 128, 1), (3,), (512, 1, 3), (7, 5), (1,), (4097,), (128,), (2, 3, 4)] * 60      # 540 tensors: 2 gather launches
    ref = [torch.nn.Parameter(torch.randn(s, device=DEV)) for s in shapes]
    ours = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    arena = ParamArena(ours)
    opt = torch.optim.Adam(ref, lr=1e-3)
    for step in range(3):
        grads = [torch.randn(s, device=DEV) * (3.0 if step == 0 else 0.01) for s in shapes]
        arena.zero_grad()
        opt.zero_grad(set_to_none=True)
        for i, (p, q, g) in enumerate(zip(ref, ours, grads)):
            if i % 11 == 5:
                continue                                    # no gradient this step
            p.grad = g.clone()
            q.grad = g.t().contiguous().t() if g.dim() == 2 else g.clone()      # one non-contiguous gradient layout
        arena.gather_grads()
        for i, gv in enumerate(arena.grad_views):
            want = ref[i].grad if ref[i].grad is not None else torch.zeros_like(ref[i])
            assert torch.equal(gv, want), i
        scale = arena.allreduce_mean()
        arena.clip_and_step(pre_scale=scale, max_norm=5.0, lr=1e-3)
        for p in ref:                                       # torch steps only parameters that have a gradient; the
            if p.grad is None:                              # reference (DDP find_unused) feeds zeros instead
                p.grad = torch.zeros_like(p)
        torch.nn.utils.clip_grad_norm_(ref, 5.0)
        opt.step()
        worst = max(rel(q, p) for p, q in zip(ref, ours))
        assert worst < 1e-5, (step, worst)


def test_param_arena_device_lr_under_cuda_graph():
    """The learning rate lives on the device (ParamArena.set_lr): a scheduler's change takes effect in a REPLAYED CUDA graph
    (the recipe's half_lr = ReduceLROnPlateau, asteroid_librimix_trainer.py:96-97); checked against torch.optim.Adam stepped
    with the same schedule."""
    from fqss_b200.parallel import ParamArena
    torch.manual_seed(0)
    ref = [torch.nn.Parameter(torch.randn(s, device=DEV)) for s in [(1000,), (33, 7)]]
    ours = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    arena = ParamArena(ours)
    arena.set_lr(1e-3)
    gs = [torch.randn_like(p) * 0.01 for p in ref]
    static_g = [g.clone() for g in gs]

    def step():
        for q, g in zip(ours, static_g):
            q.grad = g
        arena.gather_grads()
        arena.clip_and_step(pre_scale=1.0, max_norm=5.0, lr=123.0)      # the scalar is ignored once set_lr was called

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    opt = torch.optim.Adam(ref, lr=1e-3)
    lrs = [1e-3, 5e-4, 5e-4, 2.5e-4]          # the warm-up step ran for real (capturing executes nothing), then three replays
    for i, lr in enumerate(lrs):
        if i >= 1:
            arena.set_lr(lr)
            graph.replay()
        for grp in opt.param_groups:
            grp["lr"] = lr
        for p, g in zip(ref, gs):
            p.grad = g.clone()
        torch.nn.utils.clip_grad_norm_(ref, 5.0)
        opt.step()
    torch.cuda.synchronize()
    worst = max(rel(q, p) for p, q in zip(ref, ours))
    assert worst < 1e-5, worst


def test_pairwise_wsdr_module_vs_reference_and_oracle(golden):
    """S2/S3 as standalone modules (wsdr.py:46-95 and asteroid's PITLossWrapper): the pairwise matrices of the
    unmodified reference (tests/golden/loss.npz), gradients against the oracle's autograd, and the PIT search."""
    from fqss_b200.wsdr import PITLossWrapper, PairwiseWSDR, pairwise_neg_sisdr, pairwise_wsisdr
    g = golden("loss.npz")
    est, tgt, w = T(g["est"]), T(g["tgt"]), T(g["w"])
    pw_lin = PairwiseWSDR("sisdr", take_log=False)(est.to(DEV), tgt.to(DEV), w.to(DEV))
    pw_log = PairwiseWSDR("sisdr", take_log=True)(est.to(DEV), tgt.to(DEV))
    assert rel(pw_lin, T(g["pw_lin"])) < 1e-5 and rel(pw_log, T(g["pw_log"])) < 1e-5
    # gradient of a generic scalar of the matrix, against autograd through the oracle's restatement
    gen = torch.Generator().manual_seed(4)
    coeff = torch.randn(3, 2, 2, generator=gen)
    e_o = est.clone().requires_grad_(True)
    (O.pairwise_sisdr_ratio(e_o, tgt, w) * coeff).sum().backward()
    e_c = est.to(DEV).requires_grad_(True)
    (-pairwise_wsisdr(e_c, tgt.to(DEV), w.to(DEV)) * coeff.to(DEV)).sum().backward()
    assert rel(e_c.grad, e_o.grad) < 1e-4, rel(e_c.grad, e_o.grad)
    # the recipe's validation loss: PITLossWrapper(pairwise_neg_sisdr, pit_from="pw_mtx")
    pit = PITLossWrapper(pairwise_neg_sisdr, pit_from="pw_mtx")
    e_c2 = est.to(DEV).requires_grad_(True)
    loss, reordered = pit(e_c2, tgt.to(DEV), return_est=True)
    loss.backward()
    e_o2 = est.clone().requires_grad_(True)
    lo, per_b = O.neg_sisdr_db_pit(e_o2, tgt)
    lo.backward()
    assert abs(loss.item() - lo.item()) < 2e-4 and rel(e_c2.grad, e_o2.grad) < 1e-4
    assert torch.equal(reordered.cpu(), est[:, [1, 0]])            # the golden estimates are the targets, permuted
    with pytest.raises(NotImplementedError):
        PairwiseWSDR("snr")


def test_fake_quant_act_unaligned_contiguous_view():
    """A contiguous view that starts in the middle of a buffer (4-byte aligned only) is repacked, not refused (ADVICE)."""
    from fqss_b200 import ops
    base = torch.randn(4101, device=DEV)
    x = base[1:].requires_grad_(True)                 # contiguous, data_ptr % 16 == 4
    assert x.is_contiguous() and x.data_ptr() % 16 != 0
    rmin, rmax = torch.tensor([-1.5], device=DEV, requires_grad=True), torch.tensor([2.0], device=DEV, requires_grad=True)
    y = ops.FakeQuantAct.apply(x, rmin, rmax, 8)
    ref = ops.FakeQuantAct.apply(base[1:].clone().requires_grad_(True), rmin, rmax, 8)
    assert torch.equal(y, ref)
    g = torch.randn(4101, device=DEV)[1:]
    y.backward(g)
    assert x.grad is not None and torch.isfinite(x.grad).all()
