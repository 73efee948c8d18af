"""GPU: the fused small-head attention core (csrc/attention.cu; MultiheadAttentionQ, reference qat_layers.py:926-939) against
softmax(q k^T) v in float64 -- forward and all three gradients to 1e-5 -- and the quantised attention layer with the kernel
on / off (same quantiser decisions up to rounding-boundary flips)."""
import pytest
import torch

from parity_log import record

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.mark.parametrize("BH,Lq,Lk,hd", [(12, 250, 250, 16), (7, 37, 53, 8), (5, 300, 129, 32), (3, 1, 5, 16)])
def test_small_head_attention_vs_fp64(BH, Lq, Lk, hd):
    from fqss_b200.qat.qat_layers_seq import SmallHeadAttention
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(BH, L, hd, generator=g).to(DEV) for L in (Lq, Lk, Lk))
    q = q * 1.5
    leaves = [t.clone().requires_grad_(True) for t in (q, k, v)]
    o = SmallHeadAttention.apply(*leaves)
    ref_l = [t.double().requires_grad_(True) for t in (q, k, v)]
    ref = torch.softmax(ref_l[0] @ ref_l[1].transpose(1, 2), -1) @ ref_l[2]
    go = torch.randn(BH, Lq, hd, generator=g).to(DEV)
    o.backward(go)
    ref.backward(go.double())
    meas = dict(fwd=rel(o, ref), dq=rel(leaves[0].grad, ref_l[0].grad), dk=rel(leaves[1].grad, ref_l[1].grad),
                dv=rel(leaves[2].grad, ref_l[2].grad))
    record("attention/BH%d_Lq%d_Lk%d_hd%d" % (BH, Lq, Lk, hd), **meas)
    assert all(x < 1e-5 for x in meas.values()), meas


def test_mha_q_layer_native_vs_library_path():
    import torch.nn as nn
    from fqss_b200.qat import qat_layers_seq as QS
    torch.manual_seed(0)
    mha = nn.MultiheadAttention(64, 4).to(DEV)
    layer = QS.MultiheadAttentionQ(mha, gradient_based=True, weight_quant=True, act_quant=True).to(DEV)
    x = torch.randn(50, 9, 64, device=DEV)
    with torch.no_grad():
        layer(x, x, x); layer(x, x, x)
    for m in layer.modules():
        if hasattr(m, "enable_observer"):
            m.enable_observer(False)
    g = torch.randn(50, 9, 64, device=DEV)

    def run(native):
        QS.NATIVE_ATTENTION = native
        try:
            layer.zero_grad(set_to_none=True)
            xi = x.clone().requires_grad_(True)
            y = layer(xi, xi, xi)[0]
            y.backward(g)
            return y.detach(), xi.grad
        finally:
            QS.NATIVE_ATTENTION = True
    y1, g1 = run(True)
    y2, g2 = run(False)
    q = layer.activation_fake_quantize
    step = float(q.max_range - q.min_range) / 255
    d = (y1 - y2).abs() / step
    meas = dict(flip_rate=(d > 0.5).float().mean().item(), max_steps=d.max().item(), gx_rel=rel(g1, g2))
    record("attention/mha_q_layer", **meas)
    assert meas["flip_rate"] < 5e-3 and meas["max_steps"] < 1.01 and meas["gx_rel"] < 2e-2, meas
