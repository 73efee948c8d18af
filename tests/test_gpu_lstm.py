"""GPU: the register-resident integer-code LSTM recurrence (csrc/lstm.cu; LSTMQ, reference qat_layers.py:571-613) against
torch's own LSTM evaluated with the same fake-quantised weights in float64: forward 1e-5, every gradient (input, both weight
matrices, both biases, all four weight-quantiser ranges per direction) 1e-4 -- same arithmetic, different summation order."""
import pytest
import torch
import torch.nn as nn

from parity_log import record

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.mark.parametrize("T,nb,I,H,bidir", [(37, 13, 32, 32, True), (20, 8, 64, 128, True), (50, 21, 16, 64, False)])
def test_lstmq_native_recurrence_vs_torch(T, nb, I, H, bidir):
    from fqss_b200.qat import qat_layers_seq as QS
    torch.manual_seed(0)
    lstm = nn.LSTM(I, H, 1, bidirectional=bidir).to(DEV)
    layer = QS.LSTMQ(lstm, gradient_based=True, weight_quant=True, act_quant=False).to(DEV)
    x = torch.randn(T, nb, I, device=DEV)
    with torch.no_grad():
        layer(x)                                                    # first call: the weight quantisers capture their ranges
    assert layer._native_ok(x)
    g = torch.randn(T, nb, (2 if bidir else 1) * H, device=DEV)

    def run(native):
        QS.NATIVE_LSTM = native
        try:
            layer.zero_grad(set_to_none=True)
            xi = x.clone().requires_grad_(True)
            y = layer(xi)[0]
            y.backward(g)
            return y.detach(), xi.grad, {k: p.grad.clone() for k, p in layer.named_parameters() if p.grad is not None}
        finally:
            QS.NATIVE_LSTM = True

    y1, gx1, pg1 = run(True)
    # reference: torch's LSTM in float64 on the same fake-quantised weights (autograd through the weight quantisers in fp32)
    layer.zero_grad(set_to_none=True)
    xi = x.clone().requires_grad_(True)
    flat = []
    for name in lstm._flat_weights_names:
        w = getattr(lstm, name)
        flat.append((layer.weight_quantizers_dict[name](w) if name.startswith("weight") else w).double())
    D = 2 if bidir else 1
    h0 = torch.zeros(D, nb, H, dtype=torch.float64, device=DEV)
    y2 = torch._VF.lstm(xi.double(), (h0, h0.clone()), flat, True, 1, 0.0, True, bidir, False)[0]
    y2.backward(g.double())
    pg2 = {k: p.grad.clone() for k, p in layer.named_parameters() if p.grad is not None}
    meas = dict(fwd=rel(y1, y2), gx=rel(gx1, xi.grad))
    assert set(pg1) == set(pg2)
    for k in pg1:
        meas["g/" + k] = rel(pg1[k], pg2[k])
    record("lstm/native_T%d_N%d_H%d_D%d" % (T, nb, H, D), **meas)
    assert meas["fwd"] < 1e-5 and meas["gx"] < 1e-4, meas
    bad = {k: v for k, v in meas.items() if k.startswith("g/") and not (v < 2e-4 or (pg1[k[2:]] - pg2[k[2:]]).abs().max() < 1e-6)}
    assert not bad, bad
    # the library path of the same layer (cuDNN's LSTM, which runs its GEMMs in TF32 by default) agrees to ITS accuracy
    y3, gx3, _ = run(False)
    meas2 = dict(cudnn_fwd_vs_native=rel(y3, y1), cudnn_fwd_vs_fp64=rel(y3, y2), cudnn_gx_vs_native=rel(gx3, gx1))
    record("lstm/native_T%d_N%d_H%d_D%d" % (T, nb, H, D), **meas2)
    assert meas2["cudnn_fwd_vs_native"] < 2e-3 and meas2["cudnn_gx_vs_native"] < 5e-3, meas2
