"""CPU: the DPTNetQ mirror builds the reference's module tree -- same state-dict keys in the same order
and the same seeded initial values as the UNMODIFIED reference (fixture: tests/golden/make_golden_dptnet.py)."""
import numpy as np
import torch

KW = dict(n_spks=2, kernel_size=2, enc_dim=64, feature_dim=32, hidden_dim=32, layer=1, segment_size=20)
QCFG = dict(qat=True, gradient_based=True, weight_quant=True, weight_n_bits=8, act_quant=True, act_n_bits=8,
            in_quant=False, in_act_n_bits=8, out_quant=True, out_act_n_bits=8, n_splitter=2, n_combiner=2, observer=True)


def build(seed=0):
    from fqss_b200.qat.models.dptnetq import DPTNetQ
    from fqss_b200.qat.models.load_model import quantize_model
    torch.manual_seed(seed)
    return quantize_model(DPTNetQ(**KW), dict(QCFG))


def test_dptnet_state_dict_matches_reference(golden):
    g = golden("dptnet_small.npz")
    sd = build().state_dict()
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    bad = [k for k, v in sd.items() if not np.array_equal(v.cpu().numpy(), g["init/" + k])]
    assert not bad, bad[:5]


def test_dptnet_factory_and_segmentation():
    from fqss_b200.qat.models.load_model import create_model
    from fqss_b200.qat.models.dptnetq import DPTNetQ, overlap_and_add
    m = create_model({"name": "DPTNet", "n_src": 2, "kernel_size": 2})
    assert isinstance(m, DPTNetQ) and m.enc_dim == 256 and m.segment_size == 250
    # chunking followed by the overlap-add merge is the identity on the un-padded axis (each frame is covered by two chunks)
    sep = build().separator
    from fqss_b200.qat.qat_layers import Add
    sep.add = Add()                      # the float module the quantised AddQ replaced
    x = torch.randn(2, 3, 57)
    segs, rest = sep.split_feature(x, 20)
    assert segs.shape[2] == 20 and segs.shape[3] % 2 == 0
    assert torch.allclose(sep.merge_feature(segs, rest), 2 * x)
    # overlap_and_add against its definition
    s = torch.randn(2, 5, 4)
    ref = torch.zeros(2, 2 * 4 + 4)
    for f in range(5):
        ref[:, 2 * f:2 * f + 4] += s[:, f]
    assert torch.allclose(overlap_and_add(s, 2), ref)
