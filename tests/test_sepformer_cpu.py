"""CPU: the SepformerQ mirror builds the reference's module tree -- same 825 state-dict keys in the same
order, same seeded initial values (compared through the float64 fingerprints stored by tests/golden/make_golden_sepformer.py)."""
import numpy as np
import torch

KW = dict(n_spks=2, kernel_size=16, stride=8, n_filters=32, n_repeats=1, n_heads=4, chunk_size=10)
QCFG = dict(qat=True, gradient_based=True, weight_quant=True, weight_n_bits=8, act_quant=True, act_n_bits=8,
            in_quant=False, in_act_n_bits=8, out_quant=True, out_act_n_bits=8, n_splitter=2, n_combiner=2, observer=True)


def fp(t):
    t = t.detach().double().flatten().cpu()
    r = torch.randn(t.numel(), generator=torch.Generator().manual_seed(t.numel()), dtype=torch.float64)
    return np.array([t.sum().item(), t.abs().sum().item(), (t * r).sum().item()])


def build(seed=0):
    from fqss_b200.qat.models.sepformerq import SepformerQ
    from fqss_b200.qat.models.load_model import quantize_model
    torch.manual_seed(seed)
    return quantize_model(SepformerQ(**KW), dict(QCFG))


def test_sepformer_state_dict_matches_reference(golden):
    g = golden("sepformer_small.npz")
    sd = build().state_dict()
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    bad = [k for k, v in sd.items() if not np.array_equal(fp(v), g["initfp/" + k])]
    assert not bad, bad[:5]


def test_sepformer_factory_and_chunking():
    from fqss_b200.qat.models.load_model import create_model
    from fqss_b200.qat.models.sepformerq import SepformerQ
    m = create_model({"name": "Sepformer", "n_src": 2, "kernel_size": 16, "stride": 8})
    assert isinstance(m, SepformerQ) and m.enc_num_feats == 256 and m.masker.chunk_size == 250
    mk = m.masker
    x = torch.randn(2, 3, 57)
    seg, gap = mk.segmentation(x, 20)
    assert seg.shape[2] == 20 and seg.shape[3] % 2 == 0
    assert torch.allclose(mk.over_add(seg, gap), 2 * x)          # every frame is covered by exactly two chunks
