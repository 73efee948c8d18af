"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/fqss.h declares;
the drop-in module tree reproduces the reference's state_dict (keys, order, shapes, seeded values)."""
import copy
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    from fqss_b200 import build as B
    B.build()
    from fqss_b200 import _native
    return _native


def test_header_symbols_are_exported(native):
    hdr = open(os.path.join(ROOT, "include", "fqss.h")).read()
    declared = set(re.findall(r"\b(fqss_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"fqss_pw_desc", "fqss_pw_grads", "fqss_tcn_block", "fqss_tcn_block_grads", "fqss_qrange"}
    lib = native.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), "header declares %s but the library does not export it" % name
    assert set(native.EXPORTED) == declared, (set(native.EXPORTED) ^ declared)
    assert lib.fqss_abi_version() == 12
    assert lib.fqss_ws_bytes(1024) >= 4096


def test_no_cpu_fallback(native):
    from fqss_b200 import ops
    x = torch.randn(8)
    with pytest.raises(RuntimeError):
        ops.FakeQuantAct.apply(x, torch.tensor([-0.5]), torch.tensor([0.5]), 8)
    from fqss_b200.process import preprocess
    with pytest.raises(RuntimeError):
        preprocess(torch.randn(2, 1, 64), n_splitter=2)


def test_bad_arguments_report_errors(native):
    lib = native.lib()
    assert lib.fqss_fq_act_fwd(None, None, None, 16, None, None, 8, None) == -1
    assert b"fq_act_fwd" in lib.fqss_last_error()
    assert lib.fqss_conv1x1_fwd(None, 0, None, None, None, 0, 1, 1, 1, 1, None) == -1


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fqss_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py") and f != "testing.py":
                src = open(os.path.join(dp, f)).read()
                assert "fqss_oracle" not in src and "import oracle" not in src, f


def test_state_dict_matches_reference(golden):
    from fqss_b200.qat.models.convtasnetq import ConvTasNetQ
    from fqss_b200.qat.models.load_model import quantize_model
    from fqss_b200.testing import RECIPE_QUANT, SMALL_KW
    g = golden("model_small.npz")
    torch.manual_seed(0)
    m = ConvTasNetQ(**SMALL_KW)
    f = copy.deepcopy(m)
    quantize_model(m, dict(RECIPE_QUANT))
    sd = m.state_dict()
    keys = [k[5:] for k in g.files if k.startswith("init/")]
    assert list(sd.keys()) == keys
    for k in keys:
        assert np.array_equal(sd[k].cpu().numpy(), g["init/" + k]), k
    fs = f.state_dict()
    tkeys = [k[8:] for k in g.files if k.startswith("teacher/")]
    assert list(fs.keys()) == tkeys
    # a reference-trained state_dict loads strictly (and back)
    m.load_state_dict({k: torch.from_numpy(g["calib/" + k]) for k in keys}, strict=True)
    # Identity holes where the reference leaves them (qat_utils.py:273-310)
    blk = m.masker.TCN[0]
    assert isinstance(blk.shared_block[1], torch.nn.Identity) and isinstance(blk.shared_block[4], torch.nn.Identity)
    assert type(blk.shared_block[0]).__name__ == "Conv1dNlQ" and type(blk.shared_block[2]).__name__ == "GroupNormQ"
    assert type(m.encoder).__name__ == "Conv1dEncoderQ" and type(m.decoder).__name__ == "ConvTr1dDecoderQ"


def test_full_model_parameter_count():
    from fqss_b200.qat.models.convtasnetq import ConvTasNetQ
    from fqss_b200.qat.models.load_model import quantize_model
    from fqss_b200.testing import FULL_KW, RECIPE_QUANT
    m = ConvTasNetQ(**FULL_KW)
    assert sum(p.numel() for p in m.parameters()) == 5050545          # SURVEY.md 3.2 [probe]
    quantize_model(m, dict(RECIPE_QUANT))
    assert sum(p.numel() for p in m.parameters()) == 5133123      # (SURVEY.md quotes 5 115 713; the reference itself gives 5 133 123)
    assert len(m.state_dict()) == 948
