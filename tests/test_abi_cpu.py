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
    assert lib.fqss_abi_version() == 21
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


def test_ctypes_binding_matches_header(native, tmp_path):
    """The Python binding is hand-written: guard it against drift from include/fqss.h.  (1) every prototype's
    parameter count equals the ctypes signature's; (2) every struct that crosses the ABI has the same size and field
    offsets in C (gcc on the header, which must stay plain C) and in ctypes."""
    import ctypes as C
    import subprocess
    from fqss_b200 import tcn_engine as E
    hdr_path = os.path.join(ROOT, "include", "fqss.h")
    hdr = open(hdr_path).read()
    nocomment = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = dict((m.group(1), m.group(2)) for m in re.finditer(r"\b(fqss_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", nocomment))
    for name, (res, args) in native._SIGS.items():
        assert name in protos, name
        plist = [a for a in protos[name].split(",") if a.strip() and a.strip() != "void"]
        assert len(plist) == len(args), "%s: header has %d parameters, binding %d" % (name, len(plist), len(args))
    structs = {"fqss_tcn_block": E.TcnBlock, "fqss_tcn_block_grads": E.TcnBlockGrads, "fqss_qrange": E.QRange,
               "fqss_prep_item": native.PrepItem, "fqss_gather_item": native.GatherItem, "fqss_pw_desc": native.PwDesc,
               "fqss_pw_grads": native.PwGrads, "fqss_wq_item": native.WqItem}
    probes = {"fqss_tcn_block": ["no_skip", "ld", "Wc1", "q_in", "x_op", "rc1", "code3"], "fqss_tcn_block_grads": ["dY2", "dW1q", "g_q", "ws_bytes"],
              "fqss_prep_item": ["Wc", "N", "split"], "fqss_gather_item": ["offset", "numel"], "fqss_pw_desc": ["rows", "x2", "eps", "rmax"],
              "fqss_pw_grads": ["gx2", "g_beta"], "fqss_wq_item": ["rmin", "n_bits"], "fqss_qrange": ["rmax"]}
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "fqss.h"', 'int main(void) {']
    for s in structs:
        src.append('printf("%s sizeof %%zu\\n", sizeof(%s));' % (s, s))
        for f in probes[s]:
            src.append('printf("%s %s %%zu\\n", offsetof(%s, %s));' % (s, f, s, f))
    src += ['return 0;', '}']
    cfile, exe = os.path.join(str(tmp_path), "abi.c"), os.path.join(str(tmp_path), "abi")
    open(cfile, "w").write("\n".join(src))
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-I", os.path.join(ROOT, "include"), cfile, "-o", exe])
    for line in subprocess.check_output([exe], text=True).splitlines():
        s, f, v = line.split()
        want = C.sizeof(structs[s]) if f == "sizeof" else getattr(structs[s], f).offset
        assert int(v) == want, "%s.%s: C %s vs ctypes %d" % (s, f, v, want)


def test_ranges_ok_flags_crossed_or_degenerate_ranges():
    """Sync-free stand-in for the reference's per-call `assert max_range >= min_range` (qat_quant.py:238)."""
    import torch
    from fqss_b200.qat.qat_quant import GradientActivationFakeQuantize, ranges_ok
    m = torch.nn.Sequential(GradientActivationFakeQuantize(True), GradientActivationFakeQuantize(True))
    assert bool(ranges_ok(m))
    m[1].max_range.data.fill_(-1.0)
    assert not bool(ranges_ok(m))
    m[1].max_range.data.fill_(float("nan"))
    assert not bool(ranges_ok(m))


def test_skipless_block_plumbing_on_cpu():
    """Host-side plumbing of the fused engine's skip-less mode (ConvTasNetMusicQ blocks, convtasnetq_music.py:117-176): the
    block -> slot mapping names the right parameters, blocks without a skip conv leave the skip slots empty, and the
    eligibility tests refuse CPU tensors / un-calibrated models (the product path then stays on the per-layer wrappers,
    which raise on non-CUDA input: no CPU fallback anywhere)."""
    import torch
    from fqss_b200 import float_engine as FE
    from fqss_b200 import tcn_engine as E
    from fqss_b200.qat.models.convtasnetq_music import ConvTasNetMusicQ
    from fqss_b200.qat.models.load_model import quantize_model
    from fqss_b200.testing import RECIPE_QUANT
    torch.manual_seed(0)
    m = quantize_model(ConvTasNetMusicQ(sources=["a", "b"], n_filters=128, bn_chan=128, hid_chan=256, n_blocks=2, n_repeats=1),
                       dict(RECIPE_QUANT))
    blk = m.separator.network[2][0][1]
    t, dil = E.block_tensors(blk, None, True)
    assert dil == 2 and set(t) == set(E._BLOCK_SLOTS)
    assert t["W1"] is blk.net[0].conv1d.weight and t["Wdw"] is blk.net[3].net[0].conv1d.weight
    assert t["Wres"] is blk.net[3].net[3].conv1d.weight and t["slope3"] is blk.net[3].net[0].nl.weight
    assert t["qaddmin"] is blk.add.activation_fake_quantize.min_range
    for k in ("Wskip", "bskip", "wsmin", "wsmax", "qskipmin", "qskipmax", "qaddsmin", "qaddsmax", "b1", "bdw", "bres"):
        assert t[k] is None, k
    x = torch.zeros(1, 128, 64)
    assert not E.fused_eligible_noskip(m.separator, x) and not FE.noskip_eligible(m.separator, x)
    # the ctypes mirror of fqss_tcn_block carries the flag where the header puts it
    b = E.TcnBlock()
    b.no_skip = 1
    assert b.no_skip == 1 and E.TcnBlock.no_skip.offset == E.TcnBlock.split.offset + 4
