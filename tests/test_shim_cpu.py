"""CPU: the INTEGRATION.md shim (fqss_b200.shim.install) against the reference's OWN callers.

Runs only where the reference tree exists (/root/reference in the build container); on the GPU box it is skipped.  A
subprocess keeps the aliased sys.modules out of the other tests.  What is checked: with the shim installed, the
unmodified `train_env/train_utils.py:8-27` and `quantization/qat/models/load_model.py:21-74` import, build the
(student, teacher) pair from the shipped YAML, and the student is THIS repository's ConvTasNetQ with the reference's
948 state-dict keys; `enable_observer` reaches our quantisers; the unmodified `System` (mysystem.py:24-151) imports and
binds our PairwiseWSDR / PITLossWrapper; out-of-scope models fail with a message that names the scope."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("FQSS_REFERENCE_ROOT", "/root/reference")

SCRIPT = r"""
import sys, os
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, "tests", "golden"))
import numpy as np, torch, yaml
import _ref_import as R
R.install()                                   # stub finder for the optional packages absent from this image + reference on sys.path
import fqss_b200.shim as shim
shim.install()
# --- the reference's own factory code, unmodified
from train_env.train_utils import create_pretrained_model
from quantization.qat.models import load_model as LM
assert LM.__file__.startswith(%(ref)r), LM.__file__
cfg = yaml.safe_load(open(os.path.join(%(ref)r, "configs", "convtasnet_2spks_8k.yaml")))
mc = cfg["model_cfg"]
mc["model_path"] = None
torch.manual_seed(0)
model, fmodel = create_pretrained_model(mc, use_weights=False)
assert type(model).__module__ == "fqss_b200.qat.models.convtasnetq", type(model).__module__
assert type(fmodel).__module__ == "fqss_b200.qat.models.convtasnetq"
assert model.n_splitter == 2 and model.n_combiner == 2 and fmodel.n_splitter == 1
sd = model.state_dict()
assert len(sd) == 948, len(sd)
keys = [str(k) for k in np.load(os.path.join(%(root)r, "tests", "golden", "model_full_keys.npz"))["keys"]] \
    if os.path.exists(os.path.join(%(root)r, "tests", "golden", "model_full_keys.npz")) else None
if keys is not None:
    assert list(sd.keys()) == keys
from fqss_b200.qat.qat_quant import GradientActivationFakeQuantize as AQ, GradientWeightFakeQuantize as WQ
qs = [m for m in model.modules() if isinstance(m, (AQ, WQ))]
assert len(qs) == 301, len(qs)
assert all(q.observer_mode for q in qs)       # observer: True in the YAML
LM.enable_observer(model, False)              # the reference's function walks OUR quantiser classes
assert not any(q.observer_mode for q in qs)
LM.set_mac_op(model, True)
# --- the reference's Lightning system, unmodified: binds our loss modules through its star imports
import train_env.asteroid_librimix.mysystem as MS
assert MS.PairwiseWSDR.__module__ == "fqss_b200.wsdr" and MS.PITLossWrapper.__module__ == "fqss_b200.wsdr"
# --- out of scope: importable, not usable
from quantization.qat.qat_layers import EmbeddingQ, BatchNormQ, Conv2dEncoderQ
for cls in (EmbeddingQ, BatchNormQ, Conv2dEncoderQ):
    try:
        cls()
    except NotImplementedError as e:
        assert "outside the scope" in str(e)
    else:
        raise AssertionError("placeholder was usable")
# --- the sequence models through the reference's factory: our mirrors, quantised by the reference's quantize_model
for name, mod, nq in (("DPTNet", "fqss_b200.qat.models.dptnetq", None), ("Sepformer", "fqss_b200.qat.models.sepformerq", None)):
    m = LM.create_model(dict(name=name, n_src=2))
    assert type(m).__module__ == mod, type(m).__module__
    m = LM.quantize_model(m, dict(mc["quantization"]))
    assert any(isinstance(q, AQ) for q in m.modules()) and m.n_splitter == 2
print("SHIM-OK")
"""


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "quantization", "qat")), reason="reference tree not present")
def test_shim_runs_reference_callers():
    code = textwrap.dedent(SCRIPT % dict(root=ROOT, ref=REF))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SHIM-OK" in r.stdout, r.stdout[-2000:] + "\n" + r.stderr[-4000:]
