"""Import helper for the UNMODIFIED reference (only usable where /root/reference exists).

Used by tests/golden/make_golden.py and oracle/check_against_reference.py to run the real
ssi-research/FQSS code on CPU.  The reference imports several optional packages at module
top level that are not installed here and that the ConvTasNet hot path never touches
(matplotlib, torchmetrics, demucs, ...).  A meta-path finder hands out inert stub modules
for exactly those names so `import quantization.qat...` succeeds.

Nothing under tests/ or the product imports this at GPU-box run time: the reference tree
does not exist there.  It is test-fixture tooling only.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FQSS_REFERENCE_ROOT", "/root/reference")

_STUB_ROOTS = {
    "matplotlib", "torchmetrics", "demucs", "openunmix", "julius", "asteroid",
    "pytorch_lightning", "speechbrain", "hyperpyyaml", "musdb", "museval", "soundfile",
    "dora", "hydra", "omegaconf", "torchaudio", "wandb",
}


class _Inert:
    """Attribute sink: any attribute / call / decorator use returns something harmless."""

    def __init__(self, name="stub"):
        self._n = name

    def __getattr__(self, k):
        if k.startswith("__") and k.endswith("__"):
            raise AttributeError(k)
        return _Inert(self._n + "." + k)

    def __call__(self, *a, **kw):
        if len(a) == 1 and callable(a[0]) and not kw:
            return a[0]
        return _Inert(self._n + "()")

    def __mro_entries__(self, bases):
        return (object,)

    def __iter__(self):
        return iter(())


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, k):
        if k.startswith("__") and k.endswith("__"):
            raise AttributeError(k)
        return _Inert(self.__name__ + "." + k)


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "quantization", "qat"))


def install():
    """Make `import quantization.qat...`, `import process`, `import utils` resolve to the reference."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        # only stub what is genuinely missing
        for name in list(_STUB_ROOTS):
            try:
                __import__(name)
                _STUB_ROOTS.discard(name)
            except Exception:
                for k in [k for k in sys.modules if k.split(".")[0] == name]:
                    del sys.modules[k]
        sys.meta_path.insert(0, _StubFinder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def build_reference_model(model_kwargs, quant_cfg, seed=0):
    """Reference recipe: create -> deepcopy teacher -> quantize (train_env/train_utils.py:8-27)."""
    import copy
    import torch
    install()
    from quantization.qat.models.convtasnetq import ConvTasNetQ
    from quantization.qat.models import load_model as LM
    torch.manual_seed(seed)
    model = ConvTasNetQ(**model_kwargs)
    fmodel = copy.deepcopy(model)
    model = LM.quantize_model(model, quant_cfg)
    return model, fmodel, LM
