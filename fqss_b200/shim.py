"""Drop-in shim: route the reference's ConvTasNet QAT path through the B200 kernels.

    import fqss_b200.shim; fqss_b200.shim.install()      # first lines of train.py / val.py / infer.py

After `install()` the reference's own callers -- `quantization/qat/models/load_model.py:21-102`,
`train_env/train_utils.py:8-27`, `train_env/asteroid_librimix/mysystem.py:88-151` -- run UNCHANGED: their
`from quantization.qat...` / `from train_env.asteroid_librimix.wsdr import *` imports resolve to the mirrors in
`fqss_b200`.  In scope: the ConvTasNet recipes (speech, music) and the sequence models DPTNetQ / SepformerQ; the names the
reference's remaining model files (the Demucs family) import from `quantization.qat.qat_layers / qat_utils / qat_quant`
(Conv2dEncoderQ, ConvTr2dDecoderQ, EmbeddingQ, BatchNormQ, ...) exist as placeholders so that the un-shimmed `load_model.py`
(which imports all five models at its top, `load_model.py:2-6`) still imports; USING one of them raises NotImplementedError
naming the scope.
"""
import sys
import types

import torch.nn as nn

_PLACEHOLDER_CACHE = {}


def out_of_scope(module_name, name):
    """Placeholder for a reference symbol outside the ConvTasNet hot path: importable, not usable."""
    key = (module_name, name)
    if key in _PLACEHOLDER_CACHE:
        return _PLACEHOLDER_CACHE[key]
    msg = ("%s.%s is outside the scope of fqss_b200 (the fake-quantised ConvTasNet / DPTNet / Sepformer QAT paths); run that model without "
           "fqss_b200.shim.install()" % (module_name, name))
    if name[:1].isupper():
        def __init__(self, *a, **kw):
            raise NotImplementedError(msg)
        obj = type(name, (nn.Module,), {"__init__": __init__, "__doc__": msg, "__module__": module_name})
    else:
        def obj(*a, **kw):
            raise NotImplementedError(msg)
        obj.__name__ = name
        obj.__doc__ = msg
    _PLACEHOLDER_CACHE[key] = obj
    return obj


def module_getattr(module_name):
    """PEP 562 `__getattr__` for the mirror modules: unknown public names become out-of-scope placeholders."""
    def __getattr__(name):
        if name.startswith("_"):
            raise AttributeError(name)
        return out_of_scope(module_name, name)
    return __getattr__


_ALIASES = {
    "quantization.qat.qat_quant": "fqss_b200.qat.qat_quant",
    "quantization.qat.qat_layers": "fqss_b200.qat.qat_layers",
    "quantization.qat.qat_utils": "fqss_b200.qat.qat_utils",
    "quantization.qat.models.convtasnetq": "fqss_b200.qat.models.convtasnetq",
    "quantization.qat.models.convtasnetq_music": "fqss_b200.qat.models.convtasnetq_music",
    "quantization.qat.models.dptnetq": "fqss_b200.qat.models.dptnetq",
    "quantization.qat.models.sepformerq": "fqss_b200.qat.models.sepformerq",
    "train_env.asteroid_librimix.wsdr": "fqss_b200.wsdr",
}


def install(loss=True, asteroid_losses=None):
    """Alias the reference's module names to the mirrors.  Call before the first `import quantization...`.

    loss: also route `train_env.asteroid_librimix.wsdr` (PairwiseWSDR, pairwise_wsisdr; wsdr.py:46-101) to
          `fqss_b200.wsdr`, so `System.common_step` (mysystem.py:124-151) runs unmodified on the fused loss kernels.
    asteroid_losses: True also provides `asteroid.losses` (PITLossWrapper, pairwise_neg_sisdr -- the two names the recipe
          takes from it, asteroid_librimix_trainer.py:105, mysystem.py:83) from `fqss_b200.wsdr`; default (None) does so
          only when asteroid is not installed."""
    import importlib
    import importlib.util
    done = {}
    for ref_name, ours in _ALIASES.items():
        if ref_name.startswith("train_env") and not loss:
            continue
        mod = importlib.import_module(ours)
        sys.modules[ref_name] = mod
        done[ref_name] = mod
    if asteroid_losses is None:
        try:
            asteroid_losses = importlib.util.find_spec("asteroid") is None
        except (ImportError, ValueError):
            asteroid_losses = True
    if asteroid_losses and loss:
        w = importlib.import_module("fqss_b200.wsdr")
        m = types.ModuleType("asteroid.losses")
        m.PITLossWrapper, m.pairwise_neg_sisdr = w.PITLossWrapper, w.pairwise_neg_sisdr
        m.__all__ = ["PITLossWrapper", "pairwise_neg_sisdr"]
        pkg = sys.modules.get("asteroid")
        if pkg is None or not hasattr(pkg, "__path__"):
            pkg = types.ModuleType("asteroid")
            pkg.__path__ = []
            sys.modules["asteroid"] = pkg
        pkg.losses = m
        sys.modules["asteroid.losses"] = m
        done["asteroid.losses"] = m
    return done
