"""Helpers shared by tests, smoke() and bench.py: build the (student, teacher) pair the way the
reference recipe does (train_env/train_utils.py:8-27) and expose parameters to the oracle."""
import copy

import torch

RECIPE_QUANT = dict(qat=True, gradient_based=True, weight_quant=True, weight_n_bits=8, act_quant=True, act_n_bits=8,
                    in_quant=False, in_act_n_bits=8, out_quant=True, out_act_n_bits=8, n_splitter=2, n_combiner=2,
                    observer=True)          # configs/convtasnet_2spks_8k.yaml:13-26
SMALL_KW = dict(n_spks=2, kernel_size=16, stride=8, n_filters=64, bn_chan=32, hid_chan=64, n_blocks=3, n_repeats=2)
FULL_KW = dict(n_spks=2, kernel_size=16, stride=8)
# smallest model whose TCN is eligible for the fused tcgen05 engine (channel counts are multiples of 128)
FUSED_SMALL_KW = dict(n_spks=2, kernel_size=16, stride=8, n_filters=128, bn_chan=128, hid_chan=128, n_blocks=2, n_repeats=1)


def _oracle_cfg(kw):
    import fqss_oracle as O
    return O.SeparatorConfig(n_src=kw.get("n_spks", 2), kernel_size=kw["kernel_size"], stride=kw["stride"],
                             n_filters=kw.get("n_filters", 512), bn_chan=kw.get("bn_chan", 128),
                             hid_chan=kw.get("hid_chan", 512), n_blocks=kw.get("n_blocks", 8),
                             n_repeats=kw.get("n_repeats", 3))


def __getattr__(name):      # lazy: the oracle is test infrastructure, never imported by the product path
    if name == "SMALL_CFG":
        return _oracle_cfg(SMALL_KW)
    if name == "FULL_CFG":
        return _oracle_cfg(FULL_KW)
    if name == "FUSED_SMALL_CFG":
        return _oracle_cfg(FUSED_SMALL_KW)
    raise AttributeError(name)


def model_pair(kw, device, seed=0, quant_cfg=None):
    from .qat.models.convtasnetq import ConvTasNetQ
    from .qat.models.load_model import quantize_model
    torch.manual_seed(seed)
    model = ConvTasNetQ(**kw)
    fmodel = copy.deepcopy(model)
    model = quantize_model(model, dict(quant_cfg or RECIPE_QUANT))
    return model.to(device), fmodel.to(device)


def small_model_pair(device, seed=0):
    return model_pair(SMALL_KW, device, seed)


def oracle_params(module):
    import fqss_oracle as O
    return O.Params({k: v.detach().cpu().clone() for k, v in module.state_dict().items()})
