"""B200 mirror of `quantization/qat/qat_utils.py` (module surgery).

`quantize_modules(model, ['0','1'], params)` swaps the first named child for its Q-wrapper and
the remaining ones for `nn.Identity()` (so `shared_block` keeps Identity holes at indices 1 and 4);
`replace_encoderq` / `replace_decoderq` do the same for the filterbanks.  Reference: :258-332,
table :354-401.  Only the layer types the ConvTasNet recipe needs are registered; anything else
raises NotImplementedError (there is no eager fallback).
"""
import copy

import torch.nn as nn

from . import qat_layers as QL
from . import qat_layers_seq as QS
from . import qat_quant as QQ


def _resolve(root, dotted):
    mod = root
    for tok in dotted.split("."):
        mod = getattr(mod, tok)
    return mod


def _assign(root, dotted, new):
    *parents, leaf = dotted.split(".")
    mod = root
    for tok in parents:
        mod = getattr(mod, tok)
    setattr(mod, leaf, new)


def _common(p):
    return dict(gradient_based=p.get("gradient_based", True), act_quant=p.get("act_quant", True),
                act_n_bits=p.get("act_n_bits", 8))


def _weighted(p):
    d = _common(p)
    d.update(weight_quant=p.get("weight_quant", True), weight_n_bits=p.get("weight_n_bits", 8))
    return d


def quant_conv1d(conv1d, p):
    return QL.Conv1dQ(conv1d, **_weighted(p))


def quant_conv1d_nl(conv1d, nl, p):
    return QL.Conv1dNlQ(conv1d, nl, **_weighted(p))


def quant_groupnorm(gn, p):
    return QL.GroupNormQ(gn, **_common(p))


def quant_nl(nl, p):
    return QL.NlQ(nl, **_common(p))


def quant_add(add, p):
    return QL.AddQ(add, **_common(p))


def quant_sub(sub, p):
    return QL.SubQ(sub, **_common(p))


def quant_mul(mul, p):
    return QL.MulQ(mul, **_common(p))


def quant_encoderq(encoder, p):
    if isinstance(encoder[0], nn.Conv1d):
        return QL.Conv1dEncoderQ(encoder, n_splitter=p.get("n_splitter", 1), gradient_based=p.get("gradient_based", True),
                                 weight_quant=p.get("weight_quant", True), weight_n_bits=p.get("weight_n_bits", 8),
                                 act_quant=p.get("act_quant", True), act_n_bits=p.get("act_n_bits", 8),
                                 in_quant=p.get("in_quant", False), in_act_n_bits=p.get("in_act_n_bits", 8),
                                 inout_nl_quant=p.get("inout_nl_quant", False))
    raise NotImplementedError("encoder type %s is out of scope (ConvTasNet Conv1d encoder only)" % type(encoder[0]).__name__)


def quant_decoderq(decoder, p):
    if isinstance(decoder[0], nn.ConvTranspose1d):
        return QL.ConvTr1dDecoderQ(decoder, n_combiner=p.get("n_combiner", 1), gradient_based=p.get("gradient_based", True),
                                   weight_quant=p.get("weight_quant", True), weight_n_bits=p.get("weight_n_bits", 8),
                                   act_quant=p.get("act_quant", True), act_n_bits=p.get("act_n_bits", 8),
                                   inout_nl_quant=p.get("inout_nl_quant", False), out_quant=p.get("out_quant", True),
                                   out_act_n_bits=p.get("out_act_n_bits", 8), train_res_dec=bool(p.get("train_res_dec", False)))
    if isinstance(decoder[0], nn.Linear):
        return QL.LinearDecoderQ(decoder, n_combiner=p.get("n_combiner", 1), gradient_based=p.get("gradient_based", True),
                                 weight_quant=p.get("weight_quant", True), weight_n_bits=p.get("weight_n_bits", 8),
                                 act_quant=p.get("act_quant", True), act_n_bits=p.get("act_n_bits", 8),
                                 inout_nl_quant=p.get("inout_nl_quant", False), out_quant=p.get("out_quant", True),
                                 out_act_n_bits=p.get("out_act_n_bits", 8), train_res_dec=bool(p.get("train_res_dec", False)))
    raise NotImplementedError("decoder type %s is out of scope (ConvTranspose1d and Linear decoders only)" % type(decoder[0]).__name__)


def quant_layernorm(layernorm, p):
    return QL.LayerNormQ(layernorm, **_common(p))


def quant_conv2d(conv2d, p):
    return QS.Conv2dQ(conv2d, **_weighted(p))


def quant_conv2d_nl(conv2d, nl, p):
    return QS.Conv2dNlQ(conv2d, nl, **_weighted(p))


def quant_linear(linear, p):
    return QS.LinearQ(linear, **_weighted(p))


def quant_linear_nl(linear, nl, p):
    return QS.LinearNlQ(linear, nl, **_weighted(p))


def quant_lstm(lstm, p):
    return QS.LSTMQ(lstm, **_weighted(p))


def quant_mha(mha, p):
    return QS.MultiheadAttentionQ(mha, **_weighted(p))


def quant_const(const, p):
    return QS.ConstQ(const, **_common(p))


def quant_div(div, p):
    return QS.DivQ(div, **_common(p))


OP_LIST_TO_QUANTIZE_METHOD = {
    (nn.Conv1d): quant_conv1d,
    (nn.Conv1d, nn.PReLU): quant_conv1d_nl,
    (nn.Conv1d, nn.ReLU): quant_conv1d_nl,
    (nn.GroupNorm): quant_groupnorm,
    (nn.LayerNorm): quant_layernorm,
    (nn.PReLU): quant_nl,
    (nn.ReLU): quant_nl,
    (QL.Add): quant_add,
    (QL.Sub): quant_sub,
    (QL.Mul): quant_mul,
    # sequence models (DPTNetQ / SepformerQ; reference table qat_utils.py:354-401)
    (nn.Conv1d, nn.Tanh): quant_conv1d_nl,
    (nn.Conv1d, nn.Sigmoid): quant_conv1d_nl,
    (nn.Conv2d): quant_conv2d,
    (nn.Conv2d, nn.PReLU): quant_conv2d_nl,
    (nn.Conv2d, nn.ReLU): quant_conv2d_nl,
    (nn.Linear): quant_linear,
    (nn.Linear, nn.ReLU): quant_linear_nl,
    (nn.LSTM): quant_lstm,
    (nn.MultiheadAttention): quant_mha,
    (nn.Tanh): quant_nl,
    (nn.Sigmoid): quant_nl,
    (QS.Const): quant_const,
    (QS.Div): quant_div,
}


def quantize_known_modules(mod_list, params_dict):
    key = tuple(type(m) for m in mod_list)
    key = key[0] if len(key) == 1 else key
    factory = OP_LIST_TO_QUANTIZE_METHOD.get(key)
    if factory is None:
        raise NotImplementedError("Cannot quantize modules: {}".format(key))
    out = [factory(*mod_list, params_dict)]
    for _ in mod_list[1:]:
        hole = nn.Identity()
        hole.training = mod_list[0].training
        out.append(hole)
    return out


def quantize_modules(model, modules_to_quantize, params_dict={}, inplace=True, replacer_func=quantize_known_modules):
    if not inplace:
        model = copy.deepcopy(model)
    olds = [_resolve(model, name) for name in modules_to_quantize]
    news = replacer_func(olds, params_dict)
    for name, new in zip(modules_to_quantize, news):
        _assign(model, name, new)
    return model


def _replace_edge(model, names, params_dict, factory):
    mods = [_resolve(model, n) for n in names]
    _assign(model, names[0], factory(mods, params_dict))
    for n in names[1:]:
        _assign(model, n, nn.Identity())


def replace_encoderq(model, modules_to_replace, params_dict):
    _replace_edge(model, modules_to_replace, params_dict, quant_encoderq)


def replace_decoderq(model, modules_to_replace, params_dict):
    _replace_edge(model, modules_to_replace, params_dict, quant_decoderq)


# export: swap a learned quantiser for its torch-affine counterpart (qat_utils.py:246-349)
def torch_weight_quantizer(quantizer):
    return QQ.TorchWeightFakeQuantize(quantizer)


def torch_activation_quantizer(quantizer):
    return QQ.TorchActivationFakeQuantize(quantizer)


def torch_dym_activation_quantizer(quantizer):
    return QQ.TorchDymActivationFakeQuantize(quantizer)


def replace_weight_quantizer(model, module_to_replace, module):
    _assign(model, module_to_replace, torch_weight_quantizer(module))


def replace_activation_quantizer(model, module_to_replace, module):
    _assign(model, module_to_replace, torch_activation_quantizer(module))


def replace_dym_activation_quantizer(model, module_to_replace, module):
    _assign(model, module_to_replace, torch_dym_activation_quantizer(module))


# reference names outside the ConvTasNet hot path (imported by the reference's other model files): importable placeholders
# that raise NotImplementedError when used -- see fqss_b200/shim.py
from ..shim import module_getattr as _module_getattr  # noqa: E402
__getattr__ = _module_getattr("quantization.qat.qat_utils")
