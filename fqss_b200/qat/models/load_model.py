"""B200 mirror of `quantization/qat/models/load_model.py` restricted to the ConvTasNet recipes (speech and music)
(create_model :21-51, quantize_model :53-74, enable_observer :16-19, create_pretrained_model :76-102)."""
import torch

from ..qat_layers import LayerQ
from ..qat_quant import GradientActivationFakeQuantize, GradientWeightFakeQuantize
from .convtasnetq import ConvTasNetQ
from .convtasnetq_music import ConvTasNetMusicQ
from .dptnetq import DPTNetQ
from .sepformerq import SepformerQ


def set_mac_op(model, mode=False):
    for m in model.modules():
        if isinstance(m, LayerQ):
            m.do_mac_op = mode


def enable_observer(model, mode=False):
    for m in model.modules():
        if isinstance(m, (GradientWeightFakeQuantize, GradientActivationFakeQuantize)):
            m.enable_observer(mode)


def create_model(model_cfg):
    name = model_cfg["name"]
    if name == "ConvTasNet":
        return ConvTasNetQ(n_spks=model_cfg.get("n_src", 1), kernel_size=model_cfg.get("kernel_size", 32),
                           stride=model_cfg.get("stride", 16))
    if name == "DPTNet":
        return DPTNetQ(n_spks=model_cfg.get("n_src", 2), kernel_size=model_cfg.get("kernel_size", 2))
    if name == "Sepformer":
        return SepformerQ(n_spks=model_cfg.get("n_src", 2), kernel_size=model_cfg.get("kernel_size", 16),
                          stride=model_cfg.get("stride", 8))
    if name == "ConvTasNetMusic":
        return ConvTasNetMusicQ(sources=model_cfg.get("sources", ["drums", "bass", "other", "vocals"]),
                                kernel=model_cfg.get("kernel_size", 20), stride=model_cfg.get("stride", 10))
    raise NotImplementedError("fqss_b200 covers the ConvTasNet recipes (speech, music), DPTNet and Sepformer; model %r is out of scope" % name)


def quantize_model(model, quant_cfg):
    if quant_cfg.get("qat", False):
        model.set_splitter_combiner(quant_cfg.get("n_splitter", 1), quant_cfg.get("n_combiner", 1))
        model.quantize_model(gradient_based=quant_cfg.get("gradient_based", True),
                             weight_quant=quant_cfg.get("weight_quant", True),
                             weight_n_bits=quant_cfg.get("weight_n_bits", 8),
                             act_quant=quant_cfg.get("act_quant", True), act_n_bits=quant_cfg.get("act_n_bits", 8),
                             inout_nl_quant=quant_cfg.get("inout_nl_quant", False),
                             in_quant=quant_cfg.get("in_quant", False), in_act_n_bits=quant_cfg.get("in_act_n_bits", 8),
                             out_quant=quant_cfg.get("out_quant", False),
                             out_act_n_bits=quant_cfg.get("out_act_n_bits", 8))
        enable_observer(model, quant_cfg.get("observer", False))
    return model


def create_pretrained_model(model_cfg):
    model = quantize_model(create_model(model_cfg), model_cfg["quantization"])
    path = model_cfg.get("model_path", None)
    if path is None:
        return model
    if path.startswith("https"):
        sd = torch.hub.load_state_dict_from_url(path, map_location="cpu", check_hash=True)
    else:
        sd = torch.load(path)
    for key in ("state", "state_dict"):
        if key in sd:
            sd = sd[key]
            break
    try:
        model.load_state_dict(sd, strict=True)
    except Exception:
        model.load_pretrain(path)
    return model
