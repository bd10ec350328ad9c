"""antquant.quant_model: model surgery (same entry points as olive_quantization/antquant/quant_model.py).

quantize_model(model) returns a NEW model in which every nn.Conv2d / nn.Linear / HF Conv1D
(exact type match) is replaced by its quantizer wrapper; `lm_head` and `base_model` are left alone; nn.Sequential / nn.ModuleList containers are
rebuilt as nn.Sequential, as the reference does (state_dict keys are unchanged by that).
"""
import _bootstrap  # noqa: F401
import torch
import torch.nn as nn
import numpy as np
import copy
from quant_modules import TensorQuantizer, Conv2dQuantizer, LinearQuantizer, Conv1dQuantizer
from quant_utils import quant_args
import torch.distributed as dist

_WRAPPERS = {nn.Conv2d: Conv2dQuantizer, nn.Linear: LinearQuantizer}
try:
    from transformers import pytorch_utils
    _WRAPPERS[pytorch_utils.Conv1D] = Conv1dQuantizer
except Exception:        # transformers is only needed for GPT-2 style Conv1D layers
    pass
_SKIP_CHILDREN = ("base_model", "lm_head")


def _say(*a):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0:
        print(*a)


def _convert(module):
    """Recursive worker: `module` already belongs to the copy and may be modified in place."""
    wrapper = _WRAPPERS.get(type(module))
    if wrapper is not None:
        q = wrapper(**quant_args)
        q.set_param(module)
        return q
    if isinstance(module, (nn.Sequential, nn.ModuleList)):
        return nn.Sequential(*[_convert(m) for m in module.children()])
    for name, child in list(module.named_children()):
        if name in _SKIP_CHILDREN:
            continue
        setattr(module, name, _convert(child))
    return module


def quantize_model(model):
    """Recursively replace the single-precision layers of `model` (left untouched) by quantized ones."""
    return _convert(copy.deepcopy(model))


def set_first_last_layer(model):
    """Kept for the drivers that call it; the reference only collects the quantizers and does nothing."""
    return None


def _tensor_quantizers(model):
    return [m for m in model.modules() if isinstance(m, TensorQuantizer)]


def _promote(q):
    q.bit.data = torch.tensor(8, device=q.bit.device)


def _reset_calibration(qs):
    for q in qs:
        q.has_inited_quant_para.data = torch.zeros_like(q.has_inited_quant_para)


def set_8_bit_layer_l(model, layer_list):
    """Promote the layers (weight + input quantizer pairs) named by index in `layer_list` ("3,7,9") to 8 bit."""
    if layer_list == "None":
        return
    wanted = [int(x) for x in layer_list.split(',')]
    qs = _tensor_quantizers(model)
    _reset_calibration(qs)
    _say("------------- 8-bit Re-SET -------------")
    _say(len(wanted))
    assert len(wanted) > 0
    for i in range(len(qs) // 2):
        if i in wanted:
            _say(qs[2 * i].name, i)
            _say(qs[2 * i + 1].name, i)
            _promote(qs[2 * i])
            _promote(qs[2 * i + 1])
    _say("------------- 8-bit Re-SET -------------")


def set_8_bit_layer_n(model, l_num):
    """Promote `l_num` layers to 8 bit: always the last two (BERT head), then those with the largest
    calibration MSE (weight + input quantizer summed)."""
    qs = _tensor_quantizers(model)
    mses = [q.mse.item() for q in qs]
    _reset_calibration(qs)
    _say("------------- 8-bit Re-SET -------------")
    _say(l_num)
    assert l_num > 0
    n_last = 2 * 2
    for q in qs[len(qs) - n_last:]:
        _say(q.name)
        _promote(q)
    _say("------------- First and Last end -------------")
    qs, mses = qs[:len(qs) - n_last], mses[:len(mses) - n_last]
    pair_mse = np.array([mses[2 * i] + mses[2 * i + 1] for i in range(len(mses) // 2)])
    budget = (2 * l_num - n_last) // 2
    if budget > 0:
        for i in np.argsort(-pair_mse)[:budget]:
            _say(qs[2 * i].name, pair_mse[i], i)
            _say(qs[2 * i + 1].name, pair_mse[i], i)
            _promote(qs[2 * i])
            _promote(qs[2 * i + 1])
    _say("------------- 8-bit Re-SET -------------")


def load_ant_state_dict(model, checkpoint):
    """Resize every quant_grid buffer to the checkpoint's before load_state_dict(strict=True)
    (eval builds 8-bit tables, a 4-bit checkpoint holds 16 entries)."""
    for name, module in model.named_modules():
        key = name + ".quant_grid"
        if key in checkpoint:
            module.quant_grid.data = checkpoint[key]
