"""antquant.quant_modules for ANT (MICRO'22), B200 edition.

Same public names as ant_quantization/antquant/quant_modules.py -- QuantBase, Quantizer,
TensorQuantizer, Conv2dQuantizer, LinearQuantizer -- and the same names leak through
`from quant_modules import *` (logging, torch, nn, F, Tensor, np, dist, quant_cuda and the
quant_affine helpers), because the reference drivers rely on that (A/ImageNet/main.py:14-17,91).
The implementation lives in the shared core package `antq` (fused sm_100a kernels).
"""
import _bootstrap  # noqa: F401
import logging
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor
import numpy as np
import quant_cuda
import torch.distributed as dist
from quant_affine import *

from antq.layers import make_layers
from antq.quantizer import QuantBase, Quantizer, TensorQuantizer

Conv2dQuantizer, LinearQuantizer, _Conv1dQuantizer, MultiheadAttentionQuantizer = make_layers(TensorQuantizer)
for _c in (Conv2dQuantizer, LinearQuantizer, MultiheadAttentionQuantizer):
    _c.__module__ = __name__
# names BASELINE.json / papers use for the same wrappers (the reference classes are *Quantizer)
QuantConv2d, QuantLinear = Conv2dQuantizer, LinearQuantizer
