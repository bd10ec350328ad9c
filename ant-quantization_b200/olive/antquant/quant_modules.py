"""antquant.quant_modules for OliVe (ISCA'23), B200 edition.

Same public names as olive_quantization/antquant/quant_modules.py -- QuantBase, Quantizer,
TensorQuantizer, Conv1dQuantizer, Conv2dQuantizer, LinearQuantizer.  Differences from the ANT
flavour (all in antq.quantizer): grids normalised to the outlier threshold 32, abfloat
`outliers` buffer, 3-sigma alpha init, outlier-victim pair masking fused into the forward
kernel, everything under no_grad, no torch.distributed calls.
"""
import _bootstrap  # noqa: F401
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import quant_cuda

from antq.layers import make_layers
from antq.quantizer import QuantBase
from antq.quantizer import OliveQuantizer as Quantizer
from antq.quantizer import OliveTensorQuantizer as TensorQuantizer

Conv2dQuantizer, LinearQuantizer, Conv1dQuantizer, _MHA = make_layers(TensorQuantizer)
for _c in (Conv2dQuantizer, LinearQuantizer, Conv1dQuantizer):
    _c.__module__ = __name__
# names BASELINE.json / papers use for the same wrappers (the reference classes are *Quantizer)
QuantConv2d, QuantLinear = Conv2dQuantizer, LinearQuantizer
