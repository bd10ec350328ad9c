"""antquant.multihead_attention: nn.MultiheadAttention with four TensorQuantizers
(A/antquant/multihead_attention.py:486-687).  The class is built in antq.layers."""
import _bootstrap  # noqa: F401
from quant_modules import MultiheadAttentionQuantizer, TensorQuantizer  # noqa: F401
