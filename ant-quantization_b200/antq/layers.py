"""Quantized layer wrappers of antquant on top of the fused quantizer.

Each wrapper owns two TensorQuantizers and re-fake-quantizes weight and input on every
forward, exactly like the reference (A/antquant/quant_modules.py:582-646,
O/antquant/quant_modules.py:358-450, A/antquant/multihead_attention.py:486-687); the
matmul / conv itself stays a stock PyTorch op.
"""
import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F


CACHE_WEIGHTS = os.environ.get("ANTQ_CACHE_WEIGHTS", "1") != "0"
# "codes": keep the cached weight as packed 4-bit codes + alpha (0.5 byte per element instead of 2 / 4) and decode it on
# every forward (antq_decode_p4: bit-identical values, one cheap pass) -- for models whose fake-quantized fp copy does not
# fit next to the original.  Applies where antq_encode_p4 does (grids of <= 16 entries, even rows) and reproduces every
# value; other layers keep the fp tensor.
CACHE_FORMAT = os.environ.get("ANTQ_CACHE_FORMAT", "tensor")
# Opt-in: eval-mode LinearQuantizer layers with a 4-bit weight and fp16 / bf16 activations keep the
# weight as packed 4-bit codes + one alpha per output channel and run  y = x_q . dequant(W)^T + b  in ONE tcgen05 kernel
# (antq_linear_p4) instead of fake-quantizing the weight to fp16 and calling cuBLAS.  Numerically this is F.linear on
# the fake-quantized operands up to fp32 accumulation order (tests/test_gpu_gemm.py); off by default because the
# reference's F.linear rounds differently in the last bit.
FUSED_LINEAR = os.environ.get("ANTQ_FUSED_LINEAR", "0") == "1"
# With FUSED_FP8 (default on when FUSED_LINEAR is), W4A4 layers whose two grids are exact in FP8 e4m3 (every 4-bit int /
# flint / pot / float grid) feed LEVELS to the FP8 tensor cores (tcgen05.mma kind::f8f6f4, twice the 16-bit rate): the
# integer products are exact in fp32 and both scales are applied in the epilogue.
FUSED_FP8 = os.environ.get("ANTQ_FUSED_FP8", "1") == "1"
FUSED_MIN_ROWS = int(os.environ.get("ANTQ_FUSED_MIN_ROWS", "256"))


def _copy_param(t):
    return None if t is None else nn.Parameter(t.data.clone())


def make_layers(TensorQuantizer):
    """Build the wrapper classes for one flavour of TensorQuantizer."""

    class _TwoQuantizers(nn.Module):
        def __init__(self, mode=None, wbit=None, abit=None, args=None):
            super().__init__()
            assert mode is not None, 'Quantizer is not initilized!'
            op = self._op_handle()
            self.quant_weight = TensorQuantizer(mode=mode, bit=wbit, is_signed=True, is_enable=True, args=args,
                                                operator=op)
            self.quant_input = TensorQuantizer(mode=mode, bit=abit, is_signed=False, is_enable=True, args=args,
                                               operator=op, is_input=True)

        def _op_handle(self):
            return None

        def _set_weight_bias(self, src):
            self.weight = _copy_param(src.weight)
            self.bias = _copy_param(getattr(src, "bias", None))

        # Weight-quant cache: the reference re-fake-quantizes the weight on every forward
        # (A/antquant/quant_modules.py:611-617).  In EVAL mode, when nothing the result depends on has changed -- the
        # weight, alpha, the grid (+ outliers), the enable flags -- and autograd is not recording, the previous result
        # is returned: identical output, one launch and one pass over the weight less per forward.
        # Invalidation: the key holds (data_ptr, _version) of every tensor involved; train(), load_state_dict(),
        # .to()/.cuda()/.half() and invalidate_weight_cache() drop the cache outright; training mode never caches
        # (optimizers update `p.data` in place, which bumps no version counter).  An in-place `.data` edit of an
        # eval-mode module is the one thing the key cannot see: call invalidate_weight_cache() after it, or set
        # `antq.layers.CACHE_WEIGHTS = False` (ANTQ_CACHE_WEIGHTS=0) for the reference's behaviour exactly.
        def invalidate_weight_cache(self):
            self._wq_key = self._wq_val = None
            self._wc_key = self._wc_val = None

        def train(self, mode=True):
            self.invalidate_weight_cache()
            return super().train(mode)

        def _apply(self, fn, *a, **kw):
            self.invalidate_weight_cache()
            return super()._apply(fn, *a, **kw)

        def _load_from_state_dict(self, *a, **kw):
            self.invalidate_weight_cache()
            return super()._load_from_state_dict(*a, **kw)

        def _encode_for_cache(self, weight_q):
            from . import ops
            q, w = self.quant_weight, self.weight
            cb = q._codebook(w.device)
            n_out = cb.info.n_entries - cb.info.n_normal
            cols = w.numel() // w.shape[0]
            if not q.is_perchannel or cols % 2 or (cb.info.n_normal > 15 or n_out > 15 if q._ovp else cb.info.n_entries > 16):
                return None
            codes, bad = ops.encode_p4(w.detach().contiguous(), q.alpha, cb, True, ovp=q._ovp)
            if int(bad.item()) != 0:                         # some value would not be reproduced: keep the fp tensor
                return None
            return (codes, q.alpha.detach().reshape(-1).float().contiguous(), cb, q._ovp)

        def _weight_key(self):
            q, w = self.quant_weight, self.weight
            if not CACHE_WEIGHTS or self.training or not q._is_inited():
                return None
            if torch.is_grad_enabled() and (w.requires_grad or q.alpha.requires_grad) and q.flavor != "olive":
                return None                                  # QAT: the fake-quant must stay on the autograd tape
            o = getattr(q, "outliers", None)
            return (w.data_ptr(), w._version, w.dtype, w.device, tuple(w.shape),
                    q.alpha.data_ptr(), q.alpha._version, q.quant_grid.data_ptr(), q.quant_grid._version,
                    None if o is None else (o.data_ptr(), o._version),
                    q.mode, q.is_enable, q.is_enable_weight, bool(q.is_signed), q.is_perchannel)

        def _quantized(self, input):
            key = self._weight_key()
            if key is not None and key == getattr(self, "_wq_key", None):
                weight = self._wq_val
                if isinstance(weight, tuple):                # (codes, alpha, codebook, ovp): decode, bit-identical
                    from . import ops
                    weight = ops.decode_p4(weight[0], weight[1], weight[2], self.weight.shape, self.weight.dtype, True,
                                           ovp=weight[3])
            else:
                weight = self.quant_weight(self.weight, input)
                key = self._weight_key()                     # the first call calibrates: take the key afterwards
                if key is not None and weight is not self.weight:
                    self._wq_key, self._wq_val = key, weight.detach()
                    if CACHE_FORMAT == "codes":
                        self._wq_val = self._encode_for_cache(weight) or self._wq_val
                else:
                    self._wq_key = self._wq_val = None
            input = self.quant_input(input, self.weight)
            return input, weight

    class Conv2dQuantizer(_TwoQuantizers):
        """Class to quantize given convolutional layer"""

        def _op_handle(self):
            return self._conv_forward

        def set_param(self, conv):
            for k in ("in_channels", "out_channels", "kernel_size", "stride", "padding", "dilation", "groups"):
                setattr(self, k, getattr(conv, k))
            self.quant_weight.alpha.data = torch.ones([self.out_channels, 1])
            self._set_weight_bias(conv)

        def _conv_forward(self, input, weight):
            return F.conv2d(input, weight, self.bias, self.stride, self.padding, self.dilation, self.groups)

        def forward(self, input):
            input, weight = self._quantized(input)
            return self._conv_forward(input, weight)

    class LinearQuantizer(_TwoQuantizers):
        """Class to quantize given linear layer"""

        def _op_handle(self):
            return F.linear

        def set_param(self, linear):
            self.in_features = linear.in_features
            self.out_features = linear.out_features
            self.quant_weight.alpha.data = torch.ones([self.out_features, 1])
            self._set_weight_bias(linear)

        def _weight_codes(self, input):
            """(codes, alpha, codebook) of the fake-quantized weight, cached like the fp tensor (same key), or None when
            the fused kernel does not apply: OliVe outlier pairs, grids of more than 16 entries, shapes it declines,
            or a weight whose fake-quant values the codes would not reproduce bit for bit (antq_encode_p4 counts them)."""
            from . import ops
            q, w = self.quant_weight, self.weight
            if q.mode == "base" or not q.is_enable or not q.is_enable_weight or q._ovp or not w.is_cuda:
                return None
            if not q._is_inited():
                q(w, input)                                            # calibrates exactly as the unfused path would
            key = self._weight_key()
            if key is None:
                return None
            if key == getattr(self, "_wc_key", None):
                return self._wc_val
            val = None
            cb = q._codebook(w.device)
            if cb.info.n_entries <= 16 and w.shape[1] % 64 == 0 and w.shape[0] % 256 == 0 and w.dtype in (torch.float16, torch.bfloat16):
                codes, bad = ops.encode_p4(w.detach(), q.alpha, cb, True)
                if int(bad.item()) == 0:                               # one synchronisation, when the cache is filled
                    val = (codes, q.alpha.detach().reshape(-1).float().contiguous(), cb)
            self._wc_key, self._wc_val = key, val
            return val

        def forward(self, input):
            # (small batches -- fewer rows than one 256-row tile per SM wave -- are latency-bound in the fused kernel:
            #  they keep the cached fp16 weight + cuBLAS)
            if FUSED_LINEAR and not self.training and input.is_cuda and input.dtype in (torch.float16, torch.bfloat16) \
                    and input.numel() // max(input.shape[-1], 1) >= FUSED_MIN_ROWS \
                    and not (torch.is_grad_enabled() and input.requires_grad):
                pack = self._weight_codes(input)
                if pack is not None:
                    from . import ops, _lib
                    xq = self.quant_input(input, self.weight)
                    qi = self.quant_input
                    if FUSED_FP8 and xq is not input and not qi.is_perchannel and not qi._ovp and input.shape[-1] % 128 == 0:
                        xcb = qi._codebook(input.device)
                        if (xcb.info.flags & _lib.CB_PU_E4M3) and (pack[2].info.flags & _lib.CB_PU_E4M3):
                            return ops.linear_p4_fp8(xq, qi.alpha, xcb, pack[0], pack[1], pack[2], self.out_features, self.bias)
                    return ops.linear_p4(xq, pack[0], pack[1], pack[2], self.out_features, self.bias)
            input, weight = self._quantized(input)
            return F.linear(input, weight, self.bias)

    class Conv1dQuantizer(_TwoQuantizers):
        """HF GPT-2 style Conv1D (weight is [in, out]; per-'channel' scale follows dim 0 as in the reference)"""

        def _op_handle(self):
            return self._conv_forward

        def set_param(self, conv):
            self.nf = conv.nf
            self._set_weight_bias(conv)

        def _conv_forward(self, x, weight):
            size_out = x.size()[:-1] + (self.nf,)
            return torch.addmm(self.bias, x.view(-1, x.size(-1)), weight).view(size_out)

        def forward(self, input):
            input, weight = self._quantized(input)
            return self._conv_forward(input, weight)

    class MultiheadAttentionQuantizer(nn.Module):
        """Self-attention with fake-quantized in/out projections (torchvision ViT).  As in the
        reference, key and value are REPLACED by the quantized query (A/...multihead_attention.py:663-666),
        all four quantizers are signed, and the out-projection input is quantized after the head merge (:459)."""

        def __init__(self, mode=None, wbit=None, abit=None, args=None):
            super().__init__()
            assert mode is not None, 'Quantizer is not initilized!'
            kw = dict(mode=mode, is_signed=True, is_enable=True, args=args)
            self.in_quant_weight = TensorQuantizer(bit=wbit, **kw)
            self.in_quant_input = TensorQuantizer(bit=abit, is_input=True, **kw)
            self.out_quant_weight = TensorQuantizer(bit=wbit, **kw)
            self.out_quant_input = TensorQuantizer(bit=abit, is_input=True, **kw)

        def set_param(self, MA):
            self.embed_dim, self.num_heads, self.dropout = MA.embed_dim, MA.num_heads, MA.dropout
            self.kdim = MA.kdim if MA.kdim is not None else MA.embed_dim
            self.vdim = MA.vdim if MA.vdim is not None else MA.embed_dim
            self.batch_first = MA.batch_first
            self._qkv_same_embed_dim = self.kdim == self.embed_dim and self.vdim == self.embed_dim
            if not self._qkv_same_embed_dim:
                # the reference cannot build this case either (undefined `factory_kwargs`, A/...multihead_attention.py:570)
                raise NotImplementedError("MultiheadAttentionQuantizer: packed in_proj (kdim == vdim == embed_dim) only")
            self.head_dim = self.embed_dim // self.num_heads
            assert self.head_dim * self.num_heads == self.embed_dim, "embed_dim must be divisible by num_heads"
            self.add_zero_attn = MA.add_zero_attn
            self.in_quant_weight.alpha.data = torch.ones([self.embed_dim * 3, 1])
            self.out_quant_weight.alpha.data = torch.ones([self.embed_dim, 1])
            self.in_proj_weight = _copy_param(MA.in_proj_weight)
            self.in_proj_bias = _copy_param(MA.in_proj_bias)
            self.out_proj_weight = _copy_param(MA.out_proj.weight)
            self.out_proj_bias = _copy_param(MA.out_proj.bias)
            self.bias_k = _copy_param(MA.bias_k)
            self.bias_v = _copy_param(MA.bias_v)

        def forward(self, query, key=None, value=None, key_padding_mask=None, need_weights=True, attn_mask=None,
                    average_attn_weights=True):
            """A/antquant/multihead_attention.py:663-687 + the attention math of :214-480 (packed in-projection,
            bias_k / bias_v rows, the zero-attention column, bool / float masks merged with the key-padding mask)."""
            w_in = self.in_quant_weight(self.in_proj_weight)
            x = self.in_quant_input(query)                               # key and value ARE the quantized query (:665-666)
            w_out = self.out_quant_weight(self.out_proj_weight)
            batched = x.dim() == 3
            if not batched:
                x = x.unsqueeze(1)
                if key_padding_mask is not None:
                    key_padding_mask = key_padding_mask.unsqueeze(0)
            elif self.batch_first:
                x = x.transpose(0, 1)
            L, N, E = x.shape                                            # (seq, batch, embed)
            H, hd = self.num_heads, self.head_dim
            q, k, v = F.linear(x, w_in, self.in_proj_bias).chunk(3, dim=-1)
            mask = attn_mask
            if mask is not None:
                if mask.dtype == torch.uint8:
                    mask = mask.to(torch.bool)
                if mask.dim() == 2:
                    if tuple(mask.shape) != (L, L):
                        raise RuntimeError("The shape of the 2D attn_mask is %s, but should be %s." % (tuple(mask.shape), (L, L)))
                    mask = mask.unsqueeze(0)
                elif mask.dim() == 3:
                    if tuple(mask.shape) != (N * H, L, L):
                        raise RuntimeError("The shape of the 3D attn_mask is %s, but should be %s." % (tuple(mask.shape), (N * H, L, L)))
                else:
                    raise RuntimeError("attn_mask's dimension %d is not supported" % mask.dim())
            kpm = key_padding_mask
            if kpm is not None and kpm.dtype == torch.uint8:
                kpm = kpm.to(torch.bool)
            if self.bias_k is not None and self.bias_v is not None:
                k = torch.cat([k, self.bias_k.repeat(1, N, 1)])
                v = torch.cat([v, self.bias_v.repeat(1, N, 1)])
                if mask is not None:
                    mask = F.pad(mask, (0, 1))
                if kpm is not None:
                    kpm = F.pad(kpm, (0, 1))
            q = q.contiguous().view(L, N * H, hd).transpose(0, 1)
            k = k.contiguous().view(k.shape[0], N * H, hd).transpose(0, 1)
            v = v.contiguous().view(v.shape[0], N * H, hd).transpose(0, 1)
            if self.add_zero_attn:
                zeros = (N * H, 1, hd)
                k = torch.cat([k, torch.zeros(zeros, dtype=k.dtype, device=k.device)], dim=1)
                v = torch.cat([v, torch.zeros(zeros, dtype=v.dtype, device=v.device)], dim=1)
                if mask is not None:
                    mask = F.pad(mask, (0, 1))
                if kpm is not None:
                    kpm = F.pad(kpm, (0, 1))
            S = k.size(1)
            if kpm is not None:
                assert tuple(kpm.shape) == (N, S), "expecting key_padding_mask shape of %s, but got %s" % ((N, S), tuple(kpm.shape))
                kpm = kpm.view(N, 1, 1, S).expand(-1, H, -1, -1).reshape(N * H, 1, S)
                if mask is None:
                    mask = kpm
                elif mask.dtype == torch.bool:
                    mask = mask.logical_or(kpm)
                else:
                    mask = mask.masked_fill(kpm, float("-inf"))
            if mask is not None and mask.dtype == torch.bool:
                mask = torch.zeros_like(mask, dtype=q.dtype).masked_fill_(mask, float("-inf"))
            q = q / math.sqrt(hd)
            scores = torch.baddbmm(mask, q, k.transpose(-2, -1)) if mask is not None else torch.bmm(q, k.transpose(-2, -1))
            attn = F.softmax(scores, dim=-1)
            if self.training and self.dropout > 0.0:
                attn = F.dropout(attn, p=self.dropout)
            ctx = torch.bmm(attn, v).transpose(0, 1).contiguous().view(L, N, E)
            out = F.linear(self.out_quant_input(ctx), w_out, self.out_proj_bias)      # (:459-460)
            weights = None
            if need_weights:
                weights = attn.view(N, H, L, S)
                if average_attn_weights:
                    weights = weights.sum(dim=1) / H
            if not batched:
                out = out.squeeze(1)
                weights = None if weights is None else weights.squeeze(0)
            elif self.batch_first:
                out = out.transpose(0, 1)
            return out, weights

    return Conv2dQuantizer, LinearQuantizer, Conv1dQuantizer, MultiheadAttentionQuantizer
