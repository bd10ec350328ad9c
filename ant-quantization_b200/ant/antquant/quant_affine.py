"""Asymmetric uniform affine helpers.  In the reference this module is star-imported by
quant_modules but nothing calls it; it is kept so that
`from quant_modules import *` exposes the same names.  q = clamp(round(s*x - zp)), x^ = (q + zp)/s."""
import torch
from torch.autograd import Function


def clamp(input, min, max, inplace=False):
    return input.clamp_(min, max) if inplace else torch.clamp(input, min, max)


def _per_channel(t, ref):
    if isinstance(t, torch.Tensor) and t.dim() > 0 and ref.dim() in (2, 4):
        return t.view(-1, *([1] * (ref.dim() - 1)))
    return t


def linear_quantize(input, scale, zero_point, inplace=False):
    scale, zero_point = _per_channel(scale, input), _per_channel(zero_point, input)
    if inplace:
        return input.mul_(scale).sub_(zero_point).round_()
    return scale * input - zero_point


def linear_dequantize(input, scale, zero_point, inplace=False):
    scale, zero_point = _per_channel(scale, input), _per_channel(zero_point, input)
    if inplace:
        return input.add_(zero_point).div_(scale)
    return (input + zero_point) / scale


def asymmetric_linear_quantization_params(num_bits, saturation_min, saturation_max, integral_zero_point=True,
                                          signed=True):
    scale = (2 ** num_bits - 1) / torch.clamp(saturation_max - saturation_min, min=1e-8)
    zero_point = scale * saturation_min
    if integral_zero_point:
        zero_point = zero_point.round() if isinstance(zero_point, torch.Tensor) else float(round(zero_point))
    if signed:
        zero_point = zero_point + 2 ** (num_bits - 1)
    return scale, zero_point


class AsymmetricQuantFunction(Function):
    """Forward only, as in the reference."""

    @staticmethod
    def forward(ctx, x, k, x_min=None, x_max=None):
        scale, zero_point = asymmetric_linear_quantization_params(k, x_min, x_max)
        q = torch.clamp(linear_quantize(x, scale, zero_point).round(), -2 ** (k - 1), 2 ** (k - 1) - 1)
        return linear_dequantize(q, scale, zero_point)

    @staticmethod
    def backward(ctx, grad_output):
        raise NotImplementedError
