"""Drop-in for the reference's native module `quant_cuda` (A/quant/quant.cpp:26-28):

    quant(x, y) -> (z, idx)

x: 1-D CUDA tensor (float32 / float64 as in the reference; float16 / bfloat16 as an extension),
y: 1-D grid.  z[i] is the grid entry the reference scan selects for x[i] (last minimal entry,
A/quant/quant_kernel.cu:25-37).  The reference allocates `idx` but never writes it (all zeros,
same dtype as x); here it carries the selected entry's scan index, which callers that ignored
it keep ignoring.  Unlike the reference, a CPU tensor raises instead of silently returning zeros.
"""
import _bootstrap  # noqa: F401
import torch
from antq import ops


def quant(x, y):
    if not x.is_cuda:
        raise RuntimeError("quant_cuda.quant: x must be a CUDA tensor (the reference silently returns zeros here)")
    cb = ops.prepare_codebook(y.to(x.device))
    flat = x.reshape(-1)
    if flat.dtype == torch.float64:                      # `float x_v = x[idx]`: doubles are narrowed
        z, codes = ops.lut_nearest(flat.float().contiguous(), cb, want_codes=True)
        z = z.double()
    else:
        z, codes = ops.lut_nearest(flat.contiguous(), cb, want_codes=True)
    return z.view(x.shape), codes.to(x.dtype).view(x.shape)
