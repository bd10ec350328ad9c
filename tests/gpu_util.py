"""Helpers for the -m gpu parity tests: call the product path (C ABI through
ant-quantization_b200/antq) and compare BIT FOR BIT against the CPU oracle."""
import numpy as np
import torch

import antq_oracle as orc


def dev():
    return torch.device("cuda:0")


def to_np(t):
    t = t.detach().cpu()
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().view(np.uint16)
    return t.numpy()


def bits(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        return a.view(np.uint32)
    if a.dtype == np.float16:
        return a.view(np.uint16)
    return a


def assert_bit_equal(got, ref, what="", allow_zero_sign=False):
    """NaN matches NaN (any payload); everything else must have identical bits."""
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape and got.dtype == ref.dtype, (what, got.shape, ref.shape, got.dtype, ref.dtype)
    if got.dtype.kind == "f":
        ok = (bits(got) == bits(ref)) | (np.isnan(got) & np.isnan(ref))
        if allow_zero_sign:
            ok |= (got == 0) & (ref == 0)
    else:
        ok = got == ref
    if not ok.all():
        idx = np.argwhere(~ok)[:8]
        raise AssertionError("%s: %d/%d mismatches, first at %s: got %s ref %s" % (
            what, (~ok).sum(), ok.size, idx.tolist(), got[~ok][:8], ref[~ok][:8]))


def oracle_ant(x_np, alpha_np, grid_np, per_row, want_codes=False):
    return orc.ant_forward(x_np, alpha_np, grid_np, per_row, want_codes=want_codes)


def oracle_olive(x_np, alpha_np, grid_np, outl_np, per_row, no_outlier=False, want_codes=False):
    return orc.olive_forward(x_np, alpha_np, grid_np, outl_np, per_row, no_outlier=no_outlier, want_codes=want_codes)
