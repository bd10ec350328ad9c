"""The reference's OWN CUDA kernel (built unmodified from /root/reference by oracle/build_ref.py into
oracle/_ref/) as a second oracle on the GPU: it pins the C oracle and the product kernels to the code the
reference actually ships, not only to restatements of it."""
import glob
import importlib.util
import os

import numpy as np
import pytest
import torch

import antq_oracle as orc
from gpu_util import assert_bit_equal, dev, to_np

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def refk():
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "ref_quant_cuda*.so"))
    if not so:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    spec = importlib.util.spec_from_file_location("ref_quant_cuda", so[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_forward(refk, x, alpha, grid, per_row):
    """A/antquant/quant_modules.py:535-551 as the same sequence of torch ops around the reference kernel."""
    scale = alpha / torch.max(grid)
    data = (x.view(x.shape[0], -1) / scale).view(x.shape) if per_row else x / scale
    q, _ = refk.quant(data.view(-1), grid.type_as(data))
    q = q.view(x.shape)
    t = (q - data) + data
    return (t.view(t.shape[0], -1) * scale).view(x.shape) if per_row else t * scale


@pytest.mark.parametrize("kind,bit,signed", [("flint", 4, True), ("int", 4, True), ("pot", 4, False), ("float2", 4, True),
                                             ("int", 8, True), ("flint", 6, False)])
def test_reference_kernel_vs_oracle_and_product(refk, kind, bit, signed):
    import antq
    rng = np.random.default_rng(9)
    grid = orc.ant_grid(kind, bit, signed)
    g = torch.from_numpy(grid).to(dev())
    x = (rng.standard_normal(200000) * 6).astype(np.float32)
    x[:8] = [np.nan, np.inf, -np.inf, 0.0, 102395.0, 102500.0, -3e5, 1e-30]
    mids = ((np.unique(grid)[:-1].astype(np.float64) + np.unique(grid)[1:]) / 2).astype(np.float32)
    x[8:8 + mids.size] = mids                                   # exact ties
    z_ref, idx_ref = refk.quant(torch.from_numpy(x).to(dev()), g)
    assert float(idx_ref.abs().max()) == 0.0                   # the reference never writes its index output
    z_orc = orc.scan(x, grid)
    assert_bit_equal(to_np(z_ref), z_orc, "reference kernel vs C oracle", allow_zero_sign=True)
    cb = antq.prepare_codebook(g)
    z_ours = antq.lut_nearest(torch.from_numpy(x).to(dev()), cb)
    assert_bit_equal(to_np(z_ours), to_np(z_ref), "product lut_nearest vs reference kernel", allow_zero_sign=True)


@pytest.mark.parametrize("per_row", [True, False])
def test_reference_forward_sequence_vs_product(refk, per_row):
    import antq
    rng = np.random.default_rng(10)
    grid = orc.ant_grid("flint", 4, True)
    g = torch.from_numpy(grid).to(dev())
    x = torch.from_numpy((rng.standard_normal((256, 2048)) * 0.02).astype(np.float32)).to(dev())
    x[0, :4] = torch.tensor([float("nan"), float("inf"), 5.0, -5.0], device=dev())
    if per_row:
        alpha = (x.nan_to_num(0, 0, 0).abs().amax(1) * 0.9).unsqueeze(1)
    else:
        alpha = torch.tensor(0.07, device=dev())
    y_ref = ref_forward(refk, x, alpha, g, per_row)
    cb = antq.prepare_codebook(g)
    y = antq.fakequant(x, alpha.reshape(-1), cb, per_row)
    assert_bit_equal(to_np(y), to_np(y_ref), "product fakequant vs reference op sequence")
