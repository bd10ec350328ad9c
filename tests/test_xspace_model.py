"""Exhaustive check of the algorithm behind the fast CUDA kernel (numpy model,
tests/xspace_model.py) against the oracle: every fp16 bit pattern x many
scales, and randomised fp32.  CPU only."""
import numpy as np
import pytest

import antq_oracle as orc
import xspace_model as xm

f32 = np.float32
ALL_F16 = np.arange(65536, dtype=np.uint16).view(np.float16)


def scales(vmax):
    rng = np.random.default_rng(1)
    s = [0.1, 1.0, 0.5, 2.0 ** -7, 3.0, 0.0625 / vmax, 1e-3, 7.7e-3, 250.0, 6e-6, 1.0 / 3.0]
    s += list(np.exp(rng.uniform(np.log(1e-4), np.log(50.0), 12)))
    return [f32(v) for v in s]


GRIDS = [("ant", k, b, sg) for k in ("int", "flint", "pot", "float2", "apot") for b in (4,) for sg in (True, False)]
GRIDS += [("ant", "int", 3, True), ("ant", "flint", 5, True), ("ant", "int", 8, True), ("ant", "flint", 6, False),
          ("ant", "pot", 5, True)]


@pytest.mark.parametrize("tree,kind,bit,signed", GRIDS)
def test_fp16_exhaustive(tree, kind, bit, signed):
    grid = orc.ant_grid(kind, bit, signed)
    cb = xm.prepare_codebook(grid)
    assert xm.interior_exact(cb)
    gmax = grid.max()
    n_slow = 0
    for s in scales(gmax):
        alpha = f32(s * gmax)
        s_eff = f32(alpha / gmax)

        def exact(xs):
            return orc.ant_forward(xs, alpha, grid, per_row=False)
        got, slow = xm.forward_fast(ALL_F16, s_eff, cb, np.float16, exact)
        ref = orc.ant_forward(ALL_F16, alpha, grid, per_row=False)
        same = (got.view(np.uint16) == ref.view(np.uint16)) | (np.isnan(got) & np.isnan(ref))
        assert same.all(), (kind, bit, signed, s, ALL_F16[~same][:5], got[~same][:5], ref[~same][:5])
        n_slow += slow.sum()
    assert n_slow < 65536 * 23          # the fast window is actually used


@pytest.mark.parametrize("tree,kind,bit,signed", GRIDS[:6])
def test_fp32_random(tree, kind, bit, signed):
    grid = orc.ant_grid(kind, bit, signed)
    cb = xm.prepare_codebook(grid)
    gmax = grid.max()
    rng = np.random.default_rng(7)
    for s in scales(gmax)[:8]:
        alpha = f32(s * gmax)
        s_eff = f32(alpha / gmax)
        x = (rng.standard_normal(20000) * 4 * s_eff * gmax / 3).astype(f32)
        # add points hugging every threshold in x-space
        near = (cb["thr"].astype(np.float64) * float(s_eff)).astype(f32)
        for k in range(-3, 4):
            x = np.concatenate([x, xm._unord(xm._ord(near) + k)])

        def exact(xs):
            return orc.ant_forward(xs, alpha, grid, per_row=False)
        got, _ = xm.forward_fast(x, s_eff, cb, f32, exact)
        ref = orc.ant_forward(x, alpha, grid, per_row=False)
        same = (got.view(np.uint32) == ref.view(np.uint32)) | (np.isnan(got) & np.isnan(ref))
        assert same.all(), (kind, s, x[~same][:5], got[~same][:5], ref[~same][:5])


def test_olive_concat_thresholds():
    """The unsorted normal+outlier grid: ties between a normal and an outlier
    value resolve to the OUTLIER on both sides (-40 -> -48)."""
    grid = np.concatenate([orc.olive_flint_grid(4, True), orc.olive_outlier_grid(4, True)])
    cb = xm.prepare_codebook(grid)
    d = np.concatenate([np.linspace(-600, 600, 24001), [40.0, -40.0, 3.0, -3.0, 56.0, -56.0]]).astype(f32)
    rank = (d[:, None] >= cb["thr"][None, :]).sum(1)
    z = orc.scan(d, grid)
    assert np.array_equal(cb["levels"][rank], z)
    zc, codes = orc.scan(d, grid, want_codes=True)
    assert np.array_equal(cb["codes"][rank], codes)


FOLDED = [("int", 3, True), ("int", 4, True), ("int", 5, True), ("flint", 4, True), ("pot", 4, True), ("float2", 4, True)]


@pytest.mark.parametrize("kind,bit,signed", FOLDED)
def test_folded_xspace_exhaustive(kind, bit, signed):
    """Symmetric codebooks and signed int-k ('symmetric + one extra negative level', ANTQ_CB_SYMX) as the stream kernel
    runs them: magnitude thresholds X / Xn on |x| plus one signed compare for the extra level -- every fp16 pattern."""
    grid = orc.ant_grid(kind, bit, signed)
    cb = xm.prepare_codebook(grid)
    fs = xm.fold_signs(cb)
    assert fs is not None and fs["symx"] == (kind == "int") and fs["nt"] == 2 ** (bit - 1) - 1
    gmax = grid.max()
    for s in scales(gmax)[:14]:
        alpha = f32(s * gmax)
        s_eff = f32(alpha / gmax)

        def exact(xs):
            return orc.ant_forward(xs, alpha, grid, per_row=False)
        got, _ = xm.forward_fast_folded(ALL_F16, s_eff, cb, np.float16, exact)
        ref = orc.ant_forward(ALL_F16, alpha, grid, per_row=False)
        same = (got.view(np.uint16) == ref.view(np.uint16)) | (np.isnan(got) & np.isnan(ref))
        assert same.all(), (kind, bit, s, ALL_F16[~same][:5], got[~same][:5], ref[~same][:5])


@pytest.mark.parametrize("kind,bit,signed", FOLDED + [("flint", 4, False), ("int", 4, False)])
def test_dspace_chain_exhaustive(kind, bit, signed):
    """The short-row kernel's algorithm (true division, compare chain against the exact d-space thresholds, literal STE)
    on every fp16 pattern; asymmetric codebooks compare d itself."""
    grid = orc.ant_grid(kind, bit, signed)
    cb = xm.prepare_codebook(grid)
    gmax = grid.max()
    for s in scales(gmax)[:14]:
        alpha = f32(s * gmax)
        s_eff = f32(alpha / gmax)

        def exact(xs):
            return orc.ant_forward(xs, alpha, grid, per_row=False)
        got, _ = xm.forward_dspace(ALL_F16, s_eff, cb, np.float16, exact)
        ref = orc.ant_forward(ALL_F16, alpha, grid, per_row=False)
        same = (got.view(np.uint16) == ref.view(np.uint16)) | (np.isnan(got) & np.isnan(ref))
        assert same.all(), (kind, bit, s, ALL_F16[~same][:5], got[~same][:5], ref[~same][:5])
