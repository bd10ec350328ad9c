"""GPU parity of the round-2 entry points: packed 4-bit codes (antq_encode_p4 / antq_decode_p4), the fused QAT backward
(antq_fakequant_backward) and the fused calibration (antq_calibrate) -- against the oracle, against the exact kernels
they replace, and for run-to-run determinism."""
import numpy as np
import pytest
import torch

import antq_oracle as orc
from gpu_util import assert_bit_equal, dev, to_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def antq():
    import antq as m
    return m


def _cb(antq, grid, outl=None):
    return antq.prepare_codebook(torch.from_numpy(np.asarray(grid, dtype=np.float32)).to(dev()),
                                 None if outl is None else torch.from_numpy(np.asarray(outl, dtype=np.float32)).to(dev()))


# ------------------------------------------------------------------ packed codes
@pytest.mark.parametrize("dtype", ["f16", "f32"])
@pytest.mark.parametrize("kind,signed", [("flint", True), ("int", True), ("pot", True), ("flint", False), ("int", False),
                                         ("float2", True), ("apot", False)])
def test_p4_codes_ant(antq, kind, signed, dtype):
    rng = np.random.default_rng(3)
    grid = orc.ant_grid(kind, 4, signed)
    cb = _cb(antq, grid)
    rows, cols = 200, 1536
    x = (rng.standard_normal((rows, cols)) * 0.05).astype(np.float32)
    if not signed:
        x = np.abs(x)
    alpha = (np.abs(x).max(1) * rng.uniform(0.76, 1.2, rows)).astype(np.float32)          # search_mse never goes below 0.75
    if dtype == "f16":
        x = x.astype(np.float16)
    xd, ad = torch.from_numpy(x).to(dev()), torch.from_numpy(alpha).to(dev())
    codes, bad = antq.encode_p4(xd, ad, cb, True)
    assert codes.dtype == torch.uint8 and codes.numel() == rows * cols // 2
    assert int(bad.item()) == 0
    # the nibbles are the scan's code indices (oracle), low nibble = even element
    _, cref = orc.ant_forward(x, alpha, grid, per_row=True, want_codes=True)
    c = to_np(codes).reshape(rows, cols // 2)
    np.testing.assert_array_equal(c & 15, cref[:, 0::2])
    np.testing.assert_array_equal(c >> 4, cref[:, 1::2])
    # decode(encode(x)) == fakequant(x), bit for bit
    y = antq.fakequant(xd, ad, cb, True)
    z = antq.decode_p4(codes, ad, cb, x.shape, xd.dtype, True)
    assert_bit_equal(to_np(z), to_np(y), "decode vs fakequant")
    assert_bit_equal(to_np(z), orc.ant_forward(x, alpha, grid, per_row=True), "decode vs oracle")
    # per-tensor scale, and values far outside the window are REPORTED as inexact rather than silently wrong
    a0 = torch.tensor([float(np.abs(x.astype(np.float32)).max()) * 0.9], device=dev())
    codes0, bad0 = antq.encode_p4(xd, a0, cb, False)
    assert int(bad0.item()) == 0
    assert_bit_equal(to_np(antq.decode_p4(codes0, a0, cb, x.shape, xd.dtype, False)), to_np(antq.fakequant(xd, a0, cb, False)), "per-tensor")
    x2 = x.copy(); x2[0, 0] = np.nan; x2[1, 1] = np.inf
    _, bad2 = antq.encode_p4(torch.from_numpy(x2).to(dev()), ad, cb, True)
    assert int(bad2.item()) >= 2


@pytest.mark.parametrize("kind", ["flint", "int"])
def test_p4_codes_olive_pairs(antq, kind):
    """OliVe: nibble 15 marks the victim, the other nibble indexes `outliers`; decode reproduces the OVP forward."""
    rng = np.random.default_rng(5)
    grid, outl = orc.olive_grid(kind, 4, True), orc.olive_outlier_grid(4, True)
    cb = _cb(antq, grid, outl)
    rows, cols = 128, 1024
    x = rng.standard_normal((rows, cols)).astype(np.float32)
    idx = rng.integers(0, x.size, x.size // 100)
    x.reshape(-1)[idx] *= rng.choice([8.0, 20.0, 60.0], idx.size)
    x.reshape(-1)[idx[:60] ^ 1] *= 30.0                       # outliers next to outliers
    alpha = (3 * x.std(1)).astype(np.float32)
    x = x.astype(np.float16)
    xd, ad = torch.from_numpy(x).to(dev()), torch.from_numpy(alpha).to(dev())
    codes, bad = antq.encode_p4(xd, ad, cb, True, ovp=True)
    ref = orc.olive_forward(x, alpha, grid, outl, per_row=True)
    z = antq.decode_p4(codes, ad, cb, x.shape, xd.dtype, True, ovp=True)
    nbad = int(bad.item())
    mism = (to_np(z).view(np.uint16) != ref.view(np.uint16)) & ~(np.isnan(to_np(z)) & np.isnan(ref))
    assert mism.sum() <= nbad, (int(mism.sum()), nbad)            # every mismatch was counted ...
    assert nbad <= 0.001 * x.size                                   # ... and they are the rare far-out-of-window values
    c = to_np(codes)
    lo, hi = c & 15, c >> 4
    vict = (lo == 15) | (hi == 15)
    assert vict.any() and not ((lo == 15) & (hi == 15)).any()
    # a pair with a victim decodes to (outlier, 0) / (0, outlier)
    zz = to_np(z).reshape(-1, 2).astype(np.float32)
    assert (np.minimum(np.abs(zz[vict.reshape(-1)][:, 0]), np.abs(zz[vict.reshape(-1)][:, 1])) == 0).all()


# ------------------------------------------------------------------ backward
@pytest.mark.parametrize("per_row", [True, False])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_fused_backward(antq, per_row, dtype):
    g0 = torch.Generator(device="cpu").manual_seed(1)
    rows, cols = 96, 3000
    x = (torch.randn(rows, cols, generator=g0) * 0.1).to(dtype).to(dev())
    g = (torch.randn(rows, cols, generator=g0)).to(dtype).to(dev())
    grid = orc.ant_grid("flint", 4, True)
    cb = _cb(antq, grid)
    alpha = (x.float().abs().amax(1) * 0.8) if per_row else x.float().abs().max().reshape(1) * 0.8
    out = antq.fakequant(x, alpha, cb, per_row)
    gmax = float(grid.max())
    gx, ga = antq.fakequant_backward(g, x, out, alpha, gmax, per_row)
    # tensor / 0-dim CUDA tensor = IEEE division, as in the reference (`alpha / torch.max(quant_grid)`); dividing by a
    # Python / numpy scalar would silently become a multiplication by the rounded reciprocal
    gm = torch.tensor(gmax, dtype=torch.float32, device=dev())
    s = (alpha / gm).reshape(-1, 1) if per_row else alpha / gm
    gx_ref = ((g.float() * s) / s).to(dtype)                                  # autograd's mul-then-div, bit for bit
    assert torch.equal(gx.view(torch.int16 if dtype == torch.float16 else torch.int32),
                       gx_ref.view(torch.int16 if dtype == torch.float16 else torch.int32))
    qd = (out.double() - x.double()) / s.double()
    ga_ref = ((g.double() * qd).sum(1) if per_row else (g.double() * qd).sum().reshape(1)) / gmax
    rel = ((ga.double() - ga_ref).abs() / ga_ref.abs().clamp_min(1e-12)).max()
    assert float(rel) < 1e-5, float(rel)
    # deterministic: fixed-order reduction, no atomics
    gx2, ga2 = antq.fakequant_backward(g, x, out, alpha, gmax, per_row)
    assert torch.equal(ga, ga2) and torch.equal(gx, gx2)
    # partial requests
    gx3, ga3 = antq.fakequant_backward(g, x, out, alpha, gmax, per_row, need_grad_x=False)
    assert gx3 is None and torch.equal(ga3, ga)
    gx4, ga4 = antq.fakequant_backward(g, x, out, alpha, gmax, per_row, need_grad_alpha=False)
    assert ga4 is None and torch.equal(gx4, gx)


# ------------------------------------------------------------------ calibration
@pytest.mark.parametrize("per_row", [True, False])
@pytest.mark.parametrize("dtype", ["f32", "f16"])
def test_calibrate_matches_exact_sweep(antq, per_row, dtype):
    """antq_calibrate (closed-form scoring, several codebooks per launch) against antq_mse_sweep (the literal arithmetic
    per element, one codebook per launch): same alpha on >= 99 % of the rows, summed scores within 1e-5, identical type
    ranking; bit-reproducible run to run; literal scoring for the grids that are not piecewise uniform."""
    rng = np.random.default_rng(8)
    rows, cols = 192, 2304
    x = (rng.standard_normal((rows, cols)) * 0.05).astype(np.float32)
    x[:, :8] *= 5.0
    if dtype == "f16":
        x = x.astype(np.float16)
    xd = torch.from_numpy(x).to(dev())
    kinds = ["int", "flint", "pot", "float2", "apot"]
    grids = [orc.ant_grid(k, 4, True) for k in kinds]
    cbs = [_cb(antq, g) for g in grids]
    ratios = torch.tensor([i * 0.01 for i in range(75, 150)], dtype=torch.float32, device=dev())
    base = antq.absmax(xd, per_row)
    alpha, score, idx = antq.calibrate(xd, base, ratios, cbs, per_row, want_index=True)
    assert alpha.shape == (len(kinds), rows if per_row else 1)
    ncols = cols if per_row else rows * cols
    for k, cb in enumerate(cbs):
        err = antq.mse_sweep(xd, base, ratios, cb, per_row) / ncols                  # [n_cand, rows] float64
        best, _ = err.min(dim=0)
        first = (err == best.unsqueeze(0)).to(torch.int8).argmax(dim=0)
        same = (first.to(torch.int32) == idx[k]).float().mean()
        assert float(same) >= (0.99 if per_row else 1.0), (kinds[k], float(same))
        ref_score = best.sum()
        assert abs(float(score[k]) - float(ref_score)) <= 1e-5 * float(ref_score), (kinds[k], float(score[k]), float(ref_score))
        a_ref = base * ratios[first]
        agree = (alpha[k] == a_ref).float().mean()
        assert float(agree) >= (0.99 if per_row else 1.0)
    # one codebook per launch gives the same numbers as all at once
    a1, s1 = antq.calibrate(xd, base, ratios, cbs[1:2], per_row)
    assert torch.equal(a1[0], alpha[1]) and torch.equal(s1[0], score[1])
    # run-to-run determinism
    alpha2, score2 = antq.calibrate(xd, base, ratios, cbs, per_row)
    assert torch.equal(alpha, alpha2) and torch.equal(score, score2)


def test_calibrate_olive_pairs_and_dead_rows(antq):
    rng = np.random.default_rng(2)
    rows, cols = 64, 2048
    x = rng.standard_normal((rows, cols)).astype(np.float32)
    x.reshape(-1)[rng.integers(0, x.size, x.size // 100)] *= 25.0
    x[7] = 0.0                                                              # dead row: every candidate scores NaN
    xd = torch.from_numpy(x).to(dev())
    grid, outl = orc.olive_grid("flint", 4, True), orc.olive_outlier_grid(4, True)
    cb = _cb(antq, grid, outl)
    ratios = torch.tensor([i * 0.01 for i in range(75, 250, 2)], dtype=torch.float32, device=dev())
    v = xd.float()
    base = torch.maximum((v.mean(1) + 3 * v.std(1)).abs(), (v.mean(1) - 3 * v.std(1)).abs())
    alpha, score, idx = antq.calibrate(xd, base, ratios, [cb], True, ovp=True, want_index=True)
    err = antq.mse_sweep(xd, base, ratios, cb, True, ovp=True) / cols
    live = torch.ones(rows, dtype=torch.bool, device=dev()); live[7] = False
    best, _ = err[:, live].min(dim=0)
    first = (err[:, live] == best.unsqueeze(0)).to(torch.int8).argmax(dim=0)
    assert torch.equal(first.to(torch.int32), idx[0][live])                  # literal arithmetic on both sides: identical
    assert int(idx[0][7]) == -1 and float(alpha[0][7]) == float(base[7])   # the reference leaves alpha at its start value
    assert float(score[0]) >= 1e10                                          # ... and adds its 1e10 start score (A/...:297)
