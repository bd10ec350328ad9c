"""GPU parity of the closed-form (piecewise-uniform) kernels, antq_pu_stream_kernel / antq_pu_short_kernel, against the
oracle: every fp16 bit pattern x 16 scales for every grid family they serve (int 3..8 bit, unsigned 4-bit, 5/6-bit
flint / pot / float), random fp32 / fp16 / bf16 with NaN / Inf / dead rows / representable ties / ragged tails, and the
codebook analysis itself against tests/pu_model.py (whose arithmetic the CPU suite checks exhaustively)."""
import numpy as np
import pytest
import torch

import antq_oracle as orc
import pu_model as pm
from gpu_util import assert_bit_equal, dev, to_np

pytestmark = pytest.mark.gpu

SCALES = [0.1, 1.0, 0.5, 2.0 ** -7, 3.0, 1e-3, 7.7e-3, 250.0, 6e-6, 1.0 / 3.0, 0.0123, 0.37, 2.5e-2, 1.7e-4, 9.0, 41.0]
PU_GRIDS = [("int", b, sg) for b in (3, 4, 5, 6, 7, 8) for sg in (True, False)]
PU_GRIDS += [("flint", b, sg) for b in (4, 5, 6) for sg in (True, False)]
PU_GRIDS += [("pot", 4, True), ("pot", 4, False), ("float2", 4, False), ("float3", 6, True), ("float3", 5, False)]


@pytest.fixture(scope="module")
def antq():
    import antq as m
    return m


def _cb(antq, grid, outl=None):
    return antq.prepare_codebook(torch.from_numpy(np.asarray(grid, dtype=np.float32)).to(dev()),
                                 None if outl is None else torch.from_numpy(np.asarray(outl, dtype=np.float32)).to(dev()))


def _run(antq, x_np, alpha_np, cb, per_row, flags):
    x = x_np if isinstance(x_np, torch.Tensor) else torch.from_numpy(x_np)
    a = torch.from_numpy(np.ascontiguousarray(alpha_np, dtype=np.float32)).reshape(-1).to(dev())
    return antq.fakequant(x.to(dev()), a, cb, per_row, flags=flags)


def test_codebook_analysis_matches_model(antq):
    from antq import _lib
    cases = [(orc.ant_grid(k, b, s), "%s-%d-%s" % (k, b, s)) for k, b, s in PU_GRIDS]
    cases += [(orc.ant_grid("apot", 4, False), "apot-u"), (orc.ant_grid("apot", 4, True), "apot-s"),
              (np.concatenate([orc.olive_int_grid(4, True), orc.olive_outlier_grid(4, True)]), "olive int+abfloat"),
              (orc.olive_flint_grid(4, True), "olive flint"), (orc.olive_int_grid(4, False), "olive int u")]
    for grid, name in cases:
        info = _cb(antq, grid).info
        model = pm.analyze(grid)
        assert bool(info.flags & _lib.CB_PU) == (model is not None), name
        if model is not None:
            assert bool(info.flags & _lib.CB_PU_UNIFORM) == model["uniform"], name


@pytest.mark.parametrize("kind,bit,signed", PU_GRIDS)
def test_pu_exhaustive_fp16(antq, kind, bit, signed):
    """Every fp16 bit pattern x 16 scales through the closed form: one 65536-element row per scale (stream kernel), and
    the same data as 64-element rows (short kernel: one scale per 64 elements)."""
    from antq import _lib
    grid = orc.ant_grid(kind, bit, signed)
    cb = _cb(antq, grid)
    allh = np.arange(65536, dtype=np.uint16).view(np.float16)
    x = np.tile(allh, (len(SCALES), 1))
    alpha = (np.array(SCALES, dtype=np.float32) * grid.max()).astype(np.float32)
    ref = orc.ant_forward(x, alpha, grid, per_row=True)
    assert antq.fakequant_plan(torch.from_numpy(x).to(dev()), cb, True, flags=_lib.FLAG_FORCE_PU) == 4
    y = _run(antq, x, alpha, cb, True, _lib.FLAG_FORCE_PU)
    assert_bit_equal(to_np(y), ref, "stream")
    xs = x.reshape(-1, 64)
    assert antq.fakequant_plan(torch.from_numpy(xs).to(dev()), cb, True, flags=_lib.FLAG_FORCE_PU) == 5
    ys = _run(antq, xs, np.repeat(alpha, 65536 // 64), cb, True, _lib.FLAG_FORCE_PU)
    assert_bit_equal(to_np(ys).reshape(x.shape), ref, "short")
    # rows of one and two vectors: the lean kernel (window and clamp in t-space)
    for g in (8, 16):
        xl = x.reshape(-1, g)
        yl = _run(antq, xl, np.repeat(alpha, 65536 // g), cb, True, 0)
        assert_bit_equal(to_np(yl).reshape(x.shape), ref, "lean g=%d" % g)


@pytest.mark.parametrize("dtype", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("kind,bit,signed", [("int", 8, True), ("int", 8, False), ("int", 6, True), ("flint", 4, False),
                                             ("flint", 6, False), ("flint", 5, True), ("pot", 4, False), ("int", 4, True)])
def test_pu_random(antq, kind, bit, signed, dtype):
    from antq import _lib
    rng = np.random.default_rng(17 + bit)
    grid = orc.ant_grid(kind, bit, signed)
    cb = _cb(antq, grid)
    rows, cols = 96, 4096
    x = (rng.standard_normal((rows, cols)) * 0.02).astype(np.float32)
    x[rng.integers(0, rows, 50), rng.integers(0, cols, 50)] *= 30          # far outside the clip window
    x[5, 7], x[9, 100], x[11, 4095] = np.nan, np.inf, -np.inf
    x[20] = 0.0                                                           # alpha = 0 row -> NaN in the reference
    if not signed:
        x = np.abs(x)
    alpha = (np.abs(np.nan_to_num(x, nan=0, posinf=0, neginf=0)).max(1) * rng.uniform(0.5, 1.2, rows)).astype(np.float32)
    alpha[3] = np.float32(0.625 * 0.02)                                   # scales that make exact ties representable
    alpha[::7] = np.float32(0.05 * grid.max() / 8)
    alpha[40] = -1.0                                                      # a negative scale: literal arithmetic
    if dtype == "bf16":
        xt = torch.from_numpy(x).to(torch.bfloat16)
        ref = torch.from_numpy(orc.ant_forward(xt.float().numpy(), alpha, grid, per_row=True)).to(torch.bfloat16)
        for shape in ((rows, cols), (rows * cols // 32, 32)):
            a = alpha if shape[0] == rows else np.repeat(alpha, cols // 32)
            y = _run(antq, xt.reshape(shape), a, cb, True, _lib.FLAG_FORCE_PU)
            same = (y.cpu().view(torch.int16) == ref.reshape(shape).view(torch.int16)) | (y.cpu().isnan() & ref.reshape(shape).isnan())
            assert bool(same.all()), (shape, int((~same).sum()))
        return
    if dtype == "f16":
        x = x.astype(np.float16)
    ref = orc.ant_forward(x, alpha, grid, per_row=True)
    assert_bit_equal(to_np(_run(antq, x, alpha, cb, True, _lib.FLAG_FORCE_PU)), ref, "stream per-row")
    assert_bit_equal(to_np(_run(antq, x, alpha, cb, True, _lib.FLAG_FORCE_FLAT)), ref, "flat per-row")
    # scale groups of 8 / 32 / 128 elements: the short kernel
    for g in (8, 32, 128):
        xs = x.reshape(-1, g)
        a = np.repeat(alpha, cols // g)
        assert antq.fakequant_plan(torch.from_numpy(xs).to(dev()), cb, True) == 5
        assert_bit_equal(to_np(_run(antq, xs, a, cb, True, 0)).reshape(x.shape), ref, "short g=%d" % g)
    # per-tensor, ragged length: the tail goes through the literal path
    xt = x.reshape(-1)[: rows * cols - 3]
    fin = np.isfinite(xt.astype(np.float32))
    a0 = np.float32(np.abs(xt[fin].astype(np.float32)).max() * 0.9)
    assert_bit_equal(to_np(_run(antq, xt, a0, cb, False, _lib.FLAG_FORCE_PU)),
                     orc.ant_forward(xt, a0, grid, per_row=False), "per-tensor ragged")
    # in place
    xd = torch.from_numpy(x).to(dev())
    antq.fakequant(xd, torch.from_numpy(alpha).to(dev()), cb, True, out=xd, flags=_lib.FLAG_FORCE_PU)
    assert_bit_equal(to_np(xd), ref, "in place")


@pytest.mark.parametrize("dtype", ["f16", "f32", "bf16"])
@pytest.mark.parametrize("kind,bit,signed", [("int", 8, True), ("flint", 4, False), ("flint", 5, True)])
def test_pu_short_rows_ragged_shapes(antq, kind, bit, signed, dtype):
    """The tiled short-row kernel: row lengths that are not powers of two (rows straddle the 128-vector tiles, the row
    of a vector comes from the 16-bit reciprocal), a partial last tile, the longest rows the plan sends there
    (504 fp16 / 508 fp32 elements), one-vector rows, dead and NaN rows."""
    rng = np.random.default_rng(5 + bit)
    grid = orc.ant_grid(kind, bit, signed)
    cb = _cb(antq, grid)
    vec = 4 if dtype == "f32" else 8
    for cols_vec, rows in ((1, 1031), (3, 517), (5, 77), (9, 333), (25, 41), (63, 19), (127 if vec == 4 else 62, 23), (4, 2), (2, 1)):
        cols = cols_vec * vec
        x = (rng.standard_normal((rows, cols)) * 0.02).astype(np.float32)
        x[rng.integers(0, rows, 9), rng.integers(0, cols, 9)] *= 30
        if rows > 12:
            x[7] = 0.0
            x[11, 0] = np.nan
        if not signed:
            x = np.abs(x)
        if dtype == "f16":
            x = x.astype(np.float16)
        if dtype == "bf16":
            x = torch.from_numpy(x).to(torch.bfloat16).float().numpy()    # bf16-representable values, kept as fp32 for the oracle
        alpha = (np.abs(np.nan_to_num(x.astype(np.float32), nan=0)).max(1) * rng.uniform(0.5, 1.2, rows)).astype(np.float32)
        alpha[::5] = np.float32(0.05 * grid.max() / 8)                    # representable ties
        xt = torch.from_numpy(x).to(dev())
        if dtype == "bf16":
            xt = xt.to(torch.bfloat16)
        if rows > 1:
            assert antq.fakequant_plan(xt, cb, True) == 5, (cols_vec, rows)
        ref = orc.ant_forward(x, alpha, grid, per_row=True)
        if dtype == "bf16":
            y = antq.fakequant(xt, torch.from_numpy(alpha).to(dev()), cb, True).cpu()
            reft = torch.from_numpy(ref).to(torch.bfloat16)
            same = (y.view(torch.int16) == reft.view(torch.int16)) | (y.isnan() & reft.isnan())
            assert bool(same.all()), ("short rows bf16 %d x %d" % (rows, cols), int((~same).sum()))
        else:
            assert_bit_equal(to_np(_run(antq, x, alpha, cb, True, 0)), ref, "short rows %d x %d" % (rows, cols))


OLIVE = [("flint", True), ("flint", False), ("int", True), ("int", False)]


@pytest.mark.parametrize("kind,signed", OLIVE)
def test_pu_olive_pairs_exhaustive_fp16(antq, kind, signed):
    """OliVe outlier-victim pairs through the closed-form kernel (normal levels by the closed form, vectors holding an
    outlier by the pair logic): every fp16 bit pattern next to a small, a large and an outlier-sized neighbour, in both
    slots of the pair, x 8 scales."""
    from antq import _lib
    grid, outl = orc.olive_grid(kind, 4, signed), orc.olive_outlier_grid(4, signed)
    cb = _cb(antq, grid, outl)
    assert cb.info.flags & _lib.CB_PU_OVP, (kind, signed)
    allh = np.arange(65536, dtype=np.uint16).view(np.float16)
    scales = SCALES[:8]
    rows = []
    for sc in scales:
        a = np.float32(sc * grid.max())
        for nb in (0.3, -0.8, 1.04, 1.3, 3.0, -7.0):                          # neighbour, in units of alpha
            for slot in (0, 1):
                r = np.empty(2 * 65536, dtype=np.float16)
                r[slot::2] = allh
                r[1 - slot::2] = np.float16(nb * a if signed or nb > 0 else -nb * a)
                rows.append((r, a))
    x = np.stack([r for r, _ in rows])
    alpha = np.array([a for _, a in rows], dtype=np.float32)
    xt = torch.from_numpy(x).to(dev())
    assert antq.fakequant_plan(xt, cb, True, ovp=True, flags=_lib.FLAG_FORCE_PU) == 4
    ref = orc.olive_forward(x, alpha, grid, outl, per_row=True)
    y = antq.fakequant(xt, torch.from_numpy(alpha).to(dev()), cb, True, ovp=True, flags=_lib.FLAG_FORCE_PU)
    assert_bit_equal(to_np(y), ref, "olive closed form %s %s" % (kind, signed))


@pytest.mark.parametrize("dtype", ["f16", "f32", "bf16"])
@pytest.mark.parametrize("kind,signed", OLIVE)
def test_pu_olive_random(antq, kind, signed, dtype):
    """Random data with a heavy tail (a few percent of outliers: the queue of outlier vectors overflows in some CTAs), NaN /
    Inf, a dead row, a negative scale; per-row and per-tensor scales; the default plan of unsigned OliVe."""
    from antq import _lib
    rng = np.random.default_rng(3)
    grid, outl = orc.olive_grid(kind, 4, signed), orc.olive_outlier_grid(4, signed)
    cb = _cb(antq, grid, outl)
    rows, cols = 64, 8192
    x = (rng.standard_normal((rows, cols)) * 0.02).astype(np.float32)
    x[rng.random((rows, cols)) < 0.002] *= 12                                # outliers
    x[8:12][rng.random((4, cols)) < 0.2] *= 20                               # rows where a fifth of the elements are outliers
    x[3, 5], x[3, 6], x[17, 100], x[30, 4095] = np.nan, 0.01, np.inf, -np.inf
    x[20] = 0.0
    if not signed:
        x = np.abs(x)
    alpha = (3.0 * np.nanstd(np.where(np.isfinite(x), x, np.nan), axis=1) * rng.uniform(0.7, 1.3, rows)).astype(np.float32)
    alpha[20] = 0.0
    alpha[40] = -0.05
    tdt = {"f16": torch.float16, "f32": torch.float32, "bf16": torch.bfloat16}[dtype]
    xt = torch.from_numpy(x).to(tdt)
    xr = xt.float().numpy() if dtype == "bf16" else xt.numpy()
    ref = orc.olive_forward(xr, alpha, grid, outl, per_row=True)
    xd, ad = xt.to(dev()), torch.from_numpy(alpha).to(dev())
    if not signed:
        assert antq.fakequant_plan(xd, cb, True, ovp=True) == 4             # what an OPT fc2 input (post-ReLU) takes
    else:
        # signed 4-bit keeps the two-phase chain with 16-bit data (bf16: only for flint); fp32 has no packed chain
        assert antq.fakequant_plan(xd, cb, True, ovp=True) == (1 if dtype == "f16" or (dtype == "bf16" and kind == "flint") else 4)
    y = antq.fakequant(xd, ad, cb, True, ovp=True, flags=_lib.FLAG_FORCE_PU)
    if dtype == "bf16":
        reft = torch.from_numpy(ref).to(torch.bfloat16)
        same = (y.cpu().view(torch.int16) == reft.view(torch.int16)) | (y.cpu().isnan() & reft.isnan())
        assert bool(same.all()), int((~same).sum())
    else:
        assert_bit_equal(to_np(y), ref, "olive closed form per-row")
        a0 = np.float32(0.07)
        reft = orc.olive_forward(xr.reshape(-1), a0, grid, outl, per_row=False)
        yt = antq.fakequant(xd.view(-1), torch.tensor([a0], device=dev()), cb, False, ovp=True, flags=_lib.FLAG_FORCE_PU)
        assert_bit_equal(to_np(yt), reft, "olive closed form per-tensor")


@pytest.mark.parametrize("signed", [True, False])
def test_pu_olive_8bit_default_plan(antq, signed):
    """OliVe with 8-bit codebooks (255 + 254 entries signed: far beyond any compare chain) takes the closed form on its
    int-8 normal levels and the pair logic (rank search over ~500 thresholds) where an outlier sits."""
    rng = np.random.default_rng(8)
    grid, outl = orc.olive_int_grid(8, signed), orc.olive_outlier_grid(8, signed)
    cb = _cb(antq, grid, outl)
    from antq import _lib
    assert cb.info.flags & _lib.CB_PU_OVP
    rows, cols = 24, 4096
    x = (rng.standard_normal((rows, cols)) * 0.02).astype(np.float32)
    x[rng.random((rows, cols)) < 0.004] *= 14
    x[5, 9], x[7] = np.nan, 0.0
    if not signed:
        x = np.abs(x)
    x = x.astype(np.float16)
    alpha = (3.0 * np.nanstd(x.astype(np.float32), axis=1) * rng.uniform(0.8, 1.2, rows)).astype(np.float32)
    xd, ad = torch.from_numpy(x).to(dev()), torch.from_numpy(alpha).to(dev())
    assert antq.fakequant_plan(xd, cb, True, ovp=True) == 4
    ref = orc.olive_forward(x, alpha, grid, outl, per_row=True)
    assert_bit_equal(to_np(antq.fakequant(xd, ad, cb, True, ovp=True)), ref, "olive 8-bit, default plan")
    a0 = np.float32(0.06)
    reft = orc.olive_forward(x.reshape(-1), a0, grid, outl, per_row=False)
    yt = antq.fakequant(xd.view(-1), torch.tensor([a0], device=dev()), cb, False, ovp=True)
    assert_bit_equal(to_np(yt), reft, "olive 8-bit per-tensor")


def test_default_plans(antq):
    """What a model actually hits: 8-bit int weights and post-ReLU 4-bit activations take the closed form, signed 4-bit
    keeps the chain, OliVe keeps its two-phase chain."""
    x = torch.zeros(4096, 4096, dtype=torch.float16, device=dev())
    plan = lambda grid, per_row, outl=None, ovp=False, t=x: antq.fakequant_plan(t, _cb(antq, grid, outl), per_row, ovp=ovp)
    assert plan(orc.ant_grid("int", 8, True), True) == 4
    assert plan(orc.ant_grid("int", 8, False), False) == 4
    assert plan(orc.ant_grid("flint", 4, False), False) == 4
    assert plan(orc.ant_grid("flint", 6, False), False) == 4
    assert plan(orc.ant_grid("flint", 4, True), True) == 1
    assert plan(orc.ant_grid("int", 4, True), True) == 4                  # uniform grid, per-row scales: the closed form
    assert plan(orc.ant_grid("int", 4, True), False, t=x.view(1, -1)) == 1
    assert plan(orc.olive_grid("flint", 4, True), True, orc.olive_outlier_grid(4, True), True) == 1
    assert plan(orc.ant_grid("int", 8, True), True, t=x.view(-1, 64)) == 5
    assert plan(orc.ant_grid("apot", 4, False), True, t=x.view(-1, 64)) == 3


def test_pu_headline_sizes_against_flat(antq):
    """4096 x 4096 and a 16384-row tensor: closed form == generic kernel everywhere (fp16, int-8 and unsigned flint-4)."""
    from antq import _lib
    g = torch.Generator(device="cuda").manual_seed(3)
    for kind, bit, signed, shape in (("int", 8, True, (4096, 4096)), ("flint", 4, False, (4096, 4096)),
                                     ("int", 8, True, (16384, 4096))):
        grid = orc.ant_grid(kind, bit, signed)
        cb = _cb(antq, grid)
        x = (torch.randn(*shape, device=dev(), generator=g) * 0.03).to(torch.float16)
        if not signed:
            x = x.abs()
        x[::97, ::13] *= 40
        alpha = (x.float().abs().amax(1) * 0.8).contiguous()
        y = antq.fakequant(x, alpha, cb, True)
        yf = antq.fakequant(x, alpha, cb, True, flags=_lib.FLAG_FORCE_FLAT)
        assert torch.equal(y.view(torch.int16), yf.view(torch.int16)), (kind, bit, shape)
        a0 = alpha.mean().reshape(1)
        y = antq.fakequant(x, a0, cb, False)
        yf = antq.fakequant(x, a0, cb, False, flags=_lib.FLAG_FORCE_FLAT)
        assert torch.equal(y.view(torch.int16), yf.view(torch.int16)), (kind, bit, shape, "per-tensor")


@pytest.mark.parametrize("dtype", ["f16", "f32", "bf16"])
@pytest.mark.parametrize("kind,bit,signed", [("flint", 4, True), ("int", 8, True), ("flint", 4, False), ("int", 4, True)])
def test_dynamic_group_scales_single_pass(antq, kind, bit, signed, dtype):
    """antq_fakequant_dynamic: alpha = max|x| of each group * ratio computed in the same pass as the fake-quant (one HBM
    read) == the reference's per-channel path on the [numel / G, G] view with alpha = x.abs().max(1) * ratio."""
    rng = np.random.default_rng(41 + bit)
    grid = orc.ant_grid(kind, bit, signed)
    cb = _cb(antq, grid)
    n = 1 << 17
    x = (rng.standard_normal(n) * 0.05).astype(np.float32)
    x[rng.integers(0, n, 40)] *= 30
    x[1024:1024 + 256] = 0.0                      # dead groups: alpha = 0 -> NaN, as in the reference
    x[5000] = np.nan
    if not signed:
        x = np.abs(x)
    tdt = {"f16": torch.float16, "f32": torch.float32, "bf16": torch.bfloat16}[dtype]
    xt = torch.from_numpy(x).to(tdt)
    xd = xt.to(dev())
    for g in (8, 16, 32, 64, 256, 512):
        for ratio in (1.0, 0.85):
            xf = xt.float().numpy().reshape(-1, g)
            alpha = (np.abs(xf).max(1) * np.float32(ratio)).astype(np.float32)
            alpha[np.isnan(xf).any(1)] = np.nan
            ref32 = orc.ant_forward(xf, alpha, grid, per_row=True)
            ref = torch.from_numpy(ref32).to(tdt)
            y, a = antq.fakequant_dynamic(xd, cb, g, ratio=ratio, return_alpha=True)
            a = a.cpu().numpy()
            assert np.array_equal(np.isnan(a), np.isnan(alpha)) and np.array_equal(a[~np.isnan(a)], alpha[~np.isnan(alpha)]), (g, ratio)
            yc = y.cpu().reshape(-1, g)
            same = (yc.view(torch.int16 if tdt != torch.float32 else torch.int32) == ref.view(torch.int16 if tdt != torch.float32 else torch.int32)) | (yc.isnan() & ref.isnan())
            assert bool(same.all()), (g, ratio, int((~same).sum()))
