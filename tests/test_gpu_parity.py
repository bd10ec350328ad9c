"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the
committed golden vectors.  Bar: BIT-EXACT values (fp32 and fp16) and code indices."""
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


def _cb(antq, grid_np, outl_np=None):
    g = torch.from_numpy(np.ascontiguousarray(grid_np, dtype=np.float32)).to(dev())
    o = None if outl_np is None else torch.from_numpy(np.ascontiguousarray(outl_np, dtype=np.float32)).to(dev())
    return antq.prepare_codebook(g, o)


ANT_GRIDS = [(k, b, s) for k in ("int", "flint", "pot", "float2", "float3", "apot") for b in (4,) for s in (True, False)]
ANT_GRIDS += [("int", 3, True), ("int", 5, True), ("int", 6, False), ("int", 8, True), ("int", 8, False),
              ("flint", 3, True), ("flint", 5, True), ("flint", 6, True), ("flint", 6, False), ("pot", 5, True),
              ("float3", 6, True), ("float1", 4, True)]


def test_codebook_flags(antq):
    from antq import _lib
    for kind, bit, signed in ANT_GRIDS:
        grid = orc.ant_grid(kind, bit, signed)
        cb = _cb(antq, grid)
        info = cb.describe()
        assert info["n_entries"] == grid.size
        assert info["n_levels"] == np.unique(grid).size, (kind, bit, signed)
        assert info["flags"] & _lib.CB_WELLSEP, (kind, bit, signed, info)
        assert info["flags"] & _lib.CB_STE_EXACT, (kind, bit, signed, info)
        sym = bool(info["flags"] & _lib.CB_SYMMETRIC)
        assert sym == (signed and kind != "int"), (kind, bit, signed, info)
        # signed int-k: symmetric magnitudes + one extra negative level, n_mag magnitudes on both sides
        symx = bool(info["flags"] & _lib.CB_SYMX)
        assert symx == (signed and kind == "int"), (kind, bit, signed, info)
        if symx:
            assert info["n_mag"] == 2 ** (bit - 1) and info["mid"] == 2 ** (bit - 1), info
        assert info["gmax"] == grid.max()
    for kind in ("int", "flint"):
        for signed in (True, False):
            g, o = orc.olive_grid(kind, 4, signed), orc.olive_outlier_grid(4, signed)
            info = _cb(antq, g, o).describe()
            assert info["n_entries"] == g.size + o.size and info["n_normal"] == g.size
            assert info["flags"] & _lib.CB_OVP_OK and info["flags"] & _lib.CB_WELLSEP
            assert info["gmax"] == g.max() and info["vmax"] == o.max()
            assert info["ovp_index"] >= 0


def _probe(grid, n=20000, seed=0):
    rng = np.random.default_rng(seed)
    g = np.unique(grid[np.isfinite(grid)])
    mids = ((g[:-1].astype(np.float64) + g[1:]) / 2).astype(np.float32)
    around = np.concatenate([np.nextafter(mids, np.float32(np.inf)), np.nextafter(mids, np.float32(-np.inf)), mids, g])
    span = max(abs(g[0]), abs(g[-1]), 1.0)
    rnd = (rng.standard_normal(n) * span * 0.7).astype(np.float32)
    special = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 1e-42, -1e-42, 102399.0, 102401.0 + span, -102500.0 - span,
                        65535.0, 65537.0, -65537.0, 3e38, -3e38, 2 * span, -2 * span, 2.0001 * span], dtype=np.float32)
    return np.concatenate([rnd, around, special]).astype(np.float32)


@pytest.mark.parametrize("kind,bit,signed", ANT_GRIDS)
def test_lut_nearest_matches_scan(antq, kind, bit, signed):
    grid = orc.ant_grid(kind, bit, signed)
    cb = _cb(antq, grid)
    x = _probe(grid)
    z_ref, c_ref = orc.scan(x, grid, want_codes=True)
    z, c = antq.lut_nearest(torch.from_numpy(x).to(dev()), cb, want_codes=True)
    assert_bit_equal(to_np(z), z_ref, "z", allow_zero_sign=True)
    assert_bit_equal(to_np(c).astype(np.int32), c_ref, "codes")
    # fp16 input: the kernel sees the upcast value
    xh = x.astype(np.float16)
    zh = antq.lut_nearest(torch.from_numpy(xh).to(dev()), cb)
    assert_bit_equal(to_np(zh), orc.scan(xh.astype(np.float32), grid).astype(np.float16), "z16", allow_zero_sign=True)


def test_lut_nearest_unsorted_and_odd_grids(antq):
    rng = np.random.default_rng(3)
    grids = [np.concatenate([orc.olive_flint_grid(4, True), orc.olive_outlier_grid(4, True)]),
             np.concatenate([orc.olive_int_grid(8, True), orc.olive_outlier_grid(8, True)]),       # 509 entries
             rng.permutation(orc.ant_grid("flint", 5, True)),
             np.array([5.0], dtype=np.float32),
             np.array([1.0, np.nan, -1.0, 1.0, 0.5], dtype=np.float32),
             np.array([100000.0, 100000.01, -7.0], dtype=np.float32),                               # not well separated
             (rng.standard_normal(300) * 50).astype(np.float32)]
    for grid in grids:
        cb = _cb(antq, grid)
        x = _probe(grid, n=5000)
        z_ref, c_ref = orc.scan(x, grid, want_codes=True)
        z, c = antq.lut_nearest(torch.from_numpy(x).to(dev()), cb, want_codes=True)
        assert_bit_equal(to_np(z), z_ref, "z(K=%d)" % grid.size, allow_zero_sign=True)
        assert_bit_equal(to_np(c).astype(np.int32), c_ref, "codes(K=%d)" % grid.size)


def _run_ant(antq, x_np, alpha_np, grid_np, per_row, flags=0, want_codes=False):
    cb = _cb(antq, grid_np)
    x = torch.from_numpy(x_np).to(dev())
    a = torch.from_numpy(np.ascontiguousarray(alpha_np, dtype=np.float32)).to(dev())
    r = antq.fakequant(x, a, cb, per_row, flags=flags)
    if want_codes:
        # int16 code indices come from the generic kernel; its values must equal the requested kernel's
        from antq import _lib
        yf, c = antq.fakequant(x, a, cb, per_row, want_codes=True, flags=_lib.FLAG_FORCE_FLAT)
        assert_bit_equal(to_np(yf), to_np(r), "generic (codes) path vs requested path")
        return to_np(r), to_np(c).astype(np.int32)
    return to_np(r)


@pytest.mark.parametrize("path", ["auto", "flat", "rows"])
def test_golden_forward_ant(antq, golden, path):
    from antq import _lib
    flags = {"auto": 0, "flat": _lib.FLAG_FORCE_FLAT, "rows": _lib.FLAG_FORCE_ROWS}[path]
    f = golden["forward_ant"]
    n_rows_kernel = 0
    for m in golden.manifest["forward_ant"]:
        t = m["tag"]
        for lay, per_row in (("row", True), ("ten", False)):
            x, alpha, grid = f["%s_%s_x" % (lay, t)], f["%s_%s_alpha" % (lay, t)], f["%s_%s_grid" % (lay, t)]
            try:
                y = _run_ant(antq, x, alpha, grid, per_row, flags)
            except RuntimeError as e:
                if path == "rows" and "unsupported" in str(e):
                    continue        # > 31 thresholds / ragged rows: the row kernel legitimately declines
                raise
            n_rows_kernel += path == "rows"
            assert_bit_equal(y, f["%s_%s_y" % (lay, t)], "%s %s %s" % (path, lay, t))
    for kind in ("int", "flint", "pot"):
        y = _run_ant(antq, f["kat_%s_x" % kind], np.float32(1.0), f["kat_%s_grid" % kind], False, flags)
        assert_bit_equal(y, f["kat_%s_y" % kind], "kat " + kind)
    if path == "rows":
        assert n_rows_kernel >= 20


def _run_olive(antq, x_np, alpha_np, grid_np, outl_np, per_row, no_outlier, flags=0, want_codes=False):
    cb = _cb(antq, grid_np, None if no_outlier else outl_np)
    x = torch.from_numpy(x_np).to(dev())
    a = torch.from_numpy(np.ascontiguousarray(alpha_np, dtype=np.float32)).to(dev())
    r = antq.fakequant(x, a, cb, per_row, ovp=not no_outlier, flags=flags)
    if want_codes:
        from antq import _lib
        yf, c = antq.fakequant(x, a, cb, per_row, ovp=not no_outlier, want_codes=True, flags=_lib.FLAG_FORCE_FLAT)
        assert_bit_equal(to_np(yf), to_np(r), "generic (codes) path vs requested path")
        return to_np(r), to_np(c).astype(np.int32)
    return to_np(r)


@pytest.mark.parametrize("path", ["auto", "flat"])
def test_golden_forward_olive(antq, golden, path):
    from antq import _lib
    flags = {"auto": 0, "flat": _lib.FLAG_FORCE_FLAT}[path]
    f = golden["forward_olive"]
    for m in golden.manifest["forward_olive"]:
        t = m["tag"]
        for lay, per_row in (("row", True), ("ten", False)):
            k = "%s_%s_" % (lay, t)
            y = _run_olive(antq, f[k + "x"], f[k + "alpha"], f[k + "grid"], f[k + "outliers"], per_row,
                           m["no_outlier"], flags)
            assert_bit_equal(y, f[k + "y"], "%s %s %s" % (path, lay, t))
    for name in ("even", "odd"):
        y = _run_olive(antq, f["kat_%s_x" % name], np.float32(32.0), f["kat_grid"], f["kat_outliers"], False, False,
                       flags)
        assert_bit_equal(y, f["kat_%s_y" % name], "kat " + name)


SCALES = [0.1, 1.0, 0.5, 2.0 ** -7, 3.0, 1e-3, 7.7e-3, 250.0, 6e-6, 1.0 / 3.0, 0.0123, 0.37, 2.5e-2, 1.7e-4, 9.0, 41.0]


@pytest.mark.parametrize("kind,bit,signed", [("flint", 4, True), ("int", 4, True), ("pot", 4, True), ("flint", 4, False),
                                             ("int", 4, False), ("float2", 4, True), ("int", 3, True),
                                             ("flint", 5, True), ("flint", 6, True), ("apot", 4, True)])
def test_rows_kernel_exhaustive_fp16(antq, kind, bit, signed):
    """Every fp16 bit pattern x 16 scales through the row-table kernel, one row per scale."""
    from antq import _lib
    grid = orc.ant_grid(kind, bit, signed)
    allh = np.arange(65536, dtype=np.uint16).view(np.float16)
    x = np.tile(allh, (len(SCALES), 1))
    alpha = (np.array(SCALES, dtype=np.float32) * grid.max()).astype(np.float32)
    ref, cref = orc.ant_forward(x, alpha, grid, per_row=True, want_codes=True)
    y, c = _run_ant(antq, x, alpha, grid, True, _lib.FLAG_FORCE_ROWS, want_codes=True)
    assert_bit_equal(y, ref, "values")
    assert_bit_equal(c, cref, "codes")


@pytest.mark.parametrize("dtype", ["f32", "f16"])
@pytest.mark.parametrize("kind,bit,signed", [("flint", 4, True), ("int", 4, True), ("pot", 4, False), ("float3", 5, True)])
def test_rows_kernel_random(antq, kind, bit, signed, dtype):
    from antq import _lib
    rng = np.random.default_rng(11)
    grid = orc.ant_grid(kind, bit, signed)
    rows, cols = 96, 4096
    x = (rng.standard_normal((rows, cols)) * 0.02).astype(np.float32)
    x[rng.integers(0, rows, 50), rng.integers(0, cols, 50)] *= 30          # far outside the clip window
    x[5, 7], x[9, 100], x[11, 4095] = np.nan, np.inf, -np.inf
    x[20] = 0.0                                                           # alpha = 0 row -> NaN in the reference
    if not signed:
        x = np.abs(x)
    alpha = (np.abs(x).max(1) * rng.uniform(0.75, 1.2, rows)).astype(np.float32)
    alpha[3] = np.float32(0.625 * 0.02)        # scale that makes exact ties representable
    if dtype == "f16":
        x = x.astype(np.float16)
    ref, cref = orc.ant_forward(x, alpha, grid, per_row=True, want_codes=True)
    for flags in (_lib.FLAG_FORCE_ROWS, _lib.FLAG_FORCE_FLAT):
        y, c = _run_ant(antq, x, alpha, grid, True, flags, want_codes=True)
        assert_bit_equal(y, ref, "values flags=%d" % flags)
        assert_bit_equal(c, cref, "codes flags=%d" % flags)
    # per-tensor view of the same data, ragged length (tail handled by the slow path)
    xt = x.reshape(-1)[: rows * cols - 3]
    a0 = np.float32(np.abs(xt[np.isfinite(xt)].astype(np.float32)).max() * 0.9)
    reft = orc.ant_forward(xt, a0, grid, per_row=False)
    yt = _run_ant(antq, xt, a0, grid, False, _lib.FLAG_FORCE_ROWS)
    assert_bit_equal(yt, reft, "per-tensor ragged")


@pytest.mark.parametrize("dtype", ["f32", "f16"])
@pytest.mark.parametrize("kind,signed", [("int", True), ("flint", True), ("flint", False), ("int", False)])
def test_olive_rows_and_flat(antq, kind, signed, dtype):
    from antq import _lib
    rng = np.random.default_rng(5)
    grid, outl = orc.olive_grid(kind, 4, signed), orc.olive_outlier_grid(4, signed)
    rows, cols = 64, 2048
    x = rng.standard_normal((rows, cols)).astype(np.float32)
    idx = rng.integers(0, x.size, x.size // 200)
    x.reshape(-1)[idx] *= rng.choice([8.0, 20.0, 60.0, 400.0], idx.size)
    x.reshape(-1)[idx[:50] ^ 1] *= 30.0                       # outliers next to outliers
    x[2, 10], x[2, 11] = np.nan, 50.0
    if not signed:
        x = np.abs(x)
    xf = x.reshape(rows, -1)
    alpha = (3 * xf.std(1) + np.abs(xf.mean(1))).astype(np.float32)
    alpha[np.isnan(alpha)] = 1.0
    if dtype == "f16":
        x = x.astype(np.float16)
    ref, cref = orc.olive_forward(x, alpha, grid, outl, per_row=True, want_codes=True)
    for flags in (_lib.FLAG_FORCE_ROWS, _lib.FLAG_FORCE_FLAT):
        y, c = _run_olive(antq, x, alpha, grid, outl, True, False, flags, want_codes=True)
        assert_bit_equal(y, ref, "olive values flags=%d" % flags)
        assert_bit_equal(c, cref, "olive codes flags=%d" % flags)
    # per-tensor, odd numel -> torch.roll wrap-around pair; element 0 is an outlier
    xt = x.reshape(-1)[: rows * cols - 1].copy()
    xt[0] = 300.0
    a0 = np.float32(3.0)
    reft = orc.olive_forward(xt, a0, grid, outl, per_row=False)
    yt = _run_olive(antq, xt, a0, grid, outl, False, False)
    assert_bit_equal(yt, reft, "olive per-tensor odd numel")


def test_headline_shape_fp16_flint4(antq):
    """4096x4096 fp16, 4-bit flint, per-channel: the bench workload, bit-exact vs the oracle."""
    from antq import _lib
    g = torch.Generator(device="cpu").manual_seed(0)
    x = (torch.randn(4096, 4096, generator=g) * 0.02).to(torch.float16)
    grid = orc.ant_grid("flint", 4, True)
    alpha = (x.float().abs().amax(1) * 0.9).numpy()
    cb = _cb(antq, grid)
    xd = x.to(dev())
    assert antq.fakequant_plan(xd, cb, True) == 1
    y = antq.fakequant(xd, torch.from_numpy(alpha).to(dev()), cb, True)
    ref = orc.ant_forward(x.numpy(), alpha, grid, per_row=True)
    assert_bit_equal(to_np(y), ref, "headline")
    # in place
    y2 = antq.fakequant(xd, torch.from_numpy(alpha).to(dev()), cb, True, out=xd)
    assert y2.data_ptr() == xd.data_ptr()
    assert_bit_equal(to_np(xd), ref, "headline in place")


def test_bf16(antq):
    from antq import _lib
    rng = np.random.default_rng(2)
    grid = orc.ant_grid("flint", 4, True)
    x32 = (rng.standard_normal((32, 2048)) * 0.05).astype(np.float32)
    xb = torch.from_numpy(x32).to(torch.bfloat16)
    alpha = (np.abs(x32).max(1) * 0.85).astype(np.float32)
    ref32 = orc.ant_forward(xb.float().numpy(), alpha, grid, per_row=True)
    ref = torch.from_numpy(ref32).to(torch.bfloat16)        # RNE
    cb = _cb(antq, grid)
    for flags in (_lib.FLAG_FORCE_ROWS, _lib.FLAG_FORCE_FLAT):
        y = antq.fakequant(xb.to(dev()), torch.from_numpy(alpha).to(dev()), cb, True, flags=flags)
        assert torch.equal(y.cpu().view(torch.int16), ref.view(torch.int16)), flags


def test_absmax_and_sweep(antq):
    rng = np.random.default_rng(4)
    x = (rng.standard_normal((48, 1536)) * 0.1).astype(np.float32)
    xd = torch.from_numpy(x).to(dev())
    np.testing.assert_array_equal(to_np(antq.absmax(xd, True)), np.abs(x).max(1))
    np.testing.assert_array_equal(to_np(antq.absmax(xd, False)), np.abs(x).max(keepdims=True).reshape(1))
    xn = x.copy(); xn[3, 3] = np.nan
    assert np.isnan(to_np(antq.absmax(torch.from_numpy(xn).to(dev()), True))[3])
    xh = x.astype(np.float16)
    np.testing.assert_array_equal(to_np(antq.absmax(torch.from_numpy(xh).to(dev()), True)),
                                  np.abs(xh.astype(np.float32)).max(1))
    grid = orc.ant_grid("flint", 4, True)
    cb = _cb(antq, grid)
    ratios = np.array([i * 0.01 for i in range(75, 150)], dtype=np.float32)
    for per_row in (True, False):
        base = np.abs(x).max(1) if per_row else np.abs(x).max(keepdims=True).reshape(1)
        err = to_np(antq.mse_sweep(xd, torch.from_numpy(base.astype(np.float32)).to(dev()),
                                   torch.from_numpy(ratios).to(dev()), cb, per_row))
        ref = np.empty_like(err)
        for ci, r in enumerate(ratios):
            q = orc.ant_forward(x, (base * r).astype(np.float32), grid, per_row)
            e = (q.astype(np.float64) - x) ** 2
            ref[ci] = e.reshape(len(base), -1).sum(1)
        np.testing.assert_allclose(err, ref, rtol=2e-5)
        assert (err.argmin(0) == ref.argmin(0)).mean() > 0.9


def test_host_pipeline(antq):
    rng = np.random.default_rng(6)
    grid = orc.ant_grid("flint", 4, True)
    x = (rng.standard_normal((1024, 4096)) * 0.02).astype(np.float16)
    alpha = (np.abs(x.astype(np.float32)).max(1) * 0.9).astype(np.float32)
    xt = torch.from_numpy(x).pin_memory()
    out = torch.empty_like(xt).pin_memory()
    hp = antq.HostPipeline(device=0, chunk_bytes=1 << 20, n_stages=3)
    hp.fakequant(xt, out, torch.from_numpy(alpha), torch.from_numpy(grid), per_row=True)
    assert hp.last_launches >= 8
    assert_bit_equal(out.numpy(), orc.ant_forward(x, alpha, grid, per_row=True), "host per-row")
    a0 = np.float32(0.05)
    hp.fakequant(xt, out, torch.tensor([a0]), torch.from_numpy(grid), per_row=False)
    assert_bit_equal(out.numpy(), orc.ant_forward(x, a0, grid, per_row=False), "host per-tensor")
    hp.close()


def test_errors_are_loud(antq):
    grid = orc.ant_grid("int", 4, True)
    cb = _cb(antq, grid)
    with pytest.raises(RuntimeError):
        antq.fakequant(torch.zeros(4, 4), torch.ones(4), cb, True)          # CPU tensor: no fallback
    with pytest.raises(TypeError):
        antq.fakequant(torch.zeros(4, 4, dtype=torch.float64, device=dev()), torch.ones(4, device=dev()), cb, True)
    with pytest.raises(ValueError):
        antq.fakequant(torch.zeros(4, 4, device=dev()), torch.ones(3, device=dev()), cb, True)
    z = antq.fakequant(torch.zeros(0, 4, device=dev()), torch.ones(0, device=dev()), cb, True)
    assert z.numel() == 0


@pytest.mark.parametrize("dtype", ["f16", "f32", "bf16"])
@pytest.mark.parametrize("bit", [3, 4, 5, 6])
def test_stream_kernel_signed_int(antq, bit, dtype):
    """Signed int-k runs as 'symmetric magnitudes + one extra negative level' in the stream kernel (6-bit only fits
    the row-table kernels that way); values must equal the oracle and the generic flat kernel bit for bit."""
    from antq import _lib
    rng = np.random.default_rng(bit)
    grid = orc.ant_grid("int", bit, True)
    rows, cols = 200, 2048
    x = (rng.standard_normal((rows, cols)) * 0.05).astype(np.float32)
    x[rng.integers(0, rows, 30), rng.integers(0, cols, 30)] *= 40
    alpha = (np.abs(x).max(1) * rng.uniform(0.3, 1.2, rows)).astype(np.float32)     # plenty of clipping at both ends
    alpha[::9] = np.float32(0.05 * grid.max() / 8)                                   # representable ties
    cb = _cb(antq, grid)
    if dtype == "bf16":
        xt = torch.from_numpy(x).to(torch.bfloat16)
        ref = torch.from_numpy(orc.ant_forward(xt.float().numpy(), alpha, grid, per_row=True)).to(torch.bfloat16)
        y = antq.fakequant(xt.to(dev()), torch.from_numpy(alpha).to(dev()), cb, True, flags=_lib.FLAG_FORCE_ROWS)
        assert torch.equal(y.cpu().view(torch.int16), ref.view(torch.int16))
        return
    if dtype == "f16":
        x = x.astype(np.float16)
    ref = orc.ant_forward(x, alpha, grid, per_row=True)
    # FORCE_ROWS still runs the chain
    # int-k, per-row scales: the closed form (fp16 int-3 keeps its 3-threshold chain)
    assert antq.fakequant_plan(torch.from_numpy(x).to(dev()), cb, True) == (1 if bit == 3 and dtype == "f16" else 4)
    assert antq.fakequant_plan(torch.from_numpy(x).to(dev()), cb, True, flags=_lib.FLAG_FORCE_ROWS) == 1
    y = _run_ant(antq, x, alpha, grid, True, _lib.FLAG_FORCE_ROWS)
    assert_bit_equal(y, ref, "signed int-%d %s" % (bit, dtype))
    assert_bit_equal(_run_ant(antq, x, alpha, grid, True, 0), ref, "signed int-%d %s default plan" % (bit, dtype))
    yt = _run_ant(antq, x.reshape(-1), np.float32(alpha.mean()), grid, False, _lib.FLAG_FORCE_ROWS)
    assert_bit_equal(yt, orc.ant_forward(x.reshape(-1), np.float32(alpha.mean()), grid, per_row=False), "per-tensor")


@pytest.mark.parametrize("rows,cols,dtype", [(148 * 45 + 7, 512, "f16"), (9000, 1032, "f16"), (5000, 520, "f32"),
                                             (300, 10248, "f16"), (37, 512, "f16"), (3, 100000, "f16")])
def test_stream_kernel_shapes(antq, rows, cols, dtype):
    """Shapes that exercise the persistent kernel's bookkeeping: more than 32 rows per CTA (row-table ring reuse by
    the builder warps), rows that are not a whole number of chunks, fewer chunks than SMs, very long rows."""
    from antq import _lib
    rng = np.random.default_rng(rows * 7 + cols)
    grid = orc.ant_grid("flint", 4, True)
    x = (rng.standard_normal((rows, cols)) * 0.05).astype(np.float32)
    x[rng.integers(0, rows, 40), rng.integers(0, cols, 40)] *= 50           # out-of-window values -> cold fix-up pass
    x[rows // 2, cols // 3] = np.nan
    alpha = (np.abs(x).max(1) * rng.uniform(0.7, 1.3, rows)).astype(np.float32)
    alpha[np.isnan(alpha)] = 0.1
    alpha[::17] = np.float32(0.625 * 0.05)                                   # rows with representable ties
    if dtype == "f16":
        x = x.astype(np.float16)
    ref = orc.ant_forward(x, alpha, grid, per_row=True)
    y = _run_ant(antq, x, alpha, grid, True, _lib.FLAG_FORCE_ROWS)
    assert_bit_equal(y, ref, "stream kernel %dx%d %s" % (rows, cols, dtype))
    # in place
    cb = _cb(antq, grid)
    xd = torch.from_numpy(x).to(dev())
    antq.fakequant(xd, torch.from_numpy(alpha).to(dev()), cb, True, out=xd, flags=_lib.FLAG_FORCE_ROWS)
    assert_bit_equal(to_np(xd), ref, "in place")


def test_stream_kernel_olive_many_rows(antq):
    from antq import _lib
    rng = np.random.default_rng(77)
    grid, outl = orc.olive_grid("flint", 4, True), orc.olive_outlier_grid(4, True)
    rows, cols = 148 * 40, 512
    x = rng.standard_normal((rows, cols)).astype(np.float32)
    idx = rng.integers(0, x.size, x.size // 150)
    x.reshape(-1)[idx] *= rng.choice([8.0, 20.0, 60.0, 400.0], idx.size)
    alpha = (3 * x.std(1) + np.abs(x.mean(1))).astype(np.float32)
    x = x.astype(np.float16)
    ref = orc.olive_forward(x, alpha, grid, outl, per_row=True)
    y = _run_olive(antq, x, alpha, grid, outl, True, False, _lib.FLAG_FORCE_ROWS)
    assert_bit_equal(y, ref, "olive many rows")


def test_stream_kernel_dependent_launches(antq):
    """Back-to-back launches whose input is the previous launch's output (programmatic dependent launch must not let
    launch N+1 read before launch N has finished), out of place and in place, eager and inside a CUDA graph."""
    rng = np.random.default_rng(3)
    grids = [orc.ant_grid("flint", 4, True), orc.ant_grid("int", 4, True), orc.ant_grid("pot", 4, True)]
    rows, cols = 2048, 4096
    x = (rng.standard_normal((rows, cols)) * 0.05).astype(np.float16)
    alphas = [(np.abs(x.astype(np.float32)).max(1) * r).astype(np.float32) for r in (1.0, 0.8, 0.6)]
    ref = x
    for g, a in zip(grids, alphas):
        ref = orc.ant_forward(ref, a, g, per_row=True)
    cbs = [_cb(antq, g) for g in grids]
    ad = [torch.from_numpy(a).to(dev()) for a in alphas]
    xd = torch.from_numpy(x).to(dev())

    def chain(buf_in, bufs):
        cur = buf_in
        for cb, a, out in zip(cbs, ad, bufs):
            antq.fakequant(cur, a, cb, True, out=out)
            cur = out
        return cur
    for trial in range(5):
        tmp = [torch.empty_like(xd) for _ in range(3)]
        y = chain(xd, tmp)
        assert_bit_equal(to_np(y), ref, "out-of-place chain, trial %d" % trial)
        z = xd.clone()
        y = chain(z, [z, z, z])
        assert_bit_equal(to_np(y), ref, "in-place chain, trial %d" % trial)
    tmp = [torch.empty_like(xd) for _ in range(3)]
    chain(xd, tmp)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        chain(xd, tmp)
    for t in tmp:
        t.zero_()
    gr.replay()
    gr.replay()
    torch.cuda.synchronize()
    assert_bit_equal(to_np(tmp[2]), ref, "graph replay")


@pytest.mark.parametrize("dtype", ["f16", "f32"])
@pytest.mark.parametrize("kind,bit,signed,cols", [("flint", 4, True, 32), ("int", 4, True, 16), ("pot", 4, True, 8),
                                                  ("flint", 4, False, 64), ("int", 4, False, 256), ("float2", 4, True, 128),
                                                  ("int", 5, True, 32), ("flint", 3, True, 504), ("int", 3, False, 40)])
def test_short_rows_and_scale_groups(antq, kind, bit, signed, cols, dtype):
    """Rows shorter than 512 elements (group-8/16/32 scales, 1x1-conv weights) take the closed-form short kernel
    (plan 5) when the grid is piecewise uniform and the d-space chain kernel (plan 3, here forced with NO_PU) otherwise:
    both bit-exact against the oracle and the generic flat kernel, including clipped / NaN / Inf inputs and dead rows."""
    from antq import _lib
    rng = np.random.default_rng(cols * 31 + bit)
    grid = orc.ant_grid(kind, bit, signed)
    rows = 70000 // cols * 8
    x = (rng.standard_normal((rows, cols)) * 0.05).astype(np.float32)
    x[rng.integers(0, rows, 60), rng.integers(0, cols, 60)] *= 500          # far outside the window
    x[7, 3], x[11, 5], x[13, cols - 1] = np.nan, np.inf, -np.inf
    x[20] = 0.0                                                            # alpha = 0 row
    if not signed:
        x = np.abs(x)
    alpha = (np.abs(np.nan_to_num(x, nan=0.0, posinf=0.0, neginf=0.0)).max(1) * rng.uniform(0.4, 1.2, rows)).astype(np.float32)
    alpha[::5] = np.float32(0.05 * grid.max() / 8)                           # representable ties
    if dtype == "f16":
        x = x.astype(np.float16)
    cb = _cb(antq, grid)
    xd = torch.from_numpy(x).to(dev())
    assert antq.fakequant_plan(xd, cb, True) == 5
    assert antq.fakequant_plan(xd, cb, True, flags=_lib.FLAG_NO_PU) == 3
    ref = orc.ant_forward(x, alpha, grid, per_row=True)
    y = _run_ant(antq, x, alpha, grid, True, 0)
    assert_bit_equal(y, ref, "closed-form short kernel %s-%d cols=%d %s" % (kind, bit, cols, dtype))
    y = _run_ant(antq, x, alpha, grid, True, _lib.FLAG_NO_PU)
    assert_bit_equal(y, ref, "d-space short kernel %s-%d cols=%d %s" % (kind, bit, cols, dtype))
    yf = _run_ant(antq, x, alpha, grid, True, _lib.FLAG_FORCE_FLAT)
    assert_bit_equal(yf, ref, "flat kernel")
    antq.fakequant(xd, torch.from_numpy(alpha).to(dev()), cb, True, out=xd)
    assert_bit_equal(to_np(xd), ref, "in place")


@pytest.mark.parametrize("kind,signed", [("flint", True), ("int", True)])
def test_short_rows_olive(antq, kind, signed):
    rng = np.random.default_rng(9)
    grid, outl = orc.olive_grid(kind, 4, signed), orc.olive_outlier_grid(4, signed)
    rows, cols = 4096, 64
    x = rng.standard_normal((rows, cols)).astype(np.float32)
    idx = rng.integers(0, x.size, x.size // 100)
    x.reshape(-1)[idx] *= rng.choice([8.0, 20.0, 60.0, 400.0], idx.size)
    x.reshape(-1)[idx[:80] ^ 1] *= 30.0
    alpha = (3 * x.std(1) + np.abs(x.mean(1))).astype(np.float32)
    x = x.astype(np.float16)
    cb = _cb(antq, grid, outl)
    assert antq.fakequant_plan(torch.from_numpy(x).to(dev()), cb, True, ovp=True) == 3
    ref = orc.olive_forward(x, alpha, grid, outl, per_row=True)
    y = _run_olive(antq, x, alpha, grid, outl, True, False, 0)
    assert_bit_equal(y, ref, "olive short rows")


@pytest.mark.parametrize("kind,olive", [("flint", False), ("int", False), ("flint", True)])
def test_full_size_properties(antq, kind, olive):
    """BASELINE.json's largest sweep size (16384 x 16384 fp16 = 512 MB, 111 rows per CTA: row-table ring reuse) through
    size-independent properties: the stream kernel agrees with the independent generic flat kernel everywhere, with the
    oracle on a sample of rows, every output is a codebook level times the row's scale, and the op is idempotent."""
    from antq import _lib
    N = 16384
    g = torch.Generator(device="cuda").manual_seed(5)
    x = (torch.randn(N, N, device=dev(), generator=g) * 0.03).to(torch.float16)
    x[::997, ::13] *= 40                                                  # clipped / outlier values
    if olive:
        grid, outl = orc.olive_grid(kind, 4, True), orc.olive_outlier_grid(4, True)
        cb = _cb(antq, grid, outl)
        alpha = (3 * x.float().std(1)).contiguous()
    else:
        grid, outl = orc.ant_grid(kind, 4, True), None
        cb = _cb(antq, grid)
        alpha = (x.float().abs().amax(1) * 0.85).contiguous()
    assert antq.fakequant_plan(x, cb, True, ovp=olive) == (4 if kind == "int" and not olive else 1)
    y = antq.fakequant(x, alpha, cb, True, ovp=olive)
    yf = antq.fakequant(x, alpha, cb, True, ovp=olive, flags=_lib.FLAG_FORCE_FLAT)
    assert torch.equal(y.view(torch.int16), yf.view(torch.int16)), "stream kernel != flat kernel"
    del yf
    rows = torch.arange(0, N, 331, device=dev())                             # 50 rows against the oracle
    xs, als = x[rows].cpu().numpy(), alpha[rows].cpu().numpy()
    ref = orc.olive_forward(xs, als, grid, outl, per_row=True) if olive else orc.ant_forward(xs, als, grid, per_row=True)
    assert_bit_equal(to_np(y[rows]), ref, "oracle on sampled rows")
    if not olive:
        # idempotence: quantized values are fixed points of the same quantizer
        y2 = antq.fakequant(y, alpha, cb, True)
        assert torch.equal(y2.view(torch.int16), y.view(torch.int16)), "not idempotent"
        # at most 2^bit distinct values per row
        assert int(torch.unique(y[12345]).numel()) <= 16


@pytest.mark.parametrize("group", [8, 32, 128, 1024])
def test_group_scales(antq, group):
    """Group-wise scales through antq.fakequant_grouped (BASELINE.json's group-8/16/32 sweep): one alpha per `group`
    consecutive elements = the per-row path on the [numel / group, group] view."""
    rng = np.random.default_rng(group)
    grid = orc.ant_grid("flint", 4, True)
    x = (rng.standard_normal((512, 4096)) * 0.05).astype(np.float16)
    xg = x.reshape(-1, group)
    alpha = (np.abs(xg.astype(np.float32)).max(1) * 0.9).astype(np.float32)
    cb = _cb(antq, grid)
    xd = torch.from_numpy(x).to(dev())
    y = antq.fakequant_grouped(xd, torch.from_numpy(alpha).to(dev()), cb, group)
    assert y.shape == xd.shape
    ref = orc.ant_forward(xg, alpha, grid, per_row=True).reshape(x.shape)
    assert_bit_equal(to_np(y), ref, "group %d" % group)
    with pytest.raises(ValueError):
        antq.fakequant_grouped(xd[:, :4095].contiguous(), torch.from_numpy(alpha).to(dev()), cb, group)
