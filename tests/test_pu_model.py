"""Exhaustive check of the piecewise-uniform closed-form codec (tests/pu_model.py == the arithmetic of
csrc/antq_pu.cu) against the oracle: every fp16 bit pattern x many scales x every PU-eligible grid family, bf16 and
randomised fp32 hugging every threshold.  CPU only."""
import numpy as np
import pytest

import antq_oracle as orc
import pu_model as pm
import xspace_model as xm

f32 = np.float32
ALL_F16 = np.arange(65536, dtype=np.uint16).view(np.float16)


def scales(vmax, n=10):
    rng = np.random.default_rng(3)
    s = [0.1, 1.0, 2.0 ** -7, 3.0, 0.0625 / vmax, 1e-3, 250.0, 1.0 / 3.0]
    s += list(np.exp(rng.uniform(np.log(1e-4), np.log(50.0), n)))
    return [f32(v) for v in s]


def lim_of(cb):
    vmax, vmin = cb["vmax"], cb["vmin"]
    lim_pos = 2 * vmax
    lim_neg = -2 * vmin if vmin < 0 else lim_pos
    return f32(min(lim_pos, lim_neg, 65536.0))


PU_GRIDS = [("int", b, sg) for b in (3, 4, 5, 6, 7, 8) for sg in (True, False)]
PU_GRIDS += [("flint", b, sg) for b in (3, 4, 5, 6) for sg in (True, False)]
PU_GRIDS += [("pot", 3, True), ("pot", 4, True), ("pot", 4, False), ("float2", 4, True), ("float2", 4, False),
             ("float3", 6, True), ("float1", 4, False), ("float3", 5, False)]


def test_analysis_accepts_and_rejects():
    for kind, bit, signed in PU_GRIDS:
        assert pm.analyze(orc.ant_grid(kind, bit, signed)) is not None, (kind, bit, signed)
    assert pm.analyze(orc.ant_grid("int", 8, True))["uniform"]
    assert pm.analyze(orc.ant_grid("int", 8, False))["xc16"] and not pm.analyze(orc.ant_grid("int", 8, False))["xcbf"]
    assert pm.analyze(orc.ant_grid("flint", 4, True))["xcbf"]
    assert not pm.analyze(orc.ant_grid("flint", 4, True))["uniform"]
    assert pm.analyze(orc.ant_grid("apot", 4, False)) is None          # 8, 9, 12 in one octave
    # OliVe int + abfloat: 32 itself is missing from the outliers' first octave
    assert pm.analyze(np.concatenate([orc.olive_int_grid(4, True), orc.olive_outlier_grid(4, True)])) is None
    # OliVe normal grids alone are PU (no_outlier runs)
    assert pm.analyze(orc.olive_flint_grid(4, True)) is not None


@pytest.mark.parametrize("kind,bit,signed", PU_GRIDS)
def test_fp16_exhaustive(kind, bit, signed):
    grid = orc.ant_grid(kind, bit, signed)
    cb = xm.prepare_codebook(grid)
    assert xm.interior_exact(cb)
    pu = pm.analyze(grid)
    gmax = grid.max()
    flagged = 0
    for s in scales(gmax, 6 if bit >= 7 else 10):
        alpha = f32(s * gmax)
        s_eff = f32(alpha / gmax)

        def exact(xs):
            return orc.ant_forward(xs, alpha, grid, per_row=False)
        got, fl = pm.forward(ALL_F16, s_eff, pu, lim_of(cb), exact, np.float16)
        ref = orc.ant_forward(ALL_F16, alpha, grid, per_row=False)
        same = (got.view(np.uint16) == ref.view(np.uint16)) | (np.isnan(got) & np.isnan(ref))
        assert same.all(), (kind, bit, signed, s, ALL_F16[~same][:5], got[~same][:5], ref[~same][:5])
        if pu["xc16"]:                                         # the x-space clamp variant the fp16 kernels take
            got2, _ = pm.forward(ALL_F16, s_eff, pu, lim_of(cb), exact, np.float16, xclamp=True)
            same = (got2.view(np.uint16) == ref.view(np.uint16)) | (np.isnan(got2) & np.isnan(ref))
            assert same.all(), ("xclamp", kind, bit, signed, s, ALL_F16[~same][:5], got2[~same][:5], ref[~same][:5])
        # the pair decision that settles near-midpoint elements must equal the scan for EVERY in-window element
        got3, _ = pm.forward(ALL_F16, s_eff, pu, lim_of(cb), exact, np.float16, pair_all=True)
        same = (got3.view(np.uint16) == ref.view(np.uint16)) | (np.isnan(got3) & np.isnan(ref))
        assert same.all(), ("pair", kind, bit, signed, s, ALL_F16[~same][:5], got3[~same][:5], ref[~same][:5])
        # the lean variant (clamp and window on t, constant near margin for uniform grids)
        got4, _ = pm.forward(ALL_F16, s_eff, pu, lim_of(cb), exact, np.float16, lean=True)
        same = (got4.view(np.uint16) == ref.view(np.uint16)) | (np.isnan(got4) & np.isnan(ref))
        assert same.all(), ("lean", kind, bit, signed, s, ALL_F16[~same][:5], got4[~same][:5], ref[~same][:5])
        inwin = np.abs(ALL_F16.astype(f32)) <= f32(lim_of(cb) * s_eff) * f32(0.99)
        flagged += (fl & inwin).sum() / max(inwin.sum(), 1)
    assert flagged / len(scales(gmax, 6 if bit >= 7 else 10)) < 0.02          # the closed form is what runs


@pytest.mark.parametrize("kind,bit,signed", [("int", 8, True), ("int", 8, False), ("flint", 4, False), ("flint", 6, False),
                                             ("int", 5, True), ("pot", 4, True), ("float2", 4, False)])
def test_fp32_threshold_huggers(kind, bit, signed):
    grid = orc.ant_grid(kind, bit, signed)
    cb = xm.prepare_codebook(grid)
    pu = pm.analyze(grid)
    gmax = grid.max()
    rng = np.random.default_rng(11)
    for s in scales(gmax, 4):
        alpha = f32(s * gmax)
        s_eff = f32(alpha / gmax)
        x = (rng.standard_normal(30000) * s_eff * gmax / 2).astype(f32)
        near = (cb["thr"].astype(np.float64) * float(s_eff)).astype(f32)
        for k in range(-40, 41):
            x = np.concatenate([x, xm._unord(xm._ord(near) + k)])

        def exact(xs):
            return orc.ant_forward(xs, alpha, grid, per_row=False)
        got, fl = pm.forward(x, s_eff, pu, lim_of(cb), exact, f32)
        ref = orc.ant_forward(x, alpha, grid, per_row=False)
        same = (got.view(np.uint32) == ref.view(np.uint32)) | (np.isnan(got) & np.isnan(ref))
        assert same.all(), (kind, s, x[~same][:5], got[~same][:5], ref[~same][:5])
        got3, _ = pm.forward(x, s_eff, pu, lim_of(cb), exact, f32, pair_all=True)
        same = (got3.view(np.uint32) == ref.view(np.uint32)) | (np.isnan(got3) & np.isnan(ref))
        assert same.all(), ("pair", kind, s, x[~same][:5], got3[~same][:5], ref[~same][:5])
        # WITHOUT the near-midpoint redo the closed form would be wrong somewhere among the huggers -> the test bites
    # sanity: the model really is exercised on its own (few flagged among the random part)
    assert fl[:30000].mean() < 0.01


@pytest.mark.parametrize("kind,signed", [("flint", True), ("flint", False), ("int", True), ("int", False)])
def test_olive_normal_levels_below_the_outlier_threshold(kind, signed):
    """The OliVe closed-form path (ANTQ_CB_PU_OVP): grid + outliers is not piecewise uniform, but with the window cut at the
    first outlier threshold every in-window element must quantize exactly as the reference's scan over the WHOLE codebook
    does (element-wise: what a pair without an outlier gets), using only the normal levels' closed form."""
    grid, outl = orc.olive_grid(kind, 4, signed), orc.olive_outlier_grid(4, signed)
    whole = np.concatenate([grid, outl]).astype(f32)
    cbw = xm.prepare_codebook(whole)
    lev, thr = cbw["levels"], cbw["thr"]
    normal = np.abs(lev) <= 32
    lo, hi = int(np.argmax(normal)), int(len(lev) - 1 - np.argmax(normal[::-1]))
    assert normal[lo:hi + 1].all()
    tout = min(thr[hi] if hi < len(lev) - 1 else np.inf, -thr[lo - 1] if lo > 0 else np.inf)
    pu = pm.analyze(lev[lo:hi + 1])
    assert pu is not None
    gmax = grid.max()
    checked = 0
    for s in scales(gmax, 8):
        alpha = f32(s * gmax)
        s_eff = f32(alpha / gmax)

        def exact(xs):
            # the reference's OliVe forward, each value paired with a zero (never an outlier, so the value itself is
            # what the scan over grid + outliers gives at the scale alpha / max(normal grid))
            z = np.zeros(2 * xs.size, dtype=xs.dtype)
            z[::2] = xs
            return orc.olive_forward(z, alpha, grid, outl, per_row=False)[::2]
        lim = f32(min(float(lim_of(cbw)), float(tout)))
        for kw in (dict(), dict(lean=True)) + ((dict(xclamp=True),) if pu["xc16"] else ()):
            got, fl = pm.forward(ALL_F16, s_eff, pu, lim, exact, np.float16, **kw)
            ref = exact(ALL_F16)
            same = (got.view(np.uint16) == ref.view(np.uint16)) | (np.isnan(got) & np.isnan(ref))
            assert same.all(), (kind, signed, s, kw, ALL_F16[~same][:5], got[~same][:5], ref[~same][:5])
            checked += int((~fl).sum())
    assert checked > 10000                                               # the closed form, not the fallback, was what ran
