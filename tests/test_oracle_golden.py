"""Pin the CPU oracle (oracle/) to the golden vectors produced by the
unmodified reference Python (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

import antq_oracle as orc


def same(a, b):
    """Bit-level equality of float arrays, NaN == NaN, -0.0 == +0.0 allowed only if asked."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    assert a.shape == b.shape
    nan = np.isnan(a) & np.isnan(b)
    ok = nan | (a == b)
    if not ok.all():
        bad = np.argwhere(~ok)[:5]
        raise AssertionError("mismatch at %s: %s vs %s" % (bad.tolist(), a[~ok][:5], b[~ok][:5]))


def test_grids_all(golden):
    checked = 0
    for m in golden.manifest["grids"]:
        tree, kind, su, bit = m["key"].rsplit("_", 3)[0].split("_", 1)[0], None, None, None
        parts = m["key"].split("_")
        tree, kind, su, bit = parts[0], parts[1], parts[2], int(parts[3])
        signed = su == "s"
        def make():
            if tree == "ant":
                return orc.ant_grid(kind, bit, signed)
            if kind == "outlier":
                return orc.olive_outlier_grid(bit, signed)
            return orc.olive_grid(kind, bit, signed)
        if "error" in m:
            with pytest.raises((AssertionError, TypeError)):
                make()
            continue
        g = make()
        ref = golden["grids"][m["key"]]
        assert g.dtype == np.float32 and g.shape == ref.shape, m["key"]
        assert np.array_equal(g, ref, equal_nan=True), m["key"]
        nz = ref != 0          # the order of the -0.0/+0.0 pair in signed apot tables is a sort artefact
        assert np.array_equal(np.signbit(g[nz]), np.signbit(ref[nz])), m["key"]
        checked += 1
    assert checked > 100


def test_kat_survey(golden):
    f = golden["forward_ant"]
    for kind in ("int", "flint", "pot"):
        y = orc.ant_forward(f["kat_%s_x" % kind], np.float32(1.0), f["kat_%s_grid" % kind], per_row=False)
        same(y, f["kat_%s_y" % kind])
    # the literal numbers quoted in SURVEY.md 8(c)
    y = orc.ant_forward(f["kat_flint_x"], np.float32(1.0), orc.ant_grid("flint", 4, True), per_row=False)
    exp = [-1, -1, -0.5, -0.375, -0.25, 0, 0, 0.0625, 0.0625, 0.1875, 0.5, 0.5, 1, 1, 1, 1]
    np.testing.assert_allclose(y[:16], np.array(exp, dtype=np.float32), rtol=1e-6)
    assert np.isnan(y[16]) and np.isnan(y[17])


def test_forward_ant_pinned(golden):
    f = golden["forward_ant"]
    for m in golden.manifest["forward_ant"]:
        t = m["tag"]
        y = orc.ant_forward(f["row_%s_x" % t], f["row_%s_alpha" % t], f["row_%s_grid" % t], per_row=True)
        same(y, f["row_%s_y" % t])
        y = orc.ant_forward(f["ten_%s_x" % t], f["ten_%s_alpha" % t], f["ten_%s_grid" % t], per_row=False)
        same(y, f["ten_%s_y" % t])


def test_forward_olive_pinned(golden):
    f = golden["forward_olive"]
    for m in golden.manifest["forward_olive"]:
        t = m["tag"]
        for lay, per_row in (("row", True), ("ten", False)):
            y = orc.olive_forward(f["%s_%s_x" % (lay, t)], f["%s_%s_alpha" % (lay, t)], f["%s_%s_grid" % (lay, t)],
                                  f["%s_%s_outliers" % (lay, t)], per_row=per_row, no_outlier=m["no_outlier"])
            same(y, f["%s_%s_y" % (lay, t)])
    for name in ("even", "odd"):
        y = orc.olive_forward(f["kat_%s_x" % name], np.float32(32.0), f["kat_grid"], f["kat_outliers"], per_row=False)
        same(y, f["kat_%s_y" % name])
    np.testing.assert_array_equal(f["kat_odd_y"], np.array([96, 0, 2, 4, 0], dtype=np.float32))
    np.testing.assert_array_equal(
        f["kat_even_y"], np.array([0, 96, 96, 0, 96, 0, 4, 6, 48, 0, -48, 0, 384, 0, 32, 8], dtype=np.float32))


def test_scan_codes_are_last_minimum():
    g = orc.ant_grid("flint", 4, True)          # two zeros at indices 7, 8
    z, c = orc.scan(np.array([0.0, 1e-9, 4.375, -4.375, np.nan, 2e5], dtype=np.float32), g, want_codes=True)
    assert list(c[:2]) == [8, 8]
    assert z[2] == 5.0 and z[3] == -3.75       # ties go toward +inf in value
    assert c[4] == -1 and c[5] == -1 and z[4] == 0 and z[5] == 0


def test_calibration_matches_reference(golden):
    f = golden["calib"]
    n_type = 0
    for m in golden.manifest["calib"]:
        t, x = m["tag"], f[m["tag"] + "_x"]
        per_row = not m["is_input"]
        signed = m["signed"]
        if m["tree"] == "ant":
            mode = m["mode"]
            if m["bit"] > 6:
                chosen = "int"
            elif "ant-" in mode:
                chosen, _ = orc.ant_select_type(x, mode, m["bit"], signed, per_row, m["low"], m["up"])
                n_type += 1
            else:
                chosen = mode
            assert chosen == m["chosen"], t
            grid = orc.ant_grid(chosen, m["bit"], signed)
            assert np.array_equal(grid, f[t + "_grid"]), t
            _, alpha, _ = orc.ant_search_mse(x, grid, m["bit"], per_row, m["low"], m["up"])
            np.testing.assert_allclose(alpha.reshape(-1), f[t + "_alpha"].reshape(-1), rtol=1e-6, err_msg=t)
            y = orc.ant_forward(x, f[t + "_alpha"], grid, per_row)
            same(y, f[t + "_y"])
            mse = np.mean(orc._mse(y, x, per_row))
            np.testing.assert_allclose(mse, f[t + "_mse"], rtol=1e-4, err_msg=t)
        else:
            grid = f[t + "_grid"]
            y = orc.olive_forward(x, f[t + "_alpha"], grid, f[t + "_outliers"], per_row)
            same(y, f[t + "_y"])
            _, alpha = orc.olive_search_mse(x, grid, f[t + "_outliers"], per_row, m["low"], m["up"])
            np.testing.assert_allclose(alpha.reshape(-1), f[t + "_alpha"].reshape(-1), rtol=2e-5, err_msg=t)
    assert n_type >= 6


def test_fp16_definition():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((8, 64)) * 0.05).astype(np.float16)
    g = orc.ant_grid("flint", 4, True)
    a = np.abs(x.astype(np.float32)).max(1) * 0.9
    y16 = orc.ant_forward(x, a, g, per_row=True)
    y32 = orc.ant_forward(x.astype(np.float32), a, g, per_row=True)
    assert y16.dtype == np.float16
    assert np.array_equal(y16, y32.astype(np.float16))
