"""The product's codebook generators (host code) against tables dumped from the reference. CPU only."""
import numpy as np
import pytest


def test_all_reference_tables(golden):
    from antq import codebooks as cbk
    n = 0
    for m in golden.manifest["grids"]:
        tree, kind, su, bit = m["key"].split("_")
        signed, bit = su == "s", int(bit)

        def make():
            if tree == "ant":
                return cbk.ant_grid(kind, bit, signed)
            if kind == "outlier":
                return cbk.olive_outliers(bit, signed)
            return cbk.olive_grid(kind, bit, signed)
        if "error" in m:
            with pytest.raises((AssertionError, TypeError)):
                make()
            continue
        g = make().numpy()
        ref = golden["grids"][m["key"]]
        assert g.dtype == np.float32 and g.shape == ref.shape, m["key"]
        assert np.array_equal(g, ref, equal_nan=True), (m["key"], g, ref)
        n += 1
    assert n > 100


def test_float_alias_and_errors():
    from antq import codebooks as cbk
    assert np.array_equal(cbk.ant_grid("float", 4, True).numpy(), cbk.ant_grid("float3", 4, True).numpy())
    with pytest.raises(RuntimeError):
        cbk.ant_grid("bogus", 4, True)
