"""The C-ABI library loads and exports every symbol include/antq.h declares (no compute, CPU only)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "antq.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(antq_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    from antq import _lib
    assert os.path.exists(_lib.SO_PATH)
    L = ctypes.CDLL(_lib.SO_PATH)
    names = declared_symbols()
    assert len(names) >= 14, names
    for n in names:
        assert hasattr(L, n), "libantq.so does not export " + n


def test_abi_basics_without_gpu():
    from antq import _lib
    assert _lib.lib.antq_abi_version() == 1
    assert b"sm_100a" in _lib.lib.antq_build_info()
    assert _lib.lib.antq_codebook_bytes() > 4 * 512 * 4
    assert b"invalid" in _lib.lib.antq_error_string(-1)
    # argument errors are reported before any CUDA call
    assert _lib.lib.antq_fakequant(None, None, None, None, 0, 4, 4, 0, None, None, 0, None) == _lib.EINVAL
    assert _lib.lib.antq_lut_nearest(None, None, None, -1, 0, None, None) == _lib.EINVAL
    assert _lib.lib.antq_codebook_prepare(None, 0, None, 0, None, None) == _lib.EINVAL
    info = _lib.CodebookInfo(n_entries=16, n_normal=16, n_levels=15, flags=7, n_mag=8, mid=7, ovp_index=-1)
    assert _lib.lib.antq_fakequant_plan(ctypes.byref(info), 4096, 4096, _lib.F16, 0, 256, 512, None) == 1
    assert _lib.lib.antq_fakequant_plan(ctypes.byref(info), 4096, 64, _lib.F16, 0, 256, 512, None) == 3    # short rows
    assert _lib.lib.antq_fakequant_plan(ctypes.byref(info), 4096, 64, _lib.F16, 0, 256, 512, 1024) == 2  # codes wanted
    assert _lib.lib.antq_fakequant_plan(ctypes.byref(info), 4096, 61, _lib.F16, 0, 256, 512, None) == 2  # ragged rows
    assert _lib.lib.antq_fakequant_plan(ctypes.byref(info), 4096, 4097, _lib.F16, _lib.FLAG_FORCE_ROWS, 256, 512,
                                        None) == _lib.ENOTSUP
    assert _lib.lib.antq_fakequant_plan(None, 4096, 4096, _lib.F16, 0, 256, 512, None) == 2
    # piecewise-uniform grids: chain up to 7 folded thresholds, closed form beyond and for short rows
    pu7 = _lib.CodebookInfo(n_entries=16, n_normal=16, n_levels=15, flags=7 | _lib.CB_PU, n_mag=8, mid=7, ovp_index=-1)
    pu127 = _lib.CodebookInfo(n_entries=256, n_normal=256, n_levels=256, flags=3 | _lib.CB_SYMX | _lib.CB_PU | _lib.CB_PU_UNIFORM,
                              n_mag=128, mid=128, ovp_index=-1)
    plan = lambda info, r, c, fl=0, codes=None: _lib.lib.antq_fakequant_plan(ctypes.byref(info), r, c, _lib.F16, fl, 256, 512, codes)
    assert plan(pu7, 4096, 4096) == 1 and plan(pu7, 4096, 64) == 5 and plan(pu7, 4096, 64, _lib.FLAG_NO_PU) == 3
    # fp32 I/O: the closed form even for a 7-threshold grid (the chain has no packed pairs there)
    assert _lib.lib.antq_fakequant_plan(ctypes.byref(pu7), 4096, 4096, _lib.F32, 0, 256, 512, None) == 4
    assert _lib.lib.antq_fakequant_plan(ctypes.byref(pu7), 4096, 4096, _lib.F32, _lib.FLAG_NO_PU, 256, 512, None) == 1
    assert plan(pu127, 4096, 4096) == 4 and plan(pu127, 1, 12345) == 4 and plan(pu127, 4096, 32) == 5
    assert plan(pu127, 4096, 4096, _lib.FLAG_NO_PU) == 2 and plan(pu127, 4096, 4096, 0, 1024) == 2
    assert plan(pu7, 4096, 4096, _lib.FLAG_FORCE_PU) == 4 and plan(pu7, 4096, 4096, _lib.FLAG_OVP | _lib.FLAG_FORCE_PU) == _lib.ENOTSUP


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ant-quantization_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "antq_oracle" not in txt and "ref_harness" not in txt, os.path.join(dp, f)
