"""Codebook (quant_grid) generators of ANT and OliVe, written as magnitude sets.

Every generator returns the fp32 table the reference would put in `quant_grid`
(bit for bit: tests/test_codebooks.py checks all bit widths against tables dumped
from the reference).  What the reference builds with nested loops is expressed
here as a set of positive magnitudes per family:

  int     {1 .. 2^B - 1} (+ the lone -2^B when signed)      A/antquant/quant_modules.py:204-221
  pot     {2^0 .. 2^(2^B - 2)}                              A/...:189-201
  flint   octave e in [-(B-1), B-2] holds 2^mb(e) points 2^e (1 + j 2^-mb(e)),
          mb(e) = B-1+e for e < 0, B-2-e for e >= 0, plus the top value 2^(B-1)   A/...:223-278
  float   eb exponent bits, B-eb mantissa bits, first exponent row subnormal     A/...:157-187
  apot    sums of two or three power-of-two terms                               A/...:85-131
  OliVe   int / flint rescaled so the outlier threshold is 32 (no padding)       O/antquant/quant_modules.py:73-153
          abfloat outliers 2^i (1 + j 2^-mb), i in [5, 8], without 32 itself    O/...:157-179

B = value bits = bit - 1 when signed.  ANT tables are padded with one 0 up to 2^bit
entries, sorted, and multiplied by `10.0 / max`, which PyTorch evaluates as
fl32(fl32(1 / max) * 10)  (Tensor.__rtruediv__ is reciprocal-then-multiply).
"""
import itertools

import torch


def _vbits(bit, signed):
    return int(bit) - 1 if signed else int(bit)


def _mirror(mags, signed):
    vals = [0.0]
    for m in mags:
        vals.append(m)
        if signed:
            vals.append(-m)
    return vals


def _ant_finish(vals, bit, device=None):
    n = 2 ** int(bit)
    if len(vals) < n:
        vals = vals + [0.0]
    if len(vals) != n:
        raise AssertionError("codebook has %d entries, 2**bit = %d" % (len(vals), n))
    v, _ = torch.sort(torch.tensor(vals, dtype=torch.float32, device=device))
    return v * (v.max().reciprocal() * 10.0)


def int_magnitudes(B):
    return [float(i) for i in range(1, 2 ** B)]


def pot_magnitudes(B):
    return [2.0 ** i for i in range(2 ** B - 1)]


def flint_magnitudes(B):
    if B < 2:
        raise AssertionError("flint needs at least 2 value bits")
    mags = []
    for e in range(-(B - 1), B - 1):
        mb = B - 1 + e if e < 0 else B - 2 - e
        mags += [2.0 ** e * (1 + j * 2.0 ** -mb) for j in range(2 ** mb)]
    return mags + [2.0 ** (B - 1)]


def float_magnitudes(B, eb):
    mb = B - eb
    if B == 2:
        eb, mb = 2, 0
    if mb < 0:
        raise TypeError("float codebook: %d exponent bits do not fit %d value bits" % (eb, B))
    mags = [j * 2.0 ** -mb for j in range(1, 2 ** mb)]                       # subnormal row
    for i in range(1, 2 ** eb):
        mags += [2.0 ** (i - 1) * (1 + j * 2.0 ** -mb) for j in range(2 ** mb)]
    return mags


_APOT_TERMS = {   # exponents (as negative powers of two) of the two/three additive terms
    2: ([1, 2, 3], [], []),
    3: ([1, 2, 4], [3], []),
    4: ([1, 3, 5], [2, 4, 6], []),
    5: ([1, 3, 6], [2, 4, 7], [5]),
    6: ([1, 4, 7], [2, 5, 8], [3, 6, 9]),
}


def apot_values(B, signed):
    a, b, c = _APOT_TERMS.get(B, ([], [], []))
    terms = [[0.0] + [2.0 ** -k for k in t] for t in (a, b, c)]
    vals = []
    for x, y, z in itertools.product(*terms):
        vals.append(x + y + z)
        if signed:
            vals.append(-(x + y + z))
    return vals


def ant_grid(kind, bit, signed, device=None):
    """kind: int | flint | pot | float (= float3) | float1..float4 | apot."""
    B = _vbits(bit, signed)
    if kind == "int":
        vals = _mirror(int_magnitudes(B), signed)
        if signed:
            vals.append(-float(2 ** B))
    elif kind == "flint":
        vals = _mirror(flint_magnitudes(B), signed)
    elif kind == "pot":
        vals = _mirror(pot_magnitudes(B), signed)
    elif kind == "float" or (kind.startswith("float") and kind[5:].isdigit()):
        vals = _mirror(float_magnitudes(B, int(kind[5:] or 3)), signed)
    elif kind == "apot":
        vals = apot_values(B, signed)
    else:
        raise RuntimeError("Unsupported mode: " + kind)
    return _ant_finish(vals, bit, device)


def olive_grid(kind, bit, signed, device=None):
    B = _vbits(bit, signed)
    if kind == "int":
        mags, unit = int_magnitudes(B), 32 / (2 ** B)
    elif kind == "flint":
        mags, unit = flint_magnitudes(B), 32 / (2 ** (B - 1))
    else:
        raise RuntimeError("Unsupported mode: " + kind)
    v, _ = torch.sort(torch.tensor(_mirror(mags, signed), dtype=torch.float32, device=device))
    return v * unit


def olive_outliers(bit, signed, exp_bit=2, exp_base=5, device=None):
    B = _vbits(bit, signed)
    mb = B - exp_bit
    mags = [2.0 ** i * (1 + j * 2.0 ** -mb) for i in range(exp_base, exp_base + 2 ** exp_bit)
            for j in range(int(2 ** mb))][1:]
    vals = _mirror(mags, signed)[1:]
    v, _ = torch.sort(torch.tensor(vals, dtype=torch.float32, device=device))
    return v
