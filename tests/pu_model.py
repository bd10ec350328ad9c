"""numpy model of the PIECEWISE-UNIFORM (PU) closed-form codec used by antq_pu.cu -- every fp32 operation of the
kernel's fast path is one numpy float32 operation here (numpy never contracts to FMA; the kernel is built with
--fmad=false and uses explicit _rn intrinsics), so checking this model against the oracle on every fp16 bit pattern
checks the kernel's arithmetic.

Idea.  Every grid the reference generates for int / flint / pot / float (A/antquant/quant_modules.py:157-278) is
`fl32(k * c)` for integers k (c = the smallest positive level), and inside each octave [2^e, 2^(e+1)) the k's form a
uniform progression with a power-of-two step 2^(e - mb_e).  So the nearest level of d is found WITHOUT a scan:
    t  = x * kx                    kx ~ 1 / (s c), a few ulps off
    mf = (t + M_e) - M_e           M_e = 1.5 * 2^23 * step_e : round t to a multiple of step_e (the octave's magic)
    q  = fl32(clamp(mf, kmin, kmax) * c);   out = fl32(q * s)        (STE is exact inside the window)
and the result is provably the reference's unless t lies within delta_e of a midpoint.  delta_e = 2^(e - 19) covers
the error of t (<= 5 ulp), of the levels (1 ulp) and of the scan's rounded distances (1-2 ulp) with margin.  Such an
element is settled without a search (pu_vec_exact): the winner is mf or its neighbour on t's side, mf +- step_e, and
which one is the scan's own comparison of the two rounded distances on d = x / s, the upper level winning a tie
(the grid is scanned in ascending order).  Only elements outside the window (|d| beyond the STE-exact range, NaN,
Inf, dead rows) take the literal arithmetic.
"""
import numpy as np

f32 = np.float32


def _bits(a):
    return np.asarray(a, dtype=f32).view(np.uint32)


def analyze(grid):
    """Returns the PU description of a codebook, or None when the grid is not piecewise uniform."""
    lev = np.unique(np.asarray(grid, dtype=f32))
    lev = lev[~np.isnan(lev)]
    lev = np.where(lev == 0, f32(0), lev).astype(f32)
    lev = np.unique(lev)
    if not (lev == 0).any() or not (lev > 0).any():
        return None
    c = lev[lev > 0].min()
    k = np.rint(lev / c).astype(f32)
    if np.abs(k).max() >= 2 ** 20 or not np.array_equal((k * c).astype(f32), lev):
        return None
    kmin, kmax = k.min(), k.max()
    U = np.unique(np.abs(k[k != 0])).astype(np.int64)
    pos = k[k > 0].astype(np.int64)
    neg = (-k[k < 0]).astype(np.int64)
    if not np.array_equal(np.sort(pos), U[U <= kmax]) or not np.array_equal(np.sort(neg), U[U <= -kmin]):
        return None
    e_top = int(np.floor(np.log2(U.max())))
    ls = []                                                    # log2(step) per octave 0 .. e_top
    for e in range(e_top + 1):
        Ue = U[(U >= 2 ** e) & (U < 2 ** (e + 1))]
        if Ue.size == 0 or Ue[0] != 2 ** e:
            return None
        if Ue.size == 1:
            # a lone level: the octave's step is the octave itself -- except at the top, where anything beyond is
            # clamped anyway and keeping the previous octave's step keeps int-k uniform (signed int-k: {1 .. 2^B})
            step = 2 ** e if (e < e_top or e == 0) else 2 ** ls[e - 1]
        else:
            step = int(Ue[1] - Ue[0])
        if step & (step - 1) or step > 2 ** e:
            return None
        full = np.arange(2 ** e, 2 ** (e + 1), step)
        if e < e_top:
            if not np.array_equal(Ue, full):
                return None
        elif not np.array_equal(Ue, full[:Ue.size]):           # the top octave may be truncated (clamped afterwards)
            return None
        ls.append(int(np.log2(step)))
    magic = np.zeros(256, dtype=f32)
    delta = np.zeros(256, dtype=f32)
    for E in range(256):
        e = min(max(E - 127, 0), e_top)
        magic[E] = f32(1.5 * 2.0 ** (23 + ls[e]))
        delta[E] = f32(2.0 ** (e - 19))
    uniform = all(v == 0 for v in ls)
    # step of the octave each END of the grid lies in (kmin = 0: the sub-unit region, octave 0)
    oct_of = lambda k: 0 if k < 1 else min(int(np.floor(np.log2(k))), e_top)
    step_hi = f32(2.0 ** ls[oct_of(float(kmax))])
    step_lo = f32(2.0 ** ls[oct_of(-float(kmin))])
    ok = lambda lim: float(kmax) <= lim * float(step_hi) and -float(kmin) <= lim * float(step_lo)
    return dict(c=f32(c), inv_c=f32(f32(1) / c), kmin=f32(kmin), kmax=f32(kmax), magic=magic, delta=delta,
                e_top=e_top, ls=ls, uniform=uniform,
                xc_lo=f32(f32(kmin) - f32(f32(0.4) * step_lo)), xc_hi=f32(f32(kmax) + f32(f32(0.4) * step_hi)),
                xc16=ok(400.0), xcbf=ok(48.0))


def _rz16(v):
    """fp32 -> fp16 rounded toward zero."""
    h = np.float16(v)
    if abs(float(h)) > abs(float(v)):
        h = np.nextafter(h, np.float16(0))
    return h


def forward(x, s, pu, lim, exact, out_dtype, xclamp=False, pair_all=False, lean=False):
    """x: array of out_dtype; s: fp32 scale (alpha / max(grid)); lim: the codebook's exact window in d-space.
    `exact(xs)` is the literal reference arithmetic for the flagged elements.  Returns (out, flagged).
    xclamp (fp16 inputs, grids with pu["xc16"]): the clamp to [kmin, kmax] is applied to the INPUT with bounds rounded
    toward zero to fp16, as the kernel does with two packed min / max per pair, and t is not clamped afterwards.
    lean (antq_pu_lean_kernel, rows of one or two vectors): the clamp and the window test are done on t with constants of
    the codebook -- t clamped to [xc_lo, xc_hi], uniform grids flag |r| >= 0.5 - (max|k| + 1) 2^-19, in-window means
    |t| <= lim / c * 0.999 -- and the reciprocal of the scale is the approximate one (<= 1 ulp off)."""
    s = f32(s)
    xf = x.astype(f32)
    with np.errstate(all="ignore"):
        rs = f32(1) / s
        kx = f32(rs * pu["inv_c"])
        xq = xf
        if xclamp:
            sc = f32(s * pu["c"])
            lo, hi = _rz16(f32(pu["xc_lo"] * sc)), _rz16(f32(pu["xc_hi"] * sc))
            xh = x.astype(np.float16)
            xh = np.where(np.isnan(xh), lo, np.maximum(xh, lo))          # hmax2 / hmin2 return the non-NaN operand
            xq = np.minimum(xh, hi).astype(f32)
        t = (xq * kx).astype(f32)
        if lean:
            t_raw = t
            t = np.minimum(np.maximum(np.where(np.isnan(t), pu["xc_lo"], t), pu["xc_lo"]), pu["xc_hi"]).astype(f32)
        E = (_bits(t) >> 23) & 0xff
        M = pu["magic"][E]
        dl = pu["delta"][E]
        mf = ((t + M).astype(f32) - M).astype(f32)
        r = (t - mf).astype(f32)
        h = (_bits(M) - np.uint32((24 << 23) | 0x400000)).view(f32)
        near = np.abs(r) >= (h - dl).astype(f32)                  # |r| <= h always: "within delta of a midpoint"
        if lean and pu["uniform"]:
            hd_c = f32(f32(0.5) - f32(f32(max(float(pu["kmax"]), -float(pu["kmin"])) + 1.0) * f32(2.0 ** -19)))
            near = np.abs(r) >= hd_c
        mfc = mf if (xclamp or lean) else np.minimum(np.maximum(mf, pu["kmin"]), pu["kmax"]).astype(f32)
        q = (mfc * pu["c"]).astype(f32)
        o = (q * s).astype(f32)
        xl = f32(f32(f32(lim) * s) * f32(0.9990234375))
        window = np.abs(xf) <= xl                                 # False for NaN
        if lean:
            tlim = f32(f32(f32(lim) * pu["inv_c"]) * f32(0.9990234375))
            window = np.abs(t_raw) <= tlim
        row_ok = bool(s > 0) and bool(np.isfinite(s)) and bool(np.isfinite(kx)) and bool(kx > 0)
        # pu_vec_exact: the pair decision, from the UNCLAMPED input
        tp = (xf * kx).astype(f32)
        Mp = pu["magic"][(_bits(tp) >> 23) & 0xff]
        step = ((_bits(Mp) & np.uint32(0x7f800000)) - np.uint32(23 << 23)).view(f32)
        mp = ((tp + Mp).astype(f32) - Mp).astype(f32)
        rp = (tp - mp).astype(f32)
        oth = (mp + np.copysign(step, rp)).astype(f32)
        clip = lambda k: np.minimum(np.maximum(k, pu["kmin"]), pu["kmax"]).astype(f32)
        k1, k2 = clip(mp), clip(oth)
        ql, qh = (np.minimum(k1, k2) * pu["c"]).astype(f32), (np.maximum(k1, k2) * pu["c"]).astype(f32)
        d = (xf / s).astype(f32)
        pick = np.where(np.abs((d - qh).astype(f32)) <= np.abs((d - ql).astype(f32)), qh, ql).astype(f32)
        opair = (((pick - d).astype(f32) + d).astype(f32) * s).astype(f32)
    wild = ~window
    if not row_ok:
        wild = np.ones_like(wild)
    if pair_all:                                                  # test hook: the pair decision for EVERY in-window element
        near = np.ones_like(near)
    out = np.where(near, opair, o).astype(out_dtype)
    if wild.any():
        out[wild] = exact(x[wild])
    return out, near | wild
