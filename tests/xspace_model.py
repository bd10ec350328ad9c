"""numpy model of the ALGORITHM the CUDA kernels implement (test artefact only).

It mirrors, step by step, what ant-quantization_b200/csrc does on the device:

  prepare_codebook : grid -> distinct sorted levels, scan codes, and for every
                     adjacent pair the exact fp32 threshold T_r = min{d : the
                     upper level wins the reference scan}, found by bisection
                     over the fp32 number line with the scan's own rounded
                     distances and `<=` tie rule.
  row_tables       : per scale s, thresholds moved into x-space,
                     X_r = min{x in dtype : fl32(x / s) >= T_r}, and the
                     dequantised outputs O_j = fl32(level_j * s) (RNE to dtype).
  forward          : rank = #{r : x >= X_r};  out = O[rank];  anything outside
                     the "exact" window (|d| > 2*vmax, NaN, Inf) takes the
                     reference arithmetic literally.

tests/test_xspace_model.py proves on exhaustive fp16 inputs that this equals
the oracle bit for bit, which is the theory the fast kernel rests on.
"""
import numpy as np

f32 = np.float32


def _ord(x):
    """Monotone map fp32 -> int64 (total order, -0 < +0 adjacent)."""
    b = np.asarray(x, dtype=f32).view(np.int32).astype(np.int64)
    return np.where(b < 0, -(b & 0x7FFFFFFF) - 1, b)


def _unord(o):
    o = np.asarray(o, dtype=np.int64)
    b = np.where(o < 0, (-(o + 1)) | 0x80000000, o).astype(np.uint32)
    return b.view(f32)


def prepare_codebook(grid):
    g = np.asarray(grid, dtype=f32).reshape(-1)
    K = g.size
    valid = ~np.isnan(g)
    last = np.array([valid[i] and not any(g[j] == g[i] for j in range(i + 1, K)) for i in range(K)])
    idx = [i for i in range(K) if last[i]]
    idx.sort(key=lambda i: g[i])
    levels = np.array([g[i] for i in idx], dtype=f32)
    levels = np.where(levels == 0, f32(0.0), levels)          # canonical +0
    codes = np.array(idx, dtype=np.int32)
    thr = np.empty(len(idx) - 1, dtype=f32)
    for r in range(len(idx) - 1):
        lo, hi = levels[r], levels[r + 1]
        tie_hi = codes[r + 1] > codes[r]

        def hi_wins(d):
            dl, dh = np.abs(f32(d) - lo), np.abs(f32(d) - hi)
            return dh < dl or (dh == dl and tie_hi)
        a, b = int(_ord(lo)), int(_ord(hi))        # hi_wins(lo) False, hi_wins(hi) True
        while b - a > 1:
            m = (a + b) // 2
            if hi_wins(_unord(m)):
                b = m
            else:
                a = m
        thr[r] = _unord(b)
    return dict(grid=g, levels=levels, codes=codes, thr=thr, vmax=levels[-1], vmin=levels[0])


def interior_exact(cb):
    """(q - d) + d == q for every d strictly inside the grid's span."""
    L, T = cb["levels"], cb["thr"]
    ok = True
    for r in range(len(T)):
        lo, hi, t = L[r], L[r + 1], T[r]
        lo_ok = lo == 0 or (lo > 0 and t <= 2 * lo) or (lo < 0 and t <= lo / 2)
        hi_ok = hi == 0 or (hi < 0 and t >= 2 * hi) or (hi > 0 and t >= hi / 2)
        ok = ok and lo_ok and hi_ok
    return ok and cb["vmax"] > 0 and cb["vmin"] <= 0


def _step(x, up, dtype):
    x = np.asarray(x, dtype=dtype)
    return np.nextafter(x, dtype(np.inf) if up else dtype(-np.inf))


def x_threshold_exact(t, s, dtype):
    """min{x in dtype : fl32(f32(x) / s) >= t}   (s > 0 finite): monotone walk."""
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        c = dtype(f32(t) * f32(s))
        def ok(x):
            return f32(x) / f32(s) >= f32(t)
        for _ in range(6):
            p = _step(c, False, dtype)
            if np.isinf(p) or not ok(p):      # never step onto -inf: x >= -max holds for every finite x
                break
            c = p
        for _ in range(6):
            if not ok(c):
                c = _step(c, True, dtype)
            else:
                break
    return dtype(c)


N_SHORTCUT = [0, 0]


def x_threshold(t, s, dtype):
    """antq_x_threshold: division-free common case for fp16, exact walk otherwise."""
    if dtype == np.float16:
        with np.errstate(over="ignore"):
            p = f32(t) * f32(s)
        low = int(np.asarray(p, dtype=f32).view(np.uint32)) & 0x1FFF
        ap = abs(float(p))
        if 16 <= low <= 0x1FFF - 16 and 2.0 ** -13 <= ap <= 60000.0:
            c = np.float16(p)                      # RN, then fix up to round-toward-+inf
            if f32(c) < p:
                c = _step(c, True, np.float16)
            N_SHORTCUT[0] += 1
            return np.float16(c)
    N_SHORTCUT[1] += 1
    return x_threshold_exact(t, s, dtype)


def row_tables(cb, s, dtype):
    X = np.array([x_threshold(t, s, dtype) for t in cb["thr"]], dtype=dtype)
    with np.errstate(over="ignore"):
        O = (cb["levels"] * f32(s)).astype(f32).astype(dtype)
    lim = f32(2.0) * min(cb["vmax"], -cb["vmin"] if cb["vmin"] < 0 else cb["vmax"])
    # conservative x-space window: every |x| <= xlim has |fl32(x/s)| <= lim
    xlim = f32(lim) * f32(s) * f32(1 - 2.0 ** -10)
    return X, O, xlim


def forward_fast(x, s, cb, dtype, exact_fn):
    """x: 1-D array of dtype; s: fp32 scale.  exact_fn(x_subset) -> reference result."""
    x = np.asarray(x, dtype=dtype)
    out = np.empty_like(x)
    if not (np.isfinite(s) and s > 0):
        return exact_fn(x), np.ones(x.shape, bool)
    X, O, xlim = row_tables(cb, s, dtype)
    with np.errstate(invalid="ignore"):
        rank = (x[:, None] >= X[None, :]).sum(1)
        slow = ~(np.abs(x.astype(f32)) <= xlim)
    out[:] = O[rank]
    if slow.any():
        out[slow] = exact_fn(x[slow])
    return out, slow


# ---- symmetric / "symmetric + one extra negative level" (SYMX) forms used by the CUDA kernels ---------------------
def fold_signs(cb):
    """What antq_prepare_kernel derives for a codebook that is symmetric about a zero level (L odd) or symmetric except
    for ONE extra level at the negative end (L even, signed int-k):  mid = index of the zero level, magnitude thresholds
    tpos[i] (d >= 0: magnitude i+1 wins iff d >= tpos[i]) and tneg[i] (d < 0: ... iff -d >= tneg[i]), n_mag magnitudes
    present on both sides, and for SYMX the d-space threshold thr[0] below which the extra level level[0] wins."""
    L, T = cb["levels"], cb["thr"]
    n = len(L)
    mid = n >> 1
    sym = n % 2 == 1 and n >= 3 and L[mid] == 0 and all(L[mid + k] == -L[mid - k] for k in range(1, mid + 1))
    symx = (not sym) and n % 2 == 0 and n >= 4 and L[mid] == 0 and all(L[mid + k] == -L[mid - k] for k in range(1, mid))
    if not (sym or symx):
        return None
    nt = mid if sym else mid - 1
    tpos = np.array([T[mid + i] for i in range(nt)], dtype=f32)
    tneg = np.array([_unord(_ord(f32(-T[mid - i - 1])) + 1) for i in range(nt)], dtype=f32)
    return dict(sym=sym, symx=symx, mid=mid, nt=nt, tpos=tpos, tneg=tneg,
                thr_e=T[0] if symx else None, lev_e=L[0] if symx else None)


def forward_dspace(x, s, cb, dtype, exact_fn):
    """antq_short_kernel: d = fl32(x / s), compare chain against the codebook's d-space thresholds, literal STE."""
    x = np.asarray(x, dtype=dtype)
    fs = fold_signs(cb)
    with np.errstate(all="ignore"):
        d = (x.astype(f32) / f32(s)).astype(f32)
        ad = np.abs(d)
        neg = np.signbit(d)
        if fs is None:
            rank = (d[:, None] >= cb["thr"][None, :]).sum(1)
            q = cb["levels"][rank]
        else:
            t = np.where(neg[:, None], fs["tneg"][None, :], fs["tpos"][None, :])
            rank = (ad[:, None] >= t).sum(1)
            q = np.copysign(cb["levels"][fs["mid"] + rank], d).astype(f32)
            if fs["symx"]:
                q = np.where(d < fs["thr_e"], fs["lev_e"], q).astype(f32)
        out = (((q - d).astype(f32) + d).astype(f32) * f32(s)).astype(f32).astype(dtype)
        lim_idx = f32(65536.0)
        slow = ~(ad <= lim_idx)
    if slow.any():
        out[slow] = exact_fn(x[slow])
    return out, slow


def forward_fast_folded(x, s, cb, dtype, exact_fn):
    """antq_stream_kernel for SYMMETRIC / SYMX codebooks: x-space thresholds on |x| (X for x >= 0, Xn for x < 0),
    outputs O = RN(level * s) with the sign of x, and for SYMX one signed compare x < Xe for the extra level."""
    x = np.asarray(x, dtype=dtype)
    fs = fold_signs(cb)
    assert fs is not None
    if not (np.isfinite(s) and s > 0):
        return exact_fn(x), np.ones(x.shape, bool)
    X = np.array([x_threshold_exact(t, s, dtype) for t in fs["tpos"]], dtype=dtype)
    Xn = np.array([x_threshold_exact(t, s, dtype) for t in fs["tneg"]], dtype=dtype)
    with np.errstate(all="ignore"):
        O = (cb["levels"][fs["mid"]:fs["mid"] + fs["nt"] + 1] * f32(s)).astype(f32).astype(dtype)
        ax = np.abs(x)
        neg = x < 0                                   # -0 counts as positive, like HSET2.LT
        t = np.where(neg[:, None], Xn[None, :], X[None, :])
        rank = (ax[:, None] >= t).sum(1)
        out = O[rank]
        out = np.where(np.signbit(x) & (rank > 0), -out, out).astype(dtype)
        if fs["symx"]:
            Xe = x_threshold_exact(fs["thr_e"], s, dtype)
            Oe = (f32(fs["lev_e"]) * f32(s)).astype(f32).astype(dtype)
            out = np.where(x < Xe, Oe, out).astype(dtype)
        lim = f32(2.0) * min(cb["vmax"], -cb["vmin"])
        xlim = f32(lim) * f32(s) * f32(1 - 2.0 ** -10)
        slow = ~(np.abs(x.astype(f32)) <= xlim)
    if slow.any():
        out[slow] = exact_fn(x[slow])
    return out, slow
