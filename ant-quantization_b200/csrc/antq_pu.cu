// antq_pu.cu -- closed-form fake-quant for PIECEWISE-UNIFORM codebooks (ANTQ_CB_PU): int-k of every width (the
// reference forces int above 6 bits, A/antquant/quant_modules.py:482-483), unsigned 4-bit grids (every post-ReLU
// activation), 5/6-bit flint / pot / float -- everything whose compare chain would be 15 to 255 thresholds long.
//
// Same reference arithmetic as the other kernels (A/antquant/quant_modules.py:535-551, A/quant/quant_kernel.cu:25-37):
//   s = alpha / max(grid);  d = fl32(x / s);  q = scan(d);  out = fl32(((q - d) + d) * s)
// but the level is found WITHOUT a scan, a chain or a table per row (model + proof by exhaustion: tests/pu_model.py):
//   t  = x * kx                      kx = fl32(fl32(1 / s) / c), c = the grid's unit: every level is fl32(k c), k integer
//   mf = (t + M_e) - M_e             M_e = 1.5 * 2^23 * step_e rounds t to the octave's (power-of-two) step
//   q  = fl32(clamp(mf, kmin, kmax) * c);   out = RN(fl32(q * s))      ((q - d) + d == q inside the window)
// t is a few ulps off d / c, so an element whose t lies within delta_e = 2^(e - 19) of a midpoint ("near") is settled by
// comparing the two levels it lies between with the scan's own rounded distances on the true quotient (pu_elem_exact: no
// search; equals the scan for every in-window element, tests/test_pu_model.py), and an element outside the exact window,
// NaN, Inf, or in a row whose scale is not a positive finite number ("wild") takes the literal arithmetic (true division,
// the codebook's exact thresholds / the literal scan).  With 16-bit data near elements are ~1e-4 of all, except in rows
// where a tie x / s == midpoint is representable.  Where that work runs matters: a launch ends with its slowest warp,
// so near vectors are parked in a CTA-wide queue and settled by all threads after the last chunk.
//
// OliVe (ANTQ_CB_PU_OVP): the NORMAL levels are piecewise uniform and the window is cut below the first outlier
// threshold, so the closed form covers every pair without an outlier; a pair holding an out-of-window element takes the
// reference's pair logic (O/antquant/quant_modules.py:311-320) on the whole codebook.
//
// Execution shapes:
//   antq_pu_stream_kernel   rows >= 512 elements / per-tensor: the persistent pipeline of antq_stream.cu (one CTA per
//                           SM, 16 consumer warps, two private 4 KiB TMA stages each, chunk counter) -- minus the row
//                           tables, the builder warps and the prologue: a row needs three scalars.  SHORT mode: rows of
//                           16-127 vectors on uniform grids, chunked flat, with a per-chunk table of row constants.
//   antq_pu_short_kernel    shorter rows and scale groups (group-8/16/32, 1x1-conv weights): a warp owns a tile of 128
//                           vectors, issues its loads first, computes the tile's row constants once (one row per lane,
//                           shared through shared memory) and then runs the closed form.
//   antq_pu_lean_kernel     rows of one or two vectors (group-8 / 16): window and clamp in t-space, a row costs one division.
//   antq_pu_dynamic_kernel  the tile kernel with alpha = max|x| * ratio of each group computed in the kernel (one read of x).
// Bound: HBM for uniform grids (12.75 instructions per element: the chain kernel's plateau); the per-octave-table grids
// (15.7 instructions per element) are issue-bound at ~0.65 of the HBM rate (measured: profiles/r02_notes.md).
#include <stdio.h>
#include <stdlib.h>

#include "antq_common.cuh"

namespace {

#ifndef ANTQ_PU_CONSUMERS
#define ANTQ_PU_CONSUMERS 16
#endif
constexpr int kNC = ANTQ_PU_CONSUMERS;    // consumer warps per CTA
constexpr int kNS = 2 * kNC;              // two private stages per warp
constexpr int kChunkMax = 4096;
constexpr int kThreads = kNC * 32;
constexpr int kListMax = 1024;            // per-warp work list of elements / pairs to redo literally (per chunk; a chunk holds <= 1024 pairs)
constexpr int kScratch = 2176;            // per-warp scratch in bytes: that list, or (SHORT) the chunk's row table
constexpr int kQCap = 512;                // CTA-wide queue of near-midpoint vectors settled after the last chunk
struct __align__(16) PuQEntry { uint4 raw; long long v; float s, kx; };

struct PuParams {
    const void *x;
    void *out;
    const float *alpha;
    const AntqCodebook *cb;
    long long rows, cols;
    unsigned total_chunks, chunks_per_cta, chunks_rem;
    int alpha_per_row, chunks_per_row, chunk_elems, cpr_shift;
    float gmax, lim;
    int debug;
    int ovp;                                   // OliVe outlier-victim pairs (ANTQ_CB_PU_OVP codebooks)
    // short kernel
    unsigned nvec, cols_vec, cols_magic;
    int cols_shift;
};

struct PuK {                               // the codebook's closed-form constants (AntqCodebook::pu_*)
    float c, inv_c, kmin, kmax;
    float xc_lo, xc_hi;                    // x-space clamp: (kmin - 0.4 step_top) and (kmax + 0.4 step_top), in units of c
    float hd_c;                            // uniform grids behind the x-space clamp: 0.5 - (max|k| + 1) 2^-19, one constant
    float lim;                             // |d| <= lim: the closed form is proven (OVP: and no outlier level is reached)
};
__device__ __forceinline__ PuK pu_load_k(const AntqCodebook *__restrict__ cb, const float lim, const int ovp) {
    PuK k;
    k.lim = ovp ? fminf(lim, cb->pu_tout) : lim;
    k.c = cb->pu_c; k.inv_c = cb->pu_inv_c; k.kmin = cb->pu_kmin; k.kmax = cb->pu_kmax;
    // step of the octave each end of the grid lies in, from that octave's magic constant M = 1.5 * 2^23 * step
    // (kmin = 0: exponent field 0 -> the sub-unit region's entry)
    const float step_hi = __fmul_rn(cb->pu_tab[__float_as_uint(k.kmax) >> 23].x, 7.94728597e-08f);           // 1 / (1.5 * 2^23)
    const float step_lo = __fmul_rn(cb->pu_tab[(__float_as_uint(k.kmin) & 0x7fffffffu) >> 23].x, 7.94728597e-08f);
    k.xc_hi = __fadd_rn(k.kmax, __fmul_rn(0.4f, step_hi));
    k.xc_lo = __fsub_rn(k.kmin, __fmul_rn(0.4f, step_lo));
    k.hd_c = __fmaf_rn(__fadd_rn(fmaxf(k.kmax, -k.kmin), 1.0f), -1.9073486328125e-06f, 0.5f);
    return k;
}

struct PuRow {
    float s, kx, xl;
    uint32_t xl2;                     // 16-bit types: xl rounded toward zero, in both halves
    uint32_t xlo2, xhi2;              // 16-bit types: x-space clamp bounds (rounded toward zero), in both halves
    bool ok;
};

template <typename T>
__device__ __forceinline__ PuRow pu_row(float alpha, const PuParams &p, const PuK &K, bool fast_rcp) {
    PuRow r;
    const float inf = __int_as_float(0x7f800000);
    r.s = __fdiv_rn(alpha, p.gmax);                                   // scale = alpha / max(grid)
    float rs;
    if (fast_rcp) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(r.s));   // <= 1 ulp: inside delta's budget
    else rs = __fdiv_rn(1.0f, r.s);
    r.kx = __fmul_rn(rs, K.inv_c);
    r.ok = r.s > 0.0f && r.s < inf && r.kx > 0.0f && r.kx < inf;
    r.xl = __fmul_rn(__fmul_rn(K.lim, r.s), 0.9990234375f);           // conservative exact window in x-space
    r.xl2 = 0; r.xlo2 = 0; r.xhi2 = 0;
    if constexpr (sizeof(T) == 2) {
        const uint32_t b = AntqType<T>::bits(AntqType<T>::from_f32_rz(r.xl));
        r.xl2 = b | (b << 16);
        const float sc = __fmul_rn(r.s, K.c);                          // ~ 1 / kx
        const uint32_t lo = AntqType<T>::bits(AntqType<T>::from_f32_rz(__fmul_rn(K.xc_lo, sc)));
        const uint32_t hi = AntqType<T>::bits(AntqType<T>::from_f32_rz(__fmul_rn(K.xc_hi, sc)));
        r.xlo2 = lo | (lo << 16); r.xhi2 = hi | (hi << 16);
    }
    return r;
}

// The closed form for one element.  `flag` accumulates "redo me exactly".  CLAMP = false: the input was already clamped
// in x-space (packed min / max on the 16-bit pairs), so t cannot round beyond [kmin, kmax].
template <bool UNIFORM, bool CLAMP, bool HDC>
__device__ __forceinline__ float pu_quant(float xf, const PuRow &r, const PuK &K, const float2 *tab, bool &flag) {
    const float t = __fmul_rn(xf, r.kx);
    float M, hd;
    if (UNIFORM) {
        M = 12582912.0f;                                              // 1.5 * 2^23: step 1 everywhere
        // 0.5 - |t| 2^-19; behind the x-space clamp |t| <= max|k| + 0.4, so one constant (a little more conservative) does
        hd = (CLAMP || !HDC) ? __fmaf_rn(fabsf(t), -1.9073486328125e-06f, 0.5f) : K.hd_c;
    } else {
        const float2 md = tab[__float_as_uint(t) >> 23];              // sign + exponent index a 512-entry table
        M = md.x; hd = md.y;                                          // step / 2 - delta_e
    }
    const float mf = __fsub_rn(__fadd_rn(t, M), M);
    const float rr = __fsub_rn(t, mf);                                // exact; |rr| <= step / 2
    flag |= fabsf(rr) >= hd;                                          // within delta of a midpoint
    const float mc = CLAMP ? fminf(fmaxf(mf, K.kmin), K.kmax) : mf;
    return __fmul_rn(__fmul_rn(mc, K.c), r.s);
}

// Exact thresholds and levels of the codebook, staged in shared memory for the redo path (the rank search is 3-8
// dependent loads: from global memory that costs several microseconds per flagged element).
struct PuExact {
    const float *thr, *lev;           // shared memory, n_levels - 1 and n_levels entries
    int nlev;
    float win;                        // |d| <= win: the threshold search equals the scan (else the literal scan)
};

// The reference arithmetic for ONE element, literally (exact thresholds inside the proven window, else the scan).
template <typename T>
__device__ __noinline__ T pu_exact_elem(const AntqCodebook *__restrict__ cb, const PuExact X, float xf, float s) {
    const float d = __fdiv_rn(xf, s);
    float q;
    if (fabsf(d) <= X.win) {
        q = X.lev[antq_rank(X.thr, X.nlev - 1, d)];
    } else {
        int code;
        q = antq_scan_literal(cb->grid, cb->n_entries, d, code);
    }
    return AntqType<T>::from_f32_rn(antq_ste_rescale(q, d, s));
}

template <typename T> struct PuPack { typedef float2 v2; };
template <> struct PuPack<__half> {
    typedef __half2 v2;
    __device__ static __forceinline__ v2 from_u32(uint32_t u) { return *reinterpret_cast<v2 *>(&u); }
};
template <> struct PuPack<__nv_bfloat16> {
    typedef __nv_bfloat162 v2;
    __device__ static __forceinline__ v2 from_u32(uint32_t u) { return *reinterpret_cast<v2 *>(&u); }
};

template <typename V> __device__ __forceinline__ uint32_t antq_pu_u32(const V &v) { return *reinterpret_cast<const uint32_t *>(&v); }

template <typename T> struct PuIO;
template <> struct PuIO<float> {
    static constexpr int VEC = 4;
    __device__ static __forceinline__ void unpack(const uint4 r, float (&f)[4]) {
        f[0] = __uint_as_float(r.x); f[1] = __uint_as_float(r.y); f[2] = __uint_as_float(r.z); f[3] = __uint_as_float(r.w);
    }
    __device__ static __forceinline__ uint4 pack(const float (&o)[4]) {
        return make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3]));
    }
};
template <> struct PuIO<__half> {
    static constexpr int VEC = 8;
    __device__ static __forceinline__ void unpack(const uint4 r, float (&f)[8]) {
        const __half2 *h = reinterpret_cast<const __half2 *>(&r);
#pragma unroll
        for (int i = 0; i < 4; i++) { const float2 v = __half22float2(h[i]); f[2 * i] = v.x; f[2 * i + 1] = v.y; }
    }
    __device__ static __forceinline__ uint4 pack(const float (&o)[8]) {
        uint4 q;
        __half2 *h = reinterpret_cast<__half2 *>(&q);
#pragma unroll
        for (int i = 0; i < 4; i++) h[i] = __floats2half2_rn(o[2 * i], o[2 * i + 1]);
        return q;
    }
};
template <> struct PuIO<__nv_bfloat16> {
    static constexpr int VEC = 8;
    __device__ static __forceinline__ void unpack(const uint4 r, float (&f)[8]) {
        const unsigned w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; i++) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
    }
    __device__ static __forceinline__ uint4 pack(const float (&o)[8]) {
        uint4 q;
        __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&q);
#pragma unroll
        for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
        return q;
    }
};

// One element, literally, WITHOUT the rescale: the level the scan picks and the quotient it was picked for.
__device__ __forceinline__ void pu_exact_qd(const AntqCodebook *__restrict__ cb, const PuExact &X, float xf, float s, float &q, float &d) {
    d = __fdiv_rn(xf, s);
    if (fabsf(d) <= X.win) {
        q = X.lev[antq_rank(X.thr, X.nlev - 1, d)];
    } else {
        int code;
        q = antq_scan_literal(cb->grid, cb->n_entries, d, code);
    }
}
// OliVe outlier-victim pair (O/antquant/quant_modules.py:311-320): an outlier (|q| > 32) in the even slot zeroes the odd
// one, else an outlier in the odd slot zeroes the even one; then STE and rescale, each with its own quotient.
template <typename T>
__device__ __forceinline__ void pu_exact_pair(const AntqCodebook *__restrict__ cb, const PuExact &X, float x0, float x1, float s,
                                              T &o0, T &o1) {
    float q0, d0, q1, d1;
    pu_exact_qd(cb, X, x0, s, q0, d0);
    pu_exact_qd(cb, X, x1, s, q1, d1);
    const bool oe = fabsf(q0) > 32.0f, oo = fabsf(q1) > 32.0f;
    if (oe) q1 = __fmul_rn(q1, 0.0f);
    else if (oo) q0 = __fmul_rn(q0, 0.0f);
    o0 = AntqType<T>::from_f32_rn(antq_ste_rescale(q0, d0, s));
    o1 = AntqType<T>::from_f32_rn(antq_ste_rescale(q1, d1, s));
}
// A whole vector through the pair logic (queue overflow, rows with a bad scale).
template <typename T>
__device__ __noinline__ void pu_redo_vec_ovp(const AntqCodebook *__restrict__ cb, const PuExact X, const uint4 raw, const float s, T *og) {
    constexpr int VEC = PuIO<T>::VEC;
    float f[VEC];
    PuIO<T>::unpack(raw, f);
#pragma unroll
    for (int e = 0; e < VEC; e += 2) pu_exact_pair<T>(cb, X, f[e], f[e + 1], s, og[e], og[e + 1]);
}

// One 16-byte vector through the closed form.  `near` = some element lies within delta of a midpoint (settled by
// pu_vec_exact); `wild` = some element is outside the window in which the closed form is proven (|d| beyond the STE
// window, NaN, Inf: the literal path).
// XC (16-bit types): clamp the INPUT pairs with two packed min / max instead of every t with two FMNMX.
template <typename T, bool UNIFORM, bool XC, bool HDC = false>
__device__ __forceinline__ uint4 pu_vec(const uint4 raw, const PuRow &r, const PuK &K, const float2 *tab, bool &near, bool &wild) {
    constexpr int VEC = PuIO<T>::VEC;
    float f[VEC], o[VEC];
    bool fn = false, fw = false;
    if constexpr (sizeof(T) == 2) {
        typedef typename PuPack<T>::v2 v2;
        // one packed NaN-propagating max of |x| per vector against the exact window
        const v2 a = __hmax2_nan(__habs2(PuPack<T>::from_u32(raw.x)), __habs2(PuPack<T>::from_u32(raw.y)));
        const v2 b = __hmax2_nan(__habs2(PuPack<T>::from_u32(raw.z)), __habs2(PuPack<T>::from_u32(raw.w)));
        fw = __hle2_mask(__hmax2_nan(a, b), PuPack<T>::from_u32(r.xl2)) != 0xffffffffu;
        if constexpr (XC) {
            const v2 lo = PuPack<T>::from_u32(r.xlo2), hi = PuPack<T>::from_u32(r.xhi2);
            uint4 c;
            c.x = antq_pu_u32(__hmin2(__hmax2(PuPack<T>::from_u32(raw.x), lo), hi));
            c.y = antq_pu_u32(__hmin2(__hmax2(PuPack<T>::from_u32(raw.y), lo), hi));
            c.z = antq_pu_u32(__hmin2(__hmax2(PuPack<T>::from_u32(raw.z), lo), hi));
            c.w = antq_pu_u32(__hmin2(__hmax2(PuPack<T>::from_u32(raw.w), lo), hi));
            PuIO<T>::unpack(c, f);
        } else {
            PuIO<T>::unpack(raw, f);
        }
    } else {
        PuIO<T>::unpack(raw, f);
    }
#pragma unroll
    for (int e = 0; e < VEC; e++) {
        o[e] = pu_quant<UNIFORM, !(XC && sizeof(T) == 2), HDC>(f[e], r, K, tab, fn);
        if (sizeof(T) == 4) fw |= !(fabsf(f[e]) <= r.xl);             // outside the exact window, NaN, Inf
    }
    near = fn;
    wild = fw;
    return PuIO<T>::pack(o);
}

// The same vector settled exactly, element by element, WITHOUT a search: t lies between the level it rounds to (mf) and
// that level's neighbour on t's side (mf +- the octave's step), so the scan's winner is one of the two -- and which one
// is the scan's own comparison of rounded distances on d = x / s (true division), the upper level winning a tie
// (ascending grids: the later entry).  Valid for every element inside the window, near a midpoint or not; elements
// outside it get garbage here and are rewritten by the literal path.  Proof sketch: |t - d / c| is a few ulps, far
// below step / 2, so rank(d) is the index of one of the two candidates (ANTQ_CB_WELLSEP: adjacent thresholds decide).
// The two-candidate decision for one in-window element: the winning level as k (units of c), its value q = fl32(k c) and
// the true quotient d.
template <bool UNIFORM>
__device__ __forceinline__ float pu_elem_exact_k(const float xf, const float s, const float kx, const PuK &K, const float2 *tab,
                                                 float &q, float &d) {
    const float t = __fmul_rn(xf, kx);
    float M = 12582912.0f, step = 1.0f;
    if (!UNIFORM) {
        M = tab[__float_as_uint(t) >> 23].x;                          // 1.5 * 2^23 * step
        step = __uint_as_float((__float_as_uint(M) & 0x7f800000u) - (23u << 23));
    }
    const float mf = __fsub_rn(__fadd_rn(t, M), M);
    const float rr = __fsub_rn(t, mf);
    const float oth = __fadd_rn(mf, copysignf(step, rr));
    const float k1 = fminf(fmaxf(mf, K.kmin), K.kmax), k2 = fminf(fmaxf(oth, K.kmin), K.kmax);
    const float kl = fminf(k1, k2), kh = fmaxf(k1, k2);
    const float ql = __fmul_rn(kl, K.c), qh = __fmul_rn(kh, K.c);
    d = __fdiv_rn(xf, s);
    const float dl = fabsf(__fsub_rn(d, ql)), dh = fabsf(__fsub_rn(d, qh));
    const bool up = dh <= dl;
    q = up ? qh : ql;
    return up ? kh : kl;
}
template <bool UNIFORM>
__device__ __forceinline__ float pu_elem_exact(const float xf, const float s, const float kx, const PuK &K, const float2 *tab) {
    float q, d;
    (void)pu_elem_exact_k<UNIFORM>(xf, s, kx, K, tab, q, d);
    return antq_ste_rescale(q, d, s);
}
template <typename T, bool UNIFORM>
__device__ __noinline__ uint4 pu_vec_exact(const uint4 raw, const float s, const float kx, const PuK K, const float2 *tab) {
    constexpr int VEC = PuIO<T>::VEC;
    float f[VEC], o[VEC];
    PuIO<T>::unpack(raw, f);
#pragma unroll
    for (int e = 0; e < VEC; e++) o[e] = pu_elem_exact<UNIFORM>(f[e], s, kx, K, tab);
    return PuIO<T>::pack(o);
}

// Near-midpoint repair of a vector the fast path has already quantized (`q`): only the flagged ELEMENTS are settled by the
// two-candidate decision (one division each) instead of all VEC.  In the tile kernels a rare per-lane event is a frequent
// per-warp one (int-8: 0.4-0.8 % of the vectors, i.e. every 5th to 8th warp iteration), so what it costs matters: ~90
// instructions here against ~350 for pu_vec_exact.  The flags are recomputed with the most conservative margin of the
// fast paths (a superset of theirs; the decision is valid for every in-window element).
template <typename T, bool UNIFORM>
__device__ __noinline__ uint4 pu_vec_fix_near(const uint4 raw, const uint4 q, const float s, const float kx, const PuK K, const float2 *tab) {
    constexpr int VEC = PuIO<T>::VEC;
    float f[VEC];
    PuIO<T>::unpack(raw, f);
    T o[VEC];
    *reinterpret_cast<uint4 *>(o) = q;
    unsigned m = 0;
#pragma unroll
    for (int e = 0; e < VEC; e++) {
        const float tc = fminf(fmaxf(__fmul_rn(f[e], kx), K.xc_lo), K.xc_hi);
        float M = 12582912.0f, hd = K.hd_c;
        if (!UNIFORM) { const float2 md = tab[__float_as_uint(tc) >> 23]; M = md.x; hd = md.y; }
        const float mf = __fsub_rn(__fadd_rn(tc, M), M);
        m |= (fabsf(__fsub_rn(tc, mf)) >= hd ? 1u : 0u) << e;
    }
    while (m) {
        const int e = __ffs(m) - 1;
        m &= m - 1;
        float xf = f[0];
#pragma unroll
        for (int i = 1; i < VEC; i++) xf = e == i ? f[i] : xf;
        const T val = AntqType<T>::from_f32_rn(pu_elem_exact<UNIFORM>(xf, s, kx, K, tab));
#pragma unroll
        for (int i = 0; i < VEC; i++) o[i] = e == i ? val : o[i];
    }
    return *reinterpret_cast<const uint4 *>(o);
}

// Which elements of a vector need the literal path (bit e): those outside the exact window (NaN and Inf included).
template <typename T, bool UNIFORM>
__device__ __forceinline__ unsigned pu_vec_mask(const uint4 raw, const PuRow &r, const PuK &K, const float2 *tab) {
    constexpr int VEC = PuIO<T>::VEC;
    float f[VEC];
    PuIO<T>::unpack(raw, f);
    unsigned m = 0;
#pragma unroll
    for (int e = 0; e < VEC; e++) m |= (!(fabsf(f[e]) <= r.xl) ? 1u : 0u) << e;
    return m;
}

// Redo of the flagged elements of one vector (the vector itself has already been stored by this thread).
template <typename T, bool UNIFORM>
__device__ __noinline__ void pu_redo_vec(const AntqCodebook *__restrict__ cb, const PuExact X, const uint4 raw, const PuRow r,
                                         const PuK K, const float2 *tab, T *og) {
    constexpr int VEC = PuIO<T>::VEC;
    float f[VEC];
    PuIO<T>::unpack(raw, f);
    unsigned m = r.ok ? pu_vec_mask<T, UNIFORM>(raw, r, K, tab) : ((1u << VEC) - 1u);
    while (m) {
        const int e = __ffs(m) - 1;
        m &= m - 1;
        float xf = f[0];
#pragma unroll
        for (int i = 1; i < VEC; i++) xf = e == i ? f[i] : xf;
        og[e] = pu_exact_elem<T>(cb, X, xf, r.s);
    }
}

// Exact redo of one chunk, dense: the flagged ELEMENTS of the whole chunk are compacted into a per-warp list and redone
// one per lane (a row with representable ties can flag a few percent of a chunk; redoing them vector by vector inside
// their owner lane serialised a warp for tens of microseconds: profiles/r02_notes.md).  redo bit j: vector j * 32 + lane.
// Only the WILD elements come here (outside the exact window, NaN, Inf, rows with a bad scale): near-midpoint elements
// are settled by pu_vec_exact in their owner lane.
template <typename T, bool UNIFORM>
__device__ __noinline__ void pu_redo_chunk(const AntqCodebook *__restrict__ cb, const PuExact X, const uint4 *sv, T *og,
                                           const int nvec, const PuRow r, const PuK K, const float2 *tab, const unsigned redo,
                                           unsigned short *redo_list, const int lane, const bool ovp) {
    typedef AntqType<T> A;
    constexpr int VEC = A::kVec;
    // bit 8 j + e: element e of vector j * 32 + lane.  OliVe pairs: a pair with a wild element is one work item, filed
    // under its even element (its partner may become a victim, or be the outlier that makes it one).
    unsigned long long em = 0;
    if (r.ok) {
        for (int j = 0; j * 32 + lane < nvec; j++)
            if ((redo >> j) & 1u) {
                unsigned m = pu_vec_mask<T, UNIFORM>(sv[j * 32 + lane], r, K, tab);
                if (ovp) m = (m | (m >> 1)) & 0x55u;
                em |= (unsigned long long)m << (8 * j);
            }
    }
    const int cnt = __popcll(em);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const bool overflow = !r.ok || total > kListMax;
    if (!overflow) {
        int pos = incl - cnt;
        while (em) {
            const int b = __ffsll((long long)em) - 1;
            em &= em - 1;
            redo_list[pos++] = (unsigned short)((lane << 6) | b);
        }
        __syncwarp();          // list complete; the fast path's vector stores are ordered before the rewrites
        for (int i = lane; i < total; i += 32) {
            const unsigned it = redo_list[i];
            const int v = (int)((it & 63u) >> 3) * 32 + (int)(it >> 6), e = (int)(it & 7u);
            const T *xe = reinterpret_cast<const T *>(sv + v);
            T *oe = og + (long long)v * VEC;
            if (ovp) pu_exact_pair<T>(cb, X, A::to_f32(xe[e]), A::to_f32(xe[e + 1]), r.s, oe[e], oe[e + 1]);
            else oe[e] = pu_exact_elem<T>(cb, X, A::to_f32(xe[e]), r.s);
        }
        __syncwarp();          // the list is rewritten by this warp's next chunk
    } else {
        for (int j = 0; j * 32 + lane < nvec; j++) {
            if ((redo >> j) & 1u) {
                const int v = j * 32 + lane;
                if (ovp) pu_redo_vec_ovp<T>(cb, X, sv[v], r.s, og + (long long)v * VEC);
                else pu_redo_vec<T, UNIFORM>(cb, X, sv[v], r, K, tab, og + (long long)v * VEC);
            }
        }
    }
}

__device__ __forceinline__ void pu_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(antq_smem_u32(bar)) : "memory");
}

// Row constants as one 16-byte word.  A bad scale (zero, Inf, NaN) is folded into the window bound (NaN: every vector
// of the row fails the window test and takes the literal path) plus one bit for the redo.
template <typename T> __device__ __forceinline__ uint4 pu_row_pack(const PuRow &r) {
    uint4 w;
    w.x = __float_as_uint(r.s);
    w.y = __float_as_uint(r.kx);
    if constexpr (sizeof(T) == 2) {
        w.z = (r.ok ? (r.xl2 & 0xffffu) : 0x7fffu) | (r.ok ? 0x10000u : 0u);
        w.w = (r.xlo2 & 0xffffu) | (r.xhi2 << 16);
    } else {
        w.z = r.ok ? __float_as_uint(r.xl) : 0x7fc00000u;
        w.w = r.ok ? 1u : 0u;
    }
    return w;
}
template <typename T> __device__ __forceinline__ PuRow pu_row_unpack(const uint4 w) {
    PuRow r;
    r.s = __uint_as_float(w.x);
    r.kx = __uint_as_float(w.y);
    if constexpr (sizeof(T) == 2) {
        r.xl2 = __byte_perm(w.z, 0, 0x1010);
        r.ok = (w.z >> 16) != 0;
        r.xlo2 = __byte_perm(w.w, 0, 0x1010);
        r.xhi2 = __byte_perm(w.w, 0, 0x3232);
        r.xl = 0.0f;                                                  // the redo path recomputes it (pu_row_xl)
    } else {
        r.xl = __uint_as_float(w.z);
        r.ok = w.w != 0;
        r.xl2 = r.xlo2 = r.xhi2 = 0;
    }
    return r;
}
__device__ __forceinline__ float pu_row_xl(const PuK &K, float s) {
    return __fmul_rn(__fmul_rn(K.lim, s), 0.9990234375f);
}

struct PuTileRows {                          // which rows a tile touches and how a vector finds its row
    unsigned row0, nrows, vbase;             // first row, number of rows, first vector of row0
};
__device__ __forceinline__ unsigned pu_row_of(const PuParams &p, unsigned v) {
    return p.cols_shift >= 0 ? v >> p.cols_shift : v / p.cols_vec;
}
__device__ __forceinline__ PuTileRows pu_tile_rows(const PuParams &p, unsigned v0, unsigned vlast) {
    PuTileRows t;
    t.row0 = 0; t.nrows = 1; t.vbase = 0;
    if (p.alpha_per_row) {
        t.row0 = pu_row_of(p, v0);
        t.nrows = pu_row_of(p, vlast) - t.row0 + 1;
        t.vbase = t.row0 * p.cols_vec;
    }
    return t;
}
// row of vector v relative to the tile's first row: (v - vbase) / cols_vec with v - vbase < kTileVec + cols_vec <= 255
// and cols_vec <= 127: the 16-bit reciprocal (p.cols_magic = 65536 / cols_vec + 1) is exact there.
__device__ __forceinline__ unsigned pu_row_local(const PuParams &p, const PuTileRows &t, unsigned v) {
    if (!p.alpha_per_row) return 0;
    const unsigned off = v - t.vbase;
    return p.cols_shift >= 0 ? off >> p.cols_shift : (off * p.cols_magic) >> 16;
}

// ==================================================================================================
// Long rows / per-tensor: persistent CTAs, TMA-staged chunks.
// SHORT: rows shorter than a chunk (scale groups, 1x1-conv weights; 2 .. 127 vectors per row).  The tensor is chunked
// flat; the rows a chunk touches (<= kRowsPerChunk) get their constants computed once, one row per lane, into a per-warp
// table, and every vector fetches its row's constants with one LDS.128.
// ==================================================================================================
constexpr int kRowsPerChunk = 132;        // 256 vectors / 2 per row + straddle, rounded up
static_assert(kRowsPerChunk * 16 <= kScratch && kListMax * 2 <= kScratch, "per-warp scratch");
template <typename T, bool UNIFORM, bool XC, bool SHORT>
__global__ void __launch_bounds__(kThreads, 1) antq_pu_stream_kernel(const PuParams p) {
    typedef AntqType<T> A;
    constexpr int VEC = A::kVec;
    extern __shared__ __align__(128) unsigned char pu_smem[];
    float2 *tab = reinterpret_cast<float2 *>(pu_smem + (size_t)kNS * kChunkMax);          // 512 entries
    float *x_thr = reinterpret_cast<float *>(tab + 512);                                  // exact thresholds / levels (redo path)
    float *x_lev = x_thr + ANTQ_MAX_GRID;
    uint64_t *full = reinterpret_cast<uint64_t *>(x_lev + ANTQ_MAX_GRID);
    unsigned *next_k = reinterpret_cast<unsigned *>(full + kNS);
    unsigned *qcount = next_k + 1;
    PuQEntry *queue = reinterpret_cast<PuQEntry *>(next_k + 4);
    unsigned char *scratch = reinterpret_cast<unsigned char *>(queue + kQCap) + (size_t)(threadIdx.x >> 5) * kScratch;
    unsigned short *redo_list = reinterpret_cast<unsigned short *>(scratch);   // long rows: the literal pass's work list
    uint4 *rows_s = reinterpret_cast<uint4 *>(scratch);                        // SHORT: the chunk's row table (no list there)

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned c_begin = blockIdx.x * p.chunks_per_cta + min(blockIdx.x, p.chunks_rem);
    const int n = (int)p.chunks_per_cta + (blockIdx.x < p.chunks_rem ? 1 : 0);
    const unsigned cpr = (unsigned)p.chunks_per_row;
    const AntqCodebook *__restrict__ cb = p.cb;

    asm volatile("griddepcontrol.launch_dependents;");                // programmatic dependent launch, as antq_stream.cu
    if (threadIdx.x < kNS) antq_mbar_init(full + threadIdx.x, 1);
    if (threadIdx.x == 0) { *next_k = 0; *qcount = 0; }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");

    struct Geo { long long base; int nvec, tail; unsigned row; float alpha, alpha1; PuTileRows tr; };
    auto geo_of = [&](int k) {
        Geo g;
        const unsigned c = c_begin + (unsigned)k;
        g.alpha = 0.0f; g.alpha1 = 0.0f;
        g.tr.row0 = 0; g.tr.nrows = 1; g.tr.vbase = 0;
        if constexpr (SHORT) {
            const unsigned cv = (unsigned)p.chunk_elems / VEC;
            const unsigned v0 = c * cv;
            const unsigned left = p.nvec - v0;
            g.nvec = (int)(left < cv ? left : cv);
            g.tail = 0;
            g.base = (long long)v0 * VEC;
            g.tr = pu_tile_rows(p, v0, v0 + (unsigned)g.nvec - 1u);
            g.row = g.tr.row0;
        } else {
            const unsigned row = p.cpr_shift >= 0 ? c >> p.cpr_shift : c / cpr;
            const long long col0 = (long long)(c - row * cpr) * p.chunk_elems;
            const long long remain = p.cols - col0;
            const int n_el = (int)(remain < p.chunk_elems ? remain : p.chunk_elems);
            g.nvec = n_el / VEC;
            g.tail = n_el - g.nvec * VEC;
            g.base = (long long)row * p.cols + col0;
            g.row = row;
        }
        return g;
    };
    auto request = [&](Geo &g, int stage) {
        if (lane == 0) {
            const unsigned bytes = (unsigned)g.nvec * 16u;
            if (bytes) {
                antq_fence_proxy_async();
                antq_bulk_g2s(pu_smem + (size_t)stage * kChunkMax, reinterpret_cast<const T *>(p.x) + g.base, bytes,
                              full + stage);
            } else {
                pu_mbar_arrive(full + stage);
            }
        }
        // the row's alpha travels with the request: its latency hides behind the bulk copy
        if constexpr (SHORT) {
            // the first two rows of this lane (64 rows per chunk: groups of 4 vectors and longer); more are loaded late
            if ((unsigned)lane < g.tr.nrows) g.alpha = __ldg(p.alpha + (p.alpha_per_row ? g.tr.row0 + lane : 0u));
            if ((unsigned)lane + 32u < g.tr.nrows) g.alpha1 = __ldg(p.alpha + g.tr.row0 + lane + 32u);
        } else {
            g.alpha = __ldg(p.alpha + (p.alpha_per_row ? g.row : 0u));
        }
    };
    auto claim = [&]() {
        int k = 0;
        if (lane == 0) k = (int)atomicAdd(next_k, 1u);
        return __shfl_sync(0xffffffffu, k, 0);
    };

    Geo cur;
    cur.base = 0; cur.nvec = 0; cur.tail = 0; cur.row = 0; cur.alpha = 0.0f; cur.alpha1 = 0.0f;
    cur.tr.row0 = 0; cur.tr.nrows = 0; cur.tr.vbase = 0;
    int k = claim();
    if (k < n) {
        cur = geo_of(k);
        request(cur, warp);
    }
    if (!UNIFORM) {
        for (int i = threadIdx.x; i < 512; i += kThreads) tab[i] = cb->pu_tab[i & 255];
    }
    PuExact X;
    X.thr = x_thr; X.lev = x_lev; X.nlev = cb->n_levels;
    X.win = (cb->flags & ANTQ_CB_WELLSEP) ? cb->lim_idx : -1.0f;
    for (int i = threadIdx.x; i < X.nlev; i += kThreads) { x_thr[i] = cb->thr[i]; x_lev[i] = cb->level[i]; }
    const PuK K = pu_load_k(cb, p.lim, p.ovp);
    __syncthreads();

    int slot = 0;
    unsigned phases = 0;
    while (k < n) {
        const int stage = warp + slot * kNC;
        const Geo g = cur;
        const int kn = claim();
        if (kn < n) {
            cur = geo_of(kn);
            request(cur, warp + (slot ^ 1) * kNC);
        }
        PuRow r;
        if constexpr (SHORT) {
            // this chunk's row table (before waiting for the data: only the alphas are needed)
            for (unsigned i = lane, a = 0; i < g.tr.nrows; i += 32, a++) {
                const float al = a == 0 ? g.alpha : a == 1 ? g.alpha1 : __ldg(p.alpha + g.tr.row0 + i);
                rows_s[i] = pu_row_pack<T>(pu_row<T>(al, p, K, true));
            }
            r.s = 1.0f; r.kx = 1.0f; r.xl = 0.0f; r.xl2 = r.xlo2 = r.xhi2 = 0; r.ok = true;
            __syncwarp();
        } else {
            r = pu_row<T>(g.alpha, p, K, false);
        }
        const unsigned gv0 = (unsigned)(g.base / VEC);                // SHORT: flat index of the chunk's first vector
        auto row_of = [&](int v) { return pu_row_unpack<T>(rows_s[pu_row_local(p, g.tr, gv0 + (unsigned)v)]); };
        const int nvec = g.nvec;
        const uint4 *sv = reinterpret_cast<const uint4 *>(pu_smem + (size_t)stage * kChunkMax);
        T *og = reinterpret_cast<T *>(p.out) + g.base;
        uint4 *ov = reinterpret_cast<uint4 *>(og);
        antq_mbar_wait(full + stage, (phases >> slot) & 1u);
        phases ^= 1u << slot;
        unsigned redo = 0, near = 0;                                  // bit j: vector j * 32 + lane is wild / near a midpoint
        if (r.ok && !(p.debug & 2)) {
            const uint4 *sp = sv + lane;
            uint4 *op = ov + lane;
            int j = 0;
#pragma unroll 1
            for (int v = lane; v + 32 < nvec; v += 64, j += 2) {
                const uint4 r0 = sp[0], r1 = sp[32];
                bool n0, n1, w0, w1;
                uint4 q0, q1;
                if constexpr (SHORT) {
                    q0 = pu_vec<T, UNIFORM, XC, true>(r0, row_of(v), K, tab, n0, w0);
                    q1 = pu_vec<T, UNIFORM, XC, true>(r1, row_of(v + 32), K, tab, n1, w1);
                } else {
                    q0 = pu_vec<T, UNIFORM, XC, true>(r0, r, K, tab, n0, w0);
                    q1 = pu_vec<T, UNIFORM, XC, true>(r1, r, K, tab, n1, w1);
                }
                antq_stg_stream(op, q0);
                antq_stg_stream(op + 32, q1);
                redo |= ((w0 ? 1u : 0u) | (w1 ? 2u : 0u)) << j;
                near |= ((n0 ? 1u : 0u) | (n1 ? 2u : 0u)) << j;
                sp += 64; op += 64;
            }
            if (j * 32 + lane < nvec) {
                bool n0, w0;
                const uint4 q0 = SHORT ? pu_vec<T, UNIFORM, XC, true>(*sp, row_of(j * 32 + lane), K, tab, n0, w0)
                                       : pu_vec<T, UNIFORM, XC, true>(*sp, r, K, tab, n0, w0);
                antq_stg_stream(op, q0);
                redo |= (w0 ? 1u : 0u) << j;
                near |= (n0 ? 1u : 0u) << j;
            }
            // Near-midpoint vectors are NOT settled here: whatever a warp does after its last chunk is paid in full by
            // the whole launch (the launch ends with its slowest warp), and the exact pass is ~0.3 us of dependent
            // code per vector and lane.  They are parked in a CTA-wide queue (input vector, destination, row constants)
            // and settled by all 512 threads at once after the last chunk.  A vector that is also wild is settled now,
            // before the literal pass rewrites its wild elements.
            if (!(p.debug & 4)) {
                while (near) {
                    const int jj = __ffs(near) - 1;
                    near &= near - 1;
                    const int v = jj * 32 + lane;
                    float vs = r.s, vkx = r.kx;
                    if constexpr (SHORT) { const PuRow rv = row_of(v); vs = rv.s; vkx = rv.kx; }
                    const unsigned slot_q = ((redo >> jj) & 1u) ? kQCap : atomicAdd(qcount, 1u);
                    if (slot_q < kQCap) {
                        PuQEntry qe;
                        qe.raw = sv[v]; qe.v = g.base / VEC + v; qe.s = vs; qe.kx = vkx;
                        queue[slot_q] = qe;
                    } else {
                        antq_stg_stream(ov + v, pu_vec_exact<T, UNIFORM>(sv[v], vs, vkx, K, tab));
                    }
                }
            }
            // OliVe: vectors holding an outlier also go to the queue while it has room (a few per chunk at OliVe's design point
            // of < 1 % outliers: settling them here would cost every chunk ~1 us of single-warp latency); when it is full --
            // heavy-tailed data -- the rest of the chunk's pairs are settled here, densely (pu_redo_chunk).
            if (p.ovp && !(p.debug & 4) && *reinterpret_cast<volatile unsigned *>(qcount) < (unsigned)kQCap) {
                unsigned left = 0;
                while (redo) {
                    const int jj = __ffs(redo) - 1;
                    redo &= redo - 1;
                    const int v = jj * 32 + lane;
                    const unsigned m = pu_vec_mask<T, UNIFORM>(sv[v], r, K, tab);
                    unsigned pm = (m | (m >> 1)) & 0x55u;             // bit 2 p: pair p holds a wild element
                    const unsigned slot_q = atomicAdd(qcount, (unsigned)__popc(pm));
                    if (slot_q + (unsigned)__popc(pm) <= (unsigned)kQCap) {
                        const T *xe = reinterpret_cast<const T *>(sv + v);
                        unsigned k = slot_q;
                        while (pm) {                                  // one entry per PAIR: the drain gives each a thread
                            const int e = __ffs(pm) - 1;
                            pm &= pm - 1;
                            PuQEntry qe;
                            qe.raw = make_uint4(__float_as_uint(A::to_f32(xe[e])), __float_as_uint(A::to_f32(xe[e + 1])), 0u, 0u);
                            qe.v = (g.base / VEC + v) * VEC + e;      // ELEMENT index of the pair's even slot
                            qe.s = r.s; qe.kx = -1.0f;                // kx < 0: the literal pair logic on the whole codebook
                            queue[k++] = qe;
                        }
                    } else {
                        for (unsigned k = slot_q; k < (unsigned)kQCap; k++) queue[k].kx = 0.0f;   // claimed, unused: no-ops
                        left |= 1u << jj;
                    }
                }
                redo = left;
            }
        } else if (p.debug & 2) {
            for (int v = lane; v < nvec; v += 32) antq_stg_stream(ov + v, sv[v]);
        } else if (p.ovp) {
            for (int v = lane; v < nvec; v += 32) pu_redo_vec_ovp<T>(cb, X, sv[v], r.s, og + (long long)v * VEC);
        } else {
            redo = 0xffffffffu;                                       // bad scale: every vector, literally
        }
        if constexpr (SHORT) {
            // wild vectors (outside the exact window, NaN, Inf, rows with a bad scale): literally, by their owner lane
            if (!(p.debug & 4)) {
                while (redo) {
                    const int jj = __ffs(redo) - 1;
                    redo &= redo - 1;
                    const int v = jj * 32 + lane;
                    PuRow rv = row_of(v);
                    if (sizeof(T) == 2) rv.xl = pu_row_xl(K, rv.s);
                    pu_redo_vec<T, UNIFORM>(cb, X, sv[v], rv, K, tab, og + (long long)v * VEC);
                }
            }
        } else {
            if (__any_sync(0xffffffffu, redo != 0) && !(p.debug & 4))
                pu_redo_chunk<T, UNIFORM>(cb, X, sv, og, nvec, r, K, tab, redo, redo_list, lane, p.ovp != 0);
        }
        if (g.tail > 0 && lane == 0) {                                 // ragged tail of a per-tensor view
            const T *xg = reinterpret_cast<const T *>(p.x) + g.base + (long long)nvec * VEC;
            for (int e = 0; e < g.tail; e++)
                og[(long long)nvec * VEC + e] = pu_exact_elem<T>(cb, X, A::to_f32(xg[e]), r.s);
        }
        __syncwarp();
        k = kn;
        slot ^= 1;
    }
    __syncthreads();
    {
        const unsigned nq = min(*qcount, (unsigned)kQCap);
        // one thread per ELEMENT (the latency of this pass is added to the launch): 2- or 4-byte stores, ordered after
        // the owners' vector stores by the barrier above
        T *oute = reinterpret_cast<T *>(p.out);
        for (unsigned i = threadIdx.x; i < nq * VEC; i += kThreads) {
            const PuQEntry *qe = queue + i / VEC;
            const unsigned e = i % VEC;
            const T *xe = reinterpret_cast<const T *>(&qe->raw);
            if (qe->kx > 0.0f)
                oute[qe->v * VEC + e] = A::from_f32_rn(pu_elem_exact<UNIFORM>(A::to_f32(xe[e]), qe->s, qe->kx, K, tab));
        }
        if (p.ovp) {                                                  // OliVe pairs holding an outlier: one thread per pair
            for (unsigned i = threadIdx.x; i < nq; i += kThreads) {
                const PuQEntry *qe = queue + i;
                if (qe->kx < 0.0f)
                    pu_exact_pair<T>(cb, X, __uint_as_float(qe->raw.x), __uint_as_float(qe->raw.y), qe->s, oute[qe->v], oute[qe->v + 1]);
            }
        }
    }
}

// ==================================================================================================
// Short rows / scale groups.  A warp owns a tile of kTileVec consecutive 16-byte vectors (2 KiB):
//   * all of the tile's loads are issued first (VPL vectors in flight per lane: a grid-stride loop with one vector
//     per lane in flight was read-latency-bound, profiles/r02_notes.md);
//   * the rows the tile touches have their constants (scale, reciprocal, window and clamp bounds: ~30 instructions
//     with one IEEE division) computed ONCE, one row per lane, and parked in shared memory as one 16-byte word;
//     each vector then fetches its row's constants with one LDS.128 instead of recomputing them.
// ==================================================================================================
constexpr int kShortThreads = 256;
constexpr int kShortWarps = kShortThreads / 32;
#ifndef ANTQ_PU_SHORT_VPL
#define ANTQ_PU_SHORT_VPL 4
#endif
#ifndef ANTQ_PU_SHORT_CTAS
#define ANTQ_PU_SHORT_CTAS 4
#endif
constexpr int kVPL = ANTQ_PU_SHORT_VPL;                     // vectors per lane per tile
constexpr int kShortCtas = ANTQ_PU_SHORT_CTAS;              // resident CTAs per SM
constexpr int kTileVec = 32 * kVPL;

template <typename T, bool UNIFORM, bool XC>
__global__ void __launch_bounds__(kShortThreads, kShortCtas) antq_pu_short_kernel(const PuParams p) {
    typedef AntqType<T> A;
    constexpr int VEC = A::kVec;
    __shared__ float2 tab[UNIFORM ? 1 : 512];
    __shared__ float x_thr[ANTQ_MAX_GRID], x_lev[ANTQ_MAX_GRID];
    __shared__ uint4 rows_s[kShortWarps][kTileVec];
    if (!UNIFORM) {
        for (int i = threadIdx.x; i < 512; i += kShortThreads) tab[i] = p.cb->pu_tab[i & 255];
    }
    PuExact X;
    X.thr = x_thr; X.lev = x_lev; X.nlev = p.cb->n_levels;
    X.win = (p.cb->flags & ANTQ_CB_WELLSEP) ? p.cb->lim_idx : -1.0f;
    for (int i = threadIdx.x; i < X.nlev; i += kShortThreads) { x_thr[i] = p.cb->thr[i]; x_lev[i] = p.cb->level[i]; }
    __syncthreads();
    const PuK K = pu_load_k(p.cb, p.lim, p.ovp);
    const uint4 *xin = reinterpret_cast<const uint4 *>(p.x);
    uint4 *xout = reinterpret_cast<uint4 *>(p.out);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint4 *rs = rows_s[warp];
    const unsigned ntiles = (p.nvec + kTileVec - 1) / kTileVec;
    const unsigned nwarps = gridDim.x * kShortWarps;
    for (unsigned t = blockIdx.x * kShortWarps + warp; t < ntiles; t += nwarps) {
        const unsigned v0 = t * kTileVec;
        const unsigned vend = min(v0 + kTileVec, p.nvec);             // exclusive
        uint4 raw[kVPL];
#pragma unroll
        for (int j = 0; j < kVPL; j++) {
            const unsigned v = v0 + j * 32 + lane;
            raw[j] = make_uint4(0, 0, 0, 0);
            if (v < vend) raw[j] = antq_ldg_stream(xin + v);
        }
        const PuTileRows tr = pu_tile_rows(p, v0, vend - 1);
        for (unsigned i = lane; i < tr.nrows; i += 32)
            rs[i] = pu_row_pack<T>(pu_row<T>(__ldg(p.alpha + (p.alpha_per_row ? tr.row0 + i : 0u)), p, K, true));
        __syncwarp();
#pragma unroll
        for (int j = 0; j < kVPL; j++) {
            const unsigned v = v0 + j * 32 + lane;
            if (v < vend) {
                PuRow r = pu_row_unpack<T>(rs[pu_row_local(p, tr, v)]);
                bool near, wild;
                uint4 q = pu_vec<T, UNIFORM, XC>(raw[j], r, K, tab, near, wild);
                if (near) q = pu_vec_fix_near<T, UNIFORM>(raw[j], q, r.s, r.kx, K, tab);
                antq_stg_stream(xout + v, q);
                if (wild) {
                    if (sizeof(T) == 2) r.xl = pu_row_xl(K, r.s);
                    pu_redo_vec<T, UNIFORM>(p.cb, X, raw[j], r, K, tab, reinterpret_cast<T *>(p.out) + (long long)v * VEC);
                }
            }
        }
        __syncwarp();                                                 // rs is rewritten by this warp's next tile
    }
}

// ==================================================================================================
// Rows of ONE or TWO vectors (group-8 / group-16 scales in 16-bit data): every vector (pair) has its own scale, so the
// per-row constants are the bottleneck -- the x-space window and clamp bounds of the kernels above cost ~30 instructions
// per row.  Here a row costs its scale (one IEEE division), a reciprocal and a validity select; the clamp and the window
// test move to t-space, where their bounds are constants of the codebook (two FMNMX per element, one FMNMX3 tree per
// vector).  Same tile structure (four loads in flight per lane), no row table.
// ==================================================================================================
// One vector of a row that has (nearly) no other vectors to share constants with: scale, reciprocal, validity; clamp and
// window in t-space (constants of the codebook); near-midpoint vectors settled by the two-candidate decision, wild ones
// (outside the window, NaN, Inf, bad scale) literally.
template <typename T, bool UNIFORM>
__device__ __forceinline__ void pu_lean_vec(const uint4 raw, const float alpha, const PuParams &p, const PuK &K, const float tlim,
                                            const float2 *tab, const PuExact &X, uint4 *dstv, T *dste) {
    constexpr int VEC = PuIO<T>::VEC;
    const float nan = __int_as_float(0x7fc00000);
    const float s = __fdiv_rn(alpha, p.gmax);
    float rs;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(s));
    float kx = __fmul_rn(rs, K.inv_c);
    kx = (kx > 0.0f && kx < __int_as_float(0x7f800000)) ? kx : nan;   // bad scale: every t is NaN -> wild
    float f[VEC], o[VEC];
    PuIO<T>::unpack(raw, f);
    float rmax = 0.0f;
    bool near = false, is_wild = false;
#pragma unroll
    for (int e = 0; e < VEC; e++) {
        const float t_ = __fmul_rn(f[e], kx);
        is_wild |= !(fabsf(t_) <= tlim);                      // outside the exact window, NaN, Inf, bad scale
        const float tc = fminf(fmaxf(t_, K.xc_lo), K.xc_hi);  // (a NaN becomes xc_lo here: that vector is wild anyway)
        float M = 12582912.0f;
        if (!UNIFORM) M = tab[__float_as_uint(tc) >> 23].x;
        const float mf = __fsub_rn(__fadd_rn(tc, M), M);
        const float rr = fabsf(__fsub_rn(tc, mf));
        if (UNIFORM) rmax = fmaxf(rmax, rr);
        else near |= rr >= tab[__float_as_uint(tc) >> 23].y;
        o[e] = __fmul_rn(__fmul_rn(mf, K.c), s);              // tc was clamped: mf is in [kmin, kmax]
    }
    if (UNIFORM) near = rmax >= K.hd_c;
    uint4 q = PuIO<T>::pack(o);
    if (near) q = pu_vec_fix_near<T, UNIFORM>(raw, q, s, kx, K, tab);   // in-window elements settled; wild ones rewritten below
    antq_stg_stream(dstv, q);
    if (is_wild) {
        const PuRow r = pu_row<T>(alpha, p, K, true);
        pu_redo_vec<T, UNIFORM>(p.cb, X, raw, r, K, tab, dste);
    }
}

template <typename T, bool UNIFORM>
__global__ void __launch_bounds__(kShortThreads, kShortCtas) antq_pu_lean_kernel(const PuParams p) {
    typedef AntqType<T> A;
    constexpr int VEC = A::kVec;
    __shared__ float2 tab[UNIFORM ? 1 : 512];
    __shared__ float x_thr[ANTQ_MAX_GRID], x_lev[ANTQ_MAX_GRID];
    if (!UNIFORM) {
        for (int i = threadIdx.x; i < 512; i += kShortThreads) tab[i] = p.cb->pu_tab[i & 255];
    }
    PuExact X;
    X.thr = x_thr; X.lev = x_lev; X.nlev = p.cb->n_levels;
    X.win = (p.cb->flags & ANTQ_CB_WELLSEP) ? p.cb->lim_idx : -1.0f;
    for (int i = threadIdx.x; i < X.nlev; i += kShortThreads) { x_thr[i] = p.cb->thr[i]; x_lev[i] = p.cb->level[i]; }
    __syncthreads();
    const PuK K = pu_load_k(p.cb, p.lim, 0);
    const float tlim = __fmul_rn(__fmul_rn(K.lim, K.inv_c), 0.9990234375f);   // |t| <= tlim: inside the exact window
    const uint4 *xin = reinterpret_cast<const uint4 *>(p.x);
    uint4 *xout = reinterpret_cast<uint4 *>(p.out);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned ntiles = (p.nvec + kTileVec - 1) / kTileVec;
    const unsigned nwarps = gridDim.x * kShortWarps;
    for (unsigned t = blockIdx.x * kShortWarps + warp; t < ntiles; t += nwarps) {
        const unsigned v0 = t * kTileVec;
        const unsigned vend = min(v0 + kTileVec, p.nvec);
        uint4 raw[kVPL];
        float al[kVPL];
#pragma unroll
        for (int j = 0; j < kVPL; j++) {
            const unsigned v = v0 + j * 32 + lane;
            raw[j] = make_uint4(0, 0, 0, 0);
            al[j] = 0.0f;
            if (v < vend) {
                raw[j] = antq_ldg_stream(xin + v);
                al[j] = __ldg(p.alpha + (p.alpha_per_row ? v >> p.cols_shift : 0u));
            }
        }
#pragma unroll
        for (int j = 0; j < kVPL; j++) {
            const unsigned v = v0 + j * 32 + lane;
            if (v >= vend) break;
            pu_lean_vec<T, UNIFORM>(raw[j], al[j], p, K, tlim, tab, X, xout + v, reinterpret_cast<T *>(p.out) + (long long)v * VEC);
        }
    }
}

template <typename T, bool UNIFORM> int launch_lean(const PuParams &p, cudaStream_t st) {
    const long long tiles = ((long long)p.nvec + kTileVec - 1) / kTileVec;
    const long long want = (tiles + kShortWarps - 1) / kShortWarps;
    const long long cap = (long long)antq_num_sms() * kShortCtas;
    antq_pu_lean_kernel<T, UNIFORM><<<(int)(want < cap ? want : cap), kShortThreads, 0, st>>>(p);
    return (int)cudaGetLastError();
}

// ==================================================================================================
// Dynamic scale groups: alpha = max|x| over the group * ratio, computed in the same pass (ONE read of x).
// A group of cols = L * VEC elements is held by L adjacent lanes (L a power of two <= 32): local abs-max, xor-shuffle
// reduction (integer max on the fp32 bit patterns of |x|: NaN-propagating, like torch's abs().max()), then the closed form.
// ==================================================================================================
template <typename T, bool UNIFORM, bool XC>
__global__ void __launch_bounds__(kShortThreads, kShortCtas) antq_pu_dynamic_kernel(const PuParams p, float ratio, float *__restrict__ alpha_out) {
    typedef AntqType<T> A;
    constexpr int VEC = A::kVec;
    __shared__ float2 tab[UNIFORM ? 1 : 512];
    __shared__ float x_thr[ANTQ_MAX_GRID], x_lev[ANTQ_MAX_GRID];
    __shared__ uint4 rows_s[kShortWarps][kTileVec];
    if (!UNIFORM) {
        for (int i = threadIdx.x; i < 512; i += kShortThreads) tab[i] = p.cb->pu_tab[i & 255];
    }
    PuExact X;
    X.thr = x_thr; X.lev = x_lev; X.nlev = p.cb->n_levels;
    X.win = (p.cb->flags & ANTQ_CB_WELLSEP) ? p.cb->lim_idx : -1.0f;
    for (int i = threadIdx.x; i < X.nlev; i += kShortThreads) { x_thr[i] = p.cb->thr[i]; x_lev[i] = p.cb->level[i]; }
    __syncthreads();
    const PuK K = pu_load_k(p.cb, p.lim, p.ovp);
    const float tlim = __fmul_rn(__fmul_rn(K.lim, K.inv_c), 0.9990234375f);
    const uint4 *xin = reinterpret_cast<const uint4 *>(p.x);
    uint4 *xout = reinterpret_cast<uint4 *>(p.out);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint4 *rs = rows_s[warp];
    const unsigned L = p.cols_vec;                                    // lanes per group (a power of two <= 32)
    const int sh = p.cols_shift;
    const unsigned ntiles = (p.nvec + kTileVec - 1) / kTileVec;
    const unsigned nwarps = gridDim.x * kShortWarps;
    // Same tile structure as antq_pu_short_kernel: the tile's loads first, then one abs-max per group (xor shuffles over the
    // group's L lanes), then the groups' constants ONCE, one group per lane (its alpha fetched from the lane that holds the
    // group's first vector), parked in shared memory for the group's vectors.
    for (unsigned t = blockIdx.x * kShortWarps + warp; t < ntiles; t += nwarps) {
        const unsigned v0 = t * kTileVec;
        const unsigned vend = min(v0 + kTileVec, p.nvec);             // exclusive; a multiple of L (whole groups)
        uint4 raw[kVPL];
        float am[kVPL];
#pragma unroll
        for (int j = 0; j < kVPL; j++) {
            const unsigned v = v0 + j * 32 + lane;
            raw[j] = make_uint4(0, 0, 0, 0);
            if (v < vend) raw[j] = antq_ldg_stream(xin + v);
        }
#pragma unroll
        for (int j = 0; j < kVPL; j++) {
            T xv[VEC];
            *reinterpret_cast<uint4 *>(xv) = raw[j];
            unsigned m = 0;
#pragma unroll
            for (int e = 0; e < VEC; e++) {
                const unsigned b = __float_as_uint(A::to_f32(xv[e])) & 0x7fffffffu;
                m = b > m ? b : m;
            }
            for (unsigned o = 1; o < L; o <<= 1) {
                const unsigned u = __shfl_xor_sync(0xffffffffu, m, o);
                m = u > m ? u : m;
            }
            am[j] = __fmul_rn(__uint_as_float(m), ratio);             // alpha = absmax * ratio (fp32, like the torch expression)
        }
        if (L <= 2u) {
            // one or two vectors per group: no row table, the lean arithmetic (see antq_pu_lean_kernel)
#pragma unroll
            for (int j = 0; j < kVPL; j++) {
                const unsigned v = v0 + j * 32 + lane;
                if (v < vend) {
                    if (alpha_out && (v & (L - 1u)) == 0) alpha_out[v >> sh] = am[j];
                    pu_lean_vec<T, UNIFORM>(raw[j], am[j], p, K, tlim, tab, X, xout + v, reinterpret_cast<T *>(p.out) + (long long)v * VEC);
                }
            }
            continue;
        }
        const unsigned row0 = v0 >> sh, nrows = (vend - v0) >> sh;    // groups of this tile: <= kTileVec / L
        for (unsigned k = 0; k * 32u < nrows; k++) {
            const unsigned r = k * 32u + lane;                        // group r of the tile starts at vector r * L
            const unsigned src = (r << sh) & 31u, jsrc = (r << sh) >> 5;
            float al = 0.0f;
#pragma unroll
            for (int j = 0; j < kVPL; j++) {
                const float g = __shfl_sync(0xffffffffu, am[j], src);
                al = jsrc == (unsigned)j ? g : al;
            }
            if (r < nrows) {
                rs[r] = pu_row_pack<T>(pu_row<T>(al, p, K, true));
                if (alpha_out) alpha_out[row0 + r] = al;
            }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < kVPL; j++) {
            const unsigned v = v0 + j * 32 + lane;
            if (v < vend) {
                PuRow r = pu_row_unpack<T>(rs[(v - v0) >> sh]);
                bool near, wild;
                uint4 q = pu_vec<T, UNIFORM, XC>(raw[j], r, K, tab, near, wild);
                if (near) q = pu_vec_fix_near<T, UNIFORM>(raw[j], q, r.s, r.kx, K, tab);
                antq_stg_stream(xout + v, q);
                if (wild) {
                    if (sizeof(T) == 2) r.xl = pu_row_xl(K, r.s);
                    pu_redo_vec<T, UNIFORM>(p.cb, X, raw[j], r, K, tab, reinterpret_cast<T *>(p.out) + (long long)v * VEC);
                }
            }
        }
        __syncwarp();                                                 // rs is rewritten by this warp's next tile
    }
}

// ==================================================================================================
// Packed 4-bit codes by the closed form (antq_encode_p4's fast path: piecewise-uniform grids of <= 16 entries whose levels
// are exact in e4m3, no pairs).  Tile structure of antq_pu_short_kernel; the level k of an element becomes its code
// through a 256-entry table indexed by k's e4m3 byte (4-bit grids have at most 4 significant bits).  Near-midpoint
// elements: the two-candidate decision; wild elements (outside the exact window, NaN, Inf, bad scale): the literal scan,
// which is also where decoding may fail to reproduce the fake-quant value (n_inexact).
// ==================================================================================================
__device__ __forceinline__ unsigned pu_e4m3x2(float lo, float hi) {                      // two floats -> two e4m3 bytes
    unsigned short r;
    asm("{ cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2; }" : "=h"(r) : "f"(hi), "f"(lo));
    return r;
}
template <typename T, bool UNIFORM>
__global__ void __launch_bounds__(kShortThreads, 3) antq_pu_encode_kernel(const PuParams p, unsigned char *__restrict__ codes,
                                                                           unsigned int *__restrict__ n_inexact) {
    typedef AntqType<T> A;
    constexpr int VEC = A::kVec;
    __shared__ float2 tab[UNIFORM ? 1 : 512];
    __shared__ unsigned char lut[256];
    const AntqCodebook *__restrict__ cb = p.cb;
    if (!UNIFORM) {
        for (int i = threadIdx.x; i < 512; i += kShortThreads) tab[i] = cb->pu_tab[i & 255];
    }
    for (int i = threadIdx.x; i < 256; i += kShortThreads) lut[i] = 0;
    __syncthreads();
    if ((int)threadIdx.x < cb->n_levels) {
        const float k = __fdiv_rn(cb->level[threadIdx.x], cb->pu_c);
        lut[pu_e4m3x2(k, 0.0f) & 0xffu] = (unsigned char)cb->level_code[threadIdx.x];
    }
    __syncthreads();
    const PuK K = pu_load_k(cb, p.lim, 0);
    const uint4 *xin = reinterpret_cast<const uint4 *>(p.x);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned ntiles = (p.nvec + kTileVec - 1) / kTileVec;
    const unsigned nwarps = gridDim.x * kShortWarps;
    unsigned bad = 0;
    for (unsigned t = blockIdx.x * kShortWarps + warp; t < ntiles; t += nwarps) {
        const unsigned v0 = t * kTileVec;
        const unsigned vend = min(v0 + kTileVec, p.nvec);
        uint4 raw[kVPL];
#pragma unroll
        for (int j = 0; j < kVPL; j++) {
            const unsigned v = v0 + j * 32 + lane;
            raw[j] = make_uint4(0, 0, 0, 0);
            if (v < vend) raw[j] = antq_ldg_stream(xin + v);
        }
        unsigned prev_row = 0xffffffffu;
        PuRow r;
        r.s = 1.0f; r.kx = 1.0f; r.xl = 0.0f; r.xl2 = r.xlo2 = r.xhi2 = 0; r.ok = false;
#pragma unroll
        for (int j = 0; j < kVPL; j++) {
            const unsigned v = v0 + j * 32 + lane;
            if (v >= vend) break;
            const unsigned row = p.alpha_per_row ? pu_row_of(p, v) : 0u;
            if (row != prev_row) { r = pu_row<T>(__ldg(p.alpha + row), p, K, true); prev_row = row; }
            float f[VEC], kk[VEC];
            PuIO<T>::unpack(raw[j], f);
            unsigned wild = r.ok ? 0u : (1u << VEC) - 1u;
#pragma unroll
            for (int e = 0; e < VEC; e++) {
                bool near = false;
                const float t_ = __fmul_rn(f[e], r.kx);
                float M = 12582912.0f, hd = __fmaf_rn(fabsf(t_), -1.9073486328125e-06f, 0.5f);
                if (!UNIFORM) { const float2 md = tab[__float_as_uint(t_) >> 23]; M = md.x; hd = md.y; }
                const float mf = __fsub_rn(__fadd_rn(t_, M), M);
                near = fabsf(__fsub_rn(t_, mf)) >= hd;
                kk[e] = fminf(fmaxf(mf, K.kmin), K.kmax);
                if (!(fabsf(f[e]) <= r.xl)) wild |= 1u << e;          // outside the exact window, NaN, Inf
                else if (near) { float q, d; kk[e] = pu_elem_exact_k<UNIFORM>(f[e], r.s, r.kx, K, tab, q, d); }
            }
            unsigned packed = 0;
#pragma unroll
            for (int e = 0; e < VEC; e += 2) {
                const unsigned b2 = pu_e4m3x2(kk[e], kk[e + 1]);
                packed |= ((unsigned)lut[b2 & 0xffu] | ((unsigned)lut[b2 >> 8] << 4)) << (4 * e);
            }
            while (wild) {                                            // literally: the scan, and the check the codes can be trusted
                const int e = __ffs(wild) - 1;
                wild &= wild - 1;
                float xf = f[0];
#pragma unroll
                for (int i = 1; i < VEC; i++) xf = e == i ? f[i] : xf;
                const float d = __fdiv_rn(xf, r.s);
                int code;
                const float q = antq_scan_literal(cb->grid, cb->n_entries, d, code);
                const T ref = A::from_f32_rn(antq_ste_rescale(q, d, r.s));
                const T dec = A::from_f32_rn(__fmul_rn(code >= 0 ? q : 0.0f, r.s));
                if (A::bits(ref) != A::bits(dec) || code < 0 || code > 15) bad++;
                packed = (packed & ~(0xfu << (4 * e))) | ((unsigned)(code & 15) << (4 * e));
            }
            if (VEC == 8) reinterpret_cast<unsigned *>(codes)[v] = packed;
            else reinterpret_cast<unsigned short *>(codes)[v] = (unsigned short)packed;
        }
    }
    if (bad && n_inexact) atomicAdd(n_inexact, bad);
}

template <typename T, bool UNIFORM, bool XC, bool SHORT> int launch_stream(const PuParams &p, int ctas, cudaStream_t st) {
    auto kernel = antq_pu_stream_kernel<T, UNIFORM, XC, SHORT>;
    const int smem = kNS * kChunkMax + 512 * 8 + 2 * ANTQ_MAX_GRID * 4 + kNS * 8 + 16 + kQCap * (int)sizeof(PuQEntry) + kNC * kScratch;
    static unsigned long long configured = 0ull;                     // one bit per device ordinal
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 64 || !((configured >> dev) & 1ull)) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        if (dev < 64) configured |= 1ull << dev;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kernel, p);
    if (e == cudaSuccess) e = cudaGetLastError();
    return (int)e;
}

template <typename T, bool UNIFORM, bool XC> int launch_short(const PuParams &p, cudaStream_t st) {
    const long long tiles = ((long long)p.nvec + kTileVec - 1) / kTileVec;
    const long long want = (tiles + kShortWarps - 1) / kShortWarps;
    const long long cap = (long long)antq_num_sms() * kShortCtas;    // resident CTAs per SM (__launch_bounds__)
    antq_pu_short_kernel<T, UNIFORM, XC><<<(int)(want < cap ? want : cap), kShortThreads, 0, st>>>(p);
    return (int)cudaGetLastError();
}

}  // namespace



// Returns ANTQ_ENOTSUP for shapes the persistent kernel does not cover (more than 2^31 chunks).
int antq_launch_pu_stream(const void *x, void *out, const float *alpha, int alpha_per_row, long long rows, long long cols,
                          int dtype, const AntqCodebook *cb, const antq_codebook_info *info, bool ovp, cudaStream_t st) {
    const int es = dtype == ANTQ_F32 ? 4 : 2;
    static int dbg = -1;
    if (dbg < 0) { const char *e = getenv("ANTQ_DEBUG"); dbg = e ? atoi(e) : 0; }
    PuParams p = {};
    p.x = x; p.out = out; p.alpha = alpha; p.cb = cb;
    p.rows = rows; p.cols = cols;
    int chunk_bytes = kChunkMax;
    const long long want = (long long)antq_num_sms() * kNC * 2;
    while (chunk_bytes > 1024) {
        const long long ce = chunk_bytes / es;
        if (rows * ((cols + ce - 1) / ce) >= want) break;
        chunk_bytes >>= 1;
    }
    p.chunk_elems = chunk_bytes / es;
    const long long cpr = (cols + p.chunk_elems - 1) / p.chunk_elems;
    if (cpr > 0x7fffffffLL) return ANTQ_ENOTSUP;
    p.chunks_per_row = (int)cpr;
    p.cpr_shift = -1;
    if ((cpr & (cpr - 1)) == 0) {
        int sh = 0;
        while ((1LL << sh) < cpr) sh++;
        p.cpr_shift = sh;
    }
    const long long total = rows * cpr;
    if (total == 0) return 0;
    if (total > 0x7fffffffLL) return ANTQ_ENOTSUP;
    p.total_chunks = (unsigned)total;
    p.alpha_per_row = alpha_per_row;
    p.gmax = info->gmax; p.lim = info->lim;
    p.debug = dbg;
    p.ovp = ovp ? 1 : 0;
    const unsigned sms = (unsigned)antq_num_sms();
    const int ctas = (int)(p.total_chunks < sms ? p.total_chunks : sms);
    p.chunks_per_cta = p.total_chunks / (unsigned)ctas;
    p.chunks_rem = p.total_chunks % (unsigned)ctas;
    const bool uni = (info->flags & ANTQ_CB_PU_UNIFORM) != 0;
    const bool xc = dtype == ANTQ_F16 ? (info->flags & ANTQ_CB_PU_XC16) != 0 : dtype == ANTQ_BF16 ? (info->flags & ANTQ_CB_PU_XCBF) != 0 : false;
#define ANTQ_PU_GO(T, X) (uni ? launch_stream<T, true, X, false>(p, ctas, st) : launch_stream<T, false, X, false>(p, ctas, st))
    switch (dtype) {
        case ANTQ_F32: return ANTQ_PU_GO(float, false);
        case ANTQ_F16: return xc ? ANTQ_PU_GO(__half, true) : ANTQ_PU_GO(__half, false);
        case ANTQ_BF16: return xc ? ANTQ_PU_GO(__nv_bfloat16, true) : ANTQ_PU_GO(__nv_bfloat16, false);
    }
#undef ANTQ_PU_GO
    return ANTQ_EINVAL;
}

int antq_launch_pu_short(const void *x, void *out, const float *alpha, int alpha_per_row, long long rows, long long cols,
                         int dtype, const AntqCodebook *cb, const antq_codebook_info *info, cudaStream_t st) {
    const int es = dtype == ANTQ_F32 ? 4 : 2;
    const int vec = 16 / es;
    const long long n = rows * cols;
    if (n == 0) return 0;
    if (cols % vec || (n / vec) > 0x7fffffffLL) return ANTQ_ENOTSUP;
    static int dbg = -1;
    if (dbg < 0) { const char *e = getenv("ANTQ_DEBUG"); dbg = e ? atoi(e) : 0; }
    PuParams p = {};
    p.x = x; p.out = out; p.alpha = alpha; p.cb = cb;
    p.rows = rows; p.cols = cols;
    p.nvec = (unsigned)(n / vec);
    p.cols_vec = (unsigned)(cols / vec);
    if (p.cols_vec > 127u && alpha_per_row && rows > 1) return ANTQ_ENOTSUP;   // pu_row_local's 16-bit reciprocal
    if (!alpha_per_row || rows == 1) { alpha_per_row = 0; p.cols_vec = 1; }    // one scale: the row of a vector is never asked
    p.cols_magic = 65536u / p.cols_vec + 1u;
    p.cols_shift = -1;
    if ((p.cols_vec & (p.cols_vec - 1)) == 0) {
        int sh = 0;
        while ((1u << sh) < p.cols_vec) sh++;
        p.cols_shift = sh;
    }
    p.alpha_per_row = alpha_per_row;
    p.gmax = info->gmax; p.lim = info->lim;
    p.debug = dbg;
    const bool uni = (info->flags & ANTQ_CB_PU_UNIFORM) != 0;
    const bool xc = dtype == ANTQ_F16 ? (info->flags & ANTQ_CB_PU_XC16) != 0 : dtype == ANTQ_BF16 ? (info->flags & ANTQ_CB_PU_XCBF) != 0 : false;
    // Uniform grids (int-k), rows of 16 vectors and more (group-128 fp16 ...): the persistent TMA-staged kernel over the flat
    // tensor (SHORT mode: 16.4-16.8 us per 4096^2 fp16 against 17.2-19 for the tile kernel).  Shorter rows, and the
    // per-octave-table grids at any length (20.1 against 18.4-19): the tile kernel -- one row-table entry per few vectors is
    // too much bookkeeping for the persistent kernel's single CTA per SM.
    if (p.cols_vec >= 16 && uni) {
        int chunk_bytes = kChunkMax;
        const long long want = (long long)antq_num_sms() * kNC * 2;
        while (chunk_bytes > 1024 && ((long long)p.nvec * 16 + chunk_bytes - 1) / chunk_bytes < want) chunk_bytes >>= 1;
        p.chunk_elems = chunk_bytes / es;
        p.chunks_per_row = 1; p.cpr_shift = 0;
        p.total_chunks = (unsigned)(((long long)p.nvec * 16 + chunk_bytes - 1) / chunk_bytes);
        const unsigned sms = (unsigned)antq_num_sms();
        const int ctas = (int)(p.total_chunks < sms ? p.total_chunks : sms);
        p.chunks_per_cta = p.total_chunks / (unsigned)ctas;
        p.chunks_rem = p.total_chunks % (unsigned)ctas;
#define ANTQ_PU_GO(T, X) (uni ? launch_stream<T, true, X, true>(p, ctas, st) : launch_stream<T, false, X, true>(p, ctas, st))
        switch (dtype) {
            case ANTQ_F32: return ANTQ_PU_GO(float, false);
            case ANTQ_F16: return xc ? ANTQ_PU_GO(__half, true) : ANTQ_PU_GO(__half, false);
            case ANTQ_BF16: return xc ? ANTQ_PU_GO(__nv_bfloat16, true) : ANTQ_PU_GO(__nv_bfloat16, false);
        }
#undef ANTQ_PU_GO
        return ANTQ_EINVAL;
    }
    if (p.alpha_per_row && p.cols_vec <= 2 && !(dbg & 16)) {          // one or two vectors per row: the lean kernel (at four: 18.8 vs 17.5 us)
        switch (dtype) {
            case ANTQ_F32: return uni ? launch_lean<float, true>(p, st) : launch_lean<float, false>(p, st);
            case ANTQ_F16: return uni ? launch_lean<__half, true>(p, st) : launch_lean<__half, false>(p, st);
            case ANTQ_BF16: return uni ? launch_lean<__nv_bfloat16, true>(p, st) : launch_lean<__nv_bfloat16, false>(p, st);
        }
        return ANTQ_EINVAL;
    }
#define ANTQ_PU_GO(T, X) (uni ? launch_short<T, true, X>(p, st) : launch_short<T, false, X>(p, st))
    switch (dtype) {
        case ANTQ_F32: return ANTQ_PU_GO(float, false);
        case ANTQ_F16: return xc ? ANTQ_PU_GO(__half, true) : ANTQ_PU_GO(__half, false);
        case ANTQ_BF16: return xc ? ANTQ_PU_GO(__nv_bfloat16, true) : ANTQ_PU_GO(__nv_bfloat16, false);
    }
#undef ANTQ_PU_GO
    return ANTQ_EINVAL;
}

// Dynamic group scales (see antq_pu_dynamic_kernel).  ENOTSUP: grids that are not piecewise uniform, groups that do not fit
// one warp (more than 32 x 16 bytes) or whose vector count is not a power of two.
int antq_launch_pu_dynamic(const void *x, void *out, float *alpha_out, float ratio, long long rows, long long cols, int dtype,
                           const AntqCodebook *cb, const antq_codebook_info *info, cudaStream_t st) {
    const int es = dtype == ANTQ_F32 ? 4 : 2;
    const int vec = 16 / es;
    const long long n = rows * cols;
    if (n == 0) return 0;
    if (!(info->flags & ANTQ_CB_PU) || !(info->flags & ANTQ_CB_WELLSEP) || !(info->flags & ANTQ_CB_STE_EXACT)) return ANTQ_ENOTSUP;
    if (cols % vec || (n / vec) > 0x7ffffff0LL) return ANTQ_ENOTSUP;
    const long long cv = cols / vec;
    if (cv > 32 || (cv & (cv - 1))) return ANTQ_ENOTSUP;
    PuParams p = {};
    p.x = x; p.out = out; p.alpha = nullptr; p.cb = cb;
    p.rows = rows; p.cols = cols;
    p.nvec = (unsigned)(n / vec);
    p.cols_vec = (unsigned)cv;
    int sh = 0;
    while ((1u << sh) < p.cols_vec) sh++;
    p.cols_shift = sh;
    p.alpha_per_row = 1;
    p.gmax = info->gmax; p.lim = info->lim;
    const bool uni = (info->flags & ANTQ_CB_PU_UNIFORM) != 0;
    const long long tiles = ((long long)p.nvec + kTileVec - 1) / kTileVec;
    const long long want = (tiles + kShortWarps - 1) / kShortWarps;
    const long long cap = (long long)antq_num_sms() * kShortCtas;
    const int ctas = (int)(want < cap ? want : cap);
    const bool xc = dtype == ANTQ_F16 ? (info->flags & ANTQ_CB_PU_XC16) != 0 : dtype == ANTQ_BF16 ? (info->flags & ANTQ_CB_PU_XCBF) != 0 : false;
#define ANTQ_PU_GO(T, X)                                                                                      \
    do {                                                                                                      \
        if (uni) antq_pu_dynamic_kernel<T, true, X><<<ctas, kShortThreads, 0, st>>>(p, ratio, alpha_out);     \
        else antq_pu_dynamic_kernel<T, false, X><<<ctas, kShortThreads, 0, st>>>(p, ratio, alpha_out);        \
    } while (0)
    switch (dtype) {
        case ANTQ_F32: ANTQ_PU_GO(float, false); break;
        case ANTQ_F16: if (xc) ANTQ_PU_GO(__half, true); else ANTQ_PU_GO(__half, false); break;
        case ANTQ_BF16: if (xc) ANTQ_PU_GO(__nv_bfloat16, true); else ANTQ_PU_GO(__nv_bfloat16, false); break;
        default: return ANTQ_EINVAL;
    }
#undef ANTQ_PU_GO
    return (int)cudaGetLastError();
}

// antq_encode_p4's fast path.  ENOTSUP: the caller falls back to the literal encoder (antq_codes.cu).
int antq_launch_pu_encode(const void *x, unsigned char *codes, const float *alpha, int alpha_per_row, long long rows, long long cols,
                          int dtype, const AntqCodebook *cb, const antq_codebook_info *info, unsigned int *n_inexact, cudaStream_t st) {
    const int es = dtype == ANTQ_F32 ? 4 : 2;
    const int vec = 16 / es;
    const long long n = rows * cols;
    const int need = ANTQ_CB_PU | ANTQ_CB_PU_E4M3 | ANTQ_CB_WELLSEP | ANTQ_CB_STE_EXACT;
    if (!info || (info->flags & need) != need || info->n_entries > 16) return ANTQ_ENOTSUP;
    if (n == 0 || cols % vec || (n / vec) > 0x7fffffffLL || (uintptr_t)x % 16 || (uintptr_t)codes % 4) return ANTQ_ENOTSUP;
    PuParams p = {};
    p.x = x; p.alpha = alpha; p.cb = cb;
    p.rows = rows; p.cols = cols;
    p.nvec = (unsigned)(n / vec);
    p.cols_vec = (unsigned)(cols / vec);
    p.cols_shift = -1;
    if ((p.cols_vec & (p.cols_vec - 1)) == 0) {
        int sh = 0;
        while ((1u << sh) < p.cols_vec) sh++;
        p.cols_shift = sh;
    }
    p.alpha_per_row = (alpha_per_row && rows > 1) ? 1 : 0;
    p.gmax = info->gmax; p.lim = info->lim;
    const long long tiles = ((long long)p.nvec + kTileVec - 1) / kTileVec;
    const long long want = (tiles + kShortWarps - 1) / kShortWarps;
    const long long cap = (long long)antq_num_sms() * 3;
    const int ctas = (int)(want < cap ? want : cap);
    const bool uni = (info->flags & ANTQ_CB_PU_UNIFORM) != 0;
#define ANTQ_PU_GO(T)                                                                                       \
    do {                                                                                                    \
        if (uni) antq_pu_encode_kernel<T, true><<<ctas, kShortThreads, 0, st>>>(p, codes, n_inexact);       \
        else antq_pu_encode_kernel<T, false><<<ctas, kShortThreads, 0, st>>>(p, codes, n_inexact);          \
    } while (0)
    switch (dtype) {
        case ANTQ_F32: ANTQ_PU_GO(float); break;
        case ANTQ_F16: ANTQ_PU_GO(__half); break;
        case ANTQ_BF16: ANTQ_PU_GO(__nv_bfloat16); break;
        default: return ANTQ_EINVAL;
    }
#undef ANTQ_PU_GO
    return (int)cudaGetLastError();
}
