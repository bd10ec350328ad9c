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
// t is a few ulps off d / c, so an element whose t lies within delta_e = 2^(e - 19) of a midpoint -- or outside the
// exact window, NaN, Inf, or in a row whose scale is not a positive finite number -- is redone with the literal
// arithmetic (true division, the codebook's exact thresholds / the literal scan).  With 16-bit data that is ~1e-4 of
// the elements, except in rows where a tie x / s == midpoint is representable.
//
// Two execution shapes:
//   antq_pu_stream_kernel   rows >= 512 elements / per-tensor: the persistent pipeline of antq_stream.cu (one CTA per
//                           SM, 12 consumer warps, two private 4 KiB TMA stages each, chunk counter) -- minus the row
//                           tables, the builder warps and the prologue: a row needs three scalars.
//   antq_pu_short_kernel    shorter rows and scale groups (group-8/16/32, 1x1-conv weights): grid-stride over 16-byte
//                           vectors, one IEEE division per VECTOR (for s) instead of one per element.
// Bound: HBM in principle; ~17 fp32 / integer instructions per element keep the SM's issue slots ~85 % busy at the
// HBM rate (measured: profiles/r02_notes.md).
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
constexpr int kListMax = 512;             // per-warp work list of elements to redo exactly (per chunk)

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
    // short kernel
    unsigned nvec, cols_vec;
    int cols_shift;
};

struct PuK {                               // the codebook's closed-form constants (AntqCodebook::pu_*)
    float c, inv_c, kmin, kmax;
    float xc_lo, xc_hi;                    // x-space clamp: (kmin - 0.4 step_top) and (kmax + 0.4 step_top), in units of c
};
__device__ __forceinline__ PuK pu_load_k(const AntqCodebook *__restrict__ cb) {
    PuK k;
    k.c = cb->pu_c; k.inv_c = cb->pu_inv_c; k.kmin = cb->pu_kmin; k.kmax = cb->pu_kmax;
    // step of the octave each end of the grid lies in, from that octave's magic constant M = 1.5 * 2^23 * step
    // (kmin = 0: exponent field 0 -> the sub-unit region's entry)
    const float step_hi = __fmul_rn(cb->pu_tab[__float_as_uint(k.kmax) >> 23].x, 7.94728597e-08f);           // 1 / (1.5 * 2^23)
    const float step_lo = __fmul_rn(cb->pu_tab[(__float_as_uint(k.kmin) & 0x7fffffffu) >> 23].x, 7.94728597e-08f);
    k.xc_hi = __fadd_rn(k.kmax, __fmul_rn(0.4f, step_hi));
    k.xc_lo = __fsub_rn(k.kmin, __fmul_rn(0.4f, step_lo));
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
    r.xl = __fmul_rn(__fmul_rn(p.lim, r.s), 0.9990234375f);           // conservative exact window in x-space
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
template <bool UNIFORM, bool CLAMP>
__device__ __forceinline__ float pu_quant(float xf, const PuRow &r, const PuK &K, const float2 *tab, bool &flag) {
    const float t = __fmul_rn(xf, r.kx);
    float M, hd;
    if (UNIFORM) {
        M = 12582912.0f;                                              // 1.5 * 2^23: step 1 everywhere
        hd = __fmaf_rn(fabsf(t), -1.9073486328125e-06f, 0.5f);        // 0.5 - |t| 2^-19
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

// One 16-byte vector through the closed form; `flag` = some element needs the exact redo.
// XC (16-bit types): clamp the INPUT pairs with two packed min / max instead of every t with two FMNMX.
template <typename T, bool UNIFORM, bool XC>
__device__ __forceinline__ uint4 pu_vec(const uint4 raw, const PuRow &r, const PuK &K, const float2 *tab, bool &flag) {
    constexpr int VEC = PuIO<T>::VEC;
    float f[VEC], o[VEC];
    bool fl = false;
    if constexpr (sizeof(T) == 2) {
        typedef typename PuPack<T>::v2 v2;
        // one packed NaN-propagating max of |x| per vector against the exact window
        const v2 a = __hmax2_nan(__habs2(PuPack<T>::from_u32(raw.x)), __habs2(PuPack<T>::from_u32(raw.y)));
        const v2 b = __hmax2_nan(__habs2(PuPack<T>::from_u32(raw.z)), __habs2(PuPack<T>::from_u32(raw.w)));
        fl |= __hle2_mask(__hmax2_nan(a, b), PuPack<T>::from_u32(r.xl2)) != 0xffffffffu;
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
        o[e] = pu_quant<UNIFORM, !(XC && sizeof(T) == 2)>(f[e], r, K, tab, fl);
        if (sizeof(T) == 4) fl |= !(fabsf(f[e]) <= r.xl);             // outside the exact window, NaN, Inf
    }
    flag = fl;
    return PuIO<T>::pack(o);
}

// Which elements of a vector need the exact redo (bit e), recomputed with the closed form's own tests -- unrolled, in
// registers (an indexed loop over f[] would put the vector in local memory and make the redo latency-bound).
template <typename T, bool UNIFORM>
__device__ __forceinline__ unsigned pu_vec_mask(const uint4 raw, const PuRow &r, const PuK &K, const float2 *tab) {
    constexpr int VEC = PuIO<T>::VEC;
    float f[VEC];
    PuIO<T>::unpack(raw, f);
    unsigned m = 0;
#pragma unroll
    for (int e = 0; e < VEC; e++) {
        bool fl = false;
        (void)pu_quant<UNIFORM, true>(f[e], r, K, tab, fl);
        fl |= !(fabsf(f[e]) <= r.xl);
        m |= (fl ? 1u : 0u) << e;
    }
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

__device__ __forceinline__ void pu_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(antq_smem_u32(bar)) : "memory");
}

// ==================================================================================================
// Long rows / per-tensor: persistent CTAs, TMA-staged chunks.
// ==================================================================================================
template <typename T, bool UNIFORM, bool XC>
__global__ void __launch_bounds__(kThreads, 1) antq_pu_stream_kernel(const PuParams p) {
    typedef AntqType<T> A;
    constexpr int VEC = A::kVec;
    extern __shared__ __align__(128) unsigned char pu_smem[];
    float2 *tab = reinterpret_cast<float2 *>(pu_smem + (size_t)kNS * kChunkMax);          // 512 entries
    float *x_thr = reinterpret_cast<float *>(tab + 512);                                  // exact thresholds / levels (redo path)
    float *x_lev = x_thr + ANTQ_MAX_GRID;
    uint64_t *full = reinterpret_cast<uint64_t *>(x_lev + ANTQ_MAX_GRID);
    unsigned *next_k = reinterpret_cast<unsigned *>(full + kNS);
    unsigned short *redo_list = reinterpret_cast<unsigned short *>(next_k + 4) + (size_t)(threadIdx.x >> 5) * kListMax;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned c_begin = blockIdx.x * p.chunks_per_cta + min(blockIdx.x, p.chunks_rem);
    const int n = (int)p.chunks_per_cta + (blockIdx.x < p.chunks_rem ? 1 : 0);
    const unsigned cpr = (unsigned)p.chunks_per_row;
    const AntqCodebook *__restrict__ cb = p.cb;

    asm volatile("griddepcontrol.launch_dependents;");                // programmatic dependent launch, as antq_stream.cu
    if (threadIdx.x < kNS) antq_mbar_init(full + threadIdx.x, 1);
    if (threadIdx.x == 0) *next_k = 0;
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");

    struct Geo { long long base; int nvec, tail; unsigned row; float alpha; };
    auto geo_of = [&](int k) {
        Geo g;
        const unsigned c = c_begin + (unsigned)k;
        const unsigned row = p.cpr_shift >= 0 ? c >> p.cpr_shift : c / cpr;
        const long long col0 = (long long)(c - row * cpr) * p.chunk_elems;
        const long long remain = p.cols - col0;
        const int n_el = (int)(remain < p.chunk_elems ? remain : p.chunk_elems);
        g.nvec = n_el / VEC;
        g.tail = n_el - g.nvec * VEC;
        g.base = (long long)row * p.cols + col0;
        g.row = row;
        g.alpha = 0.0f;
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
        g.alpha = __ldg(p.alpha + (p.alpha_per_row ? g.row : 0u));
    };
    auto claim = [&]() {
        int k = 0;
        if (lane == 0) k = (int)atomicAdd(next_k, 1u);
        return __shfl_sync(0xffffffffu, k, 0);
    };

    Geo cur;
    cur.base = 0; cur.nvec = 0; cur.tail = 0; cur.row = 0; cur.alpha = 0.0f;
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
    const PuK K = pu_load_k(cb);
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
        const PuRow r = pu_row<T>(g.alpha, p, K, false);
        const int nvec = g.nvec;
        const uint4 *sv = reinterpret_cast<const uint4 *>(pu_smem + (size_t)stage * kChunkMax);
        T *og = reinterpret_cast<T *>(p.out) + g.base;
        uint4 *ov = reinterpret_cast<uint4 *>(og);
        antq_mbar_wait(full + stage, (phases >> slot) & 1u);
        phases ^= 1u << slot;
        unsigned redo = 0;                                            // bit j: vector j * 32 + lane needs the exact pass
        if (r.ok && !(p.debug & 2)) {
            const uint4 *sp = sv + lane;
            uint4 *op = ov + lane;
            int j = 0;
#pragma unroll 1
            for (int v = lane; v + 32 < nvec; v += 64, j += 2) {
                const uint4 r0 = sp[0], r1 = sp[32];
                bool f0, f1;
                const uint4 q0 = pu_vec<T, UNIFORM, XC>(r0, r, K, tab, f0);
                const uint4 q1 = pu_vec<T, UNIFORM, XC>(r1, r, K, tab, f1);
                antq_stg_stream(op, q0);
                antq_stg_stream(op + 32, q1);
                redo |= (f0 ? 1u : 0u) << j;
                redo |= (f1 ? 2u : 0u) << j;
                sp += 64; op += 64;
            }
            if (j * 32 + lane < nvec) {
                bool f0;
                const uint4 q0 = pu_vec<T, UNIFORM, XC>(*sp, r, K, tab, f0);
                antq_stg_stream(op, q0);
                redo |= (f0 ? 1u : 0u) << j;
            }
        } else if (p.debug & 2) {
            for (int v = lane; v < nvec; v += 32) antq_stg_stream(ov + v, sv[v]);
        } else {
            redo = 0xffffffffu;                                       // bad scale: every vector, literally
        }
        if (__any_sync(0xffffffffu, redo != 0)) {
            // Exact redo, dense: the flagged ELEMENTS of the whole chunk are compacted into a per-warp list and redone
            // one per lane (a row with representable ties can flag a few percent of a chunk; redoing them vector by
            // vector inside their owner lane serialised a warp for tens of microseconds: profiles/r02_notes.md).
            unsigned long long em = 0;                                // bit 8 j + e: element e of vector j * 32 + lane
            if (r.ok) {
                for (int j = 0; j * 32 + lane < nvec; j++)
                    if ((redo >> j) & 1u)
                        em |= (unsigned long long)pu_vec_mask<T, UNIFORM>(sv[j * 32 + lane], r, K, tab) << (8 * j);
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
                    const float xf = A::to_f32(reinterpret_cast<const T *>(sv + v)[e]);
                    og[(long long)v * VEC + e] = pu_exact_elem<T>(cb, X, xf, r.s);
                }
                __syncwarp();          // the list is rewritten by this warp's next chunk
            } else {
                for (int j = 0; j * 32 + lane < nvec; j++) {
                    if ((redo >> j) & 1u) {
                        const int v = j * 32 + lane;
                        pu_redo_vec<T, UNIFORM>(cb, X, sv[v], r, K, tab, og + (long long)v * VEC);
                    }
                }
            }
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
}

// ==================================================================================================
// Short rows / scale groups: grid-stride over 16-byte vectors.
// ==================================================================================================
constexpr int kShortThreads = 256;

template <typename T, bool UNIFORM, bool XC>
__global__ void __launch_bounds__(kShortThreads, 4) antq_pu_short_kernel(const PuParams p) {
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
    const PuK K = pu_load_k(p.cb);
    const uint4 *xin = reinterpret_cast<const uint4 *>(p.x);
    uint4 *xout = reinterpret_cast<uint4 *>(p.out);
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < p.nvec; v += stride) {
        const uint4 raw = antq_ldg_stream(xin + v);
        unsigned row = 0;
        if (p.alpha_per_row) row = p.cols_shift >= 0 ? v >> p.cols_shift : v / p.cols_vec;
        const PuRow r = pu_row<T>(__ldg(p.alpha + row), p, K, true);
        bool flag = true;
        uint4 q = raw;
        if (r.ok) q = pu_vec<T, UNIFORM, XC>(raw, r, K, tab, flag);
        antq_stg_stream(xout + v, q);
        if (flag) pu_redo_vec<T, UNIFORM>(p.cb, X, raw, r, K, tab, reinterpret_cast<T *>(p.out) + (long long)v * VEC);
    }
}

// ==================================================================================================
// Dynamic scale groups: alpha = max|x| over the group * ratio, computed in the same pass (ONE read of x).
// A group of cols = L * VEC elements is held by L adjacent lanes (L a power of two <= 32): local abs-max, xor-shuffle
// reduction (integer max on the fp32 bit patterns of |x|: NaN-propagating, like torch's abs().max()), then the closed form.
// ==================================================================================================
template <typename T, bool UNIFORM, bool XC>
__global__ void __launch_bounds__(kShortThreads, 4) antq_pu_dynamic_kernel(const PuParams p, float ratio, float *__restrict__ alpha_out) {
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
    const PuK K = pu_load_k(p.cb);
    const uint4 *xin = reinterpret_cast<const uint4 *>(p.x);
    uint4 *xout = reinterpret_cast<uint4 *>(p.out);
    const unsigned stride = gridDim.x * blockDim.x;
    const unsigned L = p.cols_vec;                                    // lanes per group
    const unsigned nvec_up = (p.nvec + 31u) & ~31u;                   // every lane of a warp runs the same trip count
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < nvec_up; v += stride) {
        const bool live = v < p.nvec;
        uint4 raw = make_uint4(0, 0, 0, 0);
        if (live) raw = antq_ldg_stream(xin + v);
        T xv[VEC];
        *reinterpret_cast<uint4 *>(xv) = raw;
        unsigned m = 0;
#pragma unroll
        for (int e = 0; e < VEC; e++) {
            const unsigned b = __float_as_uint(A::to_f32(xv[e])) & 0x7fffffffu;
            m = b > m ? b : m;
        }
        for (unsigned o = 1; o < L; o <<= 1) {
            const unsigned t = __shfl_xor_sync(0xffffffffu, m, o);
            m = t > m ? t : m;
        }
        const float alpha = __fmul_rn(__uint_as_float(m), ratio);     // alpha = absmax * ratio (fp32, like the torch expression)
        if (!live) continue;
        if (alpha_out && (v & (L - 1)) == 0) alpha_out[v >> p.cols_shift] = alpha;
        const PuRow r = pu_row<T>(alpha, p, K, true);
        bool flag = true;
        uint4 q = raw;
        if (r.ok) q = pu_vec<T, UNIFORM, XC>(raw, r, K, tab, flag);
        antq_stg_stream(xout + v, q);
        if (flag) pu_redo_vec<T, UNIFORM>(p.cb, X, raw, r, K, tab, reinterpret_cast<T *>(p.out) + (long long)v * VEC);
    }
}

template <typename T, bool UNIFORM, bool XC> int launch_stream(const PuParams &p, int ctas, cudaStream_t st) {
    auto kernel = antq_pu_stream_kernel<T, UNIFORM, XC>;
    const int smem = kNS * kChunkMax + 512 * 8 + 2 * ANTQ_MAX_GRID * 4 + kNS * 8 + 16 + kNC * kListMax * 2;
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
    const long long want = ((long long)p.nvec + kShortThreads - 1) / kShortThreads;
    const long long cap = (long long)antq_num_sms() * 8;
    antq_pu_short_kernel<T, UNIFORM, XC><<<(int)(want < cap ? want : cap), kShortThreads, 0, st>>>(p);
    return (int)cudaGetLastError();
}

}  // namespace



// Returns ANTQ_ENOTSUP for shapes the persistent kernel does not cover (more than 2^31 chunks).
int antq_launch_pu_stream(const void *x, void *out, const float *alpha, int alpha_per_row, long long rows, long long cols,
                          int dtype, const AntqCodebook *cb, const antq_codebook_info *info, cudaStream_t st) {
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
    const unsigned sms = (unsigned)antq_num_sms();
    const int ctas = (int)(p.total_chunks < sms ? p.total_chunks : sms);
    p.chunks_per_cta = p.total_chunks / (unsigned)ctas;
    p.chunks_rem = p.total_chunks % (unsigned)ctas;
    const bool uni = (info->flags & ANTQ_CB_PU_UNIFORM) != 0;
    const bool xc = dtype == ANTQ_F16 ? (info->flags & ANTQ_CB_PU_XC16) != 0 : dtype == ANTQ_BF16 ? (info->flags & ANTQ_CB_PU_XCBF) != 0 : false;
#define ANTQ_PU_GO(T, X) (uni ? launch_stream<T, true, X>(p, ctas, st) : launch_stream<T, false, X>(p, ctas, st))
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
    PuParams p = {};
    p.x = x; p.out = out; p.alpha = alpha; p.cb = cb;
    p.rows = rows; p.cols = cols;
    p.nvec = (unsigned)(n / vec);
    p.cols_vec = (unsigned)(cols / vec);
    p.cols_shift = -1;
    if ((p.cols_vec & (p.cols_vec - 1)) == 0) {
        int sh = 0;
        while ((1u << sh) < p.cols_vec) sh++;
        p.cols_shift = sh;
    }
    p.alpha_per_row = alpha_per_row;
    p.gmax = info->gmax; p.lim = info->lim;
    const bool uni = (info->flags & ANTQ_CB_PU_UNIFORM) != 0;
    const bool xc = dtype == ANTQ_F16 ? (info->flags & ANTQ_CB_PU_XC16) != 0 : dtype == ANTQ_BF16 ? (info->flags & ANTQ_CB_PU_XCBF) != 0 : false;
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
    const long long want = ((long long)p.nvec + kShortThreads - 1) / kShortThreads;
    const long long cap = (long long)antq_num_sms() * 8;
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
