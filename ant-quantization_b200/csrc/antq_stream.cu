// antq_stream.cu -- the hot kernel: fused fake-quant forward, one persistent CTA per SM.
//
// Replaces A/antquant/quant_modules.py:535-551 (+ O/...:295-330 with OVP) and the scan kernel it
// calls (A/quant/quant_kernel.cu:11-39): ~9 launches and 7 reads + 7 writes of the tensor become
// ONE launch, one read, one write.
//
// Arithmetic: x-space thresholds, exact STE window (proved bit-exact on
// exhaustive fp16 inputs in tests/test_xspace_model.py).  Per 32-bit register (two 16-bit values):
//   ALU-pipe chain   m = HSET2.BM(|x| >= X_i);  q ^= m & E_i          (E_i = O_{i+1} xor O_i)
//   FMA-pipe chain   m = HFMA2.SAT(|x|S, B_i, C_i);  q = HFMA2(m, D_i, q)   (exact, see build_tables)
// and the pairs of every vector are split between the two chains: every op class issues at most every
// other cycle, and the scheduler only reaches ~1 instruction/cycle when consecutive instructions of a
// warp go to different classes (tools/probe/pipe_probe.cu).
//
// Execution shape (B200): one persistent CTA per SM (148 on B200, queried at run time), each owning an equal contiguous range of
// chunks (<= 4 KiB pieces that never straddle a row) -- found by measurement, profiles/r01_notes.md:
//   * 12 CONSUMER warps (3 per scheduler).  Each owns two private 4 KiB stages in shared memory and
//     double-buffers: it claims the next chunk from a CTA-wide counter, puts it in flight with one
//     TMA bulk copy (cp.async.bulk -> shared memory, completion on the stage's mbarrier), then
//     computes the chunk that has already landed: LDS.128 -> chains in registers -> STG.128.  The
//     warp that frees a stage is the one that refills it, so there is no producer warp to starve and
//     no `empty` barrier; an SM keeps 12 x 4 KiB outstanding -- about the HBM latency-bandwidth
//     product -- and never floods the memory system at start-up.
//   * the per-row tables (thresholds, outputs, FMA twins) are built lane-parallel, 32/(NT+1) rows
//     per pass, into a 32-row ring in shared memory: in the PROLOGUE by every warp at once (the
//     consumers have nothing to do until their first chunk lands), afterwards by 3 BUILDER warps that
//     follow the consumers around the ring.  Consumers read them with broadcast LDS.128.
//   * programmatic dependent launch: barrier set-up of launch N+1 overlaps the tail of launch N.
//
// Out-of-window values (|d| > lim, NaN, Inf) are detected with one running packed max per register;
// the chunk is still in shared memory when the test fires, so a cold pass recomputes exactly those
// vectors with the literal reference arithmetic (also in place).
#include <stdio.h>
#include <stdlib.h>

#include "antq_common.cuh"

namespace {

#ifndef ANTQS_CONSUMERS
#define ANTQS_CONSUMERS 12
#endif
#ifndef ANTQS_BUILDERS
#define ANTQS_BUILDERS 3
#endif
#ifndef ANTQS_RING
#define ANTQS_RING 2
#endif
#ifndef ANTQS_CHUNK
#define ANTQS_CHUNK 4096
#endif
constexpr int kNC = ANTQS_CONSUMERS;      // consumer warps per CTA
constexpr int kRing = ANTQS_RING;         // private ring stages per consumer warp
constexpr int kNS = kNC * kRing;          // stages per CTA
constexpr int kChunkMax = ANTQS_CHUNK;    // bytes of tensor per stage
constexpr int kMetaWords = 12;
constexpr int kNB = ANTQS_BUILDERS;        // table-builder warps per CTA
constexpr int kThreads = (kNC + kNB) * 32;       // consumers | builders
static_assert(kChunkMax <= 4096, "two-phase OVP keeps vector indices of a chunk in one byte");
constexpr int kRT = 32;                    // row-table ring slots (power of two)

enum : uint32_t { kRowOk = 1u, kRowTies = 2u, kRowFma = 4u };

template <int NT> struct TabGeom {
    static constexpr int NTP = NT + 1;                 // table pitch in words: 4, 8, 16, 32
    static constexpr int G = 32 / NTP;                 // rows built in one pass of a warp
    static constexpr int kTabBytes = ((7 * NTP + kMetaWords) * 4 + 127) / 128 * 128;
};

struct StreamParams {
    const void *x;
    void *out;
    const float *alpha;
    const AntqCodebook *cb;
    long long rows, cols;
    unsigned total_chunks, chunks_per_cta, chunks_rem;   // CTA b owns chunks_per_cta (+1 if b < chunks_rem) chunks
    int alpha_per_row, chunks_per_row, chunk_elems;
    int cpr_shift;      // log2(chunks_per_row) when it is a power of two, else -1
    int nt_real, mid, ovp_index, n_entries;
    float gmax, lim;
    int debug;          // ANTQ_DEBUG experiments: 2 = no chain (copy through), 16 = no FMA twin, 128 = one-phase OVP
    unsigned long long *trace;   // ANTQS_TRACE builds: per-CTA timeline (tools/trace_stream.py)
};

#ifdef ANTQS_TRACE
constexpr int kTraceChunks = 64, kTraceStride = 8 + 4 * kTraceChunks;
__device__ __forceinline__ unsigned long long antqs_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define ANTQS_TR(slot) do { if (p.trace && lane == 0) p.trace[(size_t)blockIdx.x * kTraceStride + (slot)] = antqs_now(); } while (0)
#define ANTQS_TRK(k, what) do { if ((k) < kTraceChunks) ANTQS_TR(8 + 4 * (k) + (what)); } while (0)
#else
#define ANTQS_TR(slot) do { } while (0)
#define ANTQS_TRK(k, what) do { } while (0)
#endif

// ---- mbarrier helpers not in antq_common.cuh ------------------------------------------------
__device__ __forceinline__ void antqs_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(antq_smem_u32(bar)) : "memory");
}

// ---- the literal reference arithmetic for the rare vectors the chains may not handle ---------
template <typename T, bool OVP>
__device__ __noinline__ void antqs_slow_vec(const AntqCodebook *__restrict__ cb, float s, const T *xv, T *og, int n) {
    typedef AntqType<T> A;
    float q[8], d[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
        if (e < n) {
            AntqExact ex = antq_exact_quant(cb, A::to_f32(xv[e]), s);
            q[e] = ex.q; d[e] = ex.d;
        }
    }
    if (OVP) {
#pragma unroll
        for (int e = 0; e + 1 < 8; e += 2) {
            if (e + 1 < n) {
                const bool oe = fabsf(q[e]) > 32.0f, oo = fabsf(q[e + 1]) > 32.0f;
                if (oe) q[e + 1] = __fmul_rn(q[e + 1], 0.0f);
                else if (oo) q[e] = __fmul_rn(q[e], 0.0f);
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; e++)
        if (e < n) og[e] = A::from_f32_rn(antq_ste_rescale(q[e], d[e], s));
}

// Cold pass over a chunk still resident in shared memory: every vector holding an out-of-window
// element (or every vector, when the row's scale is not a positive finite number) is recomputed.
template <typename T, bool OVP>
__device__ __noinline__ void antqs_fixup_chunk(const AntqCodebook *__restrict__ cb, float s, float xlim, bool all,
                                               const uint4 *sv, T *og, int nvec, int lane) {
    typedef AntqType<T> A;
    constexpr int VEC = A::kVec;
    for (int v = lane; v < nvec; v += 32) {
        const T *xv = reinterpret_cast<const T *>(sv + v);
        bool special = all;
#pragma unroll
        for (int e = 0; e < VEC; e++) special |= !(fabsf(A::to_f32(xv[e])) <= xlim);
        if (special) antqs_slow_vec<T, OVP>(cb, s, xv, og + (long long)v * VEC, VEC);
    }
}

// Exact (X, Xn) for a 16-bit type: X = min{x in T : fl32(x / s) >= tpos}, Xn likewise for tneg (>= tpos, a few
// fp32-ulps above it).  RN_T(fl32(tpos * s)) is within one step of X (tests/xspace_model.py), so two divisions settle
// it; the walk upward is a safeguard that has never been observed to take more than one step.
template <typename T>
__device__ __forceinline__ void antqs_x_threshold16(float tpos, float tneg, float s, T *xp, T *xn) {
    typedef AntqType<T> A;
    T c = A::from_f32_rn(__fmul_rn(tpos, s));
    float qc = __fdiv_rn(A::to_f32(c), s);
    if (qc >= tpos) {
        const T p = antq_next_down(c);
        const float qp = __fdiv_rn(A::to_f32(p), s);
        if (!antq_is_inf(p) && qp >= tpos) { c = p; qc = qp; }
    } else {
#pragma unroll 1
        for (int it = 0; it < 6 && !(qc >= tpos); ++it) {
            c = antq_next_up(c);
            qc = __fdiv_rn(A::to_f32(c), s);
        }
    }
    *xp = c;
    // the next 16-bit value is thousands of fp32-ulps further, tneg only a few: one test decides
    *xn = (qc >= tneg) ? c : antq_next_up(c);
}

template <typename T> __device__ __forceinline__ uint32_t antqs_word(T v) {
    if constexpr (sizeof(T) == 2) {
        const uint32_t b = AntqType<T>::bits(v);
        return b | (b << 16);
    } else {
        return __float_as_uint(v);
    }
}

template <typename T> struct Pack2 {
    typedef float2 v2;
    __device__ static __forceinline__ v2 from_u32(uint32_t) { return make_float2(0.f, 0.f); }
};
template <> struct Pack2<__half> {
    typedef __half2 v2;
    __device__ static __forceinline__ v2 from_u32(uint32_t u) { return *reinterpret_cast<v2 *>(&u); }
    static constexpr uint32_t kOne = 0x3c003c00u;
};
template <> struct Pack2<__nv_bfloat16> {
    typedef __nv_bfloat162 v2;
    __device__ static __forceinline__ v2 from_u32(uint32_t u) { return *reinterpret_cast<v2 *>(&u); }
    static constexpr uint32_t kOne = 0x3f803f80u;
};
template <typename V> __device__ __forceinline__ uint32_t antqs_u32(const V &v) {
    return *reinterpret_cast<const uint32_t *>(&v);
}

// ==================================================================================================
// Producer side: the tables of one row, one lane per threshold.
// ==================================================================================================
struct RowTab {
    uint32_t X, Xn, T2, B, C, D, Cn;        // this lane's entry of each table
    uint32_t s_bits, flags, xlim, S2, xovp, xovpn;   // uniform over the lanes of one row group
    uint32_t Xe, Ee, De;                    // XNEG: the extra level at the negative end (x < Xe selects it)
};

//   X[i]   min{x in T : fl32(x / s) >= thr_i}   (x >= 0, or every x for an asymmetric codebook)
//   Xn[i]  the same on |x| for x < 0: differs from X only where a tie x / s == midpoint is representable
//          (ties go to the LATER grid entry: up for positive d, toward zero for negative d)
//   T2     16-bit types: E[i] = O[i+1] ^ O[i] (+ a sign marker in E[0] for symmetric codebooks), and O[0] in
//          slot NT;  fp32: O[i] = fl32(level_i * s), i <= NT
//   B, C, D, S2   FMA-pipe twin (16-bit types, tie-free rows): with P = prev(X), u = X - P (a power of two),
//          m = sat((S|x|) * (1/(uS)) - P/u) is exactly 1.0 for |x| >= X and exactly 0.0 for |x| <= P, because the
//          fused product-sum is <= 0 or >= 1 before its single rounding; D[i] = O[i+1] - O[i] must be exactly
//          representable (checked), so q <- m*D + q reproduces O[rank] without rounding.
template <typename T, int NT, bool SYM, bool OVP, bool XNEG>
__device__ __forceinline__ void build_tables(RowTab &t, const StreamParams &p, float alpha, float tpos, float tneg,
                                             float lev, float thr_e, float lev_e, int sub, int grp, bool tr) {
    typedef AntqType<T> A;
    typedef TabGeom<NT> TG;
    constexpr int NTP = TG::NTP;
    constexpr unsigned kFull = 0xffffffffu;
    const int nt_real = p.nt_real;
    const float inf = __int_as_float(0x7f800000);
    const unsigned gmask = NTP == 32 ? kFull : (((1u << (NTP & 31)) - 1u) << (grp * NTP));

    const float s = __fdiv_rn(alpha, p.gmax);                          // scale = alpha / max(grid)
    const bool row_ok = s > 0.0f && s < inf;
#ifdef ANTQS_TRACE
    const int lane = sub + grp * NTP;
    if (tr && row_ok) ANTQS_TR(5);
#endif
    T Xl = A::from_bits(A::kInf), Xnl = A::from_bits(A::kInf), Ol = A::from_bits(0);
    if (row_ok) {
        if (sub < nt_real) {
            if constexpr (sizeof(T) == 2) {
                // always the exact two-division form, inline: straight-line code on the start-up critical path
                // (the division-free shortcut saved arithmetic but put a divergent call in front of the first chunk)
                antqs_x_threshold16<T>(tpos, SYM ? tneg : tpos, s, &Xl, &Xnl);
                if (!SYM) Xnl = Xl;
            } else {
                bool near;
                Xl = antq_x_threshold<T>(tpos, s, &near);
                Xnl = Xl;
                // a negative input can land on the other side of a representable tie
                if (SYM && near) Xnl = antq_x_threshold_exact<T>(tneg, s);
            }
        }
        if (sub <= nt_real) Ol = A::from_f32_rn(__fmul_rn(lev, s));
    }
    const bool ties = (__ballot_sync(kFull, A::bits(Xl) != A::bits(Xnl)) & gmask) != 0;
#ifdef ANTQS_TRACE
    if (tr) ANTQS_TR(6);
#endif
    t.X = antqs_word<T>(Xl);
    t.Xn = antqs_word<T>(Xnl);
    t.s_bits = __float_as_uint(s);
    t.flags = (row_ok ? kRowOk : 0u) | (ties ? kRowTies : 0u);
    const float xl = __fmul_rn(__fmul_rn(p.lim, s), 0.9990234375f);    // conservative window in x-space
    t.xlim = antqs_word<T>(A::from_f32_rz(xl));
    t.S2 = 0; t.B = 0; t.C = 0; t.D = 0; t.Cn = 0;
    t.Xe = 0; t.Ee = 0; t.De = 0;
    // XNEG (signed int-k): the level below -max_common is chosen iff d < thr_e, i.e. x < Xe; its output is
    // RN(lev_e * s) = -Oe.  Every lane of the row group computes the same values.
    T Xe_t = A::from_bits(0), Oe_t = A::from_bits(0);
    if constexpr (XNEG) {
        if (row_ok) {
            if constexpr (sizeof(T) == 2) { T dummy; antqs_x_threshold16<T>(thr_e, thr_e, s, &Xe_t, &dummy); }
            else Xe_t = antq_x_threshold_exact<T>(thr_e, s);
            Oe_t = A::from_f32_rn(__fmul_rn(-lev_e, s));               // magnitude of the extra level
        }
        t.Xe = antqs_word<T>(Xe_t);
    }
    {
        const int oi = p.ovp_index;
        const bool has = OVP && oi >= 0 && oi < NT;
        const uint32_t infw = antqs_word<T>(A::from_bits(A::kInf));
        const uint32_t xo = __shfl_sync(kFull, t.X, has ? oi : 0, NTP);
        const uint32_t xon = __shfl_sync(kFull, t.Xn, has ? oi : 0, NTP);
        t.xovp = has ? xo : infw;
        t.xovpn = has ? xon : infw;
    }
    if constexpr (sizeof(T) == 2) {
        const uint32_t ob = A::bits(Ol);
        const uint32_t ob_next = __shfl_down_sync(kFull, ob, 1, NTP);
        const uint32_t o0 = __shfl_sync(kFull, ob, 0, NTP);
        uint32_t e = sub < nt_real ? (ob ^ ob_next) : 0u;
        if (SYM && sub == 0) e |= 0x8000u;                            // marker: "level is not the zero level"
        if (sub == NT) e = o0;
        t.T2 = e | (e << 16);
        if constexpr (XNEG) {
            const uint32_t otop = __shfl_sync(kFull, ob, nt_real, NTP);      // output of the largest common magnitude
            const uint32_t ee = otop ^ (uint32_t)A::bits(Oe_t);
            t.Ee = ee | (ee << 16);
        }

        // FMA-pipe twin
        bool ok = row_ok && (NT <= 15 || (OVP && !SYM && NT == 31)) && (!ties || NT <= 7 || (OVP && NT == 15)) &&
                  !(p.debug & 16);   // (two-phase OVP runs
        // its 7-threshold normal chain with the tie-aware FMA twin; the full 15-threshold chain never does)
        // S = 2^k with xlim * S in [2^13, 2^14): every in-window |x| stays finite after scaling
        int k = 13 - (int)((__float_as_uint(xl) >> 23) & 0xff) + 127;
        k = k > 15 ? 15 : (k < -14 ? -14 : k);
        const float S = __uint_as_float((unsigned)(127 + k) << 23);
        ok = ok && isfinite(A::to_f32(Ol));
        float Bv = 0.0f, Cv = -1.0f;              // padding: m = sat(0 * xs - 1) = 0
        if (sub < nt_real) {
            const T P = antq_next_down(Xl);
            if (antq_is_inf(Xl)) {
                Bv = 0.0f; Cv = -1.0f;                                   // never reached
            } else if (antq_is_inf(P)) {
                Bv = 0.0f; Cv = 1.0f;                                    // always reached
            } else {
                const float u = __fsub_rn(A::to_f32(Xl), A::to_f32(P));  // exact power of two
                const int eu = (int)((__float_as_uint(u) >> 23) & 0xff);       // u is a normal fp32
                const float Bfull = __uint_as_float((unsigned)(254 - eu) << 23);                // 1 / u
                const int eb = 254 - eu - k;
                Bv = (eb >= 1 && eb <= 254) ? __uint_as_float((unsigned)eb << 23) : 0.0f;      // 1 / (u S)
                Cv = -__fmul_rn(A::to_f32(P), Bfull);
                ok = ok && Bv >= 6.103515625e-05f && Bv <= 32768.0f && fabsf(Cv) <= 2047.0f;
                if (sizeof(T) == 2 && A::kInf == 0x7f80u) ok = ok && fabsf(Cv) <= 255.0f;  // bf16: 8-bit integers
            }
        }
        // D_i = O[i+1] - O[i] must be exactly representable (fp32 difference of two 16-bit values is exact)
        const float o_next = __shfl_down_sync(kFull, A::to_f32(Ol), 1, NTP);
        const float dex = __fsub_rn(o_next, A::to_f32(Ol));
        const T D = A::from_f32_rn(dex);
        if (sub < nt_real) ok = ok && (A::to_f32(D) == dex);
        if constexpr (XNEG) {
            const float otop = __shfl_sync(kFull, A::to_f32(Ol), nt_real, NTP);
            const float de = __fsub_rn(A::to_f32(Oe_t), otop);
            const T De_t = A::from_f32_rn(de);
            ok = ok && (A::to_f32(De_t) == de) && isfinite(A::to_f32(Oe_t));
            t.De = antqs_word<T>(De_t);
        }
        const bool fma_ok = (__ballot_sync(kFull, ok) & gmask) == gmask;
        t.D = antqs_word<T>(sub < nt_real ? D : A::from_bits(0));
        t.B = antqs_word<T>(A::from_f32_rn(Bv));
        t.C = antqs_word<T>(A::from_f32_rn(Cv));
        // negative inputs of a tie threshold compare against X + 1 ulp:  sat((|x| - X) / u) = sat(|x| / u + (C - 1))
        t.Cn = antqs_word<T>(A::from_f32_rn(A::bits(Xl) != A::bits(Xnl) ? __fsub_rn(Cv, 1.0f) : Cv));
        t.S2 = antqs_word<T>(A::from_f32_rn(S));
        if (fma_ok) t.flags |= kRowFma;
    } else {
        t.T2 = __float_as_uint(Ol);
        if constexpr (XNEG) t.Ee = __float_as_uint(-Oe_t);            // fp32: the signed output itself
    }
}

// ==================================================================================================
// Consumer side
// ==================================================================================================
template <int N> __device__ __forceinline__ void load_words(uint32_t (&dst)[N], const uint32_t *src) {
    constexpr int NV = (N + 3) / 4;
    uint4 b[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) b[i] = reinterpret_cast<const uint4 *>(src)[i];
#pragma unroll
    for (int i = 0; i < N; i++) {
        const uint4 a = b[i / 4];
        dst[i] = (i % 4 == 0) ? a.x : (i % 4 == 1) ? a.y : (i % 4 == 2) ? a.z : a.w;
    }
}

enum { kModeAlu = 0, kModeTies = 1, kModeMix = 2, kModeTiesMix = 3 };

// 16-bit element types: one 32-bit register = two elements (low half = even flat index).
//   kModeAlu      every pair on the ALU pipe
//   kModeMix      tie-free row with exact FMA twins: pairs split between the ALU and the FMA pipe
//   kModeTies     row with representable ties (negative inputs use Xn), ALU pipe only
//   kModeTiesMix  the same with FMA twins: negative inputs select C - 1 (one LOP3) and the compare / accumulate
//                 run on the FMA pipe; one pair of four stays on the ALU pipe
template <typename T, int NT, bool SYM, bool OVP, bool XNEG, int MODE, int PITCH = NT + 1> struct Chain16 {
    typedef typename Pack2<T>::v2 v2;
    static constexpr int NTP = PITCH;         // table pitch of the row-table slot (> NT + 1 for a prefix chain)
    static constexpr bool TIES = SYM && (MODE == kModeTies || MODE == kModeTiesMix);
    static constexpr bool FMA = MODE == kModeMix || MODE == kModeTiesMix;
    uint32_t X[NT + 1];                       // [NT] unused
    uint32_t E[NT + 1];                       // [NT] = O[0]
    uint32_t Xn[TIES ? NT + 1 : 1];
    uint32_t Bf[FMA ? NT + 1 : 1], Cf[FMA ? NT + 1 : 1], Df[FMA ? NT + 1 : 1];
    uint32_t Cn[(FMA && TIES) ? NT + 1 : 1];
    uint32_t S2, xovp, xovpn;
    uint32_t Xe, Ee, De;                      // XNEG: extra level at the negative end
    v2 mx;                                    // running max of |x| (NaN-propagating)

    __device__ __forceinline__ void load(const uint32_t *tab, uint32_t s2, uint32_t xo, uint32_t xon) {
        load_words<NT + 1>(X, tab);
        load_words<NT + 1>(E, tab + 2 * NTP);
        if constexpr (XNEG) {
            const uint4 me = reinterpret_cast<const uint4 *>(tab + 7 * NTP)[2];
            Xe = me.x; Ee = me.y; De = me.z;
        }
        if constexpr (TIES) load_words<NT + 1>(Xn, tab + NTP);
        if constexpr (FMA) {
            load_words<NT + 1>(Bf, tab + 3 * NTP);
            load_words<NT + 1>(Cf, tab + 4 * NTP);
            load_words<NT + 1>(Df, tab + 5 * NTP);
        }
        if constexpr (FMA && TIES) load_words<NT + 1>(Cn, tab + 6 * NTP);
        S2 = s2; xovp = xo; xovpn = xon;
        mx = Pack2<T>::from_u32(0u);
    }

    __device__ __forceinline__ uint32_t ovp_mask(v2 a2, uint32_t neg) const {
        const uint32_t t = TIES ? ((neg & xovpn) | (~neg & xovp)) : xovp;
        const uint32_t mo = __hge2_mask(a2, Pack2<T>::from_u32(t));      // element is an outlier
        const uint32_t sw = __byte_perm(mo, 0, 0x1032);                  // swap halves
        return sw & ~(mo & 0x0000ffffu);   // odd dies if even is outlier; even dies if only odd is
    }

    // ALU pipe: HSET2 + LOP3 per threshold
    __device__ __forceinline__ uint32_t pair_alu(uint32_t xb) {
        const v2 x2 = Pack2<T>::from_u32(xb);
        const v2 ab = __habs2(x2);
        const v2 a2 = SYM ? ab : x2;
        uint32_t neg = 0;
        if (TIES) neg = __hlt2_mask(x2, Pack2<T>::from_u32(0u));       // 0xffff where x < 0
        uint32_t q = SYM ? 0u : E[NT];
#pragma unroll
        for (int i = 0; i < NT; i++) {
            uint32_t t = X[i];
            if constexpr (TIES) t = (neg & Xn[i]) | (~neg & X[i]);
            const uint32_t m = __hge2_mask(a2, Pack2<T>::from_u32(t));
            q ^= m & E[i];
        }
        if (XNEG) q ^= __hlt2_mask(x2, Pack2<T>::from_u32(Xe)) & Ee;   // below the most negative common level
        if (SYM) q &= xb | 0x7fff7fffu;        // keep the marker bit (level != 0) only where x is negative
        if (OVP) q &= ~ovp_mask(a2, neg);
        return q;
    }

    // FMA pipe: saturating-FMA compare + exact accumulate
    __device__ __forceinline__ uint32_t pair_fma(uint32_t xb) {
        const v2 x2 = Pack2<T>::from_u32(xb);
        const v2 ab = __habs2(x2);
        uint32_t neg = 0;
        if (TIES) neg = __hlt2_mask(x2, Pack2<T>::from_u32(0u));
        const v2 xs = __hmul2(SYM ? ab : x2, Pack2<T>::from_u32(S2));
        v2 q = Pack2<T>::from_u32(E[NT]);
#pragma unroll
        for (int i = 0; i < NT; i++) {
            uint32_t c = Cf[i];
            if constexpr (FMA && TIES) c = (neg & Cn[i]) | (~neg & Cf[i]);
            const v2 m = __hfma2_sat(xs, Pack2<T>::from_u32(Bf[i]), Pack2<T>::from_u32(c));
            q = __hfma2(m, Pack2<T>::from_u32(Df[i]), q);
        }
        if (XNEG) q = __hfma2(__hlt2(x2, Pack2<T>::from_u32(Xe)), Pack2<T>::from_u32(De), q);
        uint32_t qb;
        if (SYM) {
            // (+-1) * q + 0: restores the sign and leaves a zero level at +0
            const v2 sg = Pack2<T>::from_u32((xb & 0x80008000u) | Pack2<T>::kOne);
            qb = antqs_u32(__hfma2(q, sg, Pack2<T>::from_u32(0u)));
        } else {
            qb = antqs_u32(q);
        }
        if (OVP) qb &= ~ovp_mask(SYM ? ab : x2, neg);
        return qb;
    }

    // max |x| of one vector (NaN-propagating)
    __device__ static __forceinline__ v2 vec_max(const uint4 r) {
        const v2 a = __hmax2_nan(__habs2(Pack2<T>::from_u32(r.x)), __habs2(Pack2<T>::from_u32(r.y)));
        return __hmax2_nan(__hmax2_nan(a, __habs2(Pack2<T>::from_u32(r.z))), __habs2(Pack2<T>::from_u32(r.w)));
    }
    template <bool MAX = true> __device__ __forceinline__ uint4 vec(const uint4 r, int debug) {
        if (debug & 2) return r;
        if constexpr (MAX) {
            // running max of |x|: written as two back-to-back pairs so that ptxas emits 3-input VHMNMX
            mx = __hmax2_nan(__hmax2_nan(mx, __habs2(Pack2<T>::from_u32(r.x))), __habs2(Pack2<T>::from_u32(r.y)));
            mx = __hmax2_nan(__hmax2_nan(mx, __habs2(Pack2<T>::from_u32(r.z))), __habs2(Pack2<T>::from_u32(r.w)));
        }
        uint4 q;
        q.x = pair_alu(r.x);
        if constexpr (MODE == kModeTiesMix) q.z = pair_fma(r.z);
        else q.z = pair_alu(r.z);
        if constexpr (FMA) {
            q.y = pair_fma(r.y);
            q.w = pair_fma(r.w);
        } else {
            q.y = pair_alu(r.y);
            q.w = pair_alu(r.w);
        }
        return q;
    }
    // true when some element this lane saw was NaN / Inf / outside the exact window
    __device__ __forceinline__ bool special(uint32_t xlim) const {
        return __hle2_mask(mx, Pack2<T>::from_u32(xlim)) != 0xffffffffu;
    }
};

// fp32: one element per register; fp32 x-space resolves ties, so negatives always use Xn.
template <int NT, bool SYM, bool OVP, bool XNEG> struct Chain32 {
    static constexpr int NTP = NT + 1;
    uint32_t X[NT + 1], Xn[SYM ? NT + 1 : 1], O[NT + 1];
    uint32_t xovp, xovpn, Xe, Oe;
    float mx;

    __device__ __forceinline__ void load(const uint32_t *tab, uint32_t, uint32_t xo, uint32_t xon) {
        load_words<NT + 1>(X, tab);
        if constexpr (SYM) load_words<NT + 1>(Xn, tab + NTP);
        load_words<NT + 1>(O, tab + 2 * NTP);
        if constexpr (XNEG) {
            const uint4 me = reinterpret_cast<const uint4 *>(tab + 7 * NTP)[2];
            Xe = me.x; Oe = me.y;
        }
        xovp = xo; xovpn = xon;
        mx = 0.0f;
    }
    __device__ __forceinline__ float one(float x, bool &outlier) {
        const float ax = fabsf(x);
        const float a = SYM ? ax : x;
        const bool neg = SYM && (__float_as_int(x) < 0);
        asm("max.NaN.f32 %0, %0, %1;" : "+f"(mx) : "f"(ax));          // NaN-propagating running max
        float q = __uint_as_float(O[0]);
        bool m0 = false;
#pragma unroll
        for (int i = 0; i < NT; i++) {
            bool m;
            if constexpr (SYM) m = a >= __uint_as_float(neg ? Xn[i] : X[i]);
            else m = a >= __uint_as_float(X[i]);
            if (i == 0) m0 = m;
            q = m ? __uint_as_float(O[i + 1]) : q;
        }
        if (SYM && m0) q = __uint_as_float(__float_as_uint(q) | (__float_as_uint(x) & 0x80000000u));
        if (XNEG) q = x < __uint_as_float(Xe) ? __uint_as_float(Oe) : q;
        if (OVP) {
            if constexpr (SYM) outlier = a >= __uint_as_float(neg ? xovpn : xovp);
            else outlier = a >= __uint_as_float(xovp);
        }
        return q;
    }
    __device__ __forceinline__ uint4 vec(const uint4 r, int debug) {
        if (debug & 2) return r;
        const float xv[4] = {__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z), __uint_as_float(r.w)};
        float qv[4];
        bool ol[4] = {false, false, false, false};
#pragma unroll
        for (int k = 0; k < 4; k++) qv[k] = one(xv[k], ol[k]);
        if (OVP) {
#pragma unroll
            for (int k = 0; k < 4; k += 2) {
                if (ol[k]) qv[k + 1] = 0.0f;
                else if (ol[k + 1]) qv[k] = 0.0f;
            }
        }
        return make_uint4(__float_as_uint(qv[0]), __float_as_uint(qv[1]), __float_as_uint(qv[2]),
                          __float_as_uint(qv[3]));
    }
    __device__ __forceinline__ bool special(uint32_t xlim) const { return !(mx <= __uint_as_float(xlim)); }
};

// One chunk from shared memory to global memory: two vectors per lane per iteration, then the remainder.
template <typename CH>
__device__ __forceinline__ bool run_chunk(CH &ch, const uint32_t *tab, const uint4 *sv, uint4 *og, int nvec, int lane,
                                          uint32_t s2, uint32_t xo, uint32_t xon, uint32_t xlim, int debug) {
    ch.load(tab, s2, xo, xon);
    const uint4 *sp = sv + lane;
    uint4 *op = og + lane;
    const int nfull = nvec >> 6;
#pragma unroll 1
    for (int it = 0; it < nfull; ++it) {
        const uint4 r0 = sp[0], r1 = sp[32];                   // LDS.128, conflict free
        const uint4 q0 = ch.vec(r0, debug);
        const uint4 q1 = ch.vec(r1, debug);
        antq_stg_stream(op, q0);
        antq_stg_stream(op + 32, q1);
        sp += 64; op += 64;
    }
#pragma unroll 1
    for (int v = (nfull << 6) + lane; v < nvec; v += 32) {
        const uint4 q0 = ch.vec(*sp, debug);
        antq_stg_stream(op, q0);
        sp += 32; op += 32;
    }
    return ch.special(xlim);
}

// OliVe 4-bit, 16-bit types: signed = 7 normal + 7 outlier thresholds on |x| (NT1 = 7, NT2 = 15), unsigned = 15 + 15
// thresholds on x (NT1 = 15, NT2 = 31).  Outliers are rare (|x| beyond ~3 sigma), but the full chain + pair masking for
// every element costs twice the normal chain.  Two phases instead:
//   1. every vector runs the NORMAL chain (exactly the ANT flint/int path; no pair can be masked if no
//      element of the vector is an outlier) and remembers whether its max |x| reaches the first outlier threshold;
//   2. the few vectors that hold an outlier are compacted into a per-warp list in shared memory and redone -- densely,
//      one vector per lane -- with the full 14-threshold chain and the outlier-victim mask, from the staged copy.
template <typename T, int NT1, int NT2, bool SYM, int MODE1, bool TIES2>
__device__ __forceinline__ bool run_chunk_ovp2(const uint32_t *tab, const uint4 *sv, uint4 *og, int nvec, int lane,
                                               uint32_t s2, uint32_t xo, uint32_t xon, uint32_t xlim,
                                               unsigned char *list, int debug) {
    typedef typename Pack2<T>::v2 v2;
    Chain16<T, NT1, SYM, false, false, MODE1, NT2 + 1> c1;
    c1.load(tab, s2, xo, xon);
    c1.E[NT1] = tab[2 * (NT2 + 1) + NT2];                           // O[0] lives in slot NT2 of the full-pitch table
    const v2 xo2 = Pack2<T>::from_u32(xo);                          // positive-side threshold: <= the negative-side one
    uint32_t pend = 0;                                              // bit j: vector j * 32 + lane holds an outlier
    const uint4 *sp = sv + lane;
    uint4 *op = og + lane;
    const int nfull = nvec >> 6;
#pragma unroll 1
    for (int it = 0; it < nfull; ++it) {
        const uint4 r0 = sp[0], r1 = sp[32];
        const v2 m0 = c1.vec_max(r0), m1 = c1.vec_max(r1);
        c1.mx = __hmax2_nan(__hmax2_nan(c1.mx, m0), m1);
        const uint4 q0 = c1.template vec<false>(r0, debug);
        const uint4 q1 = c1.template vec<false>(r1, debug);
        if (__hge2_mask(m0, xo2)) pend |= 1u << (2 * it);
        if (__hge2_mask(m1, xo2)) pend |= 2u << (2 * it);
        antq_stg_stream(op, q0);
        antq_stg_stream(op + 32, q1);
        sp += 64; op += 64;
    }
#pragma unroll 1
    for (int v = (nfull << 6) + lane; v < nvec; v += 32) {
        const uint4 r0 = *sp;
        const v2 m0 = c1.vec_max(r0);
        c1.mx = __hmax2_nan(c1.mx, m0);
        const uint4 q0 = c1.template vec<false>(r0, debug);
        if (__hge2_mask(m0, xo2)) pend |= 1u << (v >> 5);
        antq_stg_stream(op, q0);
        sp += 32; op += 32;
    }
    const bool special = c1.special(xlim);
    if (__any_sync(0xffffffffu, pend != 0) && !(debug & 2)) {
        // compact the flagged vectors of the warp into `list` (index < 256 fits a byte)
        const int cnt = __popc(pend);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        int pos = incl - cnt;
        while (pend) {
            const int j = __ffs(pend) - 1;
            pend &= pend - 1;
            list[pos++] = (unsigned char)(j * 32 + lane);
        }
        __syncwarp();          // the list is complete; phase-1 stores to these vectors are ordered before the rewrites
        Chain16<T, NT2, SYM, true, false, TIES2 ? kModeTies : kModeAlu> c2;
        c2.load(tab, s2, xo, xon);
#pragma unroll 1
        for (int i = lane; i < total; i += 32) {
            const int v = list[i];
            antq_stg_stream(og + v, c2.template vec<false>(sv[v], debug));
        }
        __syncwarp();          // the list may be rewritten by this warp's next chunk
    }
    return special;
}

__device__ __forceinline__ unsigned antqs_ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(antq_smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void antqs_st_release(unsigned *p, unsigned v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(antq_smem_u32(p)), "r"(v) : "memory");
}

template <typename T, int NT, bool SYM, bool OVP, bool XNEG>
__global__ void __launch_bounds__(kThreads, 1) antq_stream_kernel(const StreamParams p) {
    typedef AntqType<T> A;
    typedef TabGeom<NT> TG;
    constexpr int VEC = A::kVec;
    constexpr int NTP = TG::NTP;
    constexpr int G = TG::G;
    constexpr int kTabBytes = TG::kTabBytes;
    extern __shared__ __align__(128) unsigned char antqs_smem[];
    unsigned char *ring = antqs_smem + (size_t)kNS * kChunkMax;                // kRT row-table slots
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)kRT * kTabBytes);
    unsigned *built = reinterpret_cast<unsigned *>(full + kNS);               // [kRT] lap + 1 of the group in each ring slot
    unsigned *cons_row = built + kRT;                                          // [kNC] CTA-local row each consumer is on
    unsigned *next_k = cons_row + kNC;                                         // next chunk to hand out
    unsigned char *ovp_list = reinterpret_cast<unsigned char *>(next_k + 4);   // [kNC][256] two-phase OVP work lists

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // equal contiguous share of the chunk list (32-bit: the launcher refuses tensors of more than 2^31 chunks)
    const unsigned c_begin = blockIdx.x * p.chunks_per_cta + min(blockIdx.x, p.chunks_rem);
    const int n = (int)p.chunks_per_cta + (blockIdx.x < p.chunks_rem ? 1 : 0);
    const unsigned cpr = (unsigned)p.chunks_per_row;
    const unsigned row_begin = c_begin / cpr;
    const AntqCodebook *__restrict__ cb = p.cb;
    static_assert(kRT + kNC + 1 <= kThreads, "one thread per flag word at start-up");
    static_assert(kRing == 2, "each consumer warp double-buffers");

    if (warp == 0) ANTQS_TR(0);
    // Programmatic dependent launch: the next kernel in the stream may be scheduled onto SMs as our CTAs leave them ...
    asm volatile("griddepcontrol.launch_dependents;");
    if (threadIdx.x < kNS) antq_mbar_init(full + threadIdx.x, 1);
    if (threadIdx.x < kRT + kNC + 1) built[threadIdx.x] = 0;                   // built[], cons_row[], next_k are contiguous
    __syncthreads();
    // ... and everything above overlapped the tail of the previous kernel; nothing it may have written (x, alpha, the
    // codebook) or may still be reading (out) is touched before it has completed and flushed.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (warp == 0) ANTQS_TR(1);

    // ---- row tables: G rows per build pass, kept in a ring of kRT rows = GS group slots ----
    constexpr int GS = kRT / G;
    const unsigned last_row = (unsigned)(p.rows - 1);
    const unsigned row_last = (c_begin + (unsigned)n - 1u) / cpr;
    const int ngroups = p.alpha_per_row ? (int)((row_last - row_begin) / G + 1) : 1;
    // Groups [0, n_pro) are built in the PROLOGUE, one group per warp, all at once: the
    // consumers have nothing to do until the first chunk lands, and a builder warp that has to share its scheduler
    // with three busy consumers is ~4x slower per pass than at start-up (profiles/r01_notes.md).  Later groups (a CTA
    // with more than kRT rows) are made by the builder warps as the consumers free ring slots.
    const int n_pro = min(ngroups, min(GS, kNC + kNB));
    struct BuildIn { float alpha, tpos, tneg, lev, thr_e, lev_e; };
    auto build_inputs = [&](int g) {                                  // the global loads of one build pass
        const int sub = lane % NTP, grp = lane / NTP;
        const int nt_real = p.nt_real;
        const float inf = __int_as_float(0x7f800000);
        unsigned r = row_begin + (unsigned)(g * G + grp);
        r = r > last_row ? last_row : r;
        BuildIn in;
        in.alpha = __ldg(p.alpha + (p.alpha_per_row ? r : 0));
        in.tpos = sub < nt_real ? (SYM ? cb->mag_tpos[sub] : cb->thr[sub]) : inf;
        in.tneg = (SYM && sub < nt_real) ? cb->mag_tneg[sub] : in.tpos;
        in.lev = sub <= nt_real ? (SYM ? cb->level[p.mid + sub] : cb->level[sub]) : 0.0f;
        in.thr_e = XNEG ? cb->thr[0] : 0.0f;
        in.lev_e = XNEG ? cb->level[0] : 0.0f;
        return in;
    };
    auto build_group = [&](int g, const BuildIn &in, bool tr) {
        const int sub = lane % NTP, grp = lane / NTP;
        const float alpha = in.alpha, tpos = in.tpos, tneg = in.tneg, lev = in.lev;
        if (tr) ANTQS_TR(3);
        RowTab tab;
        build_tables<T, NT, SYM, OVP, XNEG>(tab, p, alpha, tpos, tneg, lev, in.thr_e, in.lev_e, sub, grp, tr);
        uint32_t *st = reinterpret_cast<uint32_t *>(ring + (size_t)((g * G + grp) & (kRT - 1)) * kTabBytes);
        st[sub] = tab.X;
        st[NTP + sub] = tab.Xn;
        st[2 * NTP + sub] = tab.T2;
        st[3 * NTP + sub] = tab.B;
        st[4 * NTP + sub] = tab.C;
        st[5 * NTP + sub] = tab.D;
        st[6 * NTP + sub] = tab.Cn;
        if (sub == 0) {
            uint4 *m = reinterpret_cast<uint4 *>(st + 7 * NTP);
            m[0] = make_uint4(tab.s_bits, tab.flags, tab.xlim, tab.S2);
            m[1] = make_uint4(tab.xovp, tab.xovpn, 0u, 0u);
            m[2] = make_uint4(tab.Xe, tab.Ee, tab.De, 0u);
        }
        __syncwarp();
        if (lane == 0) antqs_st_release(built + (g % GS), (unsigned)(g / GS) + 1u);   // release: tables visible
        if (tr) ANTQS_TR(4);
    };
    // Chunk geometry and the TMA request of chunk k into stage `stage` (lane 0 of the owning consumer warp).
    struct Geo { long long base; int nvec, tail; unsigned row; };
    auto geo_of = [&](int k) {
        Geo g;
        const unsigned c = c_begin + (unsigned)k;
        g.row = p.cpr_shift >= 0 ? c >> p.cpr_shift : c / cpr;        // chunks per row is usually a power of two
        const long long col0 = (long long)(c - g.row * cpr) * p.chunk_elems;
        const long long remain = p.cols - col0;
        const int n_el = (int)(remain < p.chunk_elems ? remain : p.chunk_elems);
        g.nvec = n_el / VEC;
        g.tail = n_el - g.nvec * VEC;
        g.base = (long long)g.row * p.cols + col0;
        return g;
    };
    auto request = [&](const Geo &g, int stage, int k) {
        if (lane == 0) {
            const unsigned bytes = (unsigned)g.nvec * 16u;
            if (bytes) {
                antq_fence_proxy_async();          // this warp's LDS reads of the stage precede the async-proxy write
                antq_bulk_g2s(antqs_smem + (size_t)stage * kChunkMax, reinterpret_cast<const T *>(p.x) + g.base, bytes,
                              full + stage);
            } else {
                antqs_mbar_arrive(full + stage);
            }
#ifdef ANTQS_TRACE
            if (p.trace && k < kTraceChunks) p.trace[(size_t)blockIdx.x * kTraceStride + 8 + 4 * k] = antqs_now();
#endif
        }
    };
    // Consumer warp w owns two private stages (w and w + kNC): the chunk after the one being computed is always in
    // flight, so an SM has kNC x 4 KiB outstanding -- the HBM latency-bandwidth product -- and no more (requesting a whole
    // deep ring at start-up makes every SM's FIRST chunk queue behind ~30 MB of other requests: profiles/r01_notes.md).
    // No issuer warp, no `empty` barriers: the warp that frees a stage is the one that refills it.
    // Chunks are handed out by a shared counter at the moment a warp REQUESTS them (one ahead of the one it computes):
    // rows with ties or out-of-window values cost more than others, and a CTA's chunk count rarely divides by kNC.
    auto claim = [&]() {
        int k = 0;
        if (lane == 0) k = (int)atomicAdd(next_k, 1u);
        return __shfl_sync(0xffffffffu, k, 0);
    };
    BuildIn pin;
    pin.alpha = 0.0f; pin.tpos = 0.0f; pin.tneg = 0.0f; pin.lev = 0.0f; pin.thr_e = 0.0f; pin.lev_e = 0.0f;
    if (warp < n_pro) pin = build_inputs(warp);    // small loads first: they would queue behind the bulk copies
    Geo cur;
    cur.base = 0; cur.nvec = 0; cur.tail = 0; cur.row = 0;
    int k = n;
    if (warp < kNC) {
        k = claim();
        if (k < n) {
            cur = geo_of(k);
            request(cur, warp, k);                 // first chunk in flight before the tables are made
        }
    }
    if (warp < n_pro) build_group(warp, pin, warp == 0);   // consumers 0.., then the builders

    if (warp >= kNC) {
        // ------------------------------ table builders ------------------------------
        // Builder b makes the tables of row groups n_pro + b, n_pro + b + kNB, ... (G rows per pass, one lane per
        // threshold) and stores them in the row-table ring; the builders run ahead of the consumers by up to kRT rows.
        const int b = warp - kNC;
        for (int g = n_pro + b; g < ngroups; g += kNB) {
            const int r0 = g * G;                                     // first CTA-local row of the group
            if (r0 + G > kRT) {
                // the slot of row r is reused by row r + kRT: wait until every consumer has moved past row r0 + G - 1 - kRT
                const unsigned need = (unsigned)(r0 + G - kRT);       // all consumers must be on a row >= need
                for (;;) {
                    unsigned v = lane < kNC ? antqs_ld_acquire(cons_row + lane) : 0xffffffffu;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
                    if (v >= need) break;
                    __nanosleep(200);
                }
            }
            build_group(g, build_inputs(g), false);
        }
    } else {
        // ------------------------------ consumers ------------------------------
        int slot = 0;                                                  // which of this warp's two stages holds chunk k
        unsigned phases = 0;                                           // bit r = parity to wait for on slot r
        while (k < n) {
            const int stage = warp + slot * kNC;
            const Geo g = cur;
            // the stage freed by the previous iteration takes the next chunk
            const int kn = claim();
            if (kn < n) {
                cur = geo_of(kn);
                request(cur, warp + (slot ^ 1) * kNC, kn);
            }
            const unsigned row = g.row;
            const unsigned rl = p.alpha_per_row ? row - row_begin : 0u;        // CTA-local row = table index
            // builders may recycle the slots of rows < rl: a plain store is enough -- every table value this warp read
            // for an earlier row has long been consumed, so there is nothing for a release fence to order
            if (lane == 0) *reinterpret_cast<volatile unsigned *>(cons_row + warp) = rl;
            {
                const unsigned g = rl / G;
                const unsigned *flag = built + (g % GS);
                const unsigned need = g / GS + 1u;
                while (antqs_ld_acquire(flag) < need) __nanosleep(100);         // only at start-up in practice
            }
            ANTQS_TRK(k, 1);
            const uint32_t *tab = reinterpret_cast<const uint32_t *>(ring + (size_t)(rl & (kRT - 1)) * kTabBytes);
            const uint4 m0 = reinterpret_cast<const uint4 *>(tab + 7 * NTP)[0];
            const uint4 m1 = reinterpret_cast<const uint4 *>(tab + 7 * NTP)[1];
            const float s = __uint_as_float(m0.x);
            const uint32_t flags = m0.y, xlim = m0.z;
            const int nvec = g.nvec, tail = g.tail;
            const long long base = g.base;
            const uint4 *sv = reinterpret_cast<const uint4 *>(antqs_smem + (size_t)stage * kChunkMax);
            T *og = reinterpret_cast<T *>(p.out) + base;
            antq_mbar_wait(full + stage, (phases >> slot) & 1u);       // the chunk has landed in shared memory
            phases ^= 1u << slot;
            ANTQS_TRK(k, 2);
            bool special = !(flags & kRowOk);
            if (nvec > 0 && !special) {
                uint4 *ov = reinterpret_cast<uint4 *>(og);
                bool done = false;
                if constexpr (sizeof(T) == 2 && OVP && SYM && NT == 15) {
                    if (p.ovp_index == 7 && !(p.debug & 128)) {
                        // signed 4-bit OliVe: normal chain for everything, full chain only for the vectors with an outlier
                        unsigned char *list = ovp_list + warp * 256;
                        if ((flags & kRowTies) && (flags & kRowFma))
                            special = run_chunk_ovp2<T, 7, 15, true, kModeTiesMix, true>(tab, sv, ov, nvec, lane, m0.w, m1.x, m1.y, xlim, list, p.debug);
                        else if (flags & kRowTies)
                            special = run_chunk_ovp2<T, 7, 15, true, kModeTies, true>(tab, sv, ov, nvec, lane, m0.w, m1.x, m1.y, xlim, list, p.debug);
                        else if (flags & kRowFma)
                            special = run_chunk_ovp2<T, 7, 15, true, kModeMix, false>(tab, sv, ov, nvec, lane, m0.w, m1.x, m1.y, xlim, list, p.debug);
                        else
                            special = run_chunk_ovp2<T, 7, 15, true, kModeAlu, false>(tab, sv, ov, nvec, lane, m0.w, m1.x, m1.y, xlim, list, p.debug);
                        done = true;
                    }
                }
                if constexpr (sizeof(T) == 2 && OVP && !SYM && NT == 31) {
                    if (p.ovp_index == 15 && !(p.debug & 128)) {
                        // unsigned 4-bit OliVe (post-ReLU inputs): 15 normal thresholds first, all 30 only where needed
                        unsigned char *list = ovp_list + warp * 256;
                        if (flags & kRowFma)
                            special = run_chunk_ovp2<T, 15, 31, false, kModeMix, false>(tab, sv, ov, nvec, lane, m0.w, m1.x, m1.y, xlim, list, p.debug);
                        else
                            special = run_chunk_ovp2<T, 15, 31, false, kModeAlu, false>(tab, sv, ov, nvec, lane, m0.w, m1.x, m1.y, xlim, list, p.debug);
                        done = true;
                    }
                }
                if (done) {
                } else if constexpr (sizeof(T) == 2) {
                    if ((flags & kRowTies) && (flags & kRowFma)) {
                        Chain16<T, NT, SYM, OVP, XNEG, (SYM && NT <= 7 ? kModeTiesMix : kModeTies)> ch;
                        special = run_chunk(ch, tab, sv, ov, nvec, lane, m0.w, m1.x, m1.y, xlim, p.debug);
                    } else if (flags & kRowTies) {
                        Chain16<T, NT, SYM, OVP, XNEG, kModeTies> ch;
                        special = run_chunk(ch, tab, sv, ov, nvec, lane, m0.w, m1.x, m1.y, xlim, p.debug);
                    } else if (flags & kRowFma) {
                        Chain16<T, NT, SYM, OVP, XNEG, (NT <= 15 ? kModeMix : kModeAlu)> ch;
                        special = run_chunk(ch, tab, sv, ov, nvec, lane, m0.w, m1.x, m1.y, xlim, p.debug);
                    } else {
                        Chain16<T, NT, SYM, OVP, XNEG, kModeAlu> ch;
                        special = run_chunk(ch, tab, sv, ov, nvec, lane, m0.w, m1.x, m1.y, xlim, p.debug);
                    }
                } else {
                    Chain32<NT, SYM, OVP, XNEG> ch;
                    special = run_chunk(ch, tab, sv, ov, nvec, lane, m0.w, m1.x, m1.y, xlim, p.debug);
                }
            }
            if (nvec > 0 && __any_sync(0xffffffffu, special)) {
                float xl;
                if constexpr (sizeof(T) == 2) xl = A::to_f32(A::from_bits((typename A::bits_t)(xlim & 0xffffu)));
                else xl = __uint_as_float(xlim);
                antqs_fixup_chunk<T, OVP>(cb, s, xl, !(flags & kRowOk), sv, og, nvec, lane);
            }
            // ragged tail (only a per-tensor view can have one: rows == 1): straight from global memory
            if (tail > 0 && lane == 0) {
                const T *xg = reinterpret_cast<const T *>(p.x) + base + (long long)nvec * VEC;
                antqs_slow_vec<T, OVP>(cb, s, xg, og + (long long)nvec * VEC, tail);
            }
            __syncwarp();                                              // every lane is done with this stage
            ANTQS_TRK(k, 3);
            k = kn;
            slot ^= 1;
        }
        if (lane == 0) antqs_st_release(cons_row + warp, 0xffffffffu);  // done: never holds a table slot again
#ifdef ANTQS_TRACE
        if (p.trace && lane == 0) atomicMax(p.trace + (size_t)blockIdx.x * kTraceStride + 2, antqs_now());
#endif
    }
}

template <typename T, int NT, bool SYM, bool OVP, bool XNEG = false>
int launch_kernel(const StreamParams &p, int ctas, cudaStream_t st) {
    auto kernel = antq_stream_kernel<T, NT, SYM, OVP, XNEG>;
    const int smem = kNS * kChunkMax + kRT * TabGeom<NT>::kTabBytes + kNS * 8 + (kRT + kNC + 4) * 4 + kNC * 256 + 16;
    // The opt-in is per device (context): one bit per device ordinal, per instantiation.
    static unsigned long long configured = 0ull;
    int dev = 0;
    cudaError_t ed = cudaGetDevice(&dev);
    if (ed != cudaSuccess) return (int)ed;
    if (dev >= 64 || !((configured >> dev) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        if (dev < 64) configured |= 1ull << dev;
    }
    static int pdl = -1;
    if (pdl < 0) { const char *v = getenv("ANTQ_PDL"); pdl = (v && atoi(v) == 0) ? 0 : 1; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, p);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess)
        fprintf(stderr, "antq: stream kernel launch failed: %s (grid %d, block %d, smem %d, chunks %u)\n",
                cudaGetErrorString(e), ctas, kThreads, smem, p.total_chunks);
    return (int)e;
}

template <typename T, bool SYM, bool OVP> int launch_nt(const StreamParams &p, int nt, int ctas, cudaStream_t st) {
#ifdef ANTQS_DEV_ONLY_NT7
    if (nt <= 7) return launch_kernel<T, 7, SYM, OVP>(p, ctas, st);
    return ANTQ_ENOTSUP;
#else
    if (nt <= 3) return launch_kernel<T, 3, SYM, OVP>(p, ctas, st);
    if (nt <= 7) return launch_kernel<T, 7, SYM, OVP>(p, ctas, st);
    if (nt <= 15) return launch_kernel<T, 15, SYM, OVP>(p, ctas, st);
    if (nt <= 31) return launch_kernel<T, 31, SYM, OVP>(p, ctas, st);
    return ANTQ_ENOTSUP;
#endif
}

// signed int-k: symmetric magnitudes plus one extra negative level
template <typename T> int launch_symx(const StreamParams &p, int nt, int ctas, cudaStream_t st) {
#ifndef ANTQS_DEV_ONLY_NT7
    if (nt <= 3) return launch_kernel<T, 3, true, false, true>(p, ctas, st);
#endif
    if (nt <= 7) return launch_kernel<T, 7, true, false, true>(p, ctas, st);
#ifndef ANTQS_DEV_ONLY_NT7
    if (nt <= 15) return launch_kernel<T, 15, true, false, true>(p, ctas, st);
    if (nt <= 31) return launch_kernel<T, 31, true, false, true>(p, ctas, st);
#endif
    return ANTQ_ENOTSUP;
}

template <typename T> int launch_t(const StreamParams &p, int nt, bool sym, bool ovp, int ctas, cudaStream_t st) {
    if (sym) return ovp ? launch_nt<T, true, true>(p, nt, ctas, st) : launch_nt<T, true, false>(p, nt, ctas, st);
    return ovp ? launch_nt<T, false, true>(p, nt, ctas, st) : launch_nt<T, false, false>(p, nt, ctas, st);
}

}  // namespace

static unsigned long long *antqs_trace_buffer = nullptr;
// Tuning builds (-DANTQS_TRACE) record a per-CTA timeline into this device buffer (148 x 264 u64); NULL = off.
extern "C" void antq_debug_stream_trace(void *device_buffer) { antqs_trace_buffer = (unsigned long long *)device_buffer; }

// Returns ANTQ_ENOTSUP for configurations this kernel does not cover (the caller falls back to the generic kernel).
int antq_launch_stream(const void *x, void *out, const float *alpha, int alpha_per_row, long long rows, long long cols,
                       int dtype, const AntqCodebook *cb, const antq_codebook_info *info, bool ovp, cudaStream_t st) {
    const bool symx = (info->flags & ANTQ_CB_SYMX) != 0 && !ovp && !(info->flags & ANTQ_CB_SYMMETRIC);
    const bool sym = (info->flags & ANTQ_CB_SYMMETRIC) != 0;
    const int nt = (sym || symx) ? info->n_mag - 1 : info->n_levels - 1;
    const int es = dtype == ANTQ_F32 ? 4 : 2;
    static int dbg = -1, chunk_env = 0;
    if (dbg < 0) {
        const char *e = getenv("ANTQ_DEBUG");
        const char *c = getenv("ANTQ_CHUNK");
        chunk_env = c ? atoi(c) : 0;
        dbg = e ? atoi(e) : 0;
    }
    StreamParams p;
    p.x = x; p.out = out; p.alpha = alpha; p.cb = cb;
    p.rows = rows; p.cols = cols;
    // chunk size: the largest power of two <= kChunkMax that still gives every consumer warp of every SM a chunk
    int chunk_bytes = kChunkMax;
    if (chunk_env >= 512 && chunk_env <= kChunkMax && (chunk_env & (chunk_env - 1)) == 0) {
        chunk_bytes = chunk_env;
    } else {
        const long long want = (long long)antq_num_sms() * kNC * 2;
        while (chunk_bytes > 1024) {
            const long long ce = chunk_bytes / es;
            if (rows * ((cols + ce - 1) / ce) >= want) break;
            chunk_bytes >>= 1;
        }
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
    p.nt_real = nt; p.mid = info->mid; p.ovp_index = info->ovp_index; p.n_entries = info->n_entries;
    p.gmax = info->gmax; p.lim = info->lim;
    p.debug = dbg;
    p.trace = antqs_trace_buffer;
    const unsigned sms = (unsigned)antq_num_sms();
    const int ctas = (int)(p.total_chunks < sms ? p.total_chunks : sms);
    p.chunks_per_cta = p.total_chunks / (unsigned)ctas;
    p.chunks_rem = p.total_chunks % (unsigned)ctas;
    if (symx) {
        switch (dtype) {
            case ANTQ_F32: return launch_symx<float>(p, nt, ctas, st);
            case ANTQ_F16: return launch_symx<__half>(p, nt, ctas, st);
            case ANTQ_BF16: return launch_symx<__nv_bfloat16>(p, nt, ctas, st);
        }
    }
    switch (dtype) {
        case ANTQ_F32: return launch_t<float>(p, nt, sym, ovp, ctas, st);
        case ANTQ_F16: return launch_t<__half>(p, nt, sym, ovp, ctas, st);
        case ANTQ_BF16: return launch_t<__nv_bfloat16>(p, nt, sym, ovp, ctas, st);
    }
    return ANTQ_EINVAL;
}
