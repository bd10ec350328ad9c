// antq_rows.cu -- the hot kernel: fused fake-quant forward with per-row tables.
//
// Replaces A/antquant/quant_modules.py:535-551 (+ O/...:295-330 with OVP) and the
// scan kernel it calls (A/quant/quant_kernel.cu:11-39): ~9 launches and 7 reads +
// 7 writes of the tensor become ONE launch, one read, one write.
//
// Arithmetic (proved bit-exact on exhaustive fp16 inputs in tests/test_xspace_model.py):
// the reference computes d = fl32(x / s), picks the grid level by a scan, and
// returns fl32(((q - d) + d) * s).  fl32(x / s) is monotone in x, so every
// d-space threshold of the prepared codebook maps to an exact x-space threshold
//      X_r = min{ x in dtype : fl32(x / s) >= thr_r }
// and inside the window |d| <= lim the STE sum is exact, so the result is simply
// O_j = RN_dtype(fl32(level_j * s)).  Per 32-bit register (two fp16 values) the work
// is one packed compare (HSET2) + one LOP3 per threshold -- no division, no lookup.
// Values outside the window, NaN/Inf and rows whose scale is not a positive finite
// number take the literal reference arithmetic (antq_slow_vec): rare, exact.
//
// Execution shape (B200): a PERSISTENT grid of 148 x 3 CTAs x 4 warps.  The tensor
// is cut into 4 KiB chunks that never straddle a row; every warp owns an equal,
// contiguous range of chunks (static balance, no tail wave) and walks it with a
// double-buffered ring of TMA bulk copies (cp.async.bulk -> shared memory, completion
// on an mbarrier), so the next chunk is in flight while the current one is computed,
// without holding registers.  When the row changes the warp rebuilds its tables: lane r
// owns threshold r, its codebook values stay in registers for the whole kernel and the
// next row's alpha is prefetched.  Half of the pairs use the ALU-pipe formulation
// (HSET2 + LOP3), the other half an exact FMA-pipe formulation (HFMA2.SAT compare +
// HFMA2 accumulate), because both pipes are half-rate and the ALU pipe alone was the
// limiter (profiles/).
#include <stdio.h>
#include <stdlib.h>

#include "antq_common.cuh"

namespace {

#ifndef ANTQ_FMA_PAIRS
#define ANTQ_FMA_PAIRS 2      // pairs per vector (of 4) that take the FMA-pipe formulation
#endif
#ifndef ANTQ_LOOP_VECS
#define ANTQ_LOOP_VECS 2      // vectors per lane per loop iteration
#endif
constexpr int kWarpsPerCta = 4;
#ifndef ANTQ_CTAS
#define ANTQ_CTAS 3
#endif
constexpr int kCtasPerSm = ANTQ_CTAS;
constexpr int kNumSms = 148;
#ifndef ANTQ_RING
#define ANTQ_RING 2      // measured best on 4096x4096 fp16 (profiles/r01_tuning.md): deeper rings front-load the
#endif                   // whole tensor's requests and delay each warp's first chunk
#ifndef ANTQ_CHUNK
#define ANTQ_CHUNK 4096
#endif
constexpr int kRing = ANTQ_RING;         // chunks in flight + in use per warp (power of two)
constexpr int kChunkBytes = ANTQ_CHUNK;  // 4096 = 256 vectors = 8 per lane
constexpr int kChunkVecs = kChunkBytes / 16;
constexpr int kVecPerLane = kChunkVecs / 32;
constexpr int kTableWords = 32;          // per table, per warp (uint32 / float)
constexpr int kWarpSmemBytes = kRing * kChunkBytes + 6 * kTableWords * 4 + 128;   // ring | X | Xn | O | B | C | D | mbarriers

template <typename T, bool OVP>
__device__ __noinline__ void antq_slow_vec(const AntqCodebook *__restrict__ cb, float s, const T *xg, T *og,
                                           int16_t *cg, int n) {
    typedef AntqType<T> A;
    float q[8], d[8];
    int c[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
        if (e < n) {
            AntqExact ex = antq_exact_quant(cb, A::to_f32(xg[e]), s);
            q[e] = ex.q; d[e] = ex.d; c[e] = ex.code;
        }
    }
    if (OVP) {
        const int K = cb->n_entries;
#pragma unroll
        for (int e = 0; e + 1 < 8; e += 2) {
            if (e + 1 < n) {
                bool oe = fabsf(q[e]) > 32.0f, oo = fabsf(q[e + 1]) > 32.0f;
                if (oe) { q[e + 1] = __fmul_rn(q[e + 1], 0.0f); c[e + 1] = K; }
                else if (oo) { q[e] = __fmul_rn(q[e], 0.0f); c[e] = K; }
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; e++) {
        if (e < n) {
            og[e] = A::from_f32_rn(antq_ste_rescale(q[e], d[e], s));
            if (cg) cg[e] = (int16_t)c[e];
        }
    }
}

// Cold pass over a chunk in which some vector was skipped by the fast path: the same window
// predicate is re-evaluated in fp32 and exactly the skipped vectors get the reference arithmetic.
// (The fast path stores nothing for them, so this also works in place.)  `all` = row scale invalid.
template <typename T, bool OVP>
__device__ __noinline__ void antq_fixup_chunk(const AntqCodebook *__restrict__ cb, float s, float xlim, bool all,
                                              const T *xg, T *og, int16_t *cg, int nvec, int lane) {
    typedef AntqType<T> A;
    constexpr int VEC = A::kVec;
    for (int v = lane; v < nvec; v += 32) {
        const T *xv = xg + (long long)v * VEC;
        bool special = all;
#pragma unroll
        for (int e = 0; e < VEC; e++) special |= !(fabsf(A::to_f32(xv[e])) <= xlim);
        if (special)
            antq_slow_vec<T, OVP>(cb, s, xv, og + (long long)v * VEC, cg ? cg + (long long)v * VEC : nullptr, VEC);
    }
}

template <bool SYM>
__device__ __forceinline__ int16_t antq_rank_to_code(const AntqCodebook *__restrict__ cb, int mid, int rank, bool neg) {
    int lvl = SYM ? (mid + (neg ? -rank : rank)) : rank;
    return (int16_t)cb->level_code[lvl];
}

// ---- packed 16-bit (fp16 / bf16) chain --------------------------------------
template <typename T> struct Pack2 {
    typedef float2 v2;     // unused for fp32
    __device__ static __forceinline__ v2 from_u32(uint32_t) { return make_float2(0.f, 0.f); }
};
template <> struct Pack2<__half> {
    typedef __half2 v2;
    __device__ static __forceinline__ v2 from_u32(uint32_t u) { return *reinterpret_cast<v2 *>(&u); }
};
template <> struct Pack2<__nv_bfloat16> {
    typedef __nv_bfloat162 v2;
    __device__ static __forceinline__ v2 from_u32(uint32_t u) { return *reinterpret_cast<v2 *>(&u); }
};

// Word a lane publishes for its table entry: 16-bit types are duplicated into both halves.
template <typename T> __device__ __forceinline__ uint32_t antq_table_word(T v) {
    if constexpr (sizeof(T) == 2) {
        const uint32_t b = AntqType<T>::bits(v);
        return b | (b << 16);
    } else {
        return __float_as_uint(v);
    }
}

template <typename T, int NT, bool SYM, bool OVP, bool CODES> struct Tables {
    static constexpr bool IS16 = sizeof(T) == 2;
    uint32_t X[NT];       // thresholds for x >= 0 (or every x when !SYM)
    uint32_t Xn[NT];      // SYM: thresholds on |x| for x < 0 -- differ from X only where a tie is representable
                          // (ties go to the LATER grid entry: up for positive d, toward zero for negative d)
    uint32_t O[NT + 1];   // dequantised outputs per level / magnitude
    uint32_t xlim, xovp, xovpn;
    // FMA-pipe twin of the compare (16-bit types): with P = prev(X), u = X - P (a power of two),
    //   m = sat((S*a) * (1/(u*S)) - P/u)  is exactly 1.0 for a >= X and exactly 0.0 for a <= P,
    // because the fused product-sum is <= 0 or >= 1 before its single rounding.  S = 2^k keeps 1/(u*S) in range.
    uint32_t Bf[IS16 ? NT : 1];   // 1/(u*S), duplicated
    uint32_t Cf[IS16 ? NT : 1];   // -P/u,    duplicated
    uint32_t Df[IS16 ? NT : 1];   // O[i+1] - O[i]: exactly representable when successive outputs are within 2x
                                  // (checked per row), so q <- m*D + q reproduces O[rank] without rounding
    uint32_t S2;

    __device__ __forceinline__ void load_fma(const uint32_t *sB, const uint32_t *sC, const uint32_t *sD, uint32_t s2) {
        if constexpr (IS16) {
            constexpr int NX = (NT + 3) / 4;
            uint4 bb[NX], bc[NX], bd[NX];
#pragma unroll
            for (int i = 0; i < NX; i++) {
                bb[i] = reinterpret_cast<const uint4 *>(sB)[i];
                bc[i] = reinterpret_cast<const uint4 *>(sC)[i];
                bd[i] = reinterpret_cast<const uint4 *>(sD)[i];
            }
#pragma unroll
            for (int i = 0; i < NT; i++) {
                const uint4 a = bb[i / 4], b = bc[i / 4], c = bd[i / 4];
                Bf[i] = (i % 4 == 0) ? a.x : (i % 4 == 1) ? a.y : (i % 4 == 2) ? a.z : a.w;
                Cf[i] = (i % 4 == 0) ? b.x : (i % 4 == 1) ? b.y : (i % 4 == 2) ? b.z : b.w;
                Df[i] = (i % 4 == 0) ? c.x : (i % 4 == 1) ? c.y : (i % 4 == 2) ? c.z : c.w;
            }
            S2 = s2;
        }
    }

    __device__ __forceinline__ void load(const uint32_t *sX, const uint32_t *sXn, const uint32_t *sO, float lim, int oi,
                                         float s) {
        // vectorised smem reads: 32-word tables, 16-byte aligned
        constexpr int NX = (NT + 3) / 4, NO = (NT + 4) / 4;
        uint4 bx[NX], bn[NX], bo[NO];
#pragma unroll
        for (int i = 0; i < NX; i++) {
            bx[i] = reinterpret_cast<const uint4 *>(sX)[i];
            bn[i] = reinterpret_cast<const uint4 *>(sXn)[i];
        }
#pragma unroll
        for (int i = 0; i < NO; i++) bo[i] = reinterpret_cast<const uint4 *>(sO)[i];
#pragma unroll
        for (int i = 0; i < NT; i++) {
            const uint4 a = bx[i / 4], b = bn[i / 4];
            X[i] = (i % 4 == 0) ? a.x : (i % 4 == 1) ? a.y : (i % 4 == 2) ? a.z : a.w;
            Xn[i] = (i % 4 == 0) ? b.x : (i % 4 == 1) ? b.y : (i % 4 == 2) ? b.z : b.w;
        }
#pragma unroll
        for (int i = 0; i <= NT; i++) {
            const uint4 a = bo[i / 4];
            O[i] = (i % 4 == 0) ? a.x : (i % 4 == 1) ? a.y : (i % 4 == 2) ? a.z : a.w;
        }
        const float xl = __fmul_rn(__fmul_rn(lim, s), 0.9990234375f);   // conservative window in x-space
        xlim = antq_table_word<T>(AntqType<T>::from_f32_rz(xl));
        const bool has = OVP && oi >= 0 && oi < NT;
        const uint32_t inf = antq_table_word<T>(AntqType<T>::from_bits(AntqType<T>::kInf));
        xovp = has ? sX[oi] : inf;
        xovpn = has ? sXn[oi] : inf;
    }
    __device__ __forceinline__ float xlim_f32() const {
        if constexpr (IS16)
            return AntqType<T>::to_f32(AntqType<T>::from_bits((typename AntqType<T>::bits_t)(xlim & 0xffffu)));
        else
            return __uint_as_float(xlim);
    }

    // 16-bit: one 32-bit register = two elements (low half = even flat index)
    template <bool TIES>
    __device__ __forceinline__ uint32_t pair(uint32_t xb, bool &special, uint32_t &ranks, uint32_t &victims) const {
        typedef typename Pack2<T>::v2 v2;
        const v2 x2 = Pack2<T>::from_u32(xb);
        const v2 ab = __habs2(x2);
        const v2 a2 = SYM ? ab : x2;
        special |= (__hle2_mask(ab, Pack2<T>::from_u32(xlim)) != 0xffffffffu);   // NaN/Inf/out-of-window
        uint32_t neg = 0;
        if (SYM && TIES) neg = __hlt2_mask(x2, Pack2<T>::from_u32(0u));          // 0xffff where x < 0
        uint32_t q = O[0], m0 = 0, rk = 0;
#pragma unroll
        for (int i = 0; i < NT; i++) {
            const uint32_t t = (SYM && TIES) ? ((neg & Xn[i]) | (~neg & X[i])) : X[i];
            const uint32_t m = __hge2_mask(a2, Pack2<T>::from_u32(t));
            if (i == 0) m0 = m;
            q = (m & O[i + 1]) | (~m & q);
            if (CODES) rk += m & 0x00010001u;
        }
        if (SYM) q |= (xb & 0x80008000u) & m0;   // restore the sign unless the level is zero
        if (OVP) {
            const uint32_t t = (SYM && TIES) ? ((neg & xovpn) | (~neg & xovp)) : xovp;
            const uint32_t mo = __hge2_mask(a2, Pack2<T>::from_u32(t));      // element is an outlier
            const uint32_t sw = __byte_perm(mo, 0, 0x1032);                  // swap halves
            const uint32_t kill = sw & ~(mo & 0x0000ffffu);  // odd dies if even is outlier; even dies if only odd is
            q &= ~kill;
            victims = kill;
        }
        ranks = rk;
        return q;
    }

    // The same pair on the FMA pipe (tie-free rows, no codes): saturating-FMA compares and an exact
    // accumulation  q <- m*D_i + q  with m in {0, 1}; only the window test, the sign and the OVP
    // mask stay on the ALU pipe.  Interleaved with pair<>() so that both half-rate pipes are busy.
    __device__ __forceinline__ uint32_t pair_fma(uint32_t xb, bool &special, uint32_t &victims) const {
        typedef typename Pack2<T>::v2 v2;
        const v2 x2 = Pack2<T>::from_u32(xb);
        const v2 ab = __habs2(x2);
        special |= (__hle2_mask(ab, Pack2<T>::from_u32(xlim)) != 0xffffffffu);
        const v2 xs = __hmul2(SYM ? ab : x2, Pack2<T>::from_u32(S2));
        v2 q = Pack2<T>::from_u32(O[0]);
#pragma unroll
        for (int i = 0; i < NT; i++) {
            const v2 m = __hfma2_sat(xs, Pack2<T>::from_u32(Bf[i]), Pack2<T>::from_u32(Cf[i]));
            q = __hfma2(m, Pack2<T>::from_u32(Df[i]), q);
        }
        uint32_t qb;
        if (SYM) {
            // (+-1) * q + 0: restores the sign and leaves a zero level at +0
            const uint32_t one = IS16 && sizeof(T) == 2 ? (AntqType<T>::kInf == 0x7c00u ? 0x3c003c00u : 0x3f803f80u) : 0u;
            const v2 sg = Pack2<T>::from_u32((xb & 0x80008000u) | one);
            const v2 r = __hfma2(q, sg, Pack2<T>::from_u32(0u));
            qb = *reinterpret_cast<const uint32_t *>(&r);
        } else {
            qb = *reinterpret_cast<const uint32_t *>(&q);
        }
        if (OVP) {
            const v2 a2 = SYM ? ab : x2;
            const uint32_t mo = __hge2_mask(a2, Pack2<T>::from_u32(xovp));
            const uint32_t sw = __byte_perm(mo, 0, 0x1032);
            const uint32_t kill = sw & ~(mo & 0x0000ffffu);
            qb &= ~kill;
            victims = kill;
        }
        return qb;
    }

    // fp32: one element per register; fp32 x-space resolves ties, so negatives always use Xn
    __device__ __forceinline__ float one(float x, bool &special, int &rank, bool &outlier) const {
        const float a = SYM ? fabsf(x) : x;
        const bool neg = SYM && (__float_as_int(x) < 0);
        special |= !(fabsf(x) <= __uint_as_float(xlim));
        float q = __uint_as_float(O[0]);
        bool m0 = false;
        int rk = 0;
#pragma unroll
        for (int i = 0; i < NT; i++) {
            const bool m = a >= __uint_as_float(neg ? Xn[i] : X[i]);
            if (i == 0) m0 = m;
            q = m ? __uint_as_float(O[i + 1]) : q;
            if (CODES) rk += m ? 1 : 0;
        }
        if (SYM && m0) q = __uint_as_float(__float_as_uint(q) | (__float_as_uint(x) & 0x80000000u));
        if (OVP) outlier = a >= __uint_as_float(neg ? xovpn : xovp);
        rank = rk;
        return q;
    }

    // One vector (8 x 16-bit or 4 x fp32 elements) of this lane; returns false if it must go to the slow path.
    template <bool TIES, bool FMA>
    __device__ __forceinline__ bool vec(const AntqCodebook *__restrict__ cb, const uint4 r, uint4 *dst, int16_t *cdst,
                                        int mid, int K, int dbg) const {
        bool special = false;
        uint4 q;
        if constexpr (IS16) {
            uint32_t rk[4] = {0, 0, 0, 0}, vi[4] = {0, 0, 0, 0};
            if (dbg & 2) {
                q = r;
            } else {
                constexpr bool F = FMA && !TIES && !CODES;
                q.x = pair<TIES>(r.x, special, rk[0], vi[0]);
                if constexpr (F && ANTQ_FMA_PAIRS >= 3) q.z = pair_fma(r.z, special, vi[2]);
                else q.z = pair<TIES>(r.z, special, rk[2], vi[2]);
                if constexpr (F) {
                    q.y = pair_fma(r.y, special, vi[1]);
                    q.w = pair_fma(r.w, special, vi[3]);
                } else {
                    q.y = pair<TIES>(r.y, special, rk[1], vi[1]);
                    q.w = pair<TIES>(r.w, special, rk[3], vi[3]);
                }
            }
            if ((dbg & 1) && q.x != 0x12345678u) return !special;
            if (!special) {
                antq_stg_stream(dst, q);
                if (CODES) {
                    const uint32_t xb[4] = {r.x, r.y, r.z, r.w};
                    __align__(16) int16_t cc[8];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        cc[2 * k] = antq_rank_to_code<SYM>(cb, mid, rk[k] & 0xffff, (xb[k] & 0x8000u) != 0);
                        cc[2 * k + 1] = antq_rank_to_code<SYM>(cb, mid, rk[k] >> 16, (xb[k] & 0x80000000u) != 0);
                        if (OVP && (vi[k] & 0xffffu)) cc[2 * k] = (int16_t)K;
                        if (OVP && (vi[k] >> 16)) cc[2 * k + 1] = (int16_t)K;
                    }
                    *reinterpret_cast<uint4 *>(cdst) = *reinterpret_cast<uint4 *>(cc);
                }
            }
        } else {
            const float xv[4] = {__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z), __uint_as_float(r.w)};
            float qv[4];
            int rk[4];
            bool ol[4] = {false, false, false, false}, vict[4] = {false, false, false, false};
#pragma unroll
            for (int k = 0; k < 4; k++) qv[k] = one(xv[k], special, rk[k], ol[k]);
            if (OVP) {
#pragma unroll
                for (int k = 0; k < 4; k += 2) {
                    vict[k + 1] = ol[k];
                    vict[k] = ol[k + 1] && !ol[k];
                    if (vict[k]) qv[k] = 0.0f;
                    if (vict[k + 1]) qv[k + 1] = 0.0f;
                }
            }
            if (!special) {
                q = make_uint4(__float_as_uint(qv[0]), __float_as_uint(qv[1]), __float_as_uint(qv[2]),
                               __float_as_uint(qv[3]));
                antq_stg_stream(dst, q);
                if (CODES) {
                    __align__(8) int16_t cc[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        cc[k] = antq_rank_to_code<SYM>(cb, mid, rk[k], (__float_as_uint(xv[k]) >> 31) != 0);
                        if (OVP && vict[k]) cc[k] = (int16_t)K;
                    }
                    *reinterpret_cast<uint2 *>(cdst) = *reinterpret_cast<uint2 *>(cc);
                }
            }
        }
        return !special;
    }

    // One chunk (<= 4 vectors per lane) from shared memory to global, as a ROLLED loop with the next
    // vector's LDS issued ahead (three code paths x four unrolled vectors overflowed the 32 KiB
    // instruction cache and the register file).  Returns true if this lane skipped a vector.
    template <bool TIES, bool FMA>
    __device__ __forceinline__ bool chunk(const AntqCodebook *__restrict__ cb, const uint4 *sv, T *og, int16_t *cg,
                                          int nvec, int lane, int mid, int K, int dbg) const {
        constexpr int VEC = AntqType<T>::kVec;
        uint4 *oout = reinterpret_cast<uint4 *>(og);
        bool skipped = false;
        constexpr int LV = ANTQ_LOOP_VECS;
        uint4 cur[LV];
#pragma unroll
        for (int u = 0; u < LV; u++) {
            cur[u] = make_uint4(0, 0, 0, 0);
            if (lane + 32 * u < nvec) cur[u] = sv[lane + 32 * u];           // LDS.128, conflict free
        }
#pragma unroll 1
        for (int v = lane; v < nvec; v += 32 * LV) {
            uint4 nxt[LV];
#pragma unroll
            for (int u = 0; u < LV; u++) {
                nxt[u] = make_uint4(0, 0, 0, 0);
                if (v + 32 * (LV + u) < nvec) nxt[u] = sv[v + 32 * (LV + u)];
            }
#pragma unroll
            for (int u = 0; u < LV; u++)
                if (u == 0 || v + 32 * u < nvec)
                    skipped |= !vec<TIES, FMA>(cb, cur[u], oout + v + 32 * u,
                                               CODES ? cg + (long long)(v + 32 * u) * VEC : nullptr, mid, K, dbg);
#pragma unroll
            for (int u = 0; u < LV; u++) cur[u] = nxt[u];
        }
        return skipped;
    }
};


struct RowsParams {
    const void *x;
    void *out;
    int16_t *codes;
    const float *alpha;
    const AntqCodebook *cb;
    long long rows, cols, total_chunks;
    int alpha_per_row, chunks_per_row, chunk_elems, total_warps;
    int nt_real, mid, ovp_index, n_entries;            // codebook header, from the host-side antq_codebook_info
    float gmax, lim;
    int debug;          // ANTQ_DEBUG experiments: 1 = no stores, 2 = no chain (copy), 4 = no table rebuild
};

// Exact (X, Xn) for a 16-bit type when the division-free shortcut could not prove them
// (p = t*s within 16 fp32-ulps of a representable value): at most three divisions.
template <typename T>
__device__ __noinline__ void antq_x_threshold16_near(float tpos, float tneg, float s, T *xp, T *xn) {
    typedef AntqType<T> A;
    T c = A::from_f32_rn(__fmul_rn(tpos, s));
    float qc = __fdiv_rn(A::to_f32(c), s);
    if (qc >= tpos) {
        const T p = antq_next_down(c);
        const float qp = __fdiv_rn(A::to_f32(p), s);
        if (!antq_is_inf(p) && qp >= tpos) { c = p; qc = qp; }
    } else {
        // RN(t*s) is within one step of the answer, but keep walking if it is not (never observed)
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

template <typename T, int NT, bool SYM, bool OVP, bool CODES>
__global__ void __launch_bounds__(kWarpsPerCta * 32, (NT <= 7 ? kCtasPerSm : (NT <= 15 ? 3 : 2)))
antq_rows_kernel(const RowsParams p) {
    typedef AntqType<T> A;
    constexpr int VEC = A::kVec;
    extern __shared__ __align__(128) unsigned char antq_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned char *wbase = antq_smem + (size_t)wib * kWarpSmemBytes;
    uint32_t *sX = reinterpret_cast<uint32_t *>(wbase + kRing * kChunkBytes);
    uint32_t *sXn = sX + kTableWords;
    uint32_t *sO = sXn + kTableWords;
    uint32_t *sB = sO + kTableWords;
    uint32_t *sC = sB + kTableWords;
    uint32_t *sD = sC + kTableWords;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(sD + kTableWords);

    const int w = blockIdx.x * kWarpsPerCta + wib;
    // equal contiguous share of the chunk list
    const long long c_begin = (p.total_chunks * w) / p.total_warps;
    const long long c_end = (p.total_chunks * (w + 1)) / p.total_warps;
    if (c_begin >= c_end) return;
    const AntqCodebook *__restrict__ cb = p.cb;

    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kRing; k++) antq_mbar_init(mbar + k, 1);
    }
    __syncwarp();

    // issue cursor (used by lane 0 only): position of the next chunk to put in flight
    long long irow = c_begin / p.chunks_per_row;
    int iidx = (int)(c_begin - irow * p.chunks_per_row);
    auto issue = [&](int slot) {
        const long long col0 = (long long)iidx * p.chunk_elems;
        const long long remain = p.cols - col0;
        const int n_el = (int)(remain < p.chunk_elems ? remain : p.chunk_elems);
        const int nvec = n_el / VEC;
        if (nvec > 0) {
            antq_fence_proxy_async();
            antq_bulk_g2s(wbase + slot * kChunkBytes, reinterpret_cast<const T *>(p.x) + irow * p.cols + col0,
                          (unsigned)nvec * 16u, mbar + slot);
        }
        if (++iidx == p.chunks_per_row) { iidx = 0; ++irow; }
    };
    // Only the FIRST chunk is requested up front; the rest of the ring is filled once it has landed.
    // Requesting ring-depth chunks from every warp at t = 0 puts most of a small tensor in the memory
    // queue at once and delays every warp's first chunk (profiles/r01_tuning.md).
    long long issued = c_begin;               // next chunk to issue (meaningful on lane 0)
    if (lane == 0) { issue(0); issued = c_begin + 1; }

    // this lane's codebook entries stay in registers for the whole kernel
    const int nt_real = p.nt_real;
    const float inf = __int_as_float(0x7f800000);
    const float tpos = lane < nt_real ? (SYM ? cb->mag_tpos[lane] : cb->thr[lane]) : inf;
    const float tneg = (SYM && lane < nt_real) ? cb->mag_tneg[lane] : tpos;
    const float lev = lane <= nt_real ? (SYM ? cb->level[p.mid + lane] : cb->level[lane]) : 0.0f;

    // compute cursor
    long long row = c_begin / p.chunks_per_row;
    int idx = (int)(c_begin - row * p.chunks_per_row);
    long long cur_row = -1;
    float alpha_next = __ldg(p.alpha + (p.alpha_per_row ? row : 0));
    float s = 0.0f;
    bool row_ok = false, row_ties = false, row_fma = false;
    Tables<T, NT, SYM, OVP, CODES> tab;
    unsigned phases = 0;                      // bit k = parity to wait for on ring slot k

    for (long long c = c_begin; c < c_end; ++c) {
        const int slot = (int)((c - c_begin) & (kRing - 1));
        // keep the ring full: the slot freed by the previous iteration receives chunk c + kRing - 1
        if (lane == 0 && c != c_begin) {
            const long long want = (c + kRing) < c_end ? (c + kRing) : c_end;
            while (issued < want) { issue((int)((issued - c_begin) & (kRing - 1))); ++issued; }
        }

        if (row != cur_row) {
            // ---- table rebuild for a new row (copies for this and the next chunks are already in flight) ----
            cur_row = row;
            const float alpha = alpha_next;
            if (p.alpha_per_row && row + 1 < p.rows) alpha_next = __ldg(p.alpha + row + 1);   // prefetch
            if (p.debug & 4) {
                s = 1.0f; row_ok = true; row_ties = false;
                sX[lane] = sXn[lane] = antq_table_word<T>(A::from_bits(A::kInf));
                sO[lane] = 0;
            } else {
                s = __fdiv_rn(alpha, p.gmax);                          // scale = alpha / max(grid)
                row_ok = s > 0.0f && s < inf;
                T Xl = A::from_bits(A::kInf), Xnl = A::from_bits(A::kInf), Ol = A::from_bits(0);
                if (row_ok) {
                    if (lane < nt_real) {
                        bool near;
                        Xl = antq_x_threshold<T>(tpos, s, &near);
                        Xnl = Xl;
                        if (SYM && near) {
                            // a negative input can land on the other side of a representable tie
                            if constexpr (sizeof(T) == 2) antq_x_threshold16_near<T>(tpos, tneg, s, &Xl, &Xnl);
                            else Xnl = antq_x_threshold_exact<T>(tneg, s);
                        }
                    }
                    if (lane <= nt_real) Ol = A::from_f32_rn(__fmul_rn(lev, s));
                }
                row_ties = __any_sync(0xffffffffu, A::bits(Xl) != A::bits(Xnl));
                sX[lane] = antq_table_word<T>(Xl);
                sXn[lane] = antq_table_word<T>(Xnl);
                sO[lane] = antq_table_word<T>(Ol);
                // FMA-pipe twin tables (16-bit types, tie-free rows)
                uint32_t s2 = 0;
                row_fma = false;
                if constexpr (sizeof(T) == 2) {
                    if (row_ok && !row_ties && !(p.debug & 16)) {
                        // S = 2^k with xlim * S in [2^13, 2^14): every in-window |x| stays finite after scaling
                        const float xl = __fmul_rn(__fmul_rn(p.lim, s), 0.9990234375f);
                        int k = 13 - (int)((__float_as_uint(xl) >> 23) & 0xff) + 127;
                        k = k > 15 ? 15 : (k < -14 ? -14 : k);
                        const float S = __uint_as_float((unsigned)(127 + k) << 23);
                        bool ok = isfinite(A::to_f32(Ol));
                        float Bv = 0.0f, Cv = -1.0f;              // padding: m = sat(0 * xs - 1) = 0
                        if (lane < nt_real) {
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
                                ok = ok && Bv >= 6.103515625e-05f && Bv <= 32768.0f && fabsf(Cv) <= 2048.0f;
                            }
                        }
                        // D_i = O[i+1] - O[i] must be exactly representable (fp32 difference of two 16-bit values is exact)
                        const float o_next = __shfl_down_sync(0xffffffffu, A::to_f32(Ol), 1);
                        const float dex = __fsub_rn(o_next, A::to_f32(Ol));
                        const T D = A::from_f32_rn(dex);
                        if (lane < nt_real) ok = ok && (A::to_f32(D) == dex);
                        sD[lane] = antq_table_word<T>(lane < nt_real ? D : A::from_bits(0));
                        sB[lane] = antq_table_word<T>(A::from_f32_rn(Bv));
                        sC[lane] = antq_table_word<T>(A::from_f32_rn(Cv));
                        s2 = antq_table_word<T>(A::from_f32_rn(S));
                        row_fma = __all_sync(0xffffffffu, ok);
                    }
                }
                __syncwarp();
                tab.load(sX, sXn, sO, p.lim, p.ovp_index, s);
                if (row_fma) tab.load_fma(sB, sC, sD, s2);
            }
            if (p.debug & 4) {
                __syncwarp();
                tab.load(sX, sXn, sO, p.lim, p.ovp_index, s);
            }
            __syncwarp();
        }

        const long long col0 = (long long)idx * p.chunk_elems;
        const long long remain = p.cols - col0;
        const int n_el = (int)(remain < p.chunk_elems ? remain : p.chunk_elems);
        const int nvec = n_el / VEC;
        const long long base = row * p.cols + col0;
        const T *xg = reinterpret_cast<const T *>(p.x) + base;
        T *og = reinterpret_cast<T *>(p.out) + base;
        int16_t *cg = CODES ? p.codes + base : nullptr;

        if (nvec > 0) {
            antq_mbar_wait(mbar + slot, (phases >> slot) & 1u);        // the chunk has landed in shared memory
            phases ^= 1u << slot;
            if (lane == 0 && c == c_begin) {                             // first chunk is in: fill the ring
                const long long want = (c + kRing) < c_end ? (c + kRing) : c_end;
                while (issued < want) { issue((int)((issued - c_begin) & (kRing - 1))); ++issued; }
            }
            const uint4 *sv = reinterpret_cast<const uint4 *>(wbase + slot * kChunkBytes);
            bool skipped = !row_ok;
            if (row_ok)
                skipped = row_ties  ? tab.template chunk<true, false>(cb, sv, og, cg, nvec, lane, p.mid, p.n_entries, p.debug)
                          : row_fma ? tab.template chunk<false, true>(cb, sv, og, cg, nvec, lane, p.mid, p.n_entries, p.debug)
                                    : tab.template chunk<false, false>(cb, sv, og, cg, nvec, lane, p.mid, p.n_entries, p.debug);
            if (__any_sync(0xffffffffu, skipped))
                antq_fixup_chunk<T, OVP>(cb, s, tab.xlim_f32(), !row_ok, xg, og, cg, nvec, lane);
        }
        // ragged tail (only a per-tensor view can have one: rows == 1)
        const int tail = n_el - nvec * VEC;
        if (tail > 0 && lane == 0)
            antq_slow_vec<T, OVP>(cb, s, xg + (long long)nvec * VEC, og + (long long)nvec * VEC,
                                  CODES ? cg + (long long)nvec * VEC : nullptr, tail);
        __syncwarp();                                                  // every lane is done with this slot
        if (++idx == p.chunks_per_row) { idx = 0; ++row; }
    }
}

template <typename K> int launch_kernel(K kernel, const RowsParams &p, cudaStream_t st) {
    const int ctas = (p.total_warps + kWarpsPerCta - 1) / kWarpsPerCta;
    const int smem = kWarpsPerCta * kWarpSmemBytes;
    if (smem > 48 * 1024) {     // opt in beyond the default dynamic-smem limit (idempotent, host-only)
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
    }
    kernel<<<dim3((unsigned)ctas), dim3(kWarpsPerCta * 32), smem, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        fprintf(stderr, "antq: rows kernel launch failed: %s (grid %d, block %d, smem %d, chunks %lld)\n",
                cudaGetErrorString(e), ctas, kWarpsPerCta * 32, smem, p.total_chunks);
    return (int)e;
}

template <typename T, int NT, bool SYM, bool OVP>
int launch_nt(const RowsParams &p, bool codes, cudaStream_t st) {
    if (codes) return launch_kernel(antq_rows_kernel<T, NT, SYM, OVP, true>, p, st);
    return launch_kernel(antq_rows_kernel<T, NT, SYM, OVP, false>, p, st);
}

template <typename T, bool SYM, bool OVP> int launch_sym(const RowsParams &p, int nt, bool codes, cudaStream_t st) {
    if (nt <= 3) return launch_nt<T, 3, SYM, OVP>(p, codes, st);
    if (nt <= 7) return launch_nt<T, 7, SYM, OVP>(p, codes, st);
    if (nt <= 15) return launch_nt<T, 15, SYM, OVP>(p, codes, st);
    if (nt <= 31) return launch_nt<T, 31, SYM, OVP>(p, codes, st);
    return ANTQ_ENOTSUP;
}

template <typename T> int launch_t(const RowsParams &p, int nt, bool sym, bool ovp, bool codes, cudaStream_t st) {
    if (sym) return ovp ? launch_sym<T, true, true>(p, nt, codes, st) : launch_sym<T, true, false>(p, nt, codes, st);
    return ovp ? launch_sym<T, false, true>(p, nt, codes, st) : launch_sym<T, false, false>(p, nt, codes, st);
}

}  // namespace

// The host knows the codebook header (antq_codebook_info, fetched once at prepare time), so kernel
// selection and the launch shape need no device round trip.
int antq_launch_rows(const void *x, void *out, int16_t *codes, const float *alpha, int alpha_per_row, long long rows,
                     long long cols, int dtype, const AntqCodebook *cb, const antq_codebook_info *info, bool ovp,
                     cudaStream_t st) {
    const bool sym = (info->flags & ANTQ_CB_SYMMETRIC) != 0;
    const int nt = sym ? info->n_mag - 1 : info->n_levels - 1;
    const int es = dtype == ANTQ_F32 ? 4 : 2;
    RowsParams p;
    p.x = x; p.out = out; p.codes = codes; p.alpha = alpha; p.cb = cb;
    p.rows = rows; p.cols = cols;
    p.chunk_elems = kChunkBytes / es;
    const long long cpr = (cols + p.chunk_elems - 1) / p.chunk_elems;
    if (cpr > 0x7fffffffLL) return ANTQ_ENOTSUP;
    p.chunks_per_row = (int)cpr;
    p.total_chunks = rows * cpr;
    if (p.total_chunks == 0) return 0;
    p.alpha_per_row = alpha_per_row;
    // persistent grid: every warp slot of the machine, fewer only when there are fewer chunks than slots
    const int regs_ctas = nt <= 7 ? kCtasPerSm : (nt <= 15 ? 3 : 2);
    long long warps = (long long)kNumSms * regs_ctas * kWarpsPerCta;
    if (warps > p.total_chunks) warps = (p.total_chunks + kWarpsPerCta - 1) / kWarpsPerCta * kWarpsPerCta;
    p.total_warps = (int)warps;
    p.nt_real = nt; p.mid = info->mid; p.ovp_index = info->ovp_index; p.n_entries = info->n_entries;
    p.gmax = info->gmax; p.lim = info->lim;
    {
        static int dbg = -1;
        if (dbg < 0) { const char *e = getenv("ANTQ_DEBUG"); dbg = e ? atoi(e) : 0; }
        p.debug = dbg;
    }
    switch (dtype) {
        case ANTQ_F32: return launch_t<float>(p, nt, sym, ovp, codes != nullptr, st);
        case ANTQ_F16: return launch_t<__half>(p, nt, sym, ovp, codes != nullptr, st);
        case ANTQ_BF16: return launch_t<__nv_bfloat16>(p, nt, sym, ovp, codes != nullptr, st);
    }
    return ANTQ_EINVAL;
}
