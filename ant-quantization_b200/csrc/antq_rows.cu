// antq_rows.cu -- the hot kernel: fused fake-quant forward with per-row tables.
//
// Replaces A/antquant/quant_modules.py:535-551 (+ O/...:295-330 with OVP) and the
// scan kernel it calls (A/quant/quant_kernel.cu:11-39): ~9 launches and 7 reads +
// 7 writes of the tensor become ONE launch, one read, one write.
//
// Idea (proved bit-exact on exhaustive fp16 inputs in tests/test_xspace_model.py):
// the reference computes d = fl32(x / s), picks the grid level by a scan, and
// returns fl32(((q - d) + d) * s).  fl32(x / s) is monotone in x, so every
// d-space threshold of the prepared codebook maps to an exact x-space threshold
//      X_r = min{ x in dtype : fl32(x / s) >= thr_r }
// and inside the window |d| <= lim the STE sum is exact, so the result is simply
// O_j = RN_dtype(fl32(level_j * s)).  One warp owns one row segment: its lanes
// build the row's (X_r, O_j) tables in parallel (lane r <-> threshold r), share
// them through shared memory, and then stream the segment with 128-bit loads:
// per 32-bit register (two fp16 values) the work is one packed compare + one
// LOP3 per threshold -- no division, no table lookup, no divergent branch.
// Values outside the window, NaN/Inf, rows whose scale is not a positive finite
// number, and rows where a positive/negative tie would differ take the literal
// reference arithmetic (antq_slow_vec) -- rare, and exact by construction.
#include "antq_common.cuh"

namespace {

constexpr int kWarpsPerCta = 4;
constexpr int kUnroll = 4;

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

// ---- packed 16-bit (fp16 / bf16) chain --------------------------------------
template <typename T> struct Pack2;
template <> struct Pack2<__half> {
    typedef __half2 v2;
    __device__ static __forceinline__ v2 from_u32(uint32_t u) { return *reinterpret_cast<v2 *>(&u); }
    __device__ static __forceinline__ uint32_t dup(__half h) {
        uint32_t b = __half_as_ushort(h);
        return b | (b << 16);
    }
};
template <> struct Pack2<__nv_bfloat16> {
    typedef __nv_bfloat162 v2;
    __device__ static __forceinline__ v2 from_u32(uint32_t u) { return *reinterpret_cast<v2 *>(&u); }
    __device__ static __forceinline__ uint32_t dup(__nv_bfloat16 h) {
        uint32_t b = __bfloat16_as_ushort(h);
        return b | (b << 16);
    }
};

template <typename T, int NT, bool SYM, bool OVP, bool CODES> struct RowTables16 {
    uint32_t X[NT];       // thresholds for x >= 0 (or for every x when !SYM), duplicated in both halves
    uint32_t Xn[NT];      // SYM: thresholds on |x| for x < 0 -- they differ from X only where an exact tie
                          // is representable (ties go to the LATER grid entry: up for +, toward 0 for -)
    uint32_t O[NT + 1];   // outputs, duplicated
    uint32_t xlim, xovp, xovpn;

    // one 32-bit register = two elements (low half = even flat index)
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
};

template <int NT, bool SYM, bool OVP, bool CODES> struct RowTables32 {
    float X[NT];
    float Xn[NT];         // SYM: thresholds on |x| for negative x (fp32 x-space resolves ties, so they differ)
    float O[NT + 1];
    float xlim, xovp, xovpn;
    __device__ __forceinline__ float one(float x, bool &special, int &rank, bool &outlier) const {
        const float a = SYM ? fabsf(x) : x;
        const bool neg = SYM && (__float_as_int(x) < 0);
        special |= !(fabsf(x) <= xlim);
        float q = O[0];
        bool m0 = false;
        int rk = 0;
#pragma unroll
        for (int i = 0; i < NT; i++) {
            const bool m = a >= (neg ? Xn[i] : X[i]);
            if (i == 0) m0 = m;
            q = m ? O[i + 1] : q;
            if (CODES) rk += m ? 1 : 0;
        }
        if (SYM && m0) q = __uint_as_float(__float_as_uint(q) | (__float_as_uint(x) & 0x80000000u));
        if (OVP) outlier = a >= (neg ? xovpn : xovp);
        rank = rk;
        return q;
    }
};

// Cold second pass over a segment in which some vector was skipped by the fast loop: the same
// window predicate is re-evaluated in fp32, and exactly the skipped vectors get the literal
// reference arithmetic.  (The fast loop stores nothing for them, so this also works in place.)
template <typename T, bool OVP>
__device__ __noinline__ void antq_fixup_pass(const AntqCodebook *__restrict__ cb, float s, float xlim, const T *xrow,
                                             T *orow, int16_t *crow, int nvec, int lane) {
    typedef AntqType<T> A;
    constexpr int VEC = A::kVec;
    for (int v = lane; v < nvec; v += 32) {
        const T *xv = xrow + (long long)v * VEC;
        bool special = false;
#pragma unroll
        for (int e = 0; e < VEC; e++) special |= !(fabsf(A::to_f32(xv[e])) <= xlim);
        if (special)
            antq_slow_vec<T, OVP>(cb, s, xv, orow + (long long)v * VEC, crow ? crow + (long long)v * VEC : nullptr,
                                  VEC);
    }
}

struct RowsParams {
    const void *x;
    void *out;
    int16_t *codes;
    const float *alpha;
    const AntqCodebook *cb;
    long long rows, cols, total_segs, total_warps;
    int alpha_per_row, segs_per_row, seg_len, segs_per_warp;
};

template <bool SYM>
__device__ __forceinline__ int16_t antq_rank_to_code(const AntqCodebook *__restrict__ cb, int rank, bool neg) {
    int lvl = SYM ? (cb->mid + (neg ? -rank : rank)) : rank;
    return (int16_t)cb->level_code[lvl];
}

// Per-row tables in registers + the streaming loop over one segment.
template <typename T, int NT, bool SYM, bool OVP, bool CODES> struct SegmentWorker;

template <typename T, int NT, bool SYM, bool OVP, bool CODES> struct SegmentWorker {   // 16-bit element types
    typedef AntqType<T> A;
    static constexpr int VEC = 8;
    RowTables16<T, NT, SYM, OVP, CODES> tab;

    __device__ __forceinline__ void load(const T *sX, const T *sXn, const T *sO, const AntqCodebook *__restrict__ cb,
                                         float s) {
#pragma unroll
        for (int i = 0; i < NT; i++) tab.X[i] = Pack2<T>::dup(sX[i]);
#pragma unroll
        for (int i = 0; i < NT; i++) tab.Xn[i] = Pack2<T>::dup(sXn[i]);
#pragma unroll
        for (int i = 0; i <= NT; i++) tab.O[i] = Pack2<T>::dup(sO[i]);
        const float xl = __fmul_rn(__fmul_rn(cb->lim, s), 0.9990234375f);   // conservative window in x-space
        tab.xlim = Pack2<T>::dup(A::from_f32_rz(xl));
        const int oi = cb->ovp_index;
        const bool has = OVP && oi >= 0 && oi < NT;
        tab.xovp = Pack2<T>::dup(has ? sX[oi] : A::from_bits(A::kInf));
        tab.xovpn = Pack2<T>::dup(has ? sXn[oi] : A::from_bits(A::kInf));
    }

    // returns true if this lane skipped at least one vector (NaN/Inf/outside the exact window)
    template <bool TIES>
    __device__ __forceinline__ bool run(const AntqCodebook *__restrict__ cb, const T *xrow, T *orow,
                                        int16_t *crow, int nvec, int lane) const {
        const uint4 *xin = reinterpret_cast<const uint4 *>(xrow);
        uint4 *oout = reinterpret_cast<uint4 *>(orow);
        const int K = cb->n_entries;
        bool any_special = false;
        for (int v0 = 0; v0 < nvec; v0 += 32 * kUnroll) {
            uint4 r[kUnroll];
#pragma unroll
            for (int j = 0; j < kUnroll; j++) {
                const int v = v0 + j * 32 + lane;
                if (v < nvec) r[j] = antq_ldg_stream(xin + v);
            }
#pragma unroll
            for (int j = 0; j < kUnroll; j++) {
                const int v = v0 + j * 32 + lane;
                if (v < nvec) {
                    bool special = false;
                    uint32_t rk[4], vi[4] = {0, 0, 0, 0};
                    uint4 q;
                    q.x = tab.template pair<TIES>(r[j].x, special, rk[0], vi[0]);
                    q.y = tab.template pair<TIES>(r[j].y, special, rk[1], vi[1]);
                    q.z = tab.template pair<TIES>(r[j].z, special, rk[2], vi[2]);
                    q.w = tab.template pair<TIES>(r[j].w, special, rk[3], vi[3]);
                    any_special |= special;
                    if (!special) {      // special vectors are left untouched for antq_fixup_pass
                        antq_stg_stream(oout + v, q);
                        if (CODES) {
                            const uint32_t xb[4] = {r[j].x, r[j].y, r[j].z, r[j].w};
                            __align__(16) int16_t cc[8];
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                cc[2 * k] = antq_rank_to_code<SYM>(cb, rk[k] & 0xffff, (xb[k] & 0x8000u) != 0);
                                cc[2 * k + 1] = antq_rank_to_code<SYM>(cb, rk[k] >> 16, (xb[k] & 0x80000000u) != 0);
                                if (OVP && (vi[k] & 0xffffu)) cc[2 * k] = (int16_t)K;
                                if (OVP && (vi[k] >> 16)) cc[2 * k + 1] = (int16_t)K;
                            }
                            *reinterpret_cast<uint4 *>(crow + (long long)v * VEC) = *reinterpret_cast<uint4 *>(cc);
                        }
                    }
                }
            }
        }
        return any_special;
    }
    __device__ __forceinline__ float xlim_f32() const {
        return A::to_f32(A::from_bits((typename A::bits_t)(tab.xlim & 0xffffu)));
    }
};

template <int NT, bool SYM, bool OVP, bool CODES> struct SegmentWorker<float, NT, SYM, OVP, CODES> {
    typedef float T;
    static constexpr int VEC = 4;
    RowTables32<NT, SYM, OVP, CODES> tab;

    __device__ __forceinline__ void load(const T *sX, const T *sXn, const T *sO, const AntqCodebook *__restrict__ cb,
                                         float s) {
#pragma unroll
        for (int i = 0; i < NT; i++) tab.X[i] = sX[i];
#pragma unroll
        for (int i = 0; i < NT; i++) tab.Xn[i] = sXn[i];
#pragma unroll
        for (int i = 0; i <= NT; i++) tab.O[i] = sO[i];
        tab.xlim = __fmul_rn(__fmul_rn(cb->lim, s), 0.9990234375f);
        const int oi = cb->ovp_index;
        const bool has = OVP && oi >= 0 && oi < NT;
        tab.xovp = has ? sX[oi] : __int_as_float(0x7f800000);
        tab.xovpn = has ? sXn[oi] : __int_as_float(0x7f800000);
    }

    // returns true if this lane skipped at least one vector (NaN/Inf/outside the exact window)
    template <bool TIES>      // fp32 x-space always resolves ties: the flag is ignored
    __device__ __forceinline__ bool run(const AntqCodebook *__restrict__ cb, const T *xrow, T *orow,
                                        int16_t *crow, int nvec, int lane) const {
        const uint4 *xin = reinterpret_cast<const uint4 *>(xrow);
        uint4 *oout = reinterpret_cast<uint4 *>(orow);
        const int K = cb->n_entries;
        bool any_special = false;
        for (int v0 = 0; v0 < nvec; v0 += 32 * kUnroll) {
            uint4 r[kUnroll];
#pragma unroll
            for (int j = 0; j < kUnroll; j++) {
                const int v = v0 + j * 32 + lane;
                if (v < nvec) r[j] = antq_ldg_stream(xin + v);
            }
#pragma unroll
            for (int j = 0; j < kUnroll; j++) {
                const int v = v0 + j * 32 + lane;
                if (v < nvec) {
                    bool special = false;
                    const float xv[4] = {__uint_as_float(r[j].x), __uint_as_float(r[j].y), __uint_as_float(r[j].z),
                                         __uint_as_float(r[j].w)};
                    float qv[4];
                    int rk[4];
                    bool ol[4] = {false, false, false, false}, vict[4] = {false, false, false, false};
#pragma unroll
                    for (int k = 0; k < 4; k++) qv[k] = tab.one(xv[k], special, rk[k], ol[k]);
                    if (OVP) {
#pragma unroll
                        for (int k = 0; k < 4; k += 2) {
                            vict[k + 1] = ol[k];
                            vict[k] = ol[k + 1] && !ol[k];
                            if (vict[k]) qv[k] = 0.0f;
                            if (vict[k + 1]) qv[k + 1] = 0.0f;
                        }
                    }
                    any_special |= special;
                    if (!special) {
                        uint4 q = {__float_as_uint(qv[0]), __float_as_uint(qv[1]), __float_as_uint(qv[2]),
                                   __float_as_uint(qv[3])};
                        antq_stg_stream(oout + v, q);
                        if (CODES) {
                            __align__(8) int16_t cc[4];
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                cc[k] = antq_rank_to_code<SYM>(cb, rk[k], (__float_as_uint(xv[k]) >> 31) != 0);
                                if (OVP && vict[k]) cc[k] = (int16_t)K;
                            }
                            *reinterpret_cast<uint2 *>(crow + (long long)v * VEC) = *reinterpret_cast<uint2 *>(cc);
                        }
                    }
                }
            }
        }
        return any_special;
    }
    __device__ __forceinline__ float xlim_f32() const { return tab.xlim; }
};

template <typename T, int NT, bool SYM, bool OVP, bool CODES>
__global__ void __launch_bounds__(kWarpsPerCta * 32, (NT <= 7 ? 6 : (NT <= 15 ? 4 : 2))) antq_rows_kernel(const RowsParams p) {
    typedef AntqType<T> A;
    constexpr int VEC = A::kVec;
    __shared__ __align__(16) T sX[kWarpsPerCta][32];
    __shared__ __align__(16) T sXn[kWarpsPerCta][32];
    __shared__ __align__(16) T sO[kWarpsPerCta][32];

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long w = (long long)blockIdx.x * kWarpsPerCta + wib;
    if (w >= p.total_warps) return;
    const AntqCodebook *__restrict__ cb = p.cb;
    const int nt_real = SYM ? cb->n_mag - 1 : cb->n_levels - 1;
    const float gmax = cb->gmax;

    SegmentWorker<T, NT, SYM, OVP, CODES> worker;
    long long cur_row = -1;
    float s = 0.0f;
    bool row_ok = false, row_ties = false;

    const long long sg0 = w * p.segs_per_warp;
    const long long sg1 = (sg0 + p.segs_per_warp) < p.total_segs ? (sg0 + p.segs_per_warp) : p.total_segs;
    for (long long sg = sg0; sg < sg1; ++sg) {
        const long long row = sg / p.segs_per_row;
        const int seg = (int)(sg - row * p.segs_per_row);
        if (row != cur_row) {
            // ---- row prologue: scale, x-space thresholds (lane r <-> threshold r), outputs ----
            cur_row = row;
            const float alpha = __ldg(p.alpha + (p.alpha_per_row ? row : 0));
            s = __fdiv_rn(alpha, gmax);                          // scale = alpha / max(grid)
            const bool s_ok = s > 0.0f && s < __int_as_float(0x7f800000);
            T Xl = A::from_bits(A::kInf), Xnl = A::from_bits(A::kInf), Ol = A::from_bits(0);
            if (s_ok) {
                if (lane < nt_real) {
                    bool near;
                    if (SYM) {
                        Xl = antq_x_threshold<T>(cb->mag_tpos[lane], s, &near);
                        // unless the shortcut proved both sides equal, a negative input can land on the
                        // other side of a representable tie: give it its own threshold
                        Xnl = near ? antq_x_threshold_exact<T>(cb->mag_tneg[lane], s) : Xl;
                    } else {
                        Xl = antq_x_threshold<T>(cb->thr[lane], s, &near);
                        Xnl = Xl;
                    }
                }
                if (lane <= nt_real) {
                    const float lv = SYM ? cb->level[cb->mid + lane] : cb->level[lane];
                    Ol = A::from_f32_rn(__fmul_rn(lv, s));
                }
            }
            row_ok = s_ok;
            row_ties = __any_sync(0xffffffffu, A::bits(Xl) != A::bits(Xnl));
            __syncwarp();
            sX[wib][lane] = Xl;
            sXn[wib][lane] = Xnl;
            sO[wib][lane] = Ol;
            __syncwarp();
            worker.load(sX[wib], sXn[wib], sO[wib], cb, s);
        }

        const long long col0 = (long long)seg * p.seg_len;
        const long long remain = p.cols - col0;
        const int n_el = (int)(remain < p.seg_len ? remain : p.seg_len);
        const int nvec = n_el / VEC;
        const long long base = row * p.cols + col0;
        const T *xrow = reinterpret_cast<const T *>(p.x) + base;
        T *orow = reinterpret_cast<T *>(p.out) + base;
        int16_t *crow = CODES ? p.codes + base : nullptr;

        if (row_ok) {
            const bool skipped = row_ties ? worker.template run<true>(cb, xrow, orow, crow, nvec, lane)
                                          : worker.template run<false>(cb, xrow, orow, crow, nvec, lane);
            if (__any_sync(0xffffffffu, skipped))
                antq_fixup_pass<T, OVP>(cb, s, worker.xlim_f32(), xrow, orow, crow, nvec, lane);
        } else {
            for (int v = lane; v < nvec; v += 32)
                antq_slow_vec<T, OVP>(cb, s, xrow + (long long)v * VEC, orow + (long long)v * VEC,
                                      CODES ? crow + (long long)v * VEC : nullptr, VEC);
        }
        // ragged tail (only a per-tensor view can have one: rows == 1)
        const int tail = n_el - nvec * VEC;
        if (tail > 0 && lane == 0)
            antq_slow_vec<T, OVP>(cb, s, xrow + (long long)nvec * VEC, orow + (long long)nvec * VEC,
                                  CODES ? crow + (long long)nvec * VEC : nullptr, tail);
    }
}

template <typename T, int NT, bool SYM, bool OVP>
int launch_nt(const RowsParams &p, bool codes, cudaStream_t st) {
    const long long ctas = (p.total_warps + kWarpsPerCta - 1) / kWarpsPerCta;
    if (ctas > 0x7fffffffLL) return ANTQ_ENOTSUP;
    dim3 grid((unsigned)ctas), block(kWarpsPerCta * 32);
    if (codes) antq_rows_kernel<T, NT, SYM, OVP, true><<<grid, block, 0, st>>>(p);
    else antq_rows_kernel<T, NT, SYM, OVP, false><<<grid, block, 0, st>>>(p);
    return (int)cudaGetLastError();
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

// nt = number of thresholds the chain needs (n_mag-1 if symmetric else n_levels-1), known to the host
// from antq_codebook_info_get at prepare time; sym / ovp likewise.
int antq_launch_rows(const void *x, void *out, int16_t *codes, const float *alpha, int alpha_per_row, long long rows,
                     long long cols, int dtype, const AntqCodebook *cb, int nt, bool sym, bool ovp,
                     cudaStream_t st) {
    const int vec = dtype == ANTQ_F32 ? 4 : 8;
    // ~16 vectors per lane per segment: enough to amortise the row prologue, small enough to balance 148 SMs
    const long long target = 32LL * vec * 16;
    long long segs = (cols + target - 1) / target;
    if (segs < 1) segs = 1;
    long long seg_len = (cols + segs - 1) / segs;
    const long long gran = 32LL * vec;
    seg_len = (seg_len + gran - 1) / gran * gran;
    segs = (cols + seg_len - 1) / seg_len;
    if (segs > 0x7fffffffLL || seg_len > 0x7fffffffLL) return ANTQ_ENOTSUP;
    RowsParams p;
    p.x = x; p.out = out; p.codes = codes; p.alpha = alpha; p.cb = cb;
    p.rows = rows; p.cols = cols; p.total_segs = rows * segs;
    p.alpha_per_row = alpha_per_row; p.segs_per_row = (int)segs; p.seg_len = (int)seg_len;
    // a warp walks `segs_per_warp` consecutive segments and rebuilds its tables only when the row
    // changes; cap the grid at ~48 warps per SM worth of work items so long rows / per-tensor
    // views reuse one prologue for many segments.
    const long long max_warps = 148LL * 48;
    long long spw = (p.total_segs + max_warps - 1) / max_warps;
    if (spw < 1) spw = 1;
    if (spw > segs) spw = segs;          // never span rows needlessly
    if (spw > 0x7fffffffLL) return ANTQ_ENOTSUP;
    p.segs_per_warp = (int)spw;
    p.total_warps = (p.total_segs + spw - 1) / spw;
    if (p.total_segs == 0) return 0;
    switch (dtype) {
        case ANTQ_F32: return launch_t<float>(p, nt, sym, ovp, codes != nullptr, st);
        case ANTQ_F16: return launch_t<__half>(p, nt, sym, ovp, codes != nullptr, st);
        case ANTQ_BF16: return launch_t<__nv_bfloat16>(p, nt, sym, ovp, codes != nullptr, st);
    }
    return ANTQ_EINVAL;
}
