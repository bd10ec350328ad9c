// antq_short.cu -- fused fake-quant for SHORT rows and scale GROUPS (cols < 512: group-8/16/32 scales, 1x1-conv and
// depthwise weights), where a table of x-space thresholds per row (antq_stream.cu) would cost more than the row itself.
//
// Same reference arithmetic, literally (A/antquant/quant_modules.py:535-551, O/...:295-330):
//   s = alpha / max(grid);  d = fl32(x / s);  q = scan(d);  out = fl32(((q - d) + d) * s)
// but the scan becomes a chain of compares in d-SPACE, where the thresholds are constants of the codebook (the exact
// bisection results of antq_prepare.cu, held in registers) instead of a binary search through shared memory, and each
// thread owns whole 16-byte vectors.  Symmetric codebooks compare |d| with the positive- or negative-side threshold
// (they differ by at most one ulp: exact ties go to the LATER grid entry, i.e. up for d > 0 and toward zero for d < 0).
// Elements outside the window in which the threshold chain provably equals the scan (|d| > lim_idx, NaN, Inf) and rows
// whose scale is not a positive finite number take the literal scan.
//
// Bound: one IEEE division + ~3 (NT + 3) fp32 instructions per ELEMENT -> instruction issue, not HBM
// (measured: profiles/r01_notes.md).
#include <stdio.h>

#include "antq_common.cuh"

namespace {

constexpr int kShortThreads = 256;
constexpr int kShortCtasPerSm = 6;

struct ShortParams {
    const void *x;
    void *out;
    const float *alpha;
    const AntqCodebook *cb;
    unsigned nvec;            // 16-byte vectors in the tensor
    unsigned cols_vec;        // vectors per row
    int cols_shift;           // log2(cols_vec) when it is a power of two, else -1
    int alpha_per_row;
    int nt_real, mid;
    float gmax;
};

template <typename T, bool OVP>
__device__ __noinline__ uint4 antq_short_slow_vec(const AntqCodebook *__restrict__ cb, float s, const uint4 raw) {
    typedef AntqType<T> A;
    constexpr int VEC = A::kVec;
    T xv[VEC], ov[VEC];
    *reinterpret_cast<uint4 *>(xv) = raw;
    float q[VEC], d[VEC];
#pragma unroll
    for (int e = 0; e < VEC; e++) {
        AntqExact ex = antq_exact_quant(cb, A::to_f32(xv[e]), s);
        q[e] = ex.q; d[e] = ex.d;
    }
    if (OVP) {
#pragma unroll
        for (int e = 0; e + 1 < VEC; e += 2) {
            const bool oe = fabsf(q[e]) > 32.0f, oo = fabsf(q[e + 1]) > 32.0f;
            if (oe) q[e + 1] = __fmul_rn(q[e + 1], 0.0f);
            else if (oo) q[e] = __fmul_rn(q[e], 0.0f);
        }
    }
#pragma unroll
    for (int e = 0; e < VEC; e++) ov[e] = A::from_f32_rn(antq_ste_rescale(q[e], d[e], s));
    return *reinterpret_cast<const uint4 *>(ov);
}

// SYM:  NT thresholds on |d| (mag_tpos / mag_tneg), levels level[mid .. mid + NT]
// XNEG: (with SYM) one extra level level[0] below the most negative common one, chosen iff d < thr[0]
// !SYM: NT thresholds thr[] on d, levels level[0 .. NT]
template <typename T, int NT, bool SYM, bool XNEG, bool OVP>
__global__ void __launch_bounds__(kShortThreads, NT <= 7 ? kShortCtasPerSm : 3) antq_short_kernel(const ShortParams p) {
    typedef AntqType<T> A;
    constexpr int VEC = A::kVec;
    const AntqCodebook *__restrict__ cb = p.cb;
    const float inf = __int_as_float(0x7f800000);
    float tp[NT], tn[SYM ? NT : 1], lv[NT + 1];
#pragma unroll
    for (int i = 0; i < NT; i++) {
        const bool real = i < p.nt_real;
        tp[i] = real ? (SYM ? cb->mag_tpos[i] : cb->thr[i]) : inf;
        if (SYM) tn[i] = real ? cb->mag_tneg[i] : inf;
    }
#pragma unroll
    for (int i = 0; i <= NT; i++) lv[i] = i <= p.nt_real ? cb->level[(SYM ? p.mid : 0) + i] : 0.0f;
    const float thr_e = XNEG ? cb->thr[0] : 0.0f, lev_e = XNEG ? cb->level[0] : 0.0f;
    const float win = cb->lim_idx;            // |d| <= win  =>  the threshold chain provably equals the scan

    const uint4 *xin = reinterpret_cast<const uint4 *>(p.x);
    uint4 *xout = reinterpret_cast<uint4 *>(p.out);
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < p.nvec; v += stride) {
        const uint4 raw = antq_ldg_stream(xin + v);
        unsigned row = 0;
        if (p.alpha_per_row) row = p.cols_shift >= 0 ? v >> p.cols_shift : v / p.cols_vec;
        const float s = __fdiv_rn(__ldg(p.alpha + row), p.gmax);          // scale = alpha / max(grid)
        T xv[VEC], ov[VEC];
        *reinterpret_cast<uint4 *>(xv) = raw;
        bool special = !(s > 0.0f && s < inf);
        float q[VEC], d[VEC];
        if (!special) {
#pragma unroll
            for (int e = 0; e < VEC; e++) {
                const float de = __fdiv_rn(A::to_f32(xv[e]), s);
                d[e] = de;
                const float ad = fabsf(de);
                special |= !(ad <= win);                              // outside the proven window, NaN, Inf
                float qe = lv[0];
                if (SYM) {
                    const bool neg = __float_as_int(de) < 0;
#pragma unroll
                    for (int i = 0; i < NT; i++) qe = ad >= (neg ? tn[i] : tp[i]) ? lv[i + 1] : qe;
                    qe = __uint_as_float(__float_as_uint(qe) | (__float_as_uint(de) & 0x80000000u));
                    if (XNEG) qe = de < thr_e ? lev_e : qe;
                } else {
#pragma unroll
                    for (int i = 0; i < NT; i++) qe = de >= tp[i] ? lv[i + 1] : qe;
                }
                q[e] = qe;
            }
        }
        uint4 res;
        if (special) {
            res = antq_short_slow_vec<T, OVP>(cb, s, raw);
        } else {
            if (OVP) {
#pragma unroll
                for (int e = 0; e + 1 < VEC; e += 2) {
                    const bool oe = fabsf(q[e]) > 32.0f, oo = fabsf(q[e + 1]) > 32.0f;
                    if (oe) q[e + 1] = __fmul_rn(q[e + 1], 0.0f);
                    else if (oo) q[e] = __fmul_rn(q[e], 0.0f);
                }
            }
#pragma unroll
            for (int e = 0; e < VEC; e++) ov[e] = A::from_f32_rn(antq_ste_rescale(q[e], d[e], s));
            res = *reinterpret_cast<const uint4 *>(ov);
        }
        antq_stg_stream(xout + v, res);
    }
}

template <typename T, int NT, bool SYM, bool XNEG, bool OVP> int launch_one(const ShortParams &p, cudaStream_t st) {
    const long long want = ((long long)p.nvec + kShortThreads - 1) / kShortThreads;
    const long long cap = (long long)antq_num_sms() * kShortCtasPerSm;
    const int ctas = (int)(want < cap ? want : cap);
    antq_short_kernel<T, NT, SYM, XNEG, OVP><<<ctas, kShortThreads, 0, st>>>(p);
    return (int)cudaGetLastError();
}

template <typename T, bool SYM, bool XNEG, bool OVP> int launch_nt(const ShortParams &p, int nt, cudaStream_t st) {
    if (nt <= 3) return launch_one<T, 3, SYM, XNEG, OVP>(p, st);
    if (nt <= 7) return launch_one<T, 7, SYM, XNEG, OVP>(p, st);
    if (nt <= 15) return launch_one<T, 15, SYM, XNEG, OVP>(p, st);
    return ANTQ_ENOTSUP;
}

template <typename T> int launch_t(const ShortParams &p, int nt, bool sym, bool symx, bool ovp, cudaStream_t st) {
    if (symx) return launch_nt<T, true, true, false>(p, nt, st);
    if (sym) return ovp ? launch_nt<T, true, false, true>(p, nt, st) : launch_nt<T, true, false, false>(p, nt, st);
    return ovp ? launch_nt<T, false, false, true>(p, nt, st) : launch_nt<T, false, false, false>(p, nt, st);
}

}  // namespace

// Thresholds after folding signs (what antq_fakequant_plan compares with 15).
int antq_short_thresholds(const antq_codebook_info *info, bool ovp) {
    const bool sym = (info->flags & ANTQ_CB_SYMMETRIC) != 0;
    const bool symx = !sym && (info->flags & ANTQ_CB_SYMX) && !ovp;
    return (sym || symx) ? info->n_mag - 1 : info->n_levels - 1;
}

int antq_launch_short(const void *x, void *out, const float *alpha, int alpha_per_row, long long rows, long long cols,
                      int dtype, const AntqCodebook *cb, const antq_codebook_info *info, bool ovp, cudaStream_t st) {
    const int es = dtype == ANTQ_F32 ? 4 : 2;
    const int vec = 16 / es;
    const long long n = rows * cols;
    if (n == 0) return 0;
    if (cols % vec || (n / vec) > 0x7fffffffLL) return ANTQ_ENOTSUP;
    const bool sym = (info->flags & ANTQ_CB_SYMMETRIC) != 0;
    const bool symx = !sym && (info->flags & ANTQ_CB_SYMX) && !ovp;
    const int nt = antq_short_thresholds(info, ovp);
    ShortParams p;
    p.x = x; p.out = out; p.alpha = alpha; p.cb = cb;
    p.nvec = (unsigned)(n / vec);
    p.cols_vec = (unsigned)(cols / vec);
    p.cols_shift = -1;
    if ((p.cols_vec & (p.cols_vec - 1)) == 0) {
        int sh = 0;
        while ((1u << sh) < p.cols_vec) sh++;
        p.cols_shift = sh;
    }
    p.alpha_per_row = alpha_per_row;
    p.nt_real = nt; p.mid = info->mid;
    p.gmax = info->gmax;
    switch (dtype) {
        case ANTQ_F32: return launch_t<float>(p, nt, sym, symx, ovp, st);
        case ANTQ_F16: return launch_t<__half>(p, nt, sym, symx, ovp, st);
        case ANTQ_BF16: return launch_t<__nv_bfloat16>(p, nt, sym, symx, ovp, st);
    }
    return ANTQ_EINVAL;
}
