// antq_common.cuh -- device-side codebook layout and exact-arithmetic helpers
// shared by every kernel in libantq.so (sm_100a only).
//
// Reference semantics being reproduced (A/ = ant_quantization/, O/ = olive_quantization/):
//   scan      A/quant/quant_kernel.cu:25-37   best = 102400, z = 0, `<=` keeps the LAST minimum
//   _forward  A/antquant/quant_modules.py:535-551   s = alpha/max(grid); d = x/s; q = scan(d);
//                                                  t = (q - d) + d; out = t * s   (all fp32, IEEE)
//   OVP       O/antquant/quant_modules.py:311-320
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/antq.h"

#define ANTQ_CB_MAGIC 0x414e5451  // "ANTQ"

// One prepared codebook, resident in device memory (antq_codebook_bytes()).
struct AntqCodebook {
    int32_t n_entries;   // K = k_normal + k_out, in scan order
    int32_t n_normal;
    int32_t n_levels;    // distinct non-NaN values, ascending
    int32_t flags;       // ANTQ_CB_*
    int32_t n_mag;       // SYMMETRIC: number of magnitudes including 0
    int32_t mid;         // SYMMETRIC: index of the zero level
    int32_t ovp_index;   // index of the first threshold whose upper level is an outlier (|v| > 32):
                         //   SYMMETRIC: into mag_tpos/mag_tneg, else into thr (positive side only); -1 = none
    int32_t magic;
    float gmax;          // max(quant_grid): the reference's scale denominator
    float vmax, vmin;    // extreme levels
    float lim;           // window |d| <= lim in which (q - d) + d == q AND the threshold search == scan
    float lim_idx;       // window |d| <= lim_idx in which the threshold search == scan (no STE claim)
    float pad_[3];
    float grid[ANTQ_MAX_GRID];          // scan order (for the literal slow path)
    float level[ANTQ_MAX_GRID];         // sorted distinct values (+0 canonical)
    int32_t level_code[ANTQ_MAX_GRID];  // scan index of each level (last occurrence)
    float thr[ANTQ_MAX_GRID];           // thr[r] = min{d : level r+1 wins the scan over level r}
    float mag_tpos[ANTQ_MAX_GRID / 2];  // SYMMETRIC: d >= 0 : magnitude k+1 wins iff  d >= mag_tpos[k]
    float mag_tneg[ANTQ_MAX_GRID / 2];  //            d <  0 : magnitude k+1 wins iff -d >= mag_tneg[k]
    // ANTQ_CB_PU (piecewise-uniform closed form, antq_pu.cu): every level is fl32(k * pu_c) for an integer k in
    // [pu_kmin, pu_kmax], and inside each octave of |k| the k's form a progression with a power-of-two step.
    float pu_c, pu_inv_c, pu_kmin, pu_kmax;
    float pu_tout;       // ANTQ_CB_PU_OVP: |d| < pu_tout never reaches an outlier level (else +inf)
    float2 pu_tab[256];                 // by biased exponent of t = d / pu_c: {1.5 * 2^23 * step, near-midpoint delta}
};

// ---- total order on fp32 bit patterns (-0 directly below +0) ---------------
__device__ __forceinline__ int antq_f2ord(float x) {
    int b = __float_as_int(x);
    return b < 0 ? -(b & 0x7fffffff) - 1 : b;
}
__device__ __forceinline__ float antq_ord2f(int o) {
    return __int_as_float(o < 0 ? ((-(o + 1)) | 0x80000000) : o);
}

// ---- the reference scan, literally (A/quant/quant_kernel.cu:25-37) ----------
__device__ __forceinline__ float antq_scan_literal(const float *__restrict__ g, int k, float x_v, int &code) {
    float sub_min = 102400.0f;
    float z_min = 0.0f;
    code = ANTQ_CODE_NONE;
    for (int i = 0; i < k; i++) {
        float gi = __ldg(g + i);
        float sub_v = fabsf(__fsub_rn(x_v, gi));
        if (sub_v <= sub_min) {
            sub_min = sub_v;
            z_min = gi;
            code = i;
        }
    }
    return z_min;
}

// rank = #{r < nt : d >= thr[r]} by branch-free binary search over a sorted array.
__device__ __forceinline__ int antq_rank(const float *__restrict__ thr, int nt, float d) {
    int lo = 0, n = nt;
    while (n > 0) {
        int half = n >> 1;
        bool ge = d >= thr[lo + half];
        lo = ge ? lo + half + 1 : lo;
        n = ge ? n - half - 1 : half;
    }
    return lo;
}

// ---- element type traits ----------------------------------------------------
template <typename T> struct AntqType;

template <> struct AntqType<float> {
    static constexpr int kVec = 4;  // elements per 16 bytes
    typedef unsigned int bits_t;
    __device__ static __forceinline__ float to_f32(float v) { return v; }
    __device__ static __forceinline__ float from_f32_rn(float v) { return v; }
    __device__ static __forceinline__ float from_f32_rz(float v) { return v; }
    __device__ static __forceinline__ float from_f32_ru(float v) { return v; }
    __device__ static __forceinline__ bits_t bits(float v) { return __float_as_uint(v); }
    __device__ static __forceinline__ float from_bits(bits_t b) { return __uint_as_float(b); }
    static constexpr bits_t kSign = 0x80000000u, kInf = 0x7f800000u;
};
template <> struct AntqType<__half> {
    static constexpr int kVec = 8;
    typedef unsigned short bits_t;
    __device__ static __forceinline__ float to_f32(__half v) { return __half2float(v); }
    __device__ static __forceinline__ __half from_f32_rn(float v) { return __float2half_rn(v); }
    __device__ static __forceinline__ __half from_f32_rz(float v) { return __float2half_rz(v); }
    __device__ static __forceinline__ __half from_f32_ru(float v) { return __float2half_ru(v); }
    static constexpr unsigned int kDropMask = 0x1fffu;          // fp32 has 13 more mantissa bits
    static constexpr float kSafeMin = 1.220703125e-4f;          // 2^-13: stay clear of fp16 subnormals
    static constexpr float kSafeMax = 60000.0f;
    __device__ static __forceinline__ bits_t bits(__half v) { return __half_as_ushort(v); }
    __device__ static __forceinline__ __half from_bits(bits_t b) { return __ushort_as_half(b); }
    static constexpr bits_t kSign = 0x8000u, kInf = 0x7c00u;
};
template <> struct AntqType<__nv_bfloat16> {
    static constexpr int kVec = 8;
    typedef unsigned short bits_t;
    __device__ static __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
    __device__ static __forceinline__ __nv_bfloat16 from_f32_rn(float v) { return __float2bfloat16_rn(v); }
    __device__ static __forceinline__ __nv_bfloat16 from_f32_rz(float v) { return __float2bfloat16_rz(v); }
    __device__ static __forceinline__ __nv_bfloat16 from_f32_ru(float v) { return __float2bfloat16_ru(v); }
    static constexpr unsigned int kDropMask = 0xffffu;          // fp32 has 16 more mantissa bits
    static constexpr float kSafeMin = 1e-30f;
    static constexpr float kSafeMax = 1e38f;
    __device__ static __forceinline__ bits_t bits(__nv_bfloat16 v) { return __bfloat16_as_ushort(v); }
    __device__ static __forceinline__ __nv_bfloat16 from_bits(bits_t b) { return __ushort_as_bfloat16(b); }
    static constexpr bits_t kSign = 0x8000u, kInf = 0x7f80u;
};

// next representable value above / below (sign-magnitude stepping; inf saturates)
template <typename T> __device__ __forceinline__ T antq_next_up(T v) {
    typedef AntqType<T> A;
    typename A::bits_t b = A::bits(v);
    typename A::bits_t mag = b & (typename A::bits_t)~A::kSign;
    if (b & A::kSign) {
        if (mag == 0) return A::from_bits(1);            // -0 -> smallest positive
        return A::from_bits((typename A::bits_t)(b - 1));  // toward zero (-inf -> -max)
    }
    if (mag == A::kInf) return v;
    return A::from_bits((typename A::bits_t)(b + 1));
}
template <typename T> __device__ __forceinline__ T antq_next_down(T v) {
    typedef AntqType<T> A;
    typename A::bits_t b = A::bits(v);
    typename A::bits_t mag = b & (typename A::bits_t)~A::kSign;
    if (b & A::kSign) {
        if (mag == A::kInf) return v;
        return A::from_bits((typename A::bits_t)(b + 1));
    }
    if (mag == 0) return A::from_bits((typename A::bits_t)(A::kSign | 1));  // +0 -> smallest negative
    return A::from_bits((typename A::bits_t)(b - 1));
}
template <typename T> __device__ __forceinline__ bool antq_is_inf(T v) {
    typedef AntqType<T> A;
    return (typename A::bits_t)(A::bits(v) & (typename A::bits_t)~A::kSign) == A::kInf;
}

// min{x in T : fl32(f32(x) / s) >= t}, for finite s > 0 and finite t.
// fl32(x / s) is monotone in x, and RN_T(fl32(t * s)) is within one T-ulp of the
// answer, so two short monotone walks settle it exactly (tests/xspace_model.py).
template <typename T> __device__ __noinline__ T antq_x_threshold_exact(float t, float s) {
    typedef AntqType<T> A;
    T c = A::from_f32_rn(__fmul_rn(t, s));
#pragma unroll 1
    for (int it = 0; it < 6; ++it) {
        T p = antq_next_down(c);
        if (antq_is_inf(p) || !(__fdiv_rn(A::to_f32(p), s) >= t)) break;
        c = p;
    }
#pragma unroll 1
    for (int it = 0; it < 6; ++it) {
        if (__fdiv_rn(A::to_f32(c), s) >= t) break;
        c = antq_next_up(c);
    }
    return c;
}

// Division-free common case for 16-bit element types.  The real boundary
// x* = (rounding boundary just below t) * s lies within a few fp32 ulps of
// p = fl32(t * s); a 16-bit type has 13 (fp16) / 16 (bf16) fewer mantissa bits, so unless
// p sits within 16 fp32-ulps of a representable 16-bit value, the answer is simply p
// rounded up.  `*near` reports that the shortcut was NOT provable and the exact walk ran
// (callers use it to decide whether positive/negative tie handling can differ).
template <typename T> __device__ __forceinline__ T antq_x_threshold(float t, float s, bool *near) {
    typedef AntqType<T> A;
    if constexpr (sizeof(T) == 2) {
        const float p = __fmul_rn(t, s);
        const unsigned int pb = __float_as_uint(p);
        const unsigned int drop = A::kDropMask;                     // mantissa bits T does not have
        const unsigned int low = pb & drop;
        const float ap = fabsf(p);
        const bool safe = low >= 16u && low <= drop - 16u && ap >= A::kSafeMin && ap <= A::kSafeMax;
        if (safe) {
            *near = false;
            return A::from_f32_ru(p);
        }
    }
    *near = true;
    return antq_x_threshold_exact<T>(t, s);
}

// 128-bit streaming loads / stores (read once, written once: keep them out of L1).
__device__ __forceinline__ uint4 antq_ldg_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void antq_stg_stream(uint4 *p, const uint4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}

// ---- TMA bulk copy (global -> shared) completing on an mbarrier ---------------------
// cp.async.bulk moves a contiguous, 16-byte aligned span without occupying registers or
// LSU issue slots while it is in flight; SASS: UBLKCP + SYNCS.
__device__ __forceinline__ uint32_t antq_smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void antq_mbar_init(uint64_t *bar, unsigned count) {
    // NOTE: no fence.mbarrier_init.release.cluster here -- ptxas lowers any cluster-scope fence to
    // CCTL.IVALL (an SM-wide L1 invalidate), which serialised every CTA start (profiles/r01_notes.md).
    // The barrier is CTA-local: fence.proxy.async (issued by the same lane before the copy) + __syncwarp suffice.
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(antq_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void antq_fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void antq_bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(antq_smem_u32(bar)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     antq_smem_u32(dst_smem)),
                 "l"(__cvta_generic_to_global(src_gmem)), "r"(bytes), "r"(antq_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void antq_mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "ANTQ_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra ANTQ_DONE;\n"
        "bra ANTQ_WAIT;\n"
        "ANTQ_DONE:\n"
        "}\n" ::"r"(antq_smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// Reference arithmetic for ONE element, literally:  d = x / s; q = scan(d).
struct AntqExact {
    float d, q;
    int code;
};
__device__ __forceinline__ AntqExact antq_exact_quant(const AntqCodebook *__restrict__ cb, float x, float s) {
    AntqExact e;
    e.d = __fdiv_rn(x, s);
    e.q = antq_scan_literal(cb->grid, cb->n_entries, e.d, e.code);
    return e;
}
// t = (q - d) + d ; out = t * s      (A/antquant/quant_modules.py:544-549)
__device__ __forceinline__ float antq_ste_rescale(float q, float d, float s) {
    return __fmul_rn(__fadd_rn(__fsub_rn(q, d), d), s);
}

int antq_launch_prepare(const float *grid, int k_normal, const float *outliers, int k_out, AntqCodebook *cb,
                        cudaStream_t stream);
// SM count of the current device (queried once per device; grids are sized in multiples of it).
int antq_num_sms();
