// antq_codes.cu -- packed 4-bit code storage ("P4") for quantized tensors: what the reference's kernel allocates and
// never writes (`tensor_idx`, A/quant/quant_kernel.cu:18,49,61), defined here so that a 4-bit weight can be kept as
// 0.5 byte per element + one fp32 scale per row and fed to the dequant-fused GEMM (rank 3).
//
// Format.  Row-major like the tensor; byte j of a row holds element 2j in its LOW nibble and element 2j + 1 in its
// HIGH nibble.  A nibble is the index into `quant_grid` of the level the reference scan selects (the LAST occurrence
// of a duplicated value, e.g. the padded second zero of a signed ANT grid).  OliVe with outliers (normal grid of at
// most 15 entries, as the 4-bit signed int / flint grids are: O/antquant/quant_modules.py:73-153): nibble 15 is the
// OUTLIER IDENTIFIER -- it marks the victim of an outlier-victim pair (value 0), and the other nibble of that byte is
// then an index into `outliers` instead of `quant_grid` (O/...:311-320: even outlier kills odd, else odd kills even).
//
//   antq_encode_p4   x, alpha -> codes       literal reference arithmetic per element (true division, exact
//                                            thresholds / literal scan); counts the elements whose fake-quant value is
//                                            NOT reproduced by decoding the code (out-of-window STE rounding, NaN, Inf)
//   antq_decode_p4   codes, alpha -> values  RN(fl32(level * s)): bit-identical to antq_fakequant wherever the
//                                            count above is zero
#include <type_traits>

#include "antq_common.cuh"

int antq_launch_pu_encode(const void *x, unsigned char *codes, const float *alpha, int alpha_per_row, long long rows, long long cols,
                          int dtype, const AntqCodebook *cb, const antq_codebook_info *info, unsigned int *n_inexact, cudaStream_t st);

namespace {

constexpr int kThreads = 256;

struct CodesParams {
    const void *x;
    unsigned char *codes;
    void *out;
    const float *alpha;
    const AntqCodebook *cb;
    long long n, cols;         // elements; elements per row
    int alpha_per_row, ovp;
    unsigned int *n_inexact;
    int cols_shift;            // log2(cols) when cols is a power of two, else -1 (decode: the row of a thread without a 64-bit division)
};

// one thread = 8 consecutive elements = 4 code bytes
template <typename T, bool OVP> __global__ void __launch_bounds__(kThreads) antq_encode_p4_kernel(const CodesParams p) {
    typedef AntqType<T> A;
    __shared__ float s_thr[32], s_lev[32];
    __shared__ int s_code[32];
    const AntqCodebook *__restrict__ cb = p.cb;
    const int nlev = cb->n_levels, K = cb->n_entries, kn = cb->n_normal;
    if (threadIdx.x < 32) {
        s_lev[threadIdx.x] = threadIdx.x < nlev ? cb->level[threadIdx.x] : 0.0f;
        s_thr[threadIdx.x] = threadIdx.x < nlev - 1 ? cb->thr[threadIdx.x] : __int_as_float(0x7f800000);
        s_code[threadIdx.x] = threadIdx.x < nlev ? cb->level_code[threadIdx.x] : 0;
    }
    __syncthreads();
    const float win = (cb->flags & ANTQ_CB_WELLSEP) ? cb->lim_idx : -1.0f;
    const float gmax = cb->gmax;
    const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i0 >= p.n) return;
    const T *x = reinterpret_cast<const T *>(p.x) + i0;
    const int cnt = (int)((p.n - i0) < 8 ? (p.n - i0) : 8);            // n is even, rows are even: cnt is even
    T xv[8];
    if (cnt == 8 && ((uintptr_t)x % 16 == 0)) {
#pragma unroll
        for (int k = 0; k < (int)(8 * sizeof(T) / 16); k++)
            reinterpret_cast<uint4 *>(xv)[k] = antq_ldg_stream(reinterpret_cast<const uint4 *>(x) + k);
    } else {
        for (int e = 0; e < cnt; e++) xv[e] = x[e];
    }
    long long row = p.alpha_per_row ? i0 / p.cols : 0;
    long long col = p.alpha_per_row ? i0 - row * p.cols : 0;
    float s = __fdiv_rn(p.alpha[row], gmax);
    float q[8];
    int c[8];
    unsigned bad = 0;
#pragma unroll
    for (int e = 0; e < 8; e++) {
        if (e < cnt) {
            if (p.alpha_per_row && col == p.cols) { col = 0; row++; s = __fdiv_rn(p.alpha[row], gmax); }
            col++;
            const float xf = A::to_f32(xv[e]);
            const float d = __fdiv_rn(xf, s);
            int code;
            float qe;
            if (fabsf(d) <= win) {
                const int rank = antq_rank(s_thr, nlev - 1, d);
                qe = s_lev[rank]; code = s_code[rank];
            } else {
                qe = antq_scan_literal(cb->grid, K, d, code);
            }
            q[e] = qe; c[e] = code;
            if (!OVP) {
                // would decoding reproduce the fake-quant value?  (with OVP the pair mask decides: antq_check_p4_ovp_kernel)
                const T ref = A::from_f32_rn(antq_ste_rescale(qe, d, s));
                const T dec = A::from_f32_rn(__fmul_rn(code >= 0 ? qe : 0.0f, s));
                if (A::bits(ref) != A::bits(dec) || code < 0) bad++;
            }
        }
    }
    unsigned packed = 0;
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
        if (e + 1 < cnt) {
            int ne = c[e], no = c[e + 1];
            if (OVP) {
                const bool oe = fabsf(q[e]) > 32.0f, oo = fabsf(q[e + 1]) > 32.0f;
                if (oe) { ne = c[e] - kn; no = 15; }
                else if (oo) { no = c[e + 1] - kn; ne = 15; }
            }
            if (ne < 0 || ne > 15 || no < 0 || no > 15) { bad++; ne &= 15; no &= 15; }
            packed |= (unsigned)(ne | (no << 4)) << (4 * e);
        }
    }
    unsigned char *dst = p.codes + i0 / 2;
    if (cnt == 8 && ((uintptr_t)dst % 4 == 0)) *reinterpret_cast<unsigned *>(dst) = packed;
    else for (int b = 0; b < cnt / 2; b++) dst[b] = (unsigned char)(packed >> (8 * b));
    if (bad && p.n_inexact) atomicAdd(p.n_inexact, bad);
}

// one thread = 16 elements = 8 code bytes
template <typename T, bool OVP> __global__ void __launch_bounds__(kThreads) antq_decode_p4_kernel(const CodesParams p) {
    typedef AntqType<T> A;
    __shared__ float s_grid[32];
    const AntqCodebook *__restrict__ cb = p.cb;
    const int kn = cb->n_normal;
    if (threadIdx.x < 32) s_grid[threadIdx.x] = threadIdx.x < cb->n_entries ? cb->grid[threadIdx.x] : 0.0f;
    __syncthreads();
    const float gmax = cb->gmax;
    const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i0 >= p.n) return;
    const int cnt = (int)((p.n - i0) < 16 ? (p.n - i0) : 16);
    const unsigned char *src = p.codes + i0 / 2;
    unsigned w[2] = {0u, 0u};
    if (cnt == 16 && ((uintptr_t)src % 8 == 0)) {
        const uint2 v = *reinterpret_cast<const uint2 *>(src);
        w[0] = v.x; w[1] = v.y;
    } else {
        for (int b = 0; b < cnt / 2; b++) w[b >> 2] |= (unsigned)src[b] << (8 * (b & 3));
    }
    long long row = 0, col = 0;
    if (p.alpha_per_row) {
        if (p.cols_shift >= 0) row = i0 >> p.cols_shift;
        else if (p.n <= 0xffffffffLL) row = (long long)((unsigned)i0 / (unsigned)p.cols);
        else row = i0 / p.cols;
        col = i0 - row * p.cols;
    }
    float s = __fdiv_rn(p.alpha[row], gmax);
    const bool one_row = !p.alpha_per_row || col + 16 <= p.cols;      // the usual case: all 16 elements share the scale
    T ov[16];
#pragma unroll
    for (int e = 0; e < 16; e += 2) {
        if (e + 1 < cnt) {
            if (!one_row && col == p.cols) { col = 0; row++; s = __fdiv_rn(p.alpha[row], gmax); }
            col += 2;                                                  // rows are even: a pair never straddles one
            const unsigned byte = (w[e >> 3] >> (4 * (e & 7))) & 0xffu;
            const int ne = byte & 15, no = byte >> 4;
            float ve, vo;
            if (OVP && no == 15) { ve = s_grid[kn + ne]; vo = 0.0f; }
            else if (OVP && ne == 15) { vo = s_grid[kn + no]; ve = 0.0f; }
            else { ve = s_grid[ne]; vo = s_grid[no]; }
            ov[e] = A::from_f32_rn(__fmul_rn(ve, s));
            ov[e + 1] = A::from_f32_rn(__fmul_rn(vo, s));
        }
    }
    T *dst = reinterpret_cast<T *>(p.out) + i0;
    if (cnt == 16 && ((uintptr_t)dst % 16 == 0)) {
        const uint4 *srcv = reinterpret_cast<const uint4 *>(ov);
#pragma unroll
        for (int k = 0; k < (int)(16 * sizeof(T) / 16); k++) antq_stg_stream(reinterpret_cast<uint4 *>(dst) + k, srcv[k]);
    } else {
        for (int e = 0; e < cnt; e++) dst[e] = ov[e];
    }
}

// The fast decoder (cols a multiple of the 16-byte vector, 16-byte aligned output, 4-byte aligned codes): a warp owns kDecU
// x 32 consecutive vectors; every load (VEC / 2 code bytes per lane) and every 16-byte store is fully coalesced, and kDecU
// of each are in flight per lane.  The one-thread-16-elements kernel above wrote 16 bytes per lane at a 32-byte stride:
// 18.8 us per 4096^2 fp16 against ~8 here (profiles/r02_notes.md).
constexpr int kDecU = 4;
template <typename T, bool OVP> __global__ void __launch_bounds__(kThreads) antq_decode_p4_fast_kernel(const CodesParams p) {
    typedef AntqType<T> A;
    constexpr int VEC = A::kVec;                                      // elements per 16-byte store: 8 or 4
    typedef typename std::conditional<VEC == 8, unsigned, unsigned short>::type code_t;    // VEC / 2 code bytes
    __shared__ float s_grid[32];
    const AntqCodebook *__restrict__ cb = p.cb;
    const int kn = cb->n_normal;
    if (threadIdx.x < 32) s_grid[threadIdx.x] = threadIdx.x < cb->n_entries ? cb->grid[threadIdx.x] : 0.0f;
    __syncthreads();
    const float gmax = cb->gmax;
    const long long nvec = p.n / VEC;
    const unsigned cols_vec = (unsigned)(p.cols / VEC);
    const long long warp_g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long v0 = warp_g * (kDecU * 32) + lane;
    const code_t *cv = reinterpret_cast<const code_t *>(p.codes);
    uint4 *ov = reinterpret_cast<uint4 *>(p.out);
    unsigned w[kDecU];
#pragma unroll
    for (int j = 0; j < kDecU; j++) {
        const long long v = v0 + j * 32;
        w[j] = v < nvec ? (unsigned)__ldg(cv + v) : 0u;
    }
#pragma unroll
    for (int j = 0; j < kDecU; j++) {
        const long long v = v0 + j * 32;
        if (v >= nvec) break;
        long long row = 0;
        if (p.alpha_per_row) row = p.cols_shift >= 0 ? (v * VEC) >> p.cols_shift : (nvec <= 0xffffffffLL ? (long long)((unsigned)v / cols_vec) : v / cols_vec);
        const float s = __fdiv_rn(__ldg(p.alpha + row), gmax);
        T o[VEC];
#pragma unroll
        for (int e = 0; e < VEC; e += 2) {
            const unsigned byte = (w[j] >> (4 * e)) & 0xffu;
            const int ne = byte & 15, no = byte >> 4;
            float ve, vo;
            if (OVP && no == 15) { ve = s_grid[kn + ne]; vo = 0.0f; }
            else if (OVP && ne == 15) { vo = s_grid[kn + no]; ve = 0.0f; }
            else { ve = s_grid[ne]; vo = s_grid[no]; }
            o[e] = A::from_f32_rn(__fmul_rn(ve, s));
            o[e + 1] = A::from_f32_rn(__fmul_rn(vo, s));
        }
        antq_stg_stream(ov + v, *reinterpret_cast<const uint4 *>(o));
    }
}

// OVP needs the reference value of every element of a pair AFTER the mask; a second small kernel keeps the encoder
// simple: it decodes nothing, it just recomputes fake-quant per pair and compares with what decoding would give.
template <typename T> __global__ void __launch_bounds__(kThreads) antq_check_p4_ovp_kernel(const CodesParams p) {
    typedef AntqType<T> A;
    const AntqCodebook *__restrict__ cb = p.cb;
    const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i0 + 1 >= p.n) return;
    const T *x = reinterpret_cast<const T *>(p.x);
    const long long row = p.alpha_per_row ? i0 / p.cols : 0;
    const float s = __fdiv_rn(p.alpha[row], cb->gmax);
    AntqExact a = antq_exact_quant(cb, A::to_f32(x[i0]), s), b = antq_exact_quant(cb, A::to_f32(x[i0 + 1]), s);
    float qa = a.q, qb = b.q, da_ = a.q, db_ = b.q;                      // decoded values
    const bool oa = fabsf(qa) > 32.0f, ob = fabsf(qb) > 32.0f;
    if (oa) { qb = __fmul_rn(qb, 0.0f); db_ = 0.0f; }
    else if (ob) { qa = __fmul_rn(qa, 0.0f); da_ = 0.0f; }
    unsigned bad = 0;
    bad += A::bits(A::from_f32_rn(antq_ste_rescale(qa, a.d, s))) != A::bits(A::from_f32_rn(__fmul_rn(da_, s))) || a.code < 0;
    bad += A::bits(A::from_f32_rn(antq_ste_rescale(qb, b.d, s))) != A::bits(A::from_f32_rn(__fmul_rn(db_, s))) || b.code < 0;
    if (bad && p.n_inexact) atomicAdd(p.n_inexact, bad);
}

inline int esize(int dtype) { return dtype == ANTQ_F32 ? 4 : (dtype == ANTQ_F16 || dtype == ANTQ_BF16) ? 2 : 0; }

}  // namespace

extern "C" {

int antq_encode_p4(const void *x, uint8_t *codes, const float *alpha, int alpha_per_row, int64_t rows, int64_t cols,
                   int dtype, const void *codebook, const antq_codebook_info *info, int flags, unsigned int *n_inexact,
                   void *stream) {
    if (esize(dtype) == 0 || rows < 0 || cols < 0 || !info) return ANTQ_EINVAL;
    const long long n = rows * cols;
    if (n == 0) return 0;
    if (!x || !codes || !alpha || !codebook) return ANTQ_EINVAL;
    const bool ovp = (flags & ANTQ_FLAG_OVP) != 0 && info->n_entries > info->n_normal;
    if ((cols & 1) || info->n_levels > 32) return ANTQ_ENOTSUP;
    if (ovp ? (info->n_normal > 15 || info->n_entries - info->n_normal > 15) : info->n_entries > 16) return ANTQ_ENOTSUP;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_inexact) {
        cudaError_t e = cudaMemsetAsync(n_inexact, 0, sizeof(unsigned int), st);
        if (e != cudaSuccess) return (int)e;
    }
    if (!ovp && !(flags & ANTQ_FLAG_NO_PU)) {                         // closed form where the grid allows it (antq_pu.cu)
        const int rc = antq_launch_pu_encode(x, codes, alpha, alpha_per_row, rows, cols, dtype, (const AntqCodebook *)codebook, info,
                                             n_inexact, st);
        if (rc != ANTQ_ENOTSUP) return rc;
    }
    CodesParams p = {};
    p.x = x; p.codes = codes; p.alpha = alpha; p.cb = (const AntqCodebook *)codebook;
    p.n = n; p.cols = cols; p.alpha_per_row = alpha_per_row && rows > 1; p.ovp = ovp; p.n_inexact = n_inexact;
    const long long ctas = (n / 8 + kThreads) / kThreads;
    if (ctas > 0x7fffffffLL) return ANTQ_ENOTSUP;
#define ANTQ_ENC(T)                                                                                        \
    do {                                                                                                   \
        if (ovp) {                                                                                         \
            CodesParams q = p; q.n_inexact = nullptr;                                                      \
            antq_encode_p4_kernel<T, true><<<(unsigned)ctas, kThreads, 0, st>>>(q);                        \
            if (n_inexact) antq_check_p4_ovp_kernel<T><<<(unsigned)((n / 2 + kThreads - 1) / kThreads), kThreads, 0, st>>>(p); \
        } else antq_encode_p4_kernel<T, false><<<(unsigned)ctas, kThreads, 0, st>>>(p);                    \
    } while (0)
    switch (dtype) {
        case ANTQ_F32: ANTQ_ENC(float); break;
        case ANTQ_F16: ANTQ_ENC(__half); break;
        case ANTQ_BF16: ANTQ_ENC(__nv_bfloat16); break;
    }
#undef ANTQ_ENC
    return (int)cudaGetLastError();
}

int antq_decode_p4(const uint8_t *codes, void *out, const float *alpha, int alpha_per_row, int64_t rows, int64_t cols,
                   int dtype, const void *codebook, const antq_codebook_info *info, int flags, void *stream) {
    if (esize(dtype) == 0 || rows < 0 || cols < 0 || !info) return ANTQ_EINVAL;
    const long long n = rows * cols;
    if (n == 0) return 0;
    if (!out || !codes || !alpha || !codebook) return ANTQ_EINVAL;
    const bool ovp = (flags & ANTQ_FLAG_OVP) != 0 && info->n_entries > info->n_normal;
    if ((cols & 1) || info->n_entries > 31) return ANTQ_ENOTSUP;
    if (ovp ? (info->n_normal > 15 || info->n_entries - info->n_normal > 15) : info->n_entries > 16) return ANTQ_ENOTSUP;
    cudaStream_t st = (cudaStream_t)stream;
    CodesParams p = {};
    p.codes = const_cast<unsigned char *>(codes); p.out = out; p.alpha = alpha; p.cb = (const AntqCodebook *)codebook;
    p.n = n; p.cols = cols; p.alpha_per_row = alpha_per_row && rows > 1; p.ovp = ovp;
    p.cols_shift = -1;
    if (cols > 0 && (cols & (cols - 1)) == 0) { int sh = 0; while ((1LL << sh) < cols) sh++; p.cols_shift = sh; }
    const int vec = 16 / esize(dtype);
    if (cols % vec == 0 && (uintptr_t)out % 16 == 0 && (uintptr_t)codes % 4 == 0) {
        const long long warps = (n / vec + kDecU * 32 - 1) / (kDecU * 32);
        const long long fctas = (warps * 32 + kThreads - 1) / kThreads;
        if (fctas <= 0x7fffffffLL) {
#define ANTQ_DECF(T)                                                                              \
    do {                                                                                          \
        if (ovp) antq_decode_p4_fast_kernel<T, true><<<(unsigned)fctas, kThreads, 0, st>>>(p);    \
        else antq_decode_p4_fast_kernel<T, false><<<(unsigned)fctas, kThreads, 0, st>>>(p);       \
    } while (0)
            switch (dtype) {
                case ANTQ_F32: ANTQ_DECF(float); break;
                case ANTQ_F16: ANTQ_DECF(__half); break;
                case ANTQ_BF16: ANTQ_DECF(__nv_bfloat16); break;
            }
#undef ANTQ_DECF
            return (int)cudaGetLastError();
        }
    }
    const long long ctas = (n / 16 + kThreads) / kThreads;
    if (ctas > 0x7fffffffLL) return ANTQ_ENOTSUP;
#define ANTQ_DEC(T)                                                                          \
    do {                                                                                     \
        if (ovp) antq_decode_p4_kernel<T, true><<<(unsigned)ctas, kThreads, 0, st>>>(p);     \
        else antq_decode_p4_kernel<T, false><<<(unsigned)ctas, kThreads, 0, st>>>(p);        \
    } while (0)
    switch (dtype) {
        case ANTQ_F32: ANTQ_DEC(float); break;
        case ANTQ_F16: ANTQ_DEC(__half); break;
        case ANTQ_BF16: ANTQ_DEC(__nv_bfloat16); break;
    }
#undef ANTQ_DEC
    return (int)cudaGetLastError();
}

}  // extern "C"
