// antq_flat.cu -- generic kernels over the FLAT tensor: every shape, alignment,
// codebook size (<= 512) and the odd-numel OVP wrap-around.  They do the
// reference arithmetic explicitly per element (true division, STE sum, rescale),
// and only replace the O(K) scan by a binary search over the prepared thresholds
// (falling back to the literal scan outside the proven window).
//
//   antq_flat_kernel<SCALE=true>   Quantizer._forward            A/antquant/quant_modules.py:535-551
//                                  OliVe _forward + OVP          O/antquant/quant_modules.py:295-330
//   antq_flat_kernel<SCALE=false>  quant_cuda.quant(x, y)        A/quant/quant_kernel.cu:11-39
//   antq_absmax_kernel             x.abs().max() / per-row max   A/antquant/quant_modules.py:473-477
//   antq_mse_sweep_kernel          search_mse candidate loop     A/antquant/quant_modules.py:299-306,317-324
#include "antq_common.cuh"

namespace {

constexpr int kFlatThreads = 256;
constexpr int kGroup = 8;   // contiguous elements per thread (pairs never straddle a group)

struct Quantized {
    float q;
    int code;
};

__device__ __forceinline__ Quantized antq_quantize_d(const AntqCodebook *__restrict__ cb, const float *s_thr,
                                                     const float *s_lev, int nlev, float win, float d,
                                                     bool want_code) {
    Quantized r;
    if (fabsf(d) <= win) {      // win < 0 disables the threshold search (codebook not well separated)
        const int rank = antq_rank(s_thr, nlev - 1, d);
        r.q = s_lev[rank];
        r.code = want_code ? cb->level_code[rank] : 0;
    } else {
        r.q = antq_scan_literal(cb->grid, cb->n_entries, d, r.code);
    }
    return r;
}

template <typename T, bool SCALE, bool OVP>
__global__ void __launch_bounds__(kFlatThreads)
antq_flat_kernel(const T *__restrict__ x, T *__restrict__ out, int16_t *__restrict__ codes,
                 const float *__restrict__ alpha, int alpha_per_row, long long n, long long cols,
                 const AntqCodebook *__restrict__ cb, int vec_ok) {
    typedef AntqType<T> A;
    __shared__ float s_thr[ANTQ_MAX_GRID];
    __shared__ float s_lev[ANTQ_MAX_GRID];
    const int nlev = cb->n_levels;
    for (int i = threadIdx.x; i < nlev; i += blockDim.x) {
        s_lev[i] = cb->level[i];
        s_thr[i] = cb->thr[i];
    }
    __syncthreads();
    const float fast = ((cb->flags & ANTQ_CB_WELLSEP) != 0 && nlev >= 1) ? cb->lim_idx : -1.0f;
    const float gmax = cb->gmax;
    const int K = cb->n_entries;

    const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * kGroup;
    if (i0 >= n) return;
    const int cnt = (int)((n - i0) < kGroup ? (n - i0) : kGroup);

    T xv[kGroup];
    if (vec_ok && cnt == kGroup) {
        const uint4 *src = reinterpret_cast<const uint4 *>(x + i0);
        uint4 *dst = reinterpret_cast<uint4 *>(xv);
#pragma unroll
        for (int k = 0; k < (int)(kGroup * sizeof(T) / 16); k++) dst[k] = antq_ldg_stream(src + k);
    } else {
#pragma unroll
        for (int e = 0; e < kGroup; e++)
            if (e < cnt) xv[e] = x[i0 + e];
    }

    long long row = 0, col = i0;
    float s = 1.0f;
    if (SCALE) {
        if (alpha_per_row) { row = i0 / cols; col = i0 - row * cols; }
        s = __fdiv_rn(alpha[alpha_per_row ? row : 0], gmax);
    }
    float q[kGroup], d[kGroup], sc[kGroup];
    int c[kGroup];
#pragma unroll
    for (int e = 0; e < kGroup; e++) {
        if (e < cnt) {
            if (SCALE && alpha_per_row && col == cols) {
                col = 0; row++;
                s = __fdiv_rn(alpha[row], gmax);
            }
            col++;
            const float xf = A::to_f32(xv[e]);
            d[e] = SCALE ? __fdiv_rn(xf, s) : xf;
            sc[e] = s;
            Quantized r = antq_quantize_d(cb, s_thr, s_lev, nlev, fast, d[e], codes != nullptr);
            q[e] = r.q; c[e] = r.code;
        }
    }
    if (OVP) {
#pragma unroll
        for (int e = 0; e < kGroup; e += 2) {
            if (e + 1 < cnt) {
                const bool oe = fabsf(q[e]) > 32.0f, oo = fabsf(q[e + 1]) > 32.0f;
                if (oe) { q[e + 1] = __fmul_rn(q[e + 1], 0.0f); c[e + 1] = K; }
                else if (oo) { q[e] = __fmul_rn(q[e], 0.0f); c[e] = K; }
            } else if (e < cnt) {
                // odd numel: torch.roll pairs the last element with element 0 (O/...:317)
                const float s0 = __fdiv_rn(alpha[0], gmax);
                const float d0 = __fdiv_rn(A::to_f32(x[0]), s0);
                Quantized r0 = antq_quantize_d(cb, s_thr, s_lev, nlev, fast, d0, false);
                if (fabsf(r0.q) > 32.0f) { q[e] = __fmul_rn(q[e], 0.0f); c[e] = K; }
            }
        }
    }
    T ov[kGroup];
#pragma unroll
    for (int e = 0; e < kGroup; e++)
        if (e < cnt) ov[e] = A::from_f32_rn(SCALE ? antq_ste_rescale(q[e], d[e], sc[e]) : q[e]);
    if (vec_ok && cnt == kGroup) {
        uint4 *dst = reinterpret_cast<uint4 *>(out + i0);
        const uint4 *src = reinterpret_cast<const uint4 *>(ov);
#pragma unroll
        for (int k = 0; k < (int)(kGroup * sizeof(T) / 16); k++) antq_stg_stream(dst + k, src[k]);
    } else {
#pragma unroll
        for (int e = 0; e < kGroup; e++)
            if (e < cnt) out[i0 + e] = ov[e];
    }
    if (codes) {
#pragma unroll
        for (int e = 0; e < kGroup; e++)
            if (e < cnt) codes[i0 + e] = (int16_t)c[e];
    }
}

// ---- abs-max -------------------------------------------------------------------
// |x| as a non-negative fp32 compares like its bit pattern, and a NaN's pattern is
// above +Inf's, so an integer max reproduces torch's NaN-propagating abs().max().
template <typename T>
__global__ void __launch_bounds__(256) antq_absmax_kernel(const T *__restrict__ x, unsigned int *__restrict__ out,
                                                          long long cols, int splits) {
    typedef AntqType<T> A;
    const long long row = blockIdx.x / splits;
    const int sp = blockIdx.x % splits;
    const long long per = (cols + splits - 1) / splits;
    const long long c0 = sp * per, c1 = (c0 + per) < cols ? (c0 + per) : cols;
    const T *xr = x + row * cols;
    unsigned int m = 0;
    for (long long c = c0 + threadIdx.x; c < c1; c += blockDim.x) {
        unsigned int b = __float_as_uint(A::to_f32(xr[c])) & 0x7fffffffu;
        m = b > m ? b : m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned int t = __shfl_xor_sync(0xffffffffu, m, o);
        m = t > m ? t : m;
    }
    __shared__ unsigned int sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int wi = 1; wi < (int)(blockDim.x >> 5); wi++) m = sm[wi] > m ? sm[wi] : m;
        atomicMax(out + row, m);
    }
}

// ---- fused alpha sweep -----------------------------------------------------------
// err[c, r] = sum over the row of (fakequant(x; alpha = base[r] * ratio[c]) - x)^2, computed with the
// reference arithmetic per element; x is read once per kCandTile candidates.
constexpr int kCandTile = 8;
constexpr int kSweepThreads = 256;

template <typename T, bool OVP>
__global__ void __launch_bounds__(kSweepThreads)
antq_mse_sweep_kernel(const T *__restrict__ x, const float *__restrict__ base_alpha, int alpha_per_row,
                      const float *__restrict__ ratios, int n_cand, double *__restrict__ err, long long rows,
                      long long cols, int chunks_per_row, const AntqCodebook *__restrict__ cb) {
    typedef AntqType<T> A;
    __shared__ float s_thr[ANTQ_MAX_GRID];
    __shared__ float s_lev[ANTQ_MAX_GRID];
    __shared__ float s_red[kSweepThreads / 32][kCandTile];
    const int nlev = cb->n_levels;
    for (int i = threadIdx.x; i < nlev; i += blockDim.x) {
        s_lev[i] = cb->level[i];
        s_thr[i] = cb->thr[i];
    }
    __syncthreads();
    const float fast = ((cb->flags & ANTQ_CB_WELLSEP) != 0 && nlev >= 1) ? cb->lim_idx : -1.0f;
    const float gmax = cb->gmax;
    const long long row = blockIdx.x / chunks_per_row;
    const int chunk = blockIdx.x % chunks_per_row;
    // chunk boundaries are even so that OVP pairs (flat index 2k, 2k+1; cols even) stay inside one chunk
    long long per = (cols + chunks_per_row - 1) / chunks_per_row;
    per = (per + 1) & ~1LL;
    const long long c0 = chunk * per, c1 = (c0 + per) < cols ? (c0 + per) : cols;
    const T *xr = x + row * cols;
    const float base = base_alpha[alpha_per_row ? row : 0];

    for (int ct = 0; ct < n_cand; ct += kCandTile) {
        float sc[kCandTile], acc[kCandTile];
#pragma unroll
        for (int k = 0; k < kCandTile; k++) {
            const float ratio = ratios[(ct + k) < n_cand ? (ct + k) : (n_cand - 1)];
            sc[k] = __fdiv_rn(__fmul_rn(base, ratio), gmax);    // new_alpha = base * (i * 0.01); scale = alpha / max
            acc[k] = 0.0f;
        }
        for (long long c = c0 + 2LL * threadIdx.x; c < c1; c += 2LL * blockDim.x) {
            const bool has2 = (c + 1) < c1;
            const float xa = A::to_f32(xr[c]);
            const float xb = has2 ? A::to_f32(xr[c + 1]) : 0.0f;
#pragma unroll
            for (int k = 0; k < kCandTile; k++) {
                const float s = sc[k];
                const float da = __fdiv_rn(xa, s), db = __fdiv_rn(xb, s);
                float qa = antq_quantize_d(cb, s_thr, s_lev, nlev, fast, da, false).q;
                float qb = antq_quantize_d(cb, s_thr, s_lev, nlev, fast, db, false).q;
                if (OVP && has2) {
                    const bool oa = fabsf(qa) > 32.0f, ob = fabsf(qb) > 32.0f;
                    if (oa) qb = __fmul_rn(qb, 0.0f);
                    else if (ob) qa = __fmul_rn(qa, 0.0f);
                }
                const float ea = __fsub_rn(antq_ste_rescale(qa, da, s), xa);
                acc[k] = __fmaf_rn(ea, ea, acc[k]);
                if (has2) {
                    const float eb = __fsub_rn(antq_ste_rescale(qb, db, s), xb);
                    acc[k] = __fmaf_rn(eb, eb, acc[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kCandTile; k++) {
            float v = acc[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5][k] = v;
        }
        __syncthreads();
        if (threadIdx.x < kCandTile && (ct + threadIdx.x) < n_cand) {
            double t = 0.0;
            for (int wi = 0; wi < kSweepThreads / 32; wi++) t += (double)s_red[wi][threadIdx.x];
            atomicAdd(err + (long long)(ct + threadIdx.x) * rows + row, t);
        }
        __syncthreads();
    }
}

template <typename T>
int launch_flat_t(const void *x, void *out, int16_t *codes, const float *alpha, int alpha_per_row, long long n,
                  long long cols, const AntqCodebook *cb, bool scale, bool ovp, cudaStream_t st) {
    if (n == 0) return 0;
    const long long groups = (n + kGroup - 1) / kGroup;
    const long long ctas = (groups + kFlatThreads - 1) / kFlatThreads;
    if (ctas > 0x7fffffffLL) return ANTQ_ENOTSUP;
    const int vec_ok = ((uintptr_t)x % 16 == 0) && ((uintptr_t)out % 16 == 0);
    dim3 grid((unsigned)ctas), block(kFlatThreads);
    const T *xi = (const T *)x;
    T *oo = (T *)out;
    if (!scale) antq_flat_kernel<T, false, false><<<grid, block, 0, st>>>(xi, oo, codes, alpha, 0, n, cols, cb, vec_ok);
    else if (ovp) antq_flat_kernel<T, true, true><<<grid, block, 0, st>>>(xi, oo, codes, alpha, alpha_per_row, n, cols, cb, vec_ok);
    else antq_flat_kernel<T, true, false><<<grid, block, 0, st>>>(xi, oo, codes, alpha, alpha_per_row, n, cols, cb, vec_ok);
    return (int)cudaGetLastError();
}

}  // namespace

int antq_launch_flat(const void *x, void *out, int16_t *codes, const float *alpha, int alpha_per_row, long long rows,
                     long long cols, int dtype, const AntqCodebook *cb, bool scale, bool ovp, cudaStream_t st) {
    const long long n = rows * cols;
    switch (dtype) {
        case ANTQ_F32: return launch_flat_t<float>(x, out, codes, alpha, alpha_per_row, n, cols, cb, scale, ovp, st);
        case ANTQ_F16: return launch_flat_t<__half>(x, out, codes, alpha, alpha_per_row, n, cols, cb, scale, ovp, st);
        case ANTQ_BF16:
            return launch_flat_t<__nv_bfloat16>(x, out, codes, alpha, alpha_per_row, n, cols, cb, scale, ovp, st);
    }
    return ANTQ_EINVAL;
}

int antq_launch_absmax(const void *x, float *out, long long rows, long long cols, int dtype, cudaStream_t st) {
    if (rows <= 0) return 0;
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * rows, st);
    if (e != cudaSuccess) return (int)e;
    if (cols == 0) return 0;
    long long splits = 1;
    const long long sms = antq_num_sms();
    if (rows < sms * 8) {
        splits = (sms * 8 + rows - 1) / rows;
        const long long max_splits = (cols + 2047) / 2048;
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
    }
    const long long ctas = rows * splits;
    if (ctas > 0x7fffffffLL) return ANTQ_ENOTSUP;
    dim3 grid((unsigned)ctas), block(256);
    unsigned int *o = reinterpret_cast<unsigned int *>(out);
    switch (dtype) {
        case ANTQ_F32: antq_absmax_kernel<float><<<grid, block, 0, st>>>((const float *)x, o, cols, (int)splits); break;
        case ANTQ_F16: antq_absmax_kernel<__half><<<grid, block, 0, st>>>((const __half *)x, o, cols, (int)splits); break;
        case ANTQ_BF16:
            antq_absmax_kernel<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16 *)x, o, cols, (int)splits);
            break;
        default: return ANTQ_EINVAL;
    }
    return (int)cudaGetLastError();
}

int antq_launch_mse_sweep(const void *x, const float *base_alpha, int alpha_per_row, const float *ratios, int n_cand,
                          double *err, long long rows, long long cols, int dtype, const AntqCodebook *cb, bool ovp,
                          cudaStream_t st) {
    if (rows <= 0 || n_cand <= 0) return 0;
    cudaError_t e = cudaMemsetAsync(err, 0, sizeof(double) * rows * n_cand, st);
    if (e != cudaSuccess) return (int)e;
    if (cols == 0) return 0;
    long long chunks = 1;
    const long long sms = antq_num_sms();
    if (rows < sms * 4) {
        chunks = (sms * 4 + rows - 1) / rows;
        const long long max_chunks = (cols + 4095) / 4096;
        if (chunks > max_chunks) chunks = max_chunks;
        if (chunks < 1) chunks = 1;
    }
    const long long ctas = rows * chunks;
    if (ctas > 0x7fffffffLL) return ANTQ_ENOTSUP;
    dim3 grid((unsigned)ctas), block(kSweepThreads);
#define ANTQ_SWEEP(T)                                                                                              \
    do {                                                                                                           \
        if (ovp) antq_mse_sweep_kernel<T, true><<<grid, block, 0, st>>>((const T *)x, base_alpha, alpha_per_row,   \
                                                                         ratios, n_cand, err, rows, cols,          \
                                                                         (int)chunks, cb);                         \
        else antq_mse_sweep_kernel<T, false><<<grid, block, 0, st>>>((const T *)x, base_alpha, alpha_per_row,      \
                                                                      ratios, n_cand, err, rows, cols, (int)chunks, \
                                                                      cb);                                         \
    } while (0)
    switch (dtype) {
        case ANTQ_F32: ANTQ_SWEEP(float); break;
        case ANTQ_F16: ANTQ_SWEEP(__half); break;
        case ANTQ_BF16: ANTQ_SWEEP(__nv_bfloat16); break;
        default: return ANTQ_EINVAL;
    }
#undef ANTQ_SWEEP
    return (int)cudaGetLastError();
}
