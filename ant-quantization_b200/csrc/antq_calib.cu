// antq_calib.cu -- fused calibration: type selection + alpha search in ONE read of the tensor.
//
// Reference (A/antquant/quant_modules.py:287-326 search_mse, :328-415 search_adaptive_numeric_type; OliVe
// O/antquant/quant_modules.py:190-256): for every candidate type, for every candidate alpha = base * (i * 0.01), run the
// whole fake-quant forward and a per-channel mean squared error -- (#types + 1) x 75..88 full passes of ~20 elementwise
// kernels each -- keep the first strictly-best alpha per channel, sum the channels' best errors per type, take the
// type with the smallest sum.
//
// Here:   antq_calibrate  =  antq_calib_score_kernel (x read once: every (codebook, candidate) pair scored from
//                            registers)  +  antq_calib_finish_kernel (per-row first-best candidate, alpha)  +
//                            antq_calib_total_kernel (sum of the rows' best errors per codebook).
// Piecewise-uniform codebooks are scored with the closed form of antq_pu.cu WITHOUT its exact redo: an element within
// a few ulps of a midpoint is equally far from both levels, so which one is taken changes its squared error by a
// relative 1e-6 -- far below the fp32 noise of the reference's own reductions (parity for calibration is
// MSE-equivalence).  Other codebooks (apot, OliVe normal + outliers with the pair mask) are scored
// with the literal arithmetic, one codebook at a time.  All sums: fp32 per lane over 32 elements, fp64 from there on,
// fixed order, no atomics -> bit-reproducible run to run.
#include "antq_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kChunk = kWarps * 1024;      // elements per work item: 32 per lane
constexpr int kMaxCb = 8;
constexpr int kMaxCand = 256;

struct CalibParams {
    const void *x;
    const float *base, *ratios;
    const AntqCodebook *cb[kMaxCb];
    int exact[kMaxCb];                     // 1: literal arithmetic (not piecewise uniform, or OVP)
    int ovp[kMaxCb];
    int n_cb, n_cand;
    long long rows, cols;                  // per-tensor: rows = 1
    int chunks_per_row, items_per_cta_loop, alpha_per_row;
    long long n_items;
    double *partial;                       // [n_partials][n_cb][n_cand]
    float *alpha_out, *mse_out;
    int *best_out;
    float *row_best;                       // [n_cb][rows] scratch: best mean error per row
};

template <typename T> __device__ __forceinline__ void load32(const T *__restrict__ xr, long long c0, long long c1, int lane,
                                                            float (&f)[32]) {
    typedef AntqType<T> A;
    // element e of lane l sits at c0 + (e / 8) * 256 + l * 8 + (e % 8): 16-byte coalesced for 16-bit types
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const long long b = c0 + j * 256 + lane * 8;
        if (b + 8 <= c1 && ((uintptr_t)(xr + b) % 16 == 0) && sizeof(T) == 2) {
            const uint4 v = antq_ldg_stream(reinterpret_cast<const uint4 *>(xr + b));
            const T *h = reinterpret_cast<const T *>(&v);
#pragma unroll
            for (int e = 0; e < 8; e++) f[j * 8 + e] = A::to_f32(h[e]);
        } else {
#pragma unroll
            for (int e = 0; e < 8; e++) f[j * 8 + e] = (b + e) < c1 ? A::to_f32(xr[b + e]) : 0.0f;   // padding: x = 0 -> error 0
        }
    }
}

// Squared error of 32 elements under one (codebook, alpha) pair: closed form.
template <bool UNIFORM>
__device__ __forceinline__ float score_pu(const float (&f)[32], float s, float kx, float c, float kmin, float kmax,
                                          const float *magic) {
    float acc = 0.0f;
#pragma unroll
    for (int e = 0; e < 32; e++) {
        const float t = __fmul_rn(f[e], kx);
        const float M = UNIFORM ? 12582912.0f : magic[__float_as_uint(t) >> 23];
        const float mf = __fsub_rn(__fadd_rn(t, M), M);
        const float q = __fmul_rn(fminf(fmaxf(mf, kmin), kmax), c);
        const float err = __fsub_rn(__fmul_rn(q, s), f[e]);
        acc = __fmaf_rn(err, err, acc);
    }
    return acc;
}

// The same with the reference arithmetic, literally (any codebook, OVP pairs = neighbouring elements of a lane).
template <bool OVP>
__device__ __noinline__ float score_exact(const float (&f)[32], float s, const AntqCodebook *__restrict__ cb, float win) {
    float acc = 0.0f;
    const int nlev = cb->n_levels;
#pragma unroll 1
    for (int e = 0; e < 32; e += 2) {
        float d[2], q[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            d[k] = __fdiv_rn(f[e + k], s);
            if (fabsf(d[k]) <= win) q[k] = cb->level[antq_rank(cb->thr, nlev - 1, d[k])];
            else { int code; q[k] = antq_scan_literal(cb->grid, cb->n_entries, d[k], code); }
        }
        if (OVP) {
            if (fabsf(q[0]) > 32.0f) q[1] = __fmul_rn(q[1], 0.0f);
            else if (fabsf(q[1]) > 32.0f) q[0] = __fmul_rn(q[0], 0.0f);
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const float err = __fsub_rn(antq_ste_rescale(q[k], d[k], s), f[e + k]);
            acc = __fmaf_rn(err, err, acc);
        }
    }
    return acc;
}

template <typename T> __global__ void __launch_bounds__(kThreads) antq_calib_score_kernel(const CalibParams p) {
    extern __shared__ __align__(16) unsigned char cal_smem[];
    double *wacc = reinterpret_cast<double *>(cal_smem);                          // [kWarps][n_cb * n_cand]
    float2 *sk = reinterpret_cast<float2 *>(wacc + (size_t)kWarps * p.n_cb * p.n_cand);    // [n_cb * n_cand] {s, kx}
    float *magic = reinterpret_cast<float *>(sk + (size_t)p.n_cb * p.n_cand);     // [n_cb][512]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int npair = p.n_cb * p.n_cand;
    for (int i = threadIdx.x; i < kWarps * npair; i += kThreads) wacc[i] = 0.0;
    for (int k = 0; k < p.n_cb; k++)
        if (!p.exact[k])
            for (int i = threadIdx.x; i < 512; i += kThreads) magic[k * 512 + i] = p.cb[k]->pu_tab[i & 255].x;
    long long cur_row = -1;
    for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const long long row = item / p.chunks_per_row;
        const long long c0 = (item - row * p.chunks_per_row) * (long long)kChunk;
        const long long c1 = (c0 + kChunk) < p.cols ? (c0 + kChunk) : p.cols;
        if (row != cur_row) {                                                     // (per-tensor: once; per-row: one item per CTA)
            __syncthreads();
            const float base = p.base[p.alpha_per_row ? row : 0];
            for (int i = threadIdx.x; i < npair; i += kThreads) {
                const int k = i / p.n_cand, c = i - k * p.n_cand;
                const float s = __fdiv_rn(__fmul_rn(base, p.ratios[c]), p.cb[k]->gmax);       // alpha = base * (i * 0.01)
                sk[i] = make_float2(s, __fmul_rn(__fdiv_rn(1.0f, s), p.cb[k]->pu_inv_c));
            }
            __syncthreads();
            cur_row = row;
        }
        const T *xr = reinterpret_cast<const T *>(p.x) + row * p.cols;
        const long long w0 = c0 + (long long)warp * 1024;
        if (w0 >= c1) continue;
        float f[32];
        load32<T>(xr, w0, (w0 + 1024) < c1 ? (w0 + 1024) : c1, lane, f);
        double *my = wacc + (size_t)warp * npair;
        for (int k = 0; k < p.n_cb; k++) {
            const AntqCodebook *__restrict__ cb = p.cb[k];
            const float c = cb->pu_c, kmin = cb->pu_kmin, kmax = cb->pu_kmax;
            const bool uni = (cb->flags & ANTQ_CB_PU_UNIFORM) != 0;
            const float win = (cb->flags & ANTQ_CB_WELLSEP) ? cb->lim_idx : -1.0f;
#pragma unroll 1
            for (int cd = 0; cd < p.n_cand; cd++) {
                const float2 v = sk[k * p.n_cand + cd];
                float a;
                if (p.exact[k]) a = p.ovp[k] ? score_exact<true>(f, v.x, cb, win) : score_exact<false>(f, v.x, cb, win);
                else if (uni) a = score_pu<true>(f, v.x, v.y, c, kmin, kmax, nullptr);
                else a = score_pu<false>(f, v.x, v.y, c, kmin, kmax, magic + k * 512);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == 0) my[k * p.n_cand + cd] += (double)a;
            }
        }
    }
    __syncthreads();
    // one partial per CTA (per-tensor) or per item (per-row: gridDim == n_items)
    double *dst = p.partial + (size_t)blockIdx.x * npair;
    for (int i = threadIdx.x; i < npair; i += kThreads) {
        double t = 0.0;
        for (int w = 0; w < kWarps; w++) t += wacc[(size_t)w * npair + i];
        dst[i] = t;
    }
}

// Per row and codebook: sum the row's partials in index order, first strictly-best candidate (the reference's
// `score < best_score` update starting from 1e10, A/...:297,322-324), alpha = base * ratio[best].
__global__ void antq_calib_finish_kernel(const CalibParams p, int partials_per_row) {
    const long long row = blockIdx.x;
    const int npair = p.n_cb * p.n_cand;
    __shared__ float s_err[kMaxCb * kMaxCand];
    for (int i = threadIdx.x; i < npair; i += blockDim.x) {
        double t = 0.0;
        for (int q = 0; q < partials_per_row; q++) t += p.partial[((size_t)row * partials_per_row + q) * npair + i];
        s_err[i] = (float)(t / (double)p.cols);                                   // mean over the row
    }
    __syncthreads();
    if (threadIdx.x < p.n_cb) {
        const int k = threadIdx.x;
        float best = 1e10f;
        int bi = -1;
        for (int c = 0; c < p.n_cand; c++) {
            const float v = s_err[k * p.n_cand + c];
            if (v < best) { best = v; bi = c; }
        }
        const float base = p.base[p.alpha_per_row ? row : 0];
        p.alpha_out[(size_t)k * p.rows + row] = bi >= 0 ? __fmul_rn(base, p.ratios[bi]) : base;
        p.row_best[(size_t)k * p.rows + row] = best;
        if (p.best_out) p.best_out[(size_t)k * p.rows + row] = bi;
    }
}

// mse_out[k] = sum over rows of the best mean error (fixed-order tree in fp64, rounded to fp32 like the reference's sum)
__global__ void antq_calib_total_kernel(const CalibParams p) {
    const int k = blockIdx.x;
    __shared__ double sm[256];
    double t = 0.0;
    for (long long r = threadIdx.x; r < p.rows; r += 256) t += (double)p.row_best[(size_t)k * p.rows + r];
    sm[threadIdx.x] = t;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) p.mse_out[k] = (float)sm[0];
}

struct Layout { long long n_items, n_partials; int chunks_per_row, partials_per_row, ctas; };

Layout layout_of(long long rows, long long cols, int per_row) {
    Layout l;
    const long long r = per_row ? rows : 1, c = per_row ? cols : rows * cols;
    l.chunks_per_row = (int)((c + kChunk - 1) / kChunk);
    l.n_items = r * l.chunks_per_row;
    if (per_row) { l.ctas = (int)l.n_items; l.partials_per_row = l.chunks_per_row; }
    else {
        const long long cap = (long long)antq_num_sms() * 2;
        l.ctas = (int)(l.n_items < cap ? l.n_items : cap);
        l.partials_per_row = l.ctas;
    }
    l.n_partials = (long long)l.ctas;
    return l;
}

}  // namespace

extern "C" {

size_t antq_calibrate_workspace_bytes(int64_t rows, int64_t cols, int alpha_per_row, int n_cand, int n_cb) {
    if (rows <= 0 || cols <= 0 || n_cand <= 0 || n_cb <= 0) return 0;
    const int per_row = alpha_per_row && rows > 1;
    const Layout l = layout_of(rows, cols, per_row);
    const long long r = per_row ? rows : 1;
    return (size_t)l.n_partials * n_cb * n_cand * sizeof(double) + (size_t)n_cb * r * sizeof(float) + 256;
}

int antq_calibrate(const void *x, int64_t rows, int64_t cols, int dtype, int alpha_per_row, const float *base_alpha,
                   const float *ratios, int n_cand, const void *const *codebooks, const antq_codebook_info *const *infos,
                   const int *flags_per_codebook, int n_cb, float *alpha_out, float *mse_out, int *best_index_out,
                   void *workspace, size_t workspace_bytes, void *stream) {
    if (rows < 0 || cols < 0 || n_cand < 1 || n_cand > kMaxCand || n_cb < 1 || n_cb > kMaxCb) return ANTQ_EINVAL;
    if (dtype != ANTQ_F32 && dtype != ANTQ_F16 && dtype != ANTQ_BF16) return ANTQ_EINVAL;
    if (!base_alpha || !ratios || !codebooks || !infos || !alpha_out || !mse_out) return ANTQ_EINVAL;
    if (rows == 0 || cols == 0) return 0;
    if (!x || !workspace) return ANTQ_EINVAL;
    const int per_row = alpha_per_row && rows > 1;
    if (workspace_bytes < antq_calibrate_workspace_bytes(rows, cols, alpha_per_row, n_cand, n_cb)) return ANTQ_EINVAL;
    if (n_cb * n_cand > kMaxCb * kMaxCand) return ANTQ_EINVAL;
    const Layout l = layout_of(rows, cols, per_row);
    if (l.n_items > 0x7fffffffLL) return ANTQ_ENOTSUP;
    CalibParams p = {};
    p.x = x; p.base = base_alpha; p.ratios = ratios;
    p.n_cb = n_cb; p.n_cand = n_cand;
    p.rows = per_row ? rows : 1;
    p.cols = per_row ? cols : rows * cols;
    p.alpha_per_row = per_row;
    p.chunks_per_row = l.chunks_per_row;
    p.n_items = l.n_items;
    for (int k = 0; k < n_cb; k++) {
        if (!codebooks[k] || !infos[k]) return ANTQ_EINVAL;
        p.cb[k] = (const AntqCodebook *)codebooks[k];
        const int fl = flags_per_codebook ? flags_per_codebook[k] : 0;
        p.ovp[k] = (fl & ANTQ_FLAG_OVP) != 0 && infos[k]->n_entries > infos[k]->n_normal;
        p.exact[k] = !(infos[k]->flags & ANTQ_CB_PU) || p.ovp[k] || (fl & ANTQ_FLAG_NO_PU);
        if (p.ovp[k] && (p.cols & 1)) return ANTQ_ENOTSUP;                 // pairs would straddle rows / wrap around
    }
    p.partial = (double *)workspace;
    p.row_best = (float *)((char *)workspace + (size_t)l.n_partials * n_cb * n_cand * sizeof(double));
    p.alpha_out = alpha_out; p.mse_out = mse_out; p.best_out = best_index_out;
    const size_t smem = (size_t)kWarps * n_cb * n_cand * sizeof(double) + (size_t)n_cb * n_cand * sizeof(float2) +
                        (size_t)n_cb * 512 * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
#define ANTQ_CAL(T)                                                                                              \
    do {                                                                                                         \
        if (smem > 48 * 1024)                                                                                    \
            e = cudaFuncSetAttribute(antq_calib_score_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e == cudaSuccess) antq_calib_score_kernel<T><<<l.ctas, kThreads, smem, st>>>(p);                     \
    } while (0)
    switch (dtype) {
        case ANTQ_F32: ANTQ_CAL(float); break;
        case ANTQ_F16: ANTQ_CAL(__half); break;
        case ANTQ_BF16: ANTQ_CAL(__nv_bfloat16); break;
    }
#undef ANTQ_CAL
    if (e != cudaSuccess) return (int)e;
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    antq_calib_finish_kernel<<<(unsigned)p.rows, 128, 0, st>>>(p, l.partials_per_row);
    antq_calib_total_kernel<<<n_cb, 256, 0, st>>>(p);
    return (int)cudaGetLastError();
}

}  // extern "C"
