// antq_bwd.cu -- backward of the ANT fake-quant forward for QAT (A/antquant/quant_modules.py:535-551 under autograd,
// driven by A/ImageNet/main.py:190-198): ONE pass over (grad_out, x, out) instead of 6-8 eager elementwise passes.
//
// Reference graph:  s = alpha / max(grid);  d = x / s;  t = (q - d).detach() + d;  out = t * s.  Autograd gives
//   grad_x     = (g * s) / s                      a mul then a div, both rounded: reproduced bit for bit (fp32)
//   grad_alpha = sum_row g * (q - d) / max(grid)  (d out/d s = t - d = q - d; the sum is per output channel, or over
//                                                  the whole tensor for a per-tensor scale)
// q - d is recovered from the saved forward result as (out - x) / s, exactly as the round-1 torch expression did, so
// the two agree to fp32 rounding (tests: grad_x bit-exact, grad_alpha <= 1e-5 relative against reference autograd).
//
// Reduction: fixed order, no atomics -> deterministic.  Each CTA owns one (row, chunk) piece, accumulates fp32 per
// thread over <= 32 elements at a time into fp64, reduces over the CTA in a fixed tree and writes one fp64 partial;
// antq_bwd_finish_kernel adds the partials of a row in index order.
#include "antq_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 8192;          // elements per CTA

struct BwdParams {
    const void *g, *x, *out;
    void *gx;
    const float *alpha;
    double *partial;                  // [rows][chunks_per_row]
    float *galpha;                    // [rows] or [1]
    long long rows, cols;             // per-tensor: rows = 1
    int chunks_per_row, alpha_per_row;
    float gmax;
};

template <typename T> __global__ void __launch_bounds__(kThreads) antq_bwd_kernel(const BwdParams p) {
    typedef AntqType<T> A;
    const long long row = blockIdx.x / p.chunks_per_row;
    const int chunk = blockIdx.x % p.chunks_per_row;
    const long long c0 = (long long)chunk * kChunk;
    const long long c1 = (c0 + kChunk) < p.cols ? (c0 + kChunk) : p.cols;
    const T *g = reinterpret_cast<const T *>(p.g) + row * p.cols;
    const T *x = reinterpret_cast<const T *>(p.x) + row * p.cols;
    const T *o = reinterpret_cast<const T *>(p.out) + row * p.cols;
    T *gx = p.gx ? reinterpret_cast<T *>(p.gx) + row * p.cols : nullptr;
    const float s = __fdiv_rn(p.alpha[p.alpha_per_row ? row : 0], p.gmax);
    double acc = 0.0;
    for (long long c = c0 + threadIdx.x; c < c1; c += kThreads * 8) {
        float part = 0.0f;
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const long long i = c + (long long)u * kThreads;
            if (i < c1) {
                const float gf = A::to_f32(g[i]);
                if (gx) gx[i] = A::from_f32_rn(__fdiv_rn(__fmul_rn(gf, s), s));
                if (p.partial) {
                    const float qd = __fdiv_rn(__fsub_rn(A::to_f32(o[i]), A::to_f32(x[i])), s);      // q - d
                    part = __fadd_rn(part, __fmul_rn(gf, qd));
                }
            }
        }
        acc += (double)part;
    }
    if (!p.partial) return;
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o2);
    __shared__ double sm[kThreads / 32];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kThreads / 32; w++) t += sm[w];
        p.partial[row * p.chunks_per_row + chunk] = t;
    }
}

__global__ void antq_bwd_finish_kernel(const double *__restrict__ partial, float *__restrict__ galpha, long long rows,
                                       int chunks_per_row, float gmax) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    double t = 0.0;
    for (int c = 0; c < chunks_per_row; c++) t += partial[row * chunks_per_row + c];
    galpha[row] = (float)(t / (double)gmax);
}

}  // namespace

extern "C" {

size_t antq_backward_workspace_bytes(int64_t rows, int64_t cols, int alpha_per_row) {
    if (rows <= 0 || cols <= 0) return 0;
    const long long r = alpha_per_row ? rows : 1, c = alpha_per_row ? cols : rows * cols;
    return (size_t)(r * ((c + kChunk - 1) / kChunk)) * sizeof(double);
}

int antq_fakequant_backward(const void *grad_out, const void *x, const void *out, const float *alpha, int alpha_per_row,
                            int64_t rows, int64_t cols, int dtype, float gmax, void *grad_x, float *grad_alpha,
                            void *workspace, size_t workspace_bytes, void *stream) {
    if (rows < 0 || cols < 0 || (dtype != ANTQ_F32 && dtype != ANTQ_F16 && dtype != ANTQ_BF16)) return ANTQ_EINVAL;
    if (rows == 0 || cols == 0) return 0;
    if (!grad_out || !alpha || (!grad_x && !grad_alpha) || (grad_alpha && (!x || !out))) return ANTQ_EINVAL;
    BwdParams p = {};
    p.g = grad_out; p.x = x; p.out = out; p.gx = grad_x; p.alpha = alpha;
    p.alpha_per_row = alpha_per_row && rows > 1;
    p.rows = p.alpha_per_row ? rows : 1;
    p.cols = p.alpha_per_row ? cols : rows * cols;
    p.gmax = gmax;
    const long long cpr = (p.cols + kChunk - 1) / kChunk;
    if (p.rows * cpr > 0x7fffffffLL) return ANTQ_ENOTSUP;
    p.chunks_per_row = (int)cpr;
    if (grad_alpha) {
        if (!workspace || workspace_bytes < antq_backward_workspace_bytes(rows, cols, alpha_per_row)) return ANTQ_EINVAL;
        p.partial = (double *)workspace;
        p.galpha = grad_alpha;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned ctas = (unsigned)(p.rows * cpr);
    switch (dtype) {
        case ANTQ_F32: antq_bwd_kernel<float><<<ctas, kThreads, 0, st>>>(p); break;
        case ANTQ_F16: antq_bwd_kernel<__half><<<ctas, kThreads, 0, st>>>(p); break;
        case ANTQ_BF16: antq_bwd_kernel<__nv_bfloat16><<<ctas, kThreads, 0, st>>>(p); break;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    if (grad_alpha)
        antq_bwd_finish_kernel<<<(unsigned)((p.rows + 127) / 128), 128, 0, st>>>(p.partial, grad_alpha, p.rows, p.chunks_per_row, gmax);
    return (int)cudaGetLastError();
}

}  // extern "C"
