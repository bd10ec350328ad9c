// antq_debug.cu -- streaming micro-benchmarks used to shape the hot kernel (tools/stream_probe.py).
// Not part of the product API (not declared in include/antq.h); kept because profiles/ cites it.
#include "antq_common.cuh"

namespace {

// mode 0: flat copy, thread t of CTA b moves vectors b*T*U + j*T + t   (j < U)  -- like an elementwise kernel
// mode 1: same addresses, read only (result folded into one conditional store)
template <int U>
__global__ void __launch_bounds__(1024) probe_flat(const uint4 *__restrict__ x, uint4 *__restrict__ out, long long nvec,
                                                   int read_only) {
    const long long base = (long long)blockIdx.x * blockDim.x * U + threadIdx.x;
    uint4 r[U];
#pragma unroll
    for (int j = 0; j < U; j++) {
        const long long v = base + (long long)j * blockDim.x;
        if (v < nvec) r[j] = antq_ldg_stream(x + v);
    }
    unsigned acc = 0;
#pragma unroll
    for (int j = 0; j < U; j++) {
        const long long v = base + (long long)j * blockDim.x;
        if (v < nvec) {
            if (read_only) acc ^= r[j].x ^ r[j].y ^ r[j].z ^ r[j].w;
            else antq_stg_stream(out + v, r[j]);
        }
    }
    if (read_only && acc == 0x9e3779b9u) out[base] = r[0];
}

// mode 2: one warp owns one contiguous span of `span_vecs` vectors (a "row") and walks it U vectors per lane at a time
template <int U>
__global__ void __launch_bounds__(1024) probe_rows(const uint4 *__restrict__ x, uint4 *__restrict__ out, long long nvec,
                                                   int span_vecs, int read_only) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long v0 = warp * span_vecs;
    if (v0 >= nvec) return;
    unsigned acc = 0;
    for (int i = 0; i < span_vecs; i += 32 * U) {
        uint4 r[U];
#pragma unroll
        for (int j = 0; j < U; j++) {
            const int o = i + j * 32 + lane;
            if (o < span_vecs && v0 + o < nvec) r[j] = antq_ldg_stream(x + v0 + o);
        }
#pragma unroll
        for (int j = 0; j < U; j++) {
            const int o = i + j * 32 + lane;
            if (o < span_vecs && v0 + o < nvec) {
                if (read_only) acc ^= r[j].x ^ r[j].y ^ r[j].z ^ r[j].w;
                else antq_stg_stream(out + v0 + o, r[j]);
            }
        }
    }
    if (read_only && acc == 0x9e3779b9u) out[v0] = make_uint4(acc, 0, 0, 0);
}

// mode 3: persistent grid-stride flat copy (CTA c handles chunks c, c+G, c+2G ... of T*U vectors)
template <int U>
__global__ void __launch_bounds__(1024) probe_persist(const uint4 *__restrict__ x, uint4 *__restrict__ out,
                                                      long long nvec, int read_only) {
    const long long chunk = (long long)blockDim.x * U;
    unsigned acc = 0;
    for (long long c0 = (long long)blockIdx.x * chunk; c0 < nvec; c0 += (long long)gridDim.x * chunk) {
        uint4 r[U];
#pragma unroll
        for (int j = 0; j < U; j++) {
            const long long v = c0 + (long long)j * blockDim.x + threadIdx.x;
            if (v < nvec) r[j] = antq_ldg_stream(x + v);
        }
#pragma unroll
        for (int j = 0; j < U; j++) {
            const long long v = c0 + (long long)j * blockDim.x + threadIdx.x;
            if (v < nvec) {
                if (read_only) acc ^= r[j].x ^ r[j].y ^ r[j].z ^ r[j].w;
                else antq_stg_stream(out + v, r[j]);
            }
        }
    }
    if (read_only && acc == 0x9e3779b9u) out[blockIdx.x] = make_uint4(acc, 0, 0, 0);
}

// mode 4: probe_rows with (unused) dynamic shared memory -> shows the effect of the smem carve-out alone
// mode 5: one warp per span: TMA bulk copy of the span into smem, mbarrier wait, LDS.128 read-back, optional store
template <int U>
__global__ void __launch_bounds__(1024) probe_tma(const uint4 *__restrict__ x, uint4 *__restrict__ out, long long nvec,
                                                  int span_vecs, int read_only, int use_tma) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long v0 = warp * span_vecs;
    if (v0 >= nvec) return;
    const int warp_bytes = span_vecs * 16 + 128;
    uint4 *sbuf = reinterpret_cast<uint4 *>(smem + (size_t)wib * warp_bytes);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + (size_t)wib * warp_bytes + span_vecs * 16);
    const int n = (int)((nvec - v0) < span_vecs ? (nvec - v0) : span_vecs);
    const uint4 *src = x + v0;
    if (use_tma) {
        if (lane == 0) {
            antq_mbar_init(mbar, 1);
            antq_fence_proxy_async();
            antq_bulk_g2s(sbuf, src, (unsigned)n * 16u, mbar);
        }
        __syncwarp();
        antq_mbar_wait(mbar, 0);
        src = sbuf;
    }
    unsigned acc = 0;
    for (int i = 0; i < n; i += 32 * U) {
        uint4 r[U];
#pragma unroll
        for (int j = 0; j < U; j++) {
            const int o = i + j * 32 + lane;
            if (o < n) r[j] = use_tma ? src[o] : antq_ldg_stream(src + o);
        }
#pragma unroll
        for (int j = 0; j < U; j++) {
            const int o = i + j * 32 + lane;
            if (o < n) {
                if (read_only) acc ^= r[j].x ^ r[j].y ^ r[j].z ^ r[j].w;
                else antq_stg_stream(out + v0 + o, r[j]);
            }
        }
    }
    if (read_only && acc == 0x9e3779b9u) out[v0] = make_uint4(acc, 0, 0, 0);
}

}  // namespace

extern "C" int antq_debug_stream(const void *x, void *out, long long nbytes, int mode, int threads, int unroll,
                                 int span_vecs, int grid_cap, int read_only, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const long long nvec = nbytes / 16;
    const uint4 *xi = (const uint4 *)x;
    uint4 *oo = (uint4 *)out;
    if (threads < 32 || threads > 1024 || (threads & 31)) return ANTQ_EINVAL;
#define ANTQ_U(K, ...)                                   \
    switch (unroll) {                                    \
        case 1: K<1> __VA_ARGS__; break;                 \
        case 2: K<2> __VA_ARGS__; break;                 \
        case 4: K<4> __VA_ARGS__; break;                 \
        case 8: K<8> __VA_ARGS__; break;                 \
        default: return ANTQ_EINVAL;                     \
    }
    if (mode == 0) {
        const long long per = (long long)threads * unroll;
        const unsigned grid = (unsigned)((nvec + per - 1) / per);
        ANTQ_U(probe_flat, <<<grid, threads, 0, st>>>(xi, oo, nvec, read_only))
    } else if (mode == 2) {
        const long long warps = (nvec + span_vecs - 1) / span_vecs;
        const int wpc = threads / 32;
        const unsigned grid = (unsigned)((warps + wpc - 1) / wpc);
        ANTQ_U(probe_rows, <<<grid, threads, 0, st>>>(xi, oo, nvec, span_vecs, read_only))
    } else if (mode == 3) {
        ANTQ_U(probe_persist, <<<(unsigned)grid_cap, threads, 0, st>>>(xi, oo, nvec, read_only))
    } else if (mode == 4 || mode == 5) {
        const long long warps = (nvec + span_vecs - 1) / span_vecs;
        const int wpc = threads / 32;
        const unsigned grid = (unsigned)((warps + wpc - 1) / wpc);
        const int smem = wpc * (span_vecs * 16 + 128);
        static int set4 = 0;
        if (smem > 48 * 1024 || !set4) {
            cudaFuncSetAttribute(probe_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            cudaFuncSetAttribute(probe_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            cudaFuncSetAttribute(probe_tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            cudaFuncSetAttribute(probe_tma<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            set4 = 1;
        }
        ANTQ_U(probe_tma, <<<grid, threads, smem, st>>>(xi, oo, nvec, span_vecs, read_only, mode == 5))
    } else {
        return ANTQ_EINVAL;
    }
#undef ANTQ_U
    return (int)cudaGetLastError();
}
