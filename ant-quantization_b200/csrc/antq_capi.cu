// antq_capi.cu -- the extern "C" surface declared in include/antq.h: argument
// checks, kernel selection, and the host-buffer pipeline.  No torch, no C++
// types across the boundary, no allocation/synchronisation in device entry points.
#include <new>
#include <stdlib.h>
#include <string.h>

#include "antq_common.cuh"

int antq_launch_stream(const void *x, void *out, const float *alpha, int alpha_per_row, long long rows, long long cols,
                       int dtype, const AntqCodebook *cb, const antq_codebook_info *info, bool ovp, cudaStream_t st);
int antq_launch_short(const void *x, void *out, const float *alpha, int alpha_per_row, long long rows, long long cols,
                      int dtype, const AntqCodebook *cb, const antq_codebook_info *info, bool ovp, cudaStream_t st);
int antq_short_thresholds(const antq_codebook_info *info, bool ovp);
int antq_launch_pu_stream(const void *x, void *out, const float *alpha, int alpha_per_row, long long rows, long long cols,
                          int dtype, const AntqCodebook *cb, const antq_codebook_info *info, bool ovp, cudaStream_t st);
int antq_launch_pu_short(const void *x, void *out, const float *alpha, int alpha_per_row, long long rows, long long cols,
                         int dtype, const AntqCodebook *cb, const antq_codebook_info *info, cudaStream_t st);
int antq_launch_pu_dynamic(const void *x, void *out, float *alpha_out, float ratio, long long rows, long long cols, int dtype,
                           const AntqCodebook *cb, const antq_codebook_info *info, cudaStream_t st);
int antq_launch_flat(const void *x, void *out, int16_t *codes, const float *alpha, int alpha_per_row, long long rows,
                     long long cols, int dtype, const AntqCodebook *cb, bool scale, bool ovp, cudaStream_t st);
int antq_launch_absmax(const void *x, float *out, long long rows, long long cols, int dtype, cudaStream_t st);
int antq_launch_mse_sweep(const void *x, const float *base_alpha, int alpha_per_row, const float *ratios, int n_cand,
                          double *err, long long rows, long long cols, int dtype, const AntqCodebook *cb, bool ovp,
                          cudaStream_t st);

int antq_num_sms() {
    static int cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] <= 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

namespace {
inline int esize(int dtype) { return dtype == ANTQ_F32 ? 4 : (dtype == ANTQ_F16 || dtype == ANTQ_BF16) ? 2 : 0; }
constexpr long long kRowsMinCols = 512;   // below this the per-row prologue costs more than it saves
}  // namespace

extern "C" {

int antq_abi_version(void) { return ANTQ_ABI_VERSION; }

const char *antq_build_info(void) {
    return "libantq sm_100a; kernels: antq_prepare_kernel antq_stream_kernel antq_pu_stream_kernel antq_pu_short_kernel antq_pu_lean_kernel antq_pu_dynamic_kernel "
           "antq_pu_encode_kernel antq_short_kernel antq_flat_kernel antq_absmax_kernel antq_mse_sweep_kernel antq_calib_score_kernel "
           "antq_encode_p4_kernel antq_decode_p4_kernel antq_decode_p4_fast_kernel antq_bwd_kernel antq_linear_p4_kernel "
           "antq_levels_e4m3_kernel; built " __DATE__;
}

const char *antq_error_string(int status) {
    if (status == 0) return "ok";
    if (status == ANTQ_EINVAL) return "antq: invalid argument";
    if (status == ANTQ_ENOTSUP) return "antq: unsupported configuration for the requested kernel";
    if (status == ANTQ_EALIGN) return "antq: pointer not aligned to the element size";
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "antq: unknown error";
}

size_t antq_codebook_bytes(void) { return sizeof(AntqCodebook); }

int antq_codebook_prepare(const float *grid, int k_normal, const float *outliers, int k_out, void *codebook,
                          void *stream) {
    if (!grid || !codebook || k_normal < 1 || k_out < 0 || k_normal + k_out > ANTQ_MAX_GRID) return ANTQ_EINVAL;
    if (k_out > 0 && !outliers) return ANTQ_EINVAL;
    return antq_launch_prepare(grid, k_normal, outliers, k_out, (AntqCodebook *)codebook, (cudaStream_t)stream);
}

int antq_lut_nearest(const void *x, void *z, int16_t *codes, int64_t n, int dtype, const void *codebook,
                     void *stream) {
    if (n < 0 || !codebook || esize(dtype) == 0) return ANTQ_EINVAL;
    if (n == 0) return 0;
    if (!x || !z) return ANTQ_EINVAL;
    if ((uintptr_t)x % esize(dtype) || (uintptr_t)z % esize(dtype)) return ANTQ_EALIGN;
    return antq_launch_flat(x, z, codes, nullptr, 0, 1, n, dtype, (const AntqCodebook *)codebook, false, false,
                            (cudaStream_t)stream);
}

int antq_fakequant_plan(const antq_codebook_info *info, int64_t rows, int64_t cols, int dtype, int flags,
                        const void *x, const void *out, const void *codes) {
    const int es = esize(dtype);
    if (es == 0 || rows < 0 || cols < 0) return ANTQ_EINVAL;
    if (flags & ANTQ_FLAG_FORCE_FLAT) return 2;
    const bool ovp = (flags & ANTQ_FLAG_OVP) != 0;
    const int vec = 16 / es;
    const bool aligned = ((uintptr_t)x % 16 == 0) && ((uintptr_t)out % 16 == 0);
    const bool exact = info && (info->flags & ANTQ_CB_WELLSEP) && (info->flags & ANTQ_CB_STE_EXACT) && aligned && !codes;
    // closed form (antq_pu.cu): any piecewise-uniform grid, no OVP
    // ... or, with outlier-victim pairs, a codebook whose NORMAL levels are (ANTQ_CB_PU_OVP): whole vectors, pairs inside rows
    const bool pu = exact && !(flags & ANTQ_FLAG_NO_PU) &&
                    (ovp ? (info->flags & ANTQ_CB_PU_OVP) && (info->flags & ANTQ_CB_OVP_OK) && cols % vec == 0
                         : (info->flags & ANTQ_CB_PU) != 0);
    const bool long_rows = (cols >= kRowsMinCols || (flags & ANTQ_FLAG_FORCE_ROWS)) && (cols % vec == 0 || rows == 1);
    bool chain = false;
    int nt = 0;
    if (exact) {
        const bool sym = (info->flags & ANTQ_CB_SYMMETRIC) != 0;
        // signed int-k grids run as "symmetric + one extra negative level" in the stream kernel
        const bool symx = !sym && (info->flags & ANTQ_CB_SYMX) && !ovp;
        nt = (sym || symx) ? info->n_mag - 1 : info->n_levels - 1;
        chain = nt >= 1 && nt <= 31 && long_rows && (!ovp || ((info->flags & ANTQ_CB_OVP_OK) && cols % 2 == 0));
    }
    if (flags & ANTQ_FLAG_FORCE_ROWS) return chain ? 1 : ANTQ_ENOTSUP;
    if (flags & ANTQ_FLAG_FORCE_TILE) return (pu && !ovp && cols % vec == 0 && (rows == 1 || cols / vec <= 127)) ? 5 : ANTQ_ENOTSUP;
    if (flags & ANTQ_FLAG_FORCE_PU) {
        if (pu && long_rows) return 4;
        if (pu && !ovp && rows > 1 && cols % vec == 0) return 5;
        return ANTQ_ENOTSUP;
    }
    // the compare chain (packed 16-bit arithmetic, two elements per instruction) wins while it is short: <= 7 thresholds
    // after folding signs = every signed 4-bit grid and OliVe's two-phase chain (13.7-15 us per 4096^2 fp16).  Beyond
    // that the closed form does (fp32 per element but independent of the number of levels, 17-20 us): unsigned 4-bit,
    // 5 to 8 bit.  Measured side by side in profiles/r02_notes.md.
    // (OliVe's signed 4-bit codebooks keep the two-phase chain: 7 thresholds in phase one, 13.2-15.1 us)
    const bool ovp2_fast = ovp && info && (info->flags & ANTQ_CB_SYMMETRIC) && nt <= 15;
    // fp32 I/O: the chain works on one element per register there (no packed pairs) and loses to the closed form even at 7
    // thresholds (31.3 vs 25.5 us per 4096^2, flint-4 signed)
    // (OliVe pairs included: its fp32 chain runs at 57-68 us, the closed form + pair logic at 26.5)
    const bool f32_pu = dtype == ANTQ_F32 && pu && long_rows;
    // bf16 I/O, uniform grids (int-k): 14.7 / 13.7 us (per-row / per-tensor) against the chain's 18.2 / 15.8
    const bool bf16_pu = dtype == ANTQ_BF16 && pu && long_rows && (info->flags & ANTQ_CB_PU_UNIFORM);   // OliVe int: 14.8 vs 16.3
    // fp16 I/O, uniform grids, per-row scales: 13.7 us against the chain's 14.7 (its SYMX path: one extra compare per pair);
    // with one scale the chain keeps a small edge (13.3 vs 13.6)
    // (int-3 and below: 3 thresholds, the chain stays -- 13.0 vs 13.9 us)
    const bool f16_int_pu = dtype == ANTQ_F16 && pu && !ovp && long_rows && rows > 1 && nt >= 6 && (info->flags & ANTQ_CB_PU_UNIFORM);
    if (chain && (nt <= 7 || !pu || ovp2_fast) && !f32_pu && !bf16_pu && !f16_int_pu) return 1;
    if (pu && long_rows) return 4;
    if (chain) return 1;
    const bool short_rows = info && !codes && aligned && rows > 1 && cols < kRowsMinCols && cols % vec == 0;
    if (short_rows && pu && !ovp) return 5;
    // short rows / scale groups of the other grids: the d-space chain kernel (<= 15 thresholds after folding signs)
    if (short_rows && (info->flags & ANTQ_CB_WELLSEP) && antq_short_thresholds(info, ovp) >= 1 &&
        antq_short_thresholds(info, ovp) <= 15 && (!ovp || (info->flags & ANTQ_CB_OVP_OK)))
        return 3;
    return 2;
}

int antq_fakequant(const void *x, void *out, int16_t *codes, const float *alpha, int alpha_per_row, int64_t rows,
                   int64_t cols, int dtype, const void *codebook, const antq_codebook_info *info, int flags,
                   void *stream) {
    const int es = esize(dtype);
    if (es == 0 || rows < 0 || cols < 0) return ANTQ_EINVAL;
    if (rows == 0 || cols == 0) return 0;
    if (!x || !out || !codebook || !alpha) return ANTQ_EINVAL;
    if ((uintptr_t)x % es || (uintptr_t)out % es || (uintptr_t)codes % 2) return ANTQ_EALIGN;
    const bool ovp = (flags & ANTQ_FLAG_OVP) != 0;
    if (ovp && x == out && ((rows * cols) & 1)) return ANTQ_EINVAL;   // wrap-around pair reads x[0]
    const int plan = antq_fakequant_plan(info, rows, cols, dtype, flags, x, out, codes);
    if (plan < 0) return plan;
    const AntqCodebook *cb = (const AntqCodebook *)codebook;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = ANTQ_ENOTSUP;
    switch (plan) {
        case 1: rc = antq_launch_stream(x, out, alpha, alpha_per_row, rows, cols, dtype, cb, info, ovp, st); break;
        case 3: rc = antq_launch_short(x, out, alpha, alpha_per_row, rows, cols, dtype, cb, info, ovp, st); break;
        case 4: rc = antq_launch_pu_stream(x, out, alpha, alpha_per_row, rows, cols, dtype, cb, info, ovp, st); break;
        case 5: rc = antq_launch_pu_short(x, out, alpha, alpha_per_row, rows, cols, dtype, cb, info, st); break;
        default: break;
    }
    if (rc != ANTQ_ENOTSUP) return rc;
    if (flags & (ANTQ_FLAG_FORCE_ROWS | ANTQ_FLAG_FORCE_PU | ANTQ_FLAG_FORCE_TILE)) return ANTQ_ENOTSUP;
    // every shape, alignment and grid: the generic kernel (also the only one that emits int16 code indices)
    return antq_launch_flat(x, out, codes, alpha, alpha_per_row, rows, cols, dtype, cb, true, ovp, st);
}

int antq_fakequant_dynamic(const void *x, void *out, float *alpha_out, float ratio, int64_t rows, int64_t cols, int dtype,
                           const void *codebook, const antq_codebook_info *info, int flags, void *stream) {
    const int es = esize(dtype);
    if (es == 0 || rows < 0 || cols < 0 || !info) return ANTQ_EINVAL;
    if (rows == 0 || cols == 0) return 0;
    if (!x || !out || !codebook) return ANTQ_EINVAL;
    if ((uintptr_t)x % 16 || (uintptr_t)out % 16) return ANTQ_EALIGN;
    if (flags & ANTQ_FLAG_OVP) return ANTQ_ENOTSUP;
    return antq_launch_pu_dynamic(x, out, alpha_out, ratio, rows, cols, dtype, (const AntqCodebook *)codebook, info,
                                  (cudaStream_t)stream);
}

int antq_absmax(const void *x, float *out, int64_t rows, int64_t cols, int dtype, void *stream) {
    const int es = esize(dtype);
    if (es == 0 || rows < 0 || cols < 0 || !out) return ANTQ_EINVAL;
    if (rows * cols > 0 && !x) return ANTQ_EINVAL;
    if ((uintptr_t)x % es) return ANTQ_EALIGN;
    return antq_launch_absmax(x, out, rows, cols, dtype, (cudaStream_t)stream);
}

int antq_mse_sweep(const void *x, const float *base_alpha, int alpha_per_row, const float *ratios, int n_cand,
                   double *err, int64_t rows, int64_t cols, int dtype, const void *codebook, int flags,
                   void *stream) {
    const int es = esize(dtype);
    if (es == 0 || rows < 0 || cols < 0 || n_cand < 0 || !codebook || !base_alpha || !ratios || !err)
        return ANTQ_EINVAL;
    if (rows * cols > 0 && !x) return ANTQ_EINVAL;
    if ((uintptr_t)x % es) return ANTQ_EALIGN;
    const bool ovp = (flags & ANTQ_FLAG_OVP) != 0;
    if (ovp && (cols & 1) && rows > 1) return ANTQ_ENOTSUP;
    return antq_launch_mse_sweep(x, base_alpha, alpha_per_row, ratios, n_cand, err, rows, cols, dtype,
                                 (const AntqCodebook *)codebook, ovp, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------
// Host-buffer pipeline: H2D -> fused kernel -> D2H in row chunks over n_stages streams.
// ---------------------------------------------------------------------------------
struct antq_host_ctx {
    int device;
    size_t chunk_bytes;
    int n_stages;
    cudaStream_t st[8];
    void *d_in[8];
    void *d_out[8];
    AntqCodebook *cb;
    float *d_grid;          // ANTQ_MAX_GRID floats
    float *d_alpha[8];      // per stage: the alpha slice of the chunk in flight on that stage
    size_t alpha_cap[8];
    int next_stage;         // stages are used round-robin ACROSS calls, so consecutive tensors overlap
    float h_grid[ANTQ_MAX_GRID];
    int h_k_normal, h_k_out;
    antq_codebook_info info;
    int last_launches;
};

int antq_host_create(antq_host_ctx **out, int device, size_t chunk_bytes, int n_stages) {
    if (!out || n_stages < 1 || n_stages > 8 || chunk_bytes < 4096) return ANTQ_EINVAL;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    antq_host_ctx *c = new (std::nothrow) antq_host_ctx();
    if (!c) return (int)cudaErrorMemoryAllocation;
    memset(c, 0, sizeof(*c));
    c->device = device; c->chunk_bytes = chunk_bytes; c->n_stages = n_stages; c->h_k_normal = -1;
    for (int i = 0; i < n_stages && e == cudaSuccess; i++) {
        e = cudaStreamCreateWithFlags(&c->st[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMalloc(&c->d_in[i], chunk_bytes);
        if (e == cudaSuccess) e = cudaMalloc(&c->d_out[i], chunk_bytes);
    }
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->cb, sizeof(AntqCodebook));
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->d_grid, sizeof(float) * ANTQ_MAX_GRID);
    if (e != cudaSuccess) { antq_host_destroy(c); return (int)e; }
    *out = c;
    return 0;
}

void antq_host_destroy(antq_host_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (int i = 0; i < 8; i++) {
        if (c->st[i]) { cudaStreamSynchronize(c->st[i]); cudaStreamDestroy(c->st[i]); }
        if (c->d_in[i]) cudaFree(c->d_in[i]);
        if (c->d_out[i]) cudaFree(c->d_out[i]);
    }
    if (c->cb) cudaFree(c->cb);
    if (c->d_grid) cudaFree(c->d_grid);
    for (int i = 0; i < 8; i++)
        if (c->d_alpha[i]) cudaFree(c->d_alpha[i]);
    delete c;
}

int antq_host_last_launches(const antq_host_ctx *c) { return c ? c->last_launches : 0; }

static int antq_host_alpha(antq_host_ctx *c, int stage, const float *alpha_host, size_t n, cudaStream_t st) {
    if (c->alpha_cap[stage] < n) {
        cudaError_t e = cudaStreamSynchronize(st);            // the old buffer may still be read by a kernel
        if (e != cudaSuccess) return (int)e;
        if (c->d_alpha[stage]) cudaFree(c->d_alpha[stage]);
        c->d_alpha[stage] = nullptr; c->alpha_cap[stage] = 0;
        size_t cap = n < 4096 ? 4096 : n;
        e = cudaMalloc((void **)&c->d_alpha[stage], sizeof(float) * cap);
        if (e != cudaSuccess) return (int)e;
        c->alpha_cap[stage] = cap;
    }
    return (int)cudaMemcpyAsync(c->d_alpha[stage], alpha_host, sizeof(float) * n, cudaMemcpyHostToDevice, st);
}

int antq_host_synchronize(antq_host_ctx *c) {
    if (!c) return ANTQ_EINVAL;
    int rc = 0;
    for (int i = 0; i < c->n_stages; i++) {
        cudaError_t e = cudaStreamSynchronize(c->st[i]);
        if (e != cudaSuccess && rc == 0) rc = (int)e;
    }
    return rc;
}

// Enqueues H2D -> kernel -> D2H for every chunk of the tensor and returns; out_host is complete after
// antq_host_synchronize().  Chunks of consecutive calls share the stage ring, so the copies of tensor i + 1
// overlap the kernels and read-backs of tensor i (both PCIe directions stay busy across calls).
int antq_host_fakequant_async(antq_host_ctx *c, const void *x_host, void *out_host, const float *alpha_host,
                              int alpha_per_row, int64_t rows, int64_t cols, int dtype, const float *grid_host,
                              int k_normal, const float *outliers_host, int k_out, int flags) {
    const int es = esize(dtype);
    if (!c || es == 0 || rows < 0 || cols < 0 || !alpha_host || !grid_host) return ANTQ_EINVAL;
    if (k_normal < 1 || k_out < 0 || k_normal + k_out > ANTQ_MAX_GRID || (k_out > 0 && !outliers_host))
        return ANTQ_EINVAL;
    c->last_launches = 0;
    if (rows == 0 || cols == 0) return 0;
    if (!x_host || !out_host) return ANTQ_EINVAL;
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return (int)e;

    // codebook: rebuild only when the grid changed (then every stage must have finished with the old one)
    float g[ANTQ_MAX_GRID];
    memset(g, 0, sizeof(g));
    memcpy(g, grid_host, sizeof(float) * k_normal);
    if (k_out) memcpy(g + k_normal, outliers_host, sizeof(float) * k_out);
    if (c->h_k_normal != k_normal || c->h_k_out != k_out || memcmp(g, c->h_grid, sizeof(g)) != 0) {
        int rc = antq_host_synchronize(c);
        if (rc) return rc;
        cudaStream_t s0 = c->st[0];
        e = cudaMemcpyAsync(c->d_grid, g, sizeof(g), cudaMemcpyHostToDevice, s0);
        if (e != cudaSuccess) return (int)e;
        rc = antq_launch_prepare(c->d_grid, k_normal, c->d_grid + k_normal, k_out, c->cb, s0);
        if (rc) return rc;
        c->last_launches++;
        rc = antq_codebook_info_get(c->cb, &c->info, s0);      // synchronises s0: the codebook is visible to all stages
        if (rc) return rc;
        c->last_launches++;
        memcpy(c->h_grid, g, sizeof(g));
        c->h_k_normal = k_normal; c->h_k_out = k_out;
    }

    const bool ovp = (flags & ANTQ_FLAG_OVP) != 0;
    const size_t row_bytes = (size_t)cols * es;
    int rc = 0;
    int stage = c->next_stage;
    if (alpha_per_row && rows > 1) {
        if (row_bytes > c->chunk_bytes) return ANTQ_ENOTSUP;
        if (ovp && (cols & 1)) return ANTQ_ENOTSUP;     // pairs would straddle chunk boundaries
        const int64_t rpc = (int64_t)(c->chunk_bytes / row_bytes);
        for (int64_t r0 = 0; r0 < rows && rc == 0; r0 += rpc, stage = (stage + 1) % c->n_stages) {
            const int64_t nr = (rows - r0) < rpc ? (rows - r0) : rpc;
            const size_t bytes = (size_t)nr * row_bytes;
            cudaStream_t st = c->st[stage];
            rc = antq_host_alpha(c, stage, alpha_host + r0, (size_t)nr, st);
            if (rc) break;
            e = cudaMemcpyAsync(c->d_in[stage], (const char *)x_host + (size_t)r0 * row_bytes, bytes,
                                cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) { rc = (int)e; break; }
            rc = antq_fakequant(c->d_in[stage], c->d_out[stage], nullptr, c->d_alpha[stage], 1, nr, cols, dtype, c->cb,
                                &c->info, flags, st);
            if (rc) break;
            c->last_launches++;
            e = cudaMemcpyAsync((char *)out_host + (size_t)r0 * row_bytes, c->d_out[stage], bytes,
                                cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) rc = (int)e;
        }
    } else {
        const int64_t n = rows * cols;
        int64_t epc = (int64_t)(c->chunk_bytes / es) & ~(int64_t)63;   // even, vector aligned
        if (ovp && (n & 1) && n > epc) return ANTQ_ENOTSUP;            // wrap-around pair needs one launch
        for (int64_t i0 = 0; i0 < n && rc == 0; i0 += epc, stage = (stage + 1) % c->n_stages) {
            const int64_t ne = (n - i0) < epc ? (n - i0) : epc;
            const size_t bytes = (size_t)ne * es;
            cudaStream_t st = c->st[stage];
            rc = antq_host_alpha(c, stage, alpha_host, 1, st);
            if (rc) break;
            e = cudaMemcpyAsync(c->d_in[stage], (const char *)x_host + (size_t)i0 * es, bytes, cudaMemcpyHostToDevice,
                                st);
            if (e != cudaSuccess) { rc = (int)e; break; }
            rc = antq_fakequant(c->d_in[stage], c->d_out[stage], nullptr, c->d_alpha[stage], 0, 1, ne, dtype, c->cb,
                                &c->info, flags, st);
            if (rc) break;
            c->last_launches++;
            e = cudaMemcpyAsync((char *)out_host + (size_t)i0 * es, c->d_out[stage], bytes, cudaMemcpyDeviceToHost,
                                st);
            if (e != cudaSuccess) rc = (int)e;
        }
    }
    c->next_stage = stage;
    return rc;
}

int antq_host_fakequant(antq_host_ctx *c, const void *x_host, void *out_host, const float *alpha_host,
                        int alpha_per_row, int64_t rows, int64_t cols, int dtype, const float *grid_host,
                        int k_normal, const float *outliers_host, int k_out, int flags) {
    int rc = antq_host_fakequant_async(c, x_host, out_host, alpha_host, alpha_per_row, rows, cols, dtype, grid_host,
                                       k_normal, outliers_host, k_out, flags);
    const int rs = c ? antq_host_synchronize(c) : 0;
    return rc ? rc : rs;
}

}  // extern "C"
