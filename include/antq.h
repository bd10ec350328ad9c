/*
 * antq.h -- C ABI of libantq.so: the B200 (sm_100a) fake-quant forward of
 * ANT / OliVe.  Plain pointers and sizes only; no torch / C++ types.
 *
 * What each entry point replaces in the reference (A/ = ant_quantization/,
 * O/ = olive_quantization/ of clevercool/ANT-Quantization @ bc84067):
 *
 *   antq_lut_nearest        quant_cuda.quant(x, y)            A/quant/quant.cpp:16-28,
 *                                                             A/quant/quant_kernel.cu:11-62
 *   antq_codebook_prepare   (new) turns the `quant_grid` (+`outliers`) buffers
 *                           A/antquant/quant_modules.py:42, O/antquant/quant_modules.py:44-45
 *                           into the device codebook the fused kernels read
 *   antq_fakequant          Quantizer._forward               A/antquant/quant_modules.py:535-551
 *                           OliVe Quantizer._forward + OVP   O/antquant/quant_modules.py:295-330
 *   antq_absmax             the abs-max alpha init           A/antquant/quant_modules.py:473-477
 *   antq_mse_sweep          search_mse's candidate loop      A/antquant/quant_modules.py:287-326,
 *                                                             O/antquant/quant_modules.py:190-233
 *   antq_calibrate          search_mse + search_adaptive_numeric_type, fused   A/antquant/quant_modules.py:287-415
 *   antq_fakequant_backward autograd of _forward (QAT)                 A/antquant/quant_modules.py:535-551
 *   antq_encode_p4 / antq_decode_p4   the never-written `tensor_idx`   A/quant/quant_kernel.cu:18,49,61
 *   antq_linear_p4          F.linear on fake-quantized operands, fused  A/antquant/quant_modules.py:642-646
 *   antq_host_*             the same forward for HOST buffers (copies inside)
 *
 * Conventions
 *   - every device entry point is asynchronous on `stream`, never allocates,
 *     never synchronises, never throws; it returns 0 on success, a positive
 *     cudaError_t, or a negative ANTQ_E* argument error.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - all pointers are device pointers unless the name ends in `_host`.
 *   - tensors are contiguous, viewed as [rows, cols]; per-tensor scale is
 *     rows = 1, cols = numel.  alpha is fp32, one value per row
 *     (alpha_per_row = 1) or a single value (alpha_per_row = 0).
 *   - dtype is the I/O element type; arithmetic is fp32 exactly as in the
 *     reference; fp16/bf16 results are the fp32 result rounded to nearest even.
 */
#ifndef ANTQ_H
#define ANTQ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ANTQ_ABI_VERSION 1
#define ANTQ_MAX_GRID 512           /* entries in grid + outliers */

/* dtype */
#define ANTQ_F32  0
#define ANTQ_F16  1
#define ANTQ_BF16 2

/* flags for antq_fakequant */
#define ANTQ_FLAG_OVP        1      /* OliVe outlier-victim pair masking on the flat tensor */
#define ANTQ_FLAG_FORCE_FLAT 2      /* testing: always take the generic flat kernel */
#define ANTQ_FLAG_FORCE_ROWS 4      /* testing: the row-table kernel (plan 1) or fail with ANTQ_ENOTSUP */
#define ANTQ_FLAG_FORCE_PU   8      /* testing: the closed-form kernels (plans 4 / 5) or fail with ANTQ_ENOTSUP */
#define ANTQ_FLAG_NO_PU     16      /* testing / A-B: never take the closed-form kernels */
#define ANTQ_FLAG_FORCE_TILE 32     /* testing / A-B: the closed-form TILE kernel (plan 5) whatever the row length, or ANTQ_ENOTSUP */

/* argument errors (negative); positive return values are cudaError_t */
#define ANTQ_EINVAL  (-1)
#define ANTQ_ENOTSUP (-2)
#define ANTQ_EALIGN  (-3)

/* codes written by the kernels: index into the (concatenated) grid of the
 * entry the reference scan selects (last minimal entry), or one of: */
#define ANTQ_CODE_NONE   (-1)       /* no entry within 102400 / NaN / Inf: the scan keeps z = 0 */
/* victims of the OVP mask get code == number of grid entries */

int antq_abi_version(void);
const char *antq_build_info(void);
const char *antq_error_string(int status);

/* Bytes of device memory one codebook occupies. */
size_t antq_codebook_bytes(void);

/* Build the codebook on the device from device-resident grid buffers.
 * grid: k_normal fp32 entries (the `quant_grid` buffer); outliers: k_out fp32
 * entries or NULL/0 (the OliVe `outliers` buffer).  The scan order is
 * grid then outliers, as torch.cat does (O/antquant/quant_modules.py:304). */
int antq_codebook_prepare(const float *grid, int k_normal, const float *outliers, int k_out,
                          void *codebook, void *stream);

/* z[i] = grid entry the reference scan selects for x[i]; codes optional (int16). */
int antq_lut_nearest(const void *x, void *z, int16_t *codes, int64_t n, int dtype,
                     const void *codebook, void *stream);

/* Introspection: copy the prepared codebook's header to the host.  This is the ONE
 * synchronising call of the device API (it waits for `stream`); call it once after
 * antq_codebook_prepare and keep the result next to the codebook pointer. */
typedef struct antq_codebook_info {
    int32_t n_entries, n_normal, n_levels, flags, n_mag, mid, ovp_index, reserved;
    float gmax, vmax, vmin, lim;
} antq_codebook_info;
int antq_codebook_info_get(const void *codebook, antq_codebook_info *info_host, void *stream);

#define ANTQ_CB_WELLSEP   1   /* threshold search is provably equal to the scan */
#define ANTQ_CB_STE_EXACT 2   /* (q - d) + d == q inside the window |d| <= lim */
#define ANTQ_CB_SYMMETRIC 4   /* levels symmetric about a zero level */
#define ANTQ_CB_OVP_OK    8   /* no outlier level (|v| > 32) on the negative side of an asymmetric grid */
#define ANTQ_CB_SYMX     16   /* symmetric about zero except for one extra level at the negative end (signed int-k);
                                 n_mag = magnitudes present on both sides, mid = index of the zero level */
#define ANTQ_CB_PU       32   /* piecewise uniform: levels are fl32(k * c), k integer, power-of-two step per octave
                                 (int / flint / pot / float of every width): the closed-form kernels apply */
#define ANTQ_CB_PU_UNIFORM 64 /* PU with one step for every octave (int-k) */
#define ANTQ_CB_PU_XC16  128  /* PU: the clamp to [kmin, kmax] may be done on the fp16 INPUT (packed min / max): the largest
                                 level is far enough from its lower midpoint for an fp16-rounded bound */
#define ANTQ_CB_PU_XCBF  256  /* the same for bf16 inputs */
#define ANTQ_CB_PU_E4M3  512  /* PU and every level / pu_c is exactly representable in FP8 e4m3 (all 4-bit int / flint / pot /
                                 float grids): the FP8 tensor-core path applies */
#define ANTQ_CB_PU_OVP   1024 /* grid + outliers (OliVe): the NORMAL levels (|v| <= 32) are piecewise uniform (UNIFORM / XC bits describe
                                 them); elements that stay below the first outlier threshold take the closed form, vectors holding
                                 an outlier take the pair logic on the whole codebook */

/* Fused scale -> nearest -> (OVP) -> STE -> rescale.  out may alias x (except OVP with odd numel).
 * `info` (host pointer, may be NULL) lets the call pick the row-table kernel
 * without touching device memory; with NULL the generic flat kernel runs. */
int antq_fakequant(const void *x, void *out, int16_t *codes, const float *alpha, int alpha_per_row,
                   int64_t rows, int64_t cols, int dtype, const void *codebook,
                   const antq_codebook_info *info, int flags, void *stream);

/* Which kernel antq_fakequant launches for these arguments:
 * 1 = antq_stream_kernel (per-row x-space threshold chain; <= 7 thresholds after folding signs, OliVe),
 * 4 = antq_pu_stream_kernel (closed form for piecewise-uniform grids: unsigned 4-bit, 5 to 8 bit),
 * 5 = antq_pu_short_kernel (the same for rows shorter than 512 elements: scale groups, 1x1-conv weights),
 * 3 = antq_short_kernel (d-space threshold chain, short rows of the grids that are not piecewise uniform),
 * 2 = antq_flat_kernel (generic; the one that emits int16 code indices), <0 = error. */
int antq_fakequant_plan(const antq_codebook_info *info, int64_t rows, int64_t cols, int dtype, int flags,
                        const void *x, const void *out, const void *codes);

/* Dynamic group scales in ONE read of x (north star: "group-wise abs-max ... with warp shuffles, encodes and decodes in
 * registers"): alpha[r] = fl32(max_c |x[r, c]| * ratio), then the fused fake-quant of row r with that alpha -- the
 * reference's per-channel path (A/antquant/quant_modules.py:473-477 + 535-551) on the [numel / G, G] view, without the
 * separate abs-max pass.  alpha_out (optional, rows floats) receives the scales.  Piecewise-uniform grids, no OVP, rows
 * of at most 32 16-byte vectors with a power-of-two vector count (group-8 ... 256 for fp16): otherwise ANTQ_ENOTSUP and
 * the caller runs antq_absmax + antq_fakequant. */
int antq_fakequant_dynamic(const void *x, void *out, float *alpha_out, float ratio, int64_t rows, int64_t cols, int dtype,
                           const void *codebook, const antq_codebook_info *info, int flags, void *stream);

/* out[r] = max_c |x[r, c]| as fp32 (rows = 1: whole tensor). */
int antq_absmax(const void *x, float *out, int64_t rows, int64_t cols, int dtype, void *stream);

/* For every candidate c in [0, n_cand): alpha_c[r] = base[r] * ratio[c]; err[c, r] =
 * sum_c (fakequant(x)[r, c] - x[r, c])^2 accumulated in fp32 per thread, fp64 across
 * threads.  One read of x per CAND_TILE candidates. */
int antq_mse_sweep(const void *x, const float *base_alpha, int alpha_per_row, const float *ratios, int n_cand,
                   double *err, int64_t rows, int64_t cols, int dtype, const void *codebook, int flags,
                   void *stream);

/* ---- packed 4-bit code storage ("P4"), the second output the reference kernel allocates and never writes
 * (`tensor_idx`, A/quant/quant_kernel.cu:18,49,61) ----
 * Byte j of a row = code of element 2j (low nibble) | code of element 2j + 1 (high nibble); a code is the index into
 * `quant_grid` of the level the scan selects.  OliVe with outliers (normal grid <= 15 entries): nibble 15 marks the
 * victim of an outlier-victim pair and the other nibble then indexes `outliers` (O/antquant/quant_modules.py:311-320).
 * cols must be even, the grid at most 16 entries (15 + 15 with outliers).
 * n_inexact (device, optional) receives the number of elements whose fake-quant value antq_decode_p4 would NOT
 * reproduce bit for bit (STE rounding of values clipped beyond twice the largest level, NaN, Inf): zero means
 * decode(encode(x)) == antq_fakequant(x) everywhere. */
int antq_encode_p4(const void *x, uint8_t *codes, const float *alpha, int alpha_per_row, int64_t rows, int64_t cols,
                   int dtype, const void *codebook, const antq_codebook_info *info, int flags, unsigned int *n_inexact,
                   void *stream);
int antq_decode_p4(const uint8_t *codes, void *out, const float *alpha, int alpha_per_row, int64_t rows, int64_t cols,
                   int dtype, const void *codebook, const antq_codebook_info *info, int flags, void *stream);

/* ---- QAT backward of Quantizer._forward under autograd (A/antquant/quant_modules.py:535-551; driver
 * A/ImageNet/main.py:190-198): one pass over (grad_out, x, out).
 *   grad_x[i]     = fl(fl(g * s) / s)                      (what autograd's mul-then-div produces; NULL to skip)
 *   grad_alpha[r] = sum_row g * (out - x) / s / max(grid)  (per row, or one value; NULL to skip)
 * Deterministic (fixed-order fp64 reduction through `workspace`, antq_backward_workspace_bytes bytes). */
size_t antq_backward_workspace_bytes(int64_t rows, int64_t cols, int alpha_per_row);
int antq_fakequant_backward(const void *grad_out, const void *x, const void *out, const float *alpha, int alpha_per_row,
                            int64_t rows, int64_t cols, int dtype, float gmax, void *grad_x, float *grad_alpha,
                            void *workspace, size_t workspace_bytes, void *stream);

/* ---- fused calibration: search_mse + search_adaptive_numeric_type in one read of the tensor
 * (A/antquant/quant_modules.py:287-326,328-415; O/antquant/quant_modules.py:190-256) ----
 * Every candidate alpha = base_alpha[row] * ratios[c] of every codebook k is scored (mean squared error of the
 * fake-quant forward per row).  Outputs, all on the device, no host synchronisation:
 *   alpha_out[k][row]   the FIRST strictly-best candidate's alpha (base itself if no candidate beats 1e10, as the
 *                       reference's `score < best_score` update leaves it)
 *   mse_out[k]          sum over rows of the best mean squared error: the reference's type-selection score
 *   best_index_out      optional [k][row] candidate index (-1: none)
 * codebooks / infos / flags_per_codebook are HOST arrays of n_cb (<= 8) entries (device codebook pointers, host
 * headers, ANTQ_FLAG_OVP per codebook).  n_cand <= 256.  Deterministic: fixed-order fp64 reductions in `workspace`. */
size_t antq_calibrate_workspace_bytes(int64_t rows, int64_t cols, int alpha_per_row, int n_cand, int n_cb);
int antq_calibrate(const void *x, int64_t rows, int64_t cols, int dtype, int alpha_per_row, const float *base_alpha,
                   const float *ratios, int n_cand, const void *const *codebooks, const antq_codebook_info *const *infos,
                   const int *flags_per_codebook, int n_cb, float *alpha_out, float *mse_out, int *best_index_out,
                   void *workspace, size_t workspace_bytes, void *stream);

/* ---- dequant-fused Linear on the tensor cores (tcgen05.mma, accumulators in tensor memory, TMA for x) ----
 * y[M, N] = x[M, K] . dequant(W)[N, K]^T + bias[N], W given as P4 codes [N, K / 2] + alpha[N] (antq_encode_p4).
 * Replaces  F.linear(quant_input(x), quant_weight(W), bias)  of LinearQuantizer.forward
 * (A/antquant/quant_modules.py:642-646; torch.addmm in O/antquant/quant_modules.py:379).
 * dtype: ANTQ_F16 or ANTQ_BF16 (x, bias, y); fp32 accumulation; K % 64 == 0, N % 256 == 0, grid <= 16 entries,
 * no outlier-victim pairs (decode those first): otherwise ANTQ_ENOTSUP and the caller keeps the unfused path. */
int antq_linear_p4(const void *x, const uint8_t *w_codes, const float *w_alpha, const void *bias, void *y, int64_t M,
                   int64_t N, int64_t K, int dtype, const void *codebook, const antq_codebook_info *info, int flags,
                   void *stream);

/* FP8 variant for W4A4: both operands travel as e4m3 LEVELS (level / smallest positive level: small integers, exact in
 * e4m3 when ANTQ_CB_PU_E4M3 is set), tcgen05.mma kind::f8f6f4 runs at twice the 16-bit rate, the integer products are
 * exact in the fp32 accumulator and both scales are applied in the epilogue:
 *   y[m, n] = (sum_k kx[m, k] kw[n, k]) * (c_x alpha_x / max(grid_x)) * (c_w alpha_w[n] / max(grid_w)) + bias[n].
 * antq_levels_e4m3 turns a fake-quantized activation (per-tensor alpha) into its level bytes.  K % 128 == 0, N % 256 == 0. */
int antq_levels_e4m3(const void *x_q, uint8_t *levels, const float *alpha, int64_t n, int dtype, const void *codebook,
                     const antq_codebook_info *info, void *stream);
int antq_linear_p4_fp8(const uint8_t *x_levels, const float *x_alpha, const void *x_codebook, const antq_codebook_info *x_info,
                       const uint8_t *w_codes, const float *w_alpha, const void *bias, void *y, int64_t M, int64_t N, int64_t K,
                       int out_dtype, const void *w_codebook, const antq_codebook_info *w_info, int flags, void *stream);

/* ---- host-buffer path (what a CPU caller of the reference would use) ---- */
typedef struct antq_host_ctx antq_host_ctx;
/* Creates streams, staging buffers (chunk_bytes per direction per stage) on `device`. */
int antq_host_create(antq_host_ctx **ctx, int device, size_t chunk_bytes, int n_stages);
void antq_host_destroy(antq_host_ctx *ctx);
/* Synchronous: returns when out_host is complete.  x_host/out_host should be pinned. */
int antq_host_fakequant(antq_host_ctx *ctx, const void *x_host, void *out_host, const float *alpha_host,
                        int alpha_per_row, int64_t rows, int64_t cols, int dtype, const float *grid_host,
                        int k_normal, const float *outliers_host, int k_out, int flags);
/* The same, without waiting: out_host is complete after antq_host_synchronize().  Consecutive calls share
 * the stage ring, so both PCIe directions stay busy across tensors.  x_host, out_host (and a pinned
 * alpha_host) must stay valid until then. */
int antq_host_fakequant_async(antq_host_ctx *ctx, const void *x_host, void *out_host, const float *alpha_host,
                              int alpha_per_row, int64_t rows, int64_t cols, int dtype, const float *grid_host,
                              int k_normal, const float *outliers_host, int k_out, int flags);
int antq_host_synchronize(antq_host_ctx *ctx);
/* Kernels launched by the most recent antq_host_fakequant call. */
int antq_host_last_launches(const antq_host_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* ANTQ_H */
