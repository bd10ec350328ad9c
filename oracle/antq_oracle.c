/*
 * antq_oracle.c -- CPU restatement of the ANT / OliVe fake-quant forward.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under ant-quantization_b200/ may link,
 * import or call this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, and only as the
 * checker or as the timed CPU baseline.
 *
 * Parity status: pinned against golden vectors produced by executing the
 * unmodified reference Python (tests/golden/make_golden.py) with a second,
 * independent pure-torch restatement of the 27-line scan kernel injected as
 * `quant_cuda`.  The reference's own native half is a CUDA kernel and cannot
 * run in the CPU-only build container; tests/test_gpu_reference_kernel.py
 * checks this file against that kernel (oracle/_ref) on the GPU box.
 *
 * Reference citations (A/ = ant_quantization/, O/ = olive_quantization/):
 *   scan            A/quant/quant_kernel.cu:20-38   (identical O/quant/...)
 *   ANT  _forward   A/antquant/quant_modules.py:535-551
 *   OliVe _forward  O/antquant/quant_modules.py:295-330 (OVP at :311-320)
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math -shared -fPIC.
 * -ffp-contract=off matters: (q - d) + d and t * s must round after every
 * operation exactly as the reference's separate PyTorch kernels do.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ANTQ_ORACLE_NO_CODE (-1)

/* A/quant/quant_kernel.cu:25-37.  Initial best distance 102400, initial
 * result 0, `<=` so the LAST minimal entry wins, fp32 arithmetic. */
static inline float scan_one(float x_v, const float *y, int y_size, int32_t *code)
{
    float sub_min = 102400.0f;
    float z_min = 0.0f;
    int32_t best = ANTQ_ORACLE_NO_CODE;
    for (int i = 0; i < y_size; i++) {
        float sub_v = fabsf(x_v - y[i]);
        if (sub_v <= sub_min) {
            sub_min = sub_v;
            z_min = y[i];
            best = i;
        }
    }
    if (code) *code = best;
    return z_min;
}

/* quant_cuda.quant(x, y) -> z (and the scan index, which the reference
 * allocates but never writes: A/quant/quant_kernel.cu:18,49,61). */
void antq_oracle_scan_f32(const float *x, int64_t n, const float *y, int y_size,
                          float *z, int32_t *codes)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        int32_t c;
        z[i] = scan_one(x[i], y, y_size, &c);
        if (codes) codes[i] = c;
    }
}

static float grid_max(const float *y, int k)
{
    /* torch.max(quant_grid): NaN-propagating max is irrelevant here, the
     * reference grids never contain NaN. */
    float m = y[0];
    for (int i = 1; i < k; i++) if (y[i] > m) m = y[i];
    return m;
}

/*
 * ANT Quantizer._forward, A/antquant/quant_modules.py:535-551.
 *   scale = alpha / max(grid)                          (:536)
 *   d     = x / scale   (per row when per-channel)     (:538-541)
 *   q     = quant_cuda.quant(d, grid)                  (:543)
 *   t     = (q - d) + d                                (:544)
 *   out   = t * scale                                  (:546-549)
 * alpha has `rows` entries when per_row != 0, else one entry.
 */
void antq_oracle_ant_forward_f32(const float *x, float *out, int32_t *codes,
                                 const float *alpha, int per_row,
                                 int64_t rows, int64_t cols,
                                 const float *grid, int k)
{
    const float gmax = grid_max(grid, k);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; r++) {
        const float scale = (per_row ? alpha[r] : alpha[0]) / gmax;
        const float *xr = x + r * cols;
        float *orow = out + r * cols;
        for (int64_t c = 0; c < cols; c++) {
            volatile float d = xr[c] / scale;
            int32_t code;
            float q = scan_one(d, grid, k, &code);
            volatile float diff = q - d;
            volatile float t = diff + d;
            orow[c] = t * scale;
            if (codes) codes[r * cols + c] = code;
        }
    }
}

/*
 * OliVe Quantizer._forward, O/antquant/quant_modules.py:295-330.
 * grid = cat(quant_grid, outliers) unless no_outlier (:303-306); the OVP
 * mask runs on the FLAT tensor (:313-320), including torch.roll's
 * wrap-around when numel is odd.  Victims receive code k_total (one past
 * the concatenated grid).
 */
void antq_oracle_olive_forward_f32(const float *x, float *out, int32_t *codes,
                                   const float *alpha, int per_row,
                                   int64_t rows, int64_t cols,
                                   const float *grid, int k_normal,
                                   const float *outliers, int k_out,
                                   int no_outlier)
{
    const int64_t n = rows * cols;
    const int k = no_outlier ? k_normal : k_normal + k_out;
    float *cat = (float *)malloc(sizeof(float) * (size_t)(k > 0 ? k : 1));
    memcpy(cat, grid, sizeof(float) * (size_t)k_normal);
    if (!no_outlier) memcpy(cat + k_normal, outliers, sizeof(float) * (size_t)k_out);
    /* scale uses max(self.quant_grid), NOT the concatenated grid (:296). */
    const float gmax = grid_max(grid, k_normal);

    float *d = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    float *q = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    int32_t *cd = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    unsigned char *mask = (unsigned char *)malloc((size_t)(n > 0 ? n : 1));

#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; r++) {
        const float scale = (per_row ? alpha[r] : alpha[0]) / gmax;
        for (int64_t c = 0; c < cols; c++) {
            int64_t i = r * cols + c;
            d[i] = x[i] / scale;
            q[i] = scan_one(d[i], cat, k, &cd[i]);
            mask[i] = fabsf(q[i]) > 32.0f;           /* :314 */
        }
    }
    if (!no_outlier && n > 0) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; i++) {
            /* victim_odd = roll(mask, 1); victim_odd[::2] = 0       (:315-316) */
            int v_odd = (i & 1) ? mask[i - 1] : 0;
            /* victim_even = roll(mask & ~victim_odd, -1); [1::2]=0  (:317-318) */
            int v_even = 0;
            if (!(i & 1)) {
                int64_t j = (i + 1 == n) ? 0 : i + 1;
                int vo_j = (j & 1) ? mask[j - 1] : 0;
                v_even = mask[j] && !vo_j;
            }
            int victim = v_odd | v_even;
            /* quant_data * (~victim): a float times a bool         (:320) */
            volatile float keep = victim ? 0.0f : 1.0f;
            q[i] = q[i] * keep;
            if (victim) cd[i] = k;
        }
    }
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; r++) {
        const float scale = (per_row ? alpha[r] : alpha[0]) / gmax;
        for (int64_t c = 0; c < cols; c++) {
            int64_t i = r * cols + c;
            volatile float diff = q[i] - d[i];       /* :323 */
            volatile float t = diff + d[i];
            out[i] = t * scale;                      /* :325-328 */
            if (codes) codes[i] = cd[i];
        }
    }
    free(cat); free(d); free(q); free(cd); free(mask);
}

/*
 * fp16 I/O definition (SURVEY.md section 2 deviations table): the reference
 * kernel dispatches float/double only, so the fp16 oracle is
 * upcast -> fp32 reference path -> round-to-nearest-even to fp16.
 */
void antq_oracle_ant_forward_f16(const _Float16 *x, _Float16 *out, int32_t *codes,
                                 const float *alpha, int per_row,
                                 int64_t rows, int64_t cols,
                                 const float *grid, int k)
{
    const float gmax = grid_max(grid, k);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; r++) {
        const float scale = (per_row ? alpha[r] : alpha[0]) / gmax;
        const _Float16 *xr = x + r * cols;
        _Float16 *orow = out + r * cols;
        for (int64_t c = 0; c < cols; c++) {
            volatile float d = (float)xr[c] / scale;
            int32_t code;
            float q = scan_one(d, grid, k, &code);
            volatile float diff = q - d;
            volatile float t = diff + d;
            volatile float o = t * scale;
            orow[c] = (_Float16)o;
            if (codes) codes[r * cols + c] = code;
        }
    }
}

#ifdef _OPENMP
#include <omp.h>
#endif
/* torchrun exports OMP_NUM_THREADS=1; the timed CPU baseline sets its thread count explicitly. */
int antq_oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

int antq_oracle_abi_version(void) { return 1; }
