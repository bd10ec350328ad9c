// antq_prepare.cu -- build the device codebook from the reference's `quant_grid`
// (+ OliVe `outliers`) buffers.  One CTA, K <= 512 threads, no host round trip.
//
// The reference scan (A/quant/quant_kernel.cu:25-37) walks the grid in order and
// keeps the LAST entry with the smallest fl32(|d - g_i|).  For two adjacent
// distinct values lo < hi the predicate "hi beats lo" is monotone in d, because
// fl32(|d - hi|) is non-increasing and fl32(|d - lo|) non-decreasing on [lo, hi]
// (rounding is monotone).  So an exact fp32 threshold thr = min{d : hi wins}
// exists; it is found here by bisection over the fp32 number line using the
// scan's own rounded distances and its tie rule (later scan index wins).
// rank(d) = #{r : d >= thr[r]} then names the level the scan returns, provided
// no NON-adjacent level can tie with the winner -- checked below (WELLSEP).
#include "antq_common.cuh"

namespace {

__device__ __forceinline__ bool hi_wins(float d, float lo, float hi, bool tie_hi) {
    float dl = fabsf(__fsub_rn(d, lo));
    float dh = fabsf(__fsub_rn(d, hi));
    return dh < dl || (dh == dl && tie_hi);
}

// Piecewise-uniform analysis (one thread; mirrors tests/pu_model.py::analyze line for line).  `ku` is scratch for the
// union of magnitudes in units of c.  Returns ANTQ_CB_PU (| ANTQ_CB_PU_UNIFORM) and fills the pu_* fields, or 0.
__device__ int antq_pu_analyze(const float *lev, const int *lcode, int L, int *ku, AntqCodebook *cb) {
    // the closed form settles near-midpoint elements by comparing the two neighbouring levels with "the upper one wins a
    // tie": that is the scan's rule (later entry wins) only when the scan meets the levels in ascending order
    for (int r = 0; r + 1 < L; r++)
        if (lcode[r] >= lcode[r + 1]) return 0;
    int zero = -1, npos = 0;
    for (int r = 0; r < L; r++) {
        if (lev[r] == 0.0f) zero = r;
        if (lev[r] > 0.0f) npos++;
    }
    if (zero < 0 || npos == 0) return 0;
    const float c = lev[zero + 1];                                   // smallest positive level (levels are sorted, distinct)
    for (int r = 0; r < L; r++) {
        const float k = rintf(__fdiv_rn(lev[r], c));
        if (!(fabsf(k) < 1048576.0f) || __fmul_rn(k, c) != lev[r]) return 0;
    }
    const int nneg = zero, np = L - 1 - zero;
    const float kmin = rintf(__fdiv_rn(lev[0], c)), kmax = rintf(__fdiv_rn(lev[L - 1], c));
    // the side reaching further defines U; the other side must be a prefix of it
    const bool pos_longer = kmax >= -kmin;
    const int nu = pos_longer ? np : nneg, nb = pos_longer ? nneg : np;
    for (int i = 0; i < nu; i++)
        ku[i] = (int)rintf(fabsf(__fdiv_rn(pos_longer ? lev[zero + 1 + i] : lev[zero - 1 - i], c)));
    for (int i = 0; i < nb; i++) {
        const int kb = (int)rintf(fabsf(__fdiv_rn(pos_longer ? lev[zero - 1 - i] : lev[zero + 1 + i], c)));
        if (i >= nu || kb != ku[i]) return 0;
    }
    if (ku[0] != 1) return 0;
    int e_top = 0;
    while ((2 << e_top) <= ku[nu - 1]) e_top++;
    if (e_top > 30) return 0;
    int ls[32];
    int i0 = 0;
    bool uniform = true;
    for (int e = 0; e <= e_top; e++) {
        int n = 0;
        while (i0 + n < nu && ku[i0 + n] < (2 << e)) n++;
        if (n == 0 || ku[i0] != (1 << e)) return 0;
        int step;
        if (n == 1) step = (e < e_top || e == 0) ? (1 << e) : (1 << ls[e - 1]);
        else step = ku[i0 + 1] - ku[i0];
        if (step <= 0 || (step & (step - 1)) || step > (1 << e)) return 0;
        const int full = (1 << e) / step;
        if (e < e_top ? n != full : n > full) return 0;
        for (int j = 0; j < n; j++)
            if (ku[i0 + j] != (1 << e) + j * step) return 0;
        int l2 = 0;
        while ((1 << l2) < step) l2++;
        ls[e] = l2;
        if (l2 != 0) uniform = false;
        i0 += n;
    }
    for (int E = 0; E < 256; E++) {
        int e = E - 127;
        e = e < 0 ? 0 : (e > e_top ? e_top : e);
        // .x = 1.5 * 2^(23 + ls): the rounding magic;  .y = step / 2 - 2^(e - 19): |t - round(t)| at or above it means
        // "within delta of a midpoint" (|t - round(t)| never exceeds step / 2)
        const float half_step = __uint_as_float((unsigned)(127 + ls[e] - 1) << 23);
        const float delta = __uint_as_float((unsigned)(127 + e - 19) << 23);
        cb->pu_tab[E] = make_float2(__uint_as_float(((unsigned)(150 + ls[e]) << 23) | 0x400000u), __fsub_rn(half_step, delta));
    }
    cb->pu_c = c;
    cb->pu_inv_c = __fdiv_rn(1.0f, c);
    cb->pu_kmin = kmin;
    cb->pu_kmax = kmax;
    // May the clamp be applied to the 16-bit input instead of to t?  The bounds (kmax + 0.4 step) / kx and (kmin - 0.4 step)
    // / kx (step = that of the octave the end lies in; kmin = 0: the sub-unit region) are rounded toward zero to the
    // input type: a relative error < 2^-10 (fp16) / 2^-7 (bf16) must keep them beyond the last midpoint.
    auto oct_of = [&](float k) { int e = 0; while (e < e_top && (float)(2 << e) <= k) e++; return k < 1.0f ? 0 : e; };
    const float step_hi = (float)(1 << ls[oct_of(kmax)]), step_lo = (float)(1 << ls[oct_of(-kmin)]);
    int xc = 0;
    if (kmax <= 400.0f * step_hi && -kmin <= 400.0f * step_lo) xc |= ANTQ_CB_PU_XC16;
    if (kmax <= 48.0f * step_hi && -kmin <= 48.0f * step_lo) xc |= ANTQ_CB_PU_XCBF;
    // Is every k exactly representable in FP8 e4m3 (1 + 3 significant bits, |k| <= 448)?  Then a 4-bit tensor can be fed
    // to the FP8 tensor cores as levels without any rounding (antq_gemm.cu).
    bool e4m3 = kmax <= 448.0f && -kmin <= 448.0f;
    for (int i = 0; i < nu && e4m3; i++) {
        int e = 0;
        while ((2 << e) <= ku[i]) e++;
        if (e > 3 && (ku[i] & ((1 << (e - 3)) - 1))) e4m3 = false;
    }
    return ANTQ_CB_PU | (uniform ? ANTQ_CB_PU_UNIFORM : 0) | xc | (e4m3 ? ANTQ_CB_PU_E4M3 : 0);
}

__global__ void __launch_bounds__(ANTQ_MAX_GRID) antq_prepare_kernel(const float *__restrict__ grid, int k_normal,
                                                                     const float *__restrict__ outliers, int k_out,
                                                                     AntqCodebook *__restrict__ cb) {
    __shared__ float g[ANTQ_MAX_GRID];
    __shared__ int keep[ANTQ_MAX_GRID];   // 1 if entry i is the last occurrence of its value
    __shared__ float lev[ANTQ_MAX_GRID];
    __shared__ int lcode[ANTQ_MAX_GRID];
    __shared__ float thr[ANTQ_MAX_GRID];
    __shared__ int s_nlev, s_flags_bad_sep, s_flags_bad_ste, s_sym_bad, s_ovp_bad;
    __shared__ float s_gmax;

    const int K = k_normal + k_out;
    const int i = threadIdx.x;
    if (i == 0) {
        s_nlev = 0; s_flags_bad_sep = 0; s_flags_bad_ste = 0; s_sym_bad = 0; s_ovp_bad = 0;
        float m = k_normal > 0 ? grid[0] : 0.0f;        // torch.max(self.quant_grid)
        for (int j = 1; j < k_normal; j++) m = (grid[j] > m || grid[j] != grid[j]) ? grid[j] : m;
        s_gmax = m;
    }
    float gi = 0.0f;
    if (i < K) {
        gi = i < k_normal ? grid[i] : outliers[i - k_normal];
        g[i] = gi;
    }
    __syncthreads();

    // distinct values: keep the LAST occurrence (it is the one the `<=` scan returns); drop NaN.
    int my_keep = 0;
    if (i < K && gi == gi) {
        my_keep = 1;
        for (int j = i + 1; j < K; j++)
            if (g[j] == gi) my_keep = 0;
    }
    if (i < K) keep[i] = my_keep;
    __syncthreads();
    if (my_keep) {
        int r = 0;
        for (int j = 0; j < K; j++)
            if (keep[j] && g[j] < gi) r++;
        lev[r] = (gi == 0.0f) ? 0.0f : gi;   // canonical +0
        lcode[r] = i;
        atomicAdd(&s_nlev, 1);
    }
    __syncthreads();
    const int L = s_nlev;

    // thresholds between adjacent levels
    if (i < L - 1) {
        float lo = lev[i], hi = lev[i + 1];
        bool tie_hi = lcode[i + 1] > lcode[i];
        long long a = antq_f2ord(lo), b = antq_f2ord(hi);   // hi_wins(lo) is false, hi_wins(hi) is true
        while (b - a > 1) {
            long long m = (a + b) >> 1;
            if (hi_wins(antq_ord2f((int)m), lo, hi, tie_hi)) b = m; else a = m;
        }
        float t = antq_ord2f((int)b);
        thr[i] = t;
        // (q - d) + d == q inside the span: |q - d| <= min(|q|, |d|) or q == 0  (Sterbenz)
        bool lo_ok = lo == 0.0f || (lo > 0.0f && t <= 2.0f * lo) || (lo < 0.0f && t <= 0.5f * lo);
        bool hi_ok = hi == 0.0f || (hi < 0.0f && t >= 2.0f * hi) || (hi > 0.0f && t >= 0.5f * hi);
        if (!(lo_ok && hi_ok)) s_flags_bad_ste = 1;
        // a non-adjacent level can only tie with the winner if two gaps differ by > 2^22
        if (i < L - 2) {
            float span = lev[i + 2] - lo;
            if (!((hi - lo) > span * 4.8e-7f) || !((lev[i + 2] - hi) > span * 4.8e-7f)) s_flags_bad_sep = 1;
        }
    }
    __syncthreads();

    if (i == 0) {
        const float vmax = L > 0 ? lev[L - 1] : 0.0f, vmin = L > 0 ? lev[0] : 0.0f;
        int flags = 0;
        // Beyond the ends the winner is the extreme level unless the runner-up has a LATER scan index
        // and their rounded distances coincide, which needs gap < ulp(distance): bound the window so
        // that cannot happen (e.g. unsigned pot: 0 and 10*2^-14 tie for d < -10240 in the reference).
        float lim_idx = 65536.0f;
        if (L >= 2) {
            if (lcode[L - 2] > lcode[L - 1]) lim_idx = fminf(lim_idx, (lev[L - 1] - lev[L - 2]) * 2097152.0f);
            if (lcode[1] > lcode[0]) lim_idx = fminf(lim_idx, (lev[1] - lev[0]) * 2097152.0f);
        }
        bool sep = !s_flags_bad_sep && L >= 1 && fabsf(vmax) <= 4096.0f && fabsf(vmin) <= 4096.0f &&
                   lim_idx >= fmaxf(fabsf(vmax), fabsf(vmin));
        if (sep) flags |= ANTQ_CB_WELLSEP;
        bool ste = !s_flags_bad_ste && vmax > 0.0f && vmin <= 0.0f;
        if (ste) flags |= ANTQ_CB_STE_EXACT;
        // symmetric about a zero level?
        bool sym = (L & 1) && L >= 3;
        int mid = L >> 1;
        if (sym) {
            if (lev[mid] != 0.0f) sym = false;
            for (int k = 1; sym && k <= mid; k++)
                if (lev[mid + k] != -lev[mid - k]) sym = false;
        }
        if (sym) flags |= ANTQ_CB_SYMMETRIC;
        // symmetric except for ONE extra level at the negative end (signed int-k: -2^(k-1) .. 2^(k-1) - 1)?
        bool symx = !sym && !(L & 1) && L >= 4 && lev[L >> 1] == 0.0f;
        for (int k = 1; symx && k < (L >> 1); k++)
            if (lev[(L >> 1) + k] != -lev[(L >> 1) - k]) symx = false;
        if (symx) { flags |= ANTQ_CB_SYMX; mid = L >> 1; }
        // OVP: outliers are levels with |v| > 32 (O/antquant/quant_modules.py:314)
        int ovp_index = -1;
        bool ovp_ok = true;
        if (sym) {
            for (int k = 0; k <= mid; k++)
                if (fabsf(lev[mid + k]) > 32.0f) { ovp_index = k - 1; break; }
        } else {
            for (int r = 0; r < L; r++) {
                if (lev[r] < -32.0f) ovp_ok = false;            // negative outliers need the symmetric path
                if (lev[r] > 32.0f && ovp_index < 0) ovp_index = r - 1;
            }
        }
        if (ovp_index < 0 && !sym) { /* no outlier level at all: OVP is a no-op */ }
        if (ovp_ok) flags |= ANTQ_CB_OVP_OK;
        cb->pu_c = 0.0f; cb->pu_inv_c = 0.0f; cb->pu_kmin = 0.0f; cb->pu_kmax = 0.0f;
        cb->pu_tout = __int_as_float(0x7f800000);
        if (sep && ste) flags |= antq_pu_analyze(lev, lcode, L, keep, cb);     // keep[] is free by now: scratch
        // OliVe: grid + outliers is not piecewise uniform, but the NORMAL levels (|v| <= 32, O/antquant/quant_modules.py:314)
        // usually are.  Then every element with thr[lo - 1] <= d < thr[hi] quantizes inside the normal grid by the closed
        // form, and only vectors holding an element beyond that window -- an outlier, which also makes its neighbour a
        // victim -- need the pair logic on the whole codebook (antq_pu.cu, OVP mode).
        // (The pu_* fields then describe the normal levels, so a whole-codebook ANTQ_CB_PU found above is withdrawn: launches
        // without pairs on an OliVe codebook take the chain / generic kernels.)
        if (sep && ste && k_out > 0 && ovp_ok) {
            int lo = 0, hi = L - 1;
            while (lo < L && fabsf(lev[lo]) > 32.0f) lo++;
            while (hi >= 0 && fabsf(lev[hi]) > 32.0f) hi--;
            bool contiguous = lo <= hi;
            for (int r = lo; r <= hi && contiguous; r++)
                if (fabsf(lev[r]) > 32.0f) contiguous = false;
            if (contiguous && (lo > 0 || hi < L - 1)) {
                const int whole = flags & (ANTQ_CB_PU | ANTQ_CB_PU_UNIFORM | ANTQ_CB_PU_XC16 | ANTQ_CB_PU_XCBF | ANTQ_CB_PU_E4M3);
                const int sub = antq_pu_analyze(lev + lo, lcode + lo, hi - lo + 1, keep, cb);
                flags &= ~whole;
                if (!(sub & ANTQ_CB_PU) && whole) flags |= antq_pu_analyze(lev, lcode, L, keep, cb);   // restore the fields
                if (sub & ANTQ_CB_PU) {
                    float tout = __int_as_float(0x7f800000);
                    if (hi < L - 1) tout = fminf(tout, thr[hi]);
                    if (lo > 0) tout = fminf(tout, -thr[lo - 1]);
                    if (tout > 0.0f) {
                        cb->pu_tout = tout;
                        flags |= ANTQ_CB_PU_OVP | (sub & ~(ANTQ_CB_PU | ANTQ_CB_PU_E4M3));
                    }
                }
            }
        }

        cb->n_entries = K;
        cb->n_normal = k_normal;
        cb->n_levels = L;
        cb->flags = flags;
        cb->n_mag = sym ? mid + 1 : (symx ? mid : 0);     // SYMX: magnitudes 0 .. mid-1 exist on both sides
        cb->mid = (sym || symx) ? mid : 0;
        cb->ovp_index = ovp_index;
        cb->gmax = s_gmax;
        cb->vmax = vmax;
        cb->vmin = vmin;
        // window in which clipped values still satisfy (q - d) + d == q  (|d| <= 2 |q|);
        // an unsigned grid (vmin == 0) is limited by its positive end only.
        float lim_pos = 2.0f * vmax;
        float lim_neg = vmin < 0.0f ? -2.0f * vmin : lim_pos;
        float lim = lim_pos < lim_neg ? lim_pos : lim_neg;
        cb->lim = lim < lim_idx ? lim : lim_idx;
        cb->lim_idx = lim_idx;
        cb->magic = ANTQ_CB_MAGIC;
    }
    if (i < ANTQ_MAX_GRID) {
        cb->grid[i] = i < K ? g[i] : 0.0f;
        cb->level[i] = i < L ? lev[i] : 0.0f;
        cb->level_code[i] = i < L ? lcode[i] : ANTQ_CODE_NONE;
        cb->thr[i] = i < L - 1 ? thr[i] : __int_as_float(0x7f800000);
    }
    // symmetric magnitude thresholds
    if (i < ANTQ_MAX_GRID / 2) {
        int mid = L >> 1;                 // index of the zero level for SYMMETRIC (L odd) and SYMX (L even) alike
        float tp = __int_as_float(0x7f800000), tn = tp;
        if ((L & 1) ? i < mid : (L >= 4 && i < mid - 1)) {
            tp = thr[mid + i];
            // negative side: -m_{i+1} is chosen while d < thr[mid-i-1], i.e. -d >= nextup(-thr)
            float nt = -thr[mid - i - 1];
            tn = antq_ord2f(antq_f2ord(nt) + 1);
        }
        cb->mag_tpos[i] = tp;
        cb->mag_tneg[i] = tn;
    }
}

__global__ void antq_cb_info_kernel(const AntqCodebook *__restrict__ cb, antq_codebook_info *__restrict__ out) {
    out->n_entries = cb->n_entries; out->n_normal = cb->n_normal; out->n_levels = cb->n_levels;
    out->flags = cb->flags; out->n_mag = cb->n_mag; out->mid = cb->mid; out->ovp_index = cb->ovp_index;
    out->reserved = cb->magic;
    out->gmax = cb->gmax; out->vmax = cb->vmax; out->vmin = cb->vmin; out->lim = cb->lim;
}

}  // namespace

int antq_launch_prepare(const float *grid, int k_normal, const float *outliers, int k_out, AntqCodebook *cb,
                        cudaStream_t stream) {
    antq_prepare_kernel<<<1, ANTQ_MAX_GRID, 0, stream>>>(grid, k_normal, outliers, k_out, cb);
    return (int)cudaGetLastError();
}

extern "C" int antq_codebook_info_get(const void *codebook, antq_codebook_info *info_host, void *stream) {
    if (!codebook || !info_host) return ANTQ_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    antq_codebook_info *tmp = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&tmp, sizeof(antq_codebook_info), st);
    if (e != cudaSuccess) return (int)e;
    antq_cb_info_kernel<<<1, 1, 0, st>>>((const AntqCodebook *)codebook, tmp);
    e = cudaMemcpyAsync(info_host, tmp, sizeof(antq_codebook_info), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFreeAsync(tmp, st);
    return (int)e;
}
