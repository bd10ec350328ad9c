// antq_gemm.cu -- dequant-fused Linear on the 5th-generation tensor cores:
//
//     y[M, N] = x[M, K] . dequant(W)[N, K]^T + bias[N]
//
// What it replaces: LinearQuantizer.forward's  F.linear(quant_input(x), quant_weight(W), bias)
// (A/antquant/quant_modules.py:642-646; torch.addmm in O/antquant/quant_modules.py:379) when the weight is held as
// packed 4-bit codes (antq_codes.cu: 0.5 byte per element + one fp32 alpha per output channel) instead of a
// re-fake-quantized fp16 / fp32 tensor.  x is the (already fake-quantized) activation, fp16 or bf16.
//
// Arithmetic.  W_q[n, k] = level[code[n, k]] * s_n with s_n = alpha_n / max(grid).  The kernel multiplies the
// UNSCALED levels (rounded once to the 16-bit operand type; every ANT 4-bit flint / pot / float level and every OliVe
// level is exact in fp16) on the tensor cores, accumulates in fp32 in tensor memory, and applies s_n and the bias in
// the epilogue:  y = fl16(acc * s_n + bias).  Against F.linear on the fake-quantized operands this differs by fp32
// accumulation order and one 16-bit rounding per weight (2^-12 relative, uncorrelated): tolerance in the test.
//
// Structure (one CTA per SM, persistent over 256 x 256 output tiles, BK = 64, kStages-deep shared-memory ring):
//   warp 4      TMA producer: x tiles [256 x 64] by cp.async.bulk.tensor.2d (128-byte swizzle) -> full[stage]
//   warps 6-13  weight decoders: thread = one of the tile's 256 output channels; 32 bytes of codes -> 64 operand values through
//               a 16-entry byte LUT held in registers (PRMT), written in the same K-major 128-byte-swizzled layout
//               -> fence.proxy.async -> full[stage]
//   warp 5      MMA issuer: one lane issues 2 x 4 tcgen05.mma (M128 N256 K16, kind::f16) per stage -- the two M-halves of
//               the 256-row x tile share the decoded W tile, which halves the decode work per flop -- into the 2 x 256
//               fp32 columns of tensor memory; tcgen05.commit frees the stage / publishes the accumulator
//   warps 0-3   epilogue: tcgen05.ld (32 lanes x 32 columns per warp), scale, bias, 16-bit stores (the accumulator fills
//               all 512 columns, so the next tile's first MMA waits for it: ~5 % of a tile)
// SASS evidence: UTCHMMA (tcgen05.mma), UTMALDG (TMA), LDTM (tcgen05.ld): profiles/r02_gemm_sass.txt.
#include <cuda.h>
#include <stdio.h>

#include "antq_common.cuh"

namespace {

// 256 x 256 output tile per CTA: two UMMA M-halves of 128 rows share one decoded W tile of 256 channels.  The x tile is
// re-read by every CTA that owns another column block, so its L2 -> SM traffic is M K (N / BN) 2 bytes: with BN = 128 that
// was 7.3 TB/s at 0.57 of the tensor peak -- the kernel was L2-bandwidth-bound, not decode- or MMA-bound (r02 notes).
constexpr int BM = 256, BN = 256, BK = 64;
constexpr int kStages = 3;
constexpr int kStageA = BM * BK * 2, kStageB = BN * BK * 2;            // bytes (16-bit operands): 32 KiB + 32 KiB
constexpr int kEpiWarps = 4, kTmaWarp = 4, kMmaWarp = 5, kDecWarp0 = 6, kDecWarps = 8;   // 2 decoder warps per scheduler
constexpr int kThreads = (kDecWarp0 + kDecWarps) * 32;                 // 448
constexpr int kTmemCols = 512;                                         // 2 (M-halves) x 256 fp32 columns: all of tensor memory
constexpr int kAccCols = 0;                                            // single accumulator buffer

struct GemmParams {
    const unsigned char *codes;       // [N, K / 2]
    const float *alpha;               // [N]
    const void *bias;                 // [N] (operand type) or NULL
    void *y;                          // [M, N]
    const AntqCodebook *cb;
    const float *x_alpha;             // FP8 mode: the activation quantizer's per-tensor alpha ...
    const AntqCodebook *x_cb;         // ... and its codebook (unit of the e4m3 levels = alpha / max(grid) * pu_c)
    int M, N, K;
    int m_tiles, n_tiles;
};

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(antq_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(antq_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(antq_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "ANTQG_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra ANTQG_DONE;\n"
        "bra ANTQG_WAIT;\n"
        "ANTQG_DONE:\n"
        "}\n" ::"r"(antq_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(antq_smem_u32(dst)), "l"(map), "r"(antq_smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
// K-major, 128-byte swizzle, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor: start >> 4 | LBO << 16 |
// SBO << 32 | version 1 << 46 | SWIZZLE_128B (2) << 61)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
template <bool FP8>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (FP8) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n"
            "}\n" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
            "}\n" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(antq_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <typename T> struct Op16;
template <> struct Op16<__half> {
    static constexpr uint32_t kFormat = 0;                      // cute::UMMA::F16F32Format::F16
    __device__ static __forceinline__ uint32_t bits(float v) { return __half_as_ushort(__float2half_rn(v)); }
    __device__ static __forceinline__ float to_f32(__half v) { return __half2float(v); }
    __device__ static __forceinline__ uint32_t pack(float a, float b) {
        const __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<const uint32_t *>(&h);
    }
};
template <> struct Op16<__nv_bfloat16> {
    static constexpr uint32_t kFormat = 1;                      // BF16
    __device__ static __forceinline__ uint32_t bits(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }
    __device__ static __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
    __device__ static __forceinline__ uint32_t pack(float a, float b) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<const uint32_t *>(&h);
    }
};

// Four 4-bit codes (the low 16 bits of `c`) -> four 16-bit operand values, through the 16-entry byte LUTs
// lo[0..3] / hi[0..3] (entry e's low / high byte sits in byte e % 4 of register e / 4).  PRMT is an 8-entry byte LUT
// with four lookups per instruction: one PRMT per half of the table, a third picks the half by bit 3 of each code.
__device__ __forceinline__ void lut4(uint32_t c, const uint32_t (&lo)[4], const uint32_t (&hi)[4], uint32_t &out01, uint32_t &out23) {
    const uint32_t sel = c & 0x7777u;
    const uint32_t pick = 0x3210u | ((c & 0x8888u) >> 1);
    const uint32_t l = __byte_perm(__byte_perm(lo[0], lo[1], sel), __byte_perm(lo[2], lo[3], sel), pick);
    const uint32_t h = __byte_perm(__byte_perm(hi[0], hi[1], sel), __byte_perm(hi[2], hi[3], sel), pick);
    out01 = __byte_perm(l, h, 0x5140u);
    out23 = __byte_perm(l, h, 0x7362u);
}

// The same for ONE byte per code (e4m3 operands): four codes -> four bytes.
__device__ __forceinline__ uint32_t lut4_b(uint32_t c, const uint32_t (&tb)[4]) {
    const uint32_t sel = c & 0x7777u;
    const uint32_t pick = 0x3210u | ((c & 0x8888u) >> 1);
    return __byte_perm(__byte_perm(tb[0], tb[1], sel), __byte_perm(tb[2], tb[3], sel), pick);
}

// float -> e4m3 bits (round to nearest even, saturating): only used on values the codebook analysis proved exact
__device__ __forceinline__ uint32_t e4m3_bits(float v) {
    unsigned short r;
    asm("{ .reg .b16 t; cvt.rn.satfinite.e4m3x2.f32 t, %1, %2; mov.b16 %0, t; }" : "=h"(r) : "f"(0.0f), "f"(v));
    return (uint32_t)(r & 0xffu);
}

// FP8 = false: x is T (fp16 / bf16), W decoded to T, tcgen05.mma kind::f16, BK = 64 elements.
// FP8 = true : x is e4m3 LEVELS (one byte per element, antq_levels_e4m3), W decoded to e4m3 levels, kind::f8f6f4 at twice
//              the tensor rate, BK = 128 elements; the products are small integers, exact in the fp32 accumulator;
//              both scales are applied in the epilogue.  Byte geometry (128-byte rows, 32 bytes per MMA k-step) is the same.
template <typename T, bool FP8>
__global__ void __launch_bounds__(kThreads, 1) antq_linear_p4_kernel(const __grid_constant__ CUtensorMap tmap_x, const GemmParams p) {
    extern __shared__ __align__(1024) unsigned char gsm_raw[];
    // the 128-byte swizzle pattern is a function of the shared-memory ADDRESS: tiles must start 1024-byte aligned
    unsigned char *gsm = gsm_raw + ((1024u - (antq_smem_u32(gsm_raw) & 1023u)) & 1023u);
    unsigned char *smem_a = gsm;                                           // kStages x 16 KiB, 1024-byte aligned
    unsigned char *smem_b = gsm + kStages * kStageA;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_b + kStages * kStageB);
    uint64_t *empty = full + kStages;
    uint64_t *tmem_full = empty + kStages;                                 // [2]
    uint64_t *tmem_empty = tmem_full + 2;                                  // [2]
    uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    float *s_scale = reinterpret_cast<float *>(tmem_base_slot + 4);        // [2][BN] alpha / max(grid) of the tile's channels
    float *s_bias = s_scale + 2 * BN;                                      // [2][BN] (only buffer 0 is used with one accumulator)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int BKE = FP8 ? 2 * BK : BK;                                  // elements of K per stage (128 bytes per row)
    const int num_kb = p.K / BKE;
    const int num_tiles = p.m_tiles * p.n_tiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(full + s, 1 + kDecWarps); mbar_init(empty + s, 1); }   // one arrival per decoder WARP
        for (int a = 0; a < 2; a++) { mbar_init(tmem_full + a, 1); mbar_init(tmem_empty + a, kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kTmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(antq_smem_u32(tmem_base_slot)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == kTmaWarp) {
        // ------------------------------ TMA producer (x tiles) ------------------------------
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
            int stage = 0;
            unsigned phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int m_blk = t % p.m_tiles;
                for (int kb = 0; kb < num_kb; kb++) {
                    mbar_wait(empty + stage, phase ^ 1u);
                    mbar_expect_tx(full + stage, kStageA);
                    tma_load_2d(smem_a + stage * kStageA, &tmap_x, full + stage, kb * BKE, m_blk * BM);
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ------------------------------ MMA issuer ------------------------------
        if (lane == 0) {
            // cute::UMMA::InstrDescriptor: D = F32 (1 << 4), A / B format << 7 / << 10, K-major both, N >> 3 << 17, M >> 4 << 24
            const uint32_t fmt = FP8 ? 0u : Op16<T>::kFormat;                      // f8f6f4: 0 = E4M3; f16: 0 = F16, 1 = BF16
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(128 >> 4) << 24);
            int stage = 0;
            unsigned phase = 0;
            int it = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, it++) {
                const int acc = 0;
                mbar_wait(tmem_empty + acc, (it & 1) ^ 1u);                        // the epilogue has drained the accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccCols);
                for (int kb = 0; kb < num_kb; kb++) {
                    mbar_wait(full + stage, phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t adesc = umma_desc(antq_smem_u32(smem_a + stage * kStageA));
                    const uint64_t bdesc = umma_desc(antq_smem_u32(smem_b + stage * kStageB));
#pragma unroll
                    for (int h = 0; h < 2; h++)                                     // two M-halves reuse the decoded W tile
#pragma unroll
                        for (int j = 0; j < 4; j++)                                 // 32 bytes of K per instruction (16 x 16-bit / 32 x 8-bit)
                            umma<FP8>(d_tmem + (uint32_t)(h * BN), adesc + (uint64_t)(h * (128 * 128 / 16) + 2 * j),
                                      bdesc + (uint64_t)(2 * j), idesc, (kb | j) != 0);
                    umma_commit(empty + stage);                                     // frees the stage when these MMAs retire
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
                umma_commit(tmem_full + acc);                                       // accumulator complete
            }
        }
    } else if (warp >= kDecWarp0) {
        // ------------------------------ weight decoders ------------------------------
        // thread = one of the tile's 256 output channels: 32 bytes of codes -> 8 x 16-byte operand chunks per k-block
        const int n_local = threadIdx.x - kDecWarp0 * 32;
        uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};                        // FP8: lo[] is the one byte table
#pragma unroll
        for (int e = 0; e < 16; e++) {
            uint32_t b = 0;
            if (e < p.cb->n_entries) b = FP8 ? e4m3_bits(__fdiv_rn(p.cb->grid[e], p.cb->pu_c)) : Op16<T>::bits(p.cb->grid[e]);
            lo[e >> 2] |= (b & 0xffu) << (8 * (e & 3));
            hi[e >> 2] |= (b >> 8) << (8 * (e & 3));
        }
        const int row_off = (n_local >> 3) * 1024 + (n_local & 7) * 128;
        const int sw = n_local & 7;
        int stage = 0;
        unsigned phase = 0;
        const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const long long total_it = (long long)my_tiles * num_kb;
        int f_tile = blockIdx.x, f_kb = 0;                                          // position of the next fetch
        constexpr int NC = FP8 ? 4 : 2;                                             // 16-byte loads of codes per k-block
        struct Codes { uint4 v[NC]; };
        auto fetch = [&](Codes &c) {
            if (f_tile < num_tiles) {
                const int n_blk = f_tile / p.m_tiles;
                const uint4 *src = reinterpret_cast<const uint4 *>(p.codes + (size_t)(n_blk * BN + n_local) * (size_t)(p.K / 2) + f_kb * (BKE / 2));
#pragma unroll
                for (int i = 0; i < NC; i++) c.v[i] = __ldg(src + i);
                if (++f_kb == num_kb) { f_kb = 0; f_tile += gridDim.x; }
            }
        };
        // kPrefetch k-blocks of codes are in flight while the current one is decoded: a global load issued and consumed in
        // the same iteration exposed ~1 us of latency per stage (profiles/r02_notes.md)
        constexpr int kPrefetch = FP8 ? 2 : 3;
        Codes pf[kPrefetch];
#pragma unroll
        for (int u = 0; u < kPrefetch; u++) {
#pragma unroll
            for (int i = 0; i < NC; i++) pf[u].v[i] = make_uint4(0, 0, 0, 0);
            fetch(pf[u]);
        }
        for (long long it = 0; it < total_it; it += kPrefetch) {
#pragma unroll
            for (int u = 0; u < kPrefetch; u++) {
                if (it + u < total_it) {
                    const Codes c = pf[u];
                    fetch(pf[u]);                                                   // the block kPrefetch ahead
                    mbar_wait(empty + stage, phase ^ 1u);
                    unsigned char *dst = smem_b + stage * kStageB + row_off;
                    const uint32_t *w = reinterpret_cast<const uint32_t *>(&c);
#pragma unroll
                    for (int j = 0; j < 8; j++) {                                   // one 16-byte chunk of the operand row
                        uint4 v;
                        if constexpr (FP8) {                                        // 16 codes = two 32-bit words -> 16 bytes
                            v.x = lut4_b(w[2 * j], lo); v.y = lut4_b(w[2 * j] >> 16, lo);
                            v.z = lut4_b(w[2 * j + 1], lo); v.w = lut4_b(w[2 * j + 1] >> 16, lo);
                        } else {                                                    // 8 codes = one 32-bit word -> 8 x 16 bit
                            lut4(w[j], lo, hi, v.x, v.y);
                            lut4(w[j] >> 16, lo, hi, v.z, v.w);
                        }
                        *reinterpret_cast<uint4 *>(dst + ((j ^ sw) << 4)) = v;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy stores -> tensor-core reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive(full + stage);                       // one arrival per warp
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else {
        // ------------------------------ epilogue (warps 0-3 = tensor-memory lanes 32 w .. 32 w + 31) ------------------------------
        const float gmax = p.cb->gmax;
        // FP8: the accumulator holds sum k_x k_w; unit = (c_w s_w[n]) (c_x s_x)
        float unit = 1.0f;
        if constexpr (FP8) unit = __fmul_rn(__fmul_rn(__fdiv_rn(p.x_alpha[0], p.x_cb->gmax), p.x_cb->pu_c), p.cb->pu_c);
        int it = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, it++) {
            const int acc = 0;
            const int m_blk = t % p.m_tiles, n_blk = t / p.m_tiles;
            float *sc = s_scale + acc * BN, *bs = s_bias + acc * BN;
            for (int c = threadIdx.x; c < BN; c += kEpiWarps * 32) {               // 128 epilogue threads, 256 channels
                const int n = n_blk * BN + c;
                sc[c] = FP8 ? __fmul_rn(__fdiv_rn(p.alpha[n], gmax), unit) : __fdiv_rn(p.alpha[n], gmax);
                bs[c] = p.bias ? Op16<T>::to_f32(reinterpret_cast<const T *>(p.bias)[n]) : 0.0f;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");                          // the epilogue warps only
            mbar_wait(tmem_full + acc, it & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int h = 0; h < 2; h++) {
                const int m = m_blk * BM + h * 128 + warp * 32 + lane;
                T *yrow = reinterpret_cast<T *>(p.y) + (size_t)m * p.N + (size_t)n_blk * BN;
#pragma unroll 1
                for (int c = 0; c < BN; c += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * kAccCols + h * BN + c), v);
                    if (m < p.M) {
#pragma unroll
                        for (int q = 0; q < 32; q += 8) {
                            uint4 o;
                            uint32_t *ow = reinterpret_cast<uint32_t *>(&o);
#pragma unroll
                            for (int e = 0; e < 8; e += 2)
                                ow[e >> 1] = Op16<T>::pack(__fmaf_rn(__uint_as_float(v[q + e]), sc[c + q + e], bs[c + q + e]),
                                                           __fmaf_rn(__uint_as_float(v[q + e + 1]), sc[c + q + e + 1], bs[c + q + e + 1]));
                            *reinterpret_cast<uint4 *>(yrow + c + q) = o;
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + acc);
            asm volatile("bar.sync 1, 128;" ::: "memory");                          // sc / bs of this accumulator may be rewritten
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kTmaWarp)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

template <typename T, bool FP8> int launch(const CUtensorMap &map, const GemmParams &p, int ctas, cudaStream_t st) {
    const int smem = kStages * (kStageA + kStageB) + (2 * kStages + 4) * 8 + 16 + 4 * BN * 4 + 1024;
    static unsigned long long configured = 0ull;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 64 || !((configured >> dev) & 1ull)) {
        e = cudaFuncSetAttribute(antq_linear_p4_kernel<T, FP8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        if (dev < 64) configured |= 1ull << dev;
    }
    antq_linear_p4_kernel<T, FP8><<<ctas, kThreads, smem, st>>>(map, p);
    return (int)cudaGetLastError();
}

}  // namespace

extern "C" int antq_linear_p4(const void *x, const uint8_t *w_codes, const float *w_alpha, const void *bias, void *y, int64_t M,
                              int64_t N, int64_t K, int dtype, const void *codebook, const antq_codebook_info *info, int flags,
                              void *stream) {
    if (M < 0 || N < 0 || K < 0 || !info) return ANTQ_EINVAL;
    if (M == 0 || N == 0) return 0;
    if (!x || !w_codes || !w_alpha || !y || !codebook) return ANTQ_EINVAL;
    if (dtype != ANTQ_F16 && dtype != ANTQ_BF16) return ANTQ_ENOTSUP;
    if ((flags & ANTQ_FLAG_OVP) && info->n_entries > info->n_normal) return ANTQ_ENOTSUP;      // pair bytes: decode first
    if (info->n_entries > 16 || K % BK || N % BN || K == 0 || M > 0x7fffffffLL / 2 || N > 0x7fffffffLL / 2) return ANTQ_ENOTSUP;
    if ((uintptr_t)x % 16 || (uintptr_t)y % 16 || (uintptr_t)w_codes % 16) return ANTQ_EALIGN;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return ANTQ_ENOTSUP;
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&map, dtype == ANTQ_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                           const_cast<void *>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ANTQ_EINVAL;
    GemmParams p;
    p.x_alpha = nullptr; p.x_cb = nullptr;
    p.codes = w_codes; p.alpha = w_alpha; p.bias = bias; p.y = y; p.cb = (const AntqCodebook *)codebook;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.m_tiles = (int)((M + BM - 1) / BM);
    p.n_tiles = (int)(N / BN);
    const long long tiles = (long long)p.m_tiles * p.n_tiles;
    const int sms = antq_num_sms();
    const int ctas = (int)(tiles < sms ? tiles : sms);
    return dtype == ANTQ_F16 ? launch<__half, false>(map, p, ctas, (cudaStream_t)stream)
                             : launch<__nv_bfloat16, false>(map, p, ctas, (cudaStream_t)stream);
}

// ---- FP8 variant: both operands as e4m3 LEVELS (integers in units of each grid's smallest level) ----
namespace {
template <typename T> __global__ void antq_levels_e4m3_kernel(const T *__restrict__ xq, unsigned char *__restrict__ out, const float *__restrict__ alpha,
                                                              const AntqCodebook *__restrict__ cb, long long n) {
    typedef AntqType<T> A;
    // x_q = RN_T(fl32(k c s)) -> k = rint(x_q / (c s)): |k| <= 448 and T has >= 8 significant bits, so rint is exact
    const float kx = __fdiv_rn(1.0f, __fmul_rn(__fdiv_rn(alpha[0], cb->gmax), cb->pu_c));
    const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i0 >= n) return;
    if (i0 + 16 <= n && ((uintptr_t)(xq + i0) % 16 == 0) && ((uintptr_t)(out + i0) % 16 == 0) && sizeof(T) == 2) {
        T v[16];
        reinterpret_cast<uint4 *>(v)[0] = antq_ldg_stream(reinterpret_cast<const uint4 *>(xq + i0));
        reinterpret_cast<uint4 *>(v)[1] = antq_ldg_stream(reinterpret_cast<const uint4 *>(xq + i0) + 1);
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int e = 0; e < 16; e++) w[e >> 2] |= e4m3_bits(rintf(__fmul_rn(A::to_f32(v[e]), kx))) << (8 * (e & 3));
        *reinterpret_cast<uint4 *>(out + i0) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
        for (long long i = i0; i < n && i < i0 + 16; i++) out[i] = (unsigned char)e4m3_bits(rintf(__fmul_rn(A::to_f32(xq[i]), kx)));
    }
}
}  // namespace

extern "C" int antq_levels_e4m3(const void *x_q, uint8_t *levels, const float *alpha, int64_t n, int dtype, const void *codebook,
                                const antq_codebook_info *info, void *stream) {
    if (n < 0 || !info) return ANTQ_EINVAL;
    if (n == 0) return 0;
    if (!x_q || !levels || !alpha || !codebook) return ANTQ_EINVAL;
    if (!(info->flags & ANTQ_CB_PU_E4M3)) return ANTQ_ENOTSUP;
    const unsigned ctas = (unsigned)((n / 16 + 256) / 256);
    const AntqCodebook *cb = (const AntqCodebook *)codebook;
    switch (dtype) {
        case ANTQ_F16: antq_levels_e4m3_kernel<__half><<<ctas, 256, 0, (cudaStream_t)stream>>>((const __half *)x_q, levels, alpha, cb, n); break;
        case ANTQ_BF16: antq_levels_e4m3_kernel<__nv_bfloat16><<<ctas, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)x_q, levels, alpha, cb, n); break;
        case ANTQ_F32: antq_levels_e4m3_kernel<float><<<ctas, 256, 0, (cudaStream_t)stream>>>((const float *)x_q, levels, alpha, cb, n); break;
        default: return ANTQ_EINVAL;
    }
    return (int)cudaGetLastError();
}

extern "C" int antq_linear_p4_fp8(const uint8_t *x_levels, const float *x_alpha, const void *x_codebook, const antq_codebook_info *x_info,
                                  const uint8_t *w_codes, const float *w_alpha, const void *bias, void *y, int64_t M, int64_t N, int64_t K,
                                  int out_dtype, const void *w_codebook, const antq_codebook_info *w_info, int flags, void *stream) {
    if (M < 0 || N < 0 || K < 0 || !w_info || !x_info) return ANTQ_EINVAL;
    if (M == 0 || N == 0) return 0;
    if (!x_levels || !x_alpha || !x_codebook || !w_codes || !w_alpha || !y || !w_codebook) return ANTQ_EINVAL;
    if (out_dtype != ANTQ_F16 && out_dtype != ANTQ_BF16) return ANTQ_ENOTSUP;
    if ((flags & ANTQ_FLAG_OVP) && w_info->n_entries > w_info->n_normal) return ANTQ_ENOTSUP;
    if (!(w_info->flags & ANTQ_CB_PU_E4M3) || !(x_info->flags & ANTQ_CB_PU_E4M3)) return ANTQ_ENOTSUP;   // levels must be exact in e4m3
    if (w_info->n_entries > 16 || K % (2 * BK) || N % BN || K == 0 || M > 0x7fffffffLL / 2 || N > 0x7fffffffLL / 2) return ANTQ_ENOTSUP;
    if ((uintptr_t)x_levels % 16 || (uintptr_t)y % 16 || (uintptr_t)w_codes % 16) return ANTQ_EALIGN;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return ANTQ_ENOTSUP;
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)K};
    const cuuint32_t box[2] = {(cuuint32_t)(2 * BK), (cuuint32_t)BM};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t *>(x_levels), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ANTQ_EINVAL;
    GemmParams p;
    p.codes = w_codes; p.alpha = w_alpha; p.bias = bias; p.y = y; p.cb = (const AntqCodebook *)w_codebook;
    p.x_alpha = x_alpha; p.x_cb = (const AntqCodebook *)x_codebook;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.m_tiles = (int)((M + BM - 1) / BM);
    p.n_tiles = (int)(N / BN);
    const long long tiles = (long long)p.m_tiles * p.n_tiles;
    const int sms = antq_num_sms();
    const int ctas = (int)(tiles < sms ? tiles : sms);
    return out_dtype == ANTQ_F16 ? launch<__half, true>(map, p, ctas, (cudaStream_t)stream)
                                 : launch<__nv_bfloat16, true>(map, p, ctas, (cudaStream_t)stream);
}
