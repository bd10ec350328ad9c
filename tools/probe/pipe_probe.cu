// Per-instruction issue rate of the packed-half ops the fake-quant chains are made of (B200, sm_100a).
// One CTA per SM is enough: 4 warps per scheduler, 8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu && ./pipe_probe
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CH 8
#define ITERS 2048

template <int OP> __device__ __forceinline__ void step(unsigned (&a)[CH], unsigned b, unsigned c) {
#pragma unroll
    for (int i = 0; i < CH; i++) {
        unsigned r;
        if (OP == 0) asm volatile("lop3.b32 %0, %1, %2, %3, 0x78;" : "=r"(r) : "r"(a[i]), "r"(b), "r"(c));
        if (OP == 1) asm volatile("set.ge.u32.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a[i]), "r"(b));
        if (OP == 2) asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a[i]), "r"(b), "r"(c));
        if (OP == 3) asm volatile("fma.rn.sat.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a[i]), "r"(b), "r"(c));
        if (OP == 4) asm volatile("max.NaN.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a[i]), "r"(b));
        if (OP == 5) asm volatile("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a[i]), "r"(b));
        if (OP == 6) asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a[i]), "r"(b), "r"(c));
        if (OP == 7) asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(a[i]), "r"(b));
        if (OP == 8) asm volatile("set.ge.f16x2.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a[i]), "r"(b));
        if (OP == 9) asm volatile("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a[i]), "r"(b));
        if (OP == 10) asm volatile("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a[i]), "r"(b), "r"(c));
        if (OP == 11) asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a[i]), "r"(b));
        if (OP == 12) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a[i]), "r"(b), "r"(c));
        if (OP == 13) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=r"(r) : "r"(a[i]), "r"(b), "r"(c));
        if (OP == 14) asm volatile("abs.f16x2 %0, %1;" : "=r"(r) : "r"(a[i]));
        if (OP == 15) asm volatile("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a[i]), "r"(b));
        a[i] = r;
    }
}
// mixes: A then B per chain (one of each per "threshold")
template <int OPA, int OPB> __device__ __forceinline__ void step2(unsigned (&a)[CH], unsigned (&q)[CH], unsigned b, unsigned c) {
#pragma unroll
    for (int i = 0; i < CH; i++) {
        unsigned m;
        if (OPA == 1) asm volatile("set.ge.u32.f16x2 %0, %1, %2;" : "=r"(m) : "r"(a[i]), "r"(b));
        if (OPA == 3) asm volatile("fma.rn.sat.f16x2 %0, %1, %2, %3;" : "=r"(m) : "r"(a[i]), "r"(b), "r"(c));
        if (OPB == 0) asm volatile("lop3.b32 %0, %1, %2, %0, 0x78;" : "+r"(q[i]) : "r"(m), "r"(c));
        if (OPB == 2) asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(q[i]) : "r"(m), "r"(c));
    }
}

template <int OP> __global__ void k1(unsigned *out, unsigned b, unsigned c, long long *cyc) {
    unsigned a[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) a[i] = threadIdx.x * 7 + i;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) step<OP>(a, b, c);
    long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OPA, int OPB, int HALF> __global__ void k2(unsigned *out, unsigned b, unsigned c, long long *cyc) {
    unsigned a[CH], q[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { a[i] = threadIdx.x * 7 + i; q[i] = i; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (HALF) {   // half of the chains ALU-style, half FMA-style, in the same warp
            unsigned a1[CH / 2], q1[CH / 2], a2[CH / 2], q2[CH / 2];
            for (int i = 0; i < CH / 2; i++) { a1[i] = a[i]; q1[i] = q[i]; a2[i] = a[i + CH / 2]; q2[i] = q[i + CH / 2]; }
#pragma unroll
            for (int i = 0; i < CH / 2; i++) {
                unsigned m, m2;
                asm volatile("set.ge.u32.f16x2 %0, %1, %2;" : "=r"(m) : "r"(a1[i]), "r"(b));
                asm volatile("fma.rn.sat.f16x2 %0, %1, %2, %3;" : "=r"(m2) : "r"(a2[i]), "r"(b), "r"(c));
                asm volatile("lop3.b32 %0, %1, %2, %0, 0x78;" : "+r"(q1[i]) : "r"(m), "r"(c));
                asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(q2[i]) : "r"(m2), "r"(c));
            }
            for (int i = 0; i < CH / 2; i++) { q[i] = q1[i]; q[i + CH / 2] = q2[i]; }
        } else {
            step2<OPA, OPB>(a, q, b, c);
        }
    }
    long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s ^= q[i] ^ a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}


template <int A, int B, int C, int D> __global__ void kmix(unsigned *out, unsigned b, unsigned c, long long *cyc) {
    unsigned a0[CH], a1[CH], a2[CH], a3[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { a0[i] = threadIdx.x * 7 + i; a1[i] = a0[i] + 1; a2[i] = a0[i] + 2; a3[i] = a0[i] + 3; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        step<A>(a0, b, c);
        if (B >= 0) step<(B >= 0 ? B : 0)>(a1, b, c);
        if (C >= 0) step<(C >= 0 ? C : 0)>(a2, b, c);
        if (D >= 0) step<(D >= 0 ? D : 0)>(a3, b, c);
    }
    long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s ^= a0[i] ^ a1[i] ^ a2[i] ^ a3[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename K> void run(const char *name, K kern, int threads, int per_iter) {
    unsigned *out; long long *cyc, h;
    cudaMalloc(&out, 4 * 1024 * 4); cudaMalloc(&cyc, 64);
    kern<<<1, threads>>>(out, 0x3c003c00u, 0x38003800u, cyc);
    kern<<<1, threads>>>(out, 0x3c003c00u, 0x38003800u, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double warps_per_smsp = threads / 128.0;
    const double inst = (double)ITERS * per_iter * warps_per_smsp;      // warp-instructions per scheduler
    printf("%-28s threads %4d  cycles %9lld  warp-inst/cycle/SMSP %.3f  (cycles per warp-inst %.2f)\n", name, threads, h,
           inst / h, h / inst);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int threads : {512}) {
        run("LOP3", k1<0>, threads, CH);
        run("HSET2.BM (set.u32.f16x2)", k1<1>, threads, CH);
        run("HSET2 (set.f16x2.f16x2)", k1<8>, threads, CH);
        run("HFMA2", k1<2>, threads, CH);
        run("HFMA2.SAT", k1<3>, threads, CH);
        run("HFMA2.RELU", k1<10>, threads, CH);
        run("HMNMX2.NAN", k1<4>, threads, CH);
        run("HMNMX2 (min)", k1<15>, threads, CH);
        run("HMUL2", k1<5>, threads, CH);
        run("HADD2", k1<11>, threads, CH);
        run("PRMT", k1<6>, threads, CH);
        run("IADD", k1<7>, threads, CH);
        run("VIMNMX.U16x2", k1<9>, threads, CH);
        run("IMAD", k1<12>, threads, CH);
        run("FFMA", k1<13>, threads, CH);
        run("abs.f16x2", k1<14>, threads, CH);
        run("HSET2.BM + LOP3 (ALU chain)", k2<1, 0, 0>, threads, 2 * CH);
        run("HFMA2.SAT + HFMA2 (FMA chain)", k2<3, 2, 0>, threads, 2 * CH);
        run("half ALU chain, half FMA chain", k2<1, 0, 1>, threads, 2 * CH);
        run("mix LOP3+PRMT", kmix<0, 6, -1, -1>, threads, 2 * CH);
        run("mix LOP3+IADD", kmix<0, 7, -1, -1>, threads, 2 * CH);
        run("mix LOP3+VIMNMX", kmix<0, 9, -1, -1>, threads, 2 * CH);
        run("mix HSET2+HMNMX2", kmix<1, 4, -1, -1>, threads, 2 * CH);
        run("mix HSET2+LOP3", kmix<1, 0, -1, -1>, threads, 2 * CH);
        run("mix HMNMX2+LOP3", kmix<4, 0, -1, -1>, threads, 2 * CH);
        run("mix HFMA2+HMUL2", kmix<2, 5, -1, -1>, threads, 2 * CH);
        run("mix HFMA2+HFMA2.SAT", kmix<2, 3, -1, -1>, threads, 2 * CH);
        run("mix HFMA2.SAT+HFMA2.RELU", kmix<3, 10, -1, -1>, threads, 2 * CH);
        run("mix HFMA2+HADD2", kmix<2, 11, -1, -1>, threads, 2 * CH);
        run("mix HFMA2+LOP3", kmix<2, 0, -1, -1>, threads, 2 * CH);
        run("mix HFMA2+HSET2", kmix<2, 1, -1, -1>, threads, 2 * CH);
        run("mix HFMA2.SAT+HSET2", kmix<3, 1, -1, -1>, threads, 2 * CH);
        run("mix HFMA2+IMAD", kmix<2, 12, -1, -1>, threads, 2 * CH);
        run("mix HFMA2+FFMA", kmix<2, 13, -1, -1>, threads, 2 * CH);
        run("mix LOP3+HSET2+HFMA2", kmix<0, 1, 2, -1>, threads, 3 * CH);
        run("mix LOP3+HSET2+HFMA2+HFMA2.SAT", kmix<0, 1, 2, 3>, threads, 4 * CH);
        run("mix LOP3+PRMT+HFMA2+HFMA2.SAT", kmix<0, 6, 2, 3>, threads, 4 * CH);
        run("mix LOP3+HSET2+HMNMX2+HFMA2", kmix<0, 1, 4, 2>, threads, 4 * CH);
    }
    return 0;
}
