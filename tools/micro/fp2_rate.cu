// Microbenchmark: issue rate of the packed FP32 forms (add/mul/fma.f32x2, sm_100+) against the scalar ones.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float x, float y) {
    unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, float a, float b, long long* cyc) {
    float r[16];
    unsigned long long p[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 0.001f + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = pk(r[2 * i], r[2 * i + 1]);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) r[i] = r[i] + r[(i + 5) & 15];                                  // 16 FADD
            if (MODE == 1) r[i] = fmaf(r[i], r[(i + 5) & 15], r[(i + 9) & 15]);             // 16 FFMA 3-reg
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 2) p[i] = add2(p[i], p[(i + 3) & 7]);                               // 8 FADD2 (= 16 adds)
            if (MODE == 3) p[i] = fma2(p[i], p[(i + 3) & 7], p[(i + 5) & 7]);               // 8 FFMA2
            if (MODE == 4) p[i] = mul2(p[i], p[(i + 3) & 7]);
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += (float)(p[i] & 0xffff);
    if (s == 12345.678f) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE> void run(const char* name, double flops_per_iter, int threads = 512) {
    float* d; long long* c; cudaMalloc(&d, 64); cudaMalloc(&c, 8);
    const int iters = 4096;
    k<MODE><<<148, threads>>>(d, iters, 0.999f, 0.001f, c);
    k<MODE><<<148, threads>>>(d, iters, 0.999f, 0.001f, c);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-24s threads %4d cycles %8lld  lane-ops/cycle/SM %.1f (%s)\n", name, threads, h,
           flops_per_iter * iters * threads / h, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    run<0>("FADD x16", 16); run<1>("FFMA 3-reg x16", 16);
    run<2>("FADD2 x8", 16); run<3>("FFMA2 x8", 16); run<4>("FMUL2 x8", 16);
    run<0>("FADD x16", 16, 128); run<2>("FADD2 x8", 16, 128); run<3>("FFMA2 x8", 16, 128);
    return 0;
}
