// Microbenchmark: issue rate of FP32 instruction forms on sm_100a (per SM, warp-instr/cycle).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, float a, float b, long long* cyc) {
    float r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 0.001f + i;
    float c0 = a + threadIdx.x, c1 = b + threadIdx.x * 2.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) r[i] = fmaf(r[i], a, b);                 // const-bank operands
            if (MODE == 1) r[i] = fmaf(r[i], c0, c1);               // 3 registers, 2 shared
            if (MODE == 2) r[i] = fmaf(r[i], r[(i + 5) & 15], r[(i + 9) & 15]);   // 3 distinct registers
            if (MODE == 3) r[i] = r[i] + r[(i + 5) & 15];           // FADD 2 distinct registers
            if (MODE == 4) r[i] = r[i] * r[(i + 5) & 15];           // FMUL 2 distinct
            if (MODE == 5) r[i] = r[i] + c0;                        // FADD reg + shared reg
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i];
    if (s == 12345.678f) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE> void run(const char* name, int threads = 512) {
    float* d; long long* c; cudaMalloc(&d, 64); cudaMalloc(&c, 8);
    const int iters = 4096;
    k<MODE><<<148, threads>>>(d, iters, 0.999f, 0.001f, c);
    k<MODE><<<148, threads>>>(d, iters, 0.999f, 0.001f, c);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    double winstr = 16.0 * iters * (threads / 32);    // warp-instr per SM
    printf("%-28s threads %4d cycles %lld  warp-instr/cycle/SM %.2f  (per SMSP %.2f)\n", name, threads, h, winstr / h, winstr / h / 4);
}
int main() {
    run<0>("FFMA r,c[],c[]");
    run<1>("FFMA r,rS,rS");
    run<2>("FFMA r,r,r distinct");
    run<3>("FADD r,r distinct");
    run<4>("FMUL r,r distinct");
    run<5>("FADD r,rS");
    run<3>("FADD r,r distinct", 128);
    run<3>("FADD r,r distinct", 256);
    run<2>("FFMA r,r,r distinct", 128);
    run<1>("FFMA r,rS,rS", 128);
    run<1>("FFMA r,rS,rS", 256);
    return 0;
}
