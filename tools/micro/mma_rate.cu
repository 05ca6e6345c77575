// Microbenchmark: legacy mma.sync throughput on sm_100a (TF32 m16n8k8, BF16 m16n8k16).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, long long* cyc) {
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b[2] = {threadIdx.x * 3u, threadIdx.x * 5u};
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 1234.5f) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE> void run(const char* name, int macs, int threads) {
    float* d; long long* c; cudaMalloc(&d, 64); cudaMalloc(&c, 8);
    const int iters = 2048;
    k<MODE><<<148 * 2, threads>>>(d, iters, c);
    k<MODE><<<148 * 2, threads>>>(d, iters, c);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    double mma_per_sm = 8.0 * iters * (threads / 32) * 2;     // 2 CTAs per SM
    printf("%-22s threads/CTA %d: %.3f mma/cycle/SM = %.0f MAC/cycle/SM (%s)\n", name, threads, mma_per_sm / h, mma_per_sm / h * macs,
           cudaGetErrorString(cudaGetLastError()));
}
int main() {
    run<0>("mma.sync tf32 m16n8k8", 16 * 8 * 8, 256);
    run<1>("mma.sync bf16 m16n8k16", 16 * 8 * 16, 256);
    run<0>("mma.sync tf32 m16n8k8", 16 * 8 * 8, 128);
    return 0;
}
