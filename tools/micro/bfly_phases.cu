// Where does one radix-16 DIF butterfly spend its latency?  clock64() around its phases.
#include <cstdio>
#include "../../vae-teb_b200/csrc/scat_core.cuh"
using namespace tebscat;

__global__ void __launch_bounds__(512, 1) k(long long* out, int logB, int reps) {
    extern __shared__ float2 S[];
    float2* twA = S + 20000; float2* twB = twA + kTwA;
    for (int i = threadIdx.x; i < 20000 + kTwA + kTwB; i += blockDim.x) S[i] = make_float2(0.001f * i, 1.f);
    __syncthreads();
    constexpr int LOGR = 4, R = 16;
    const int u = threadIdx.x;
    long long acc[5] = {0, 0, 0, 0, 0};
    for (int rep = 0; rep < reps; ++rep) {
        const int logs = logB - LOGR;
        const int i0 = u & ((1 << logs) - 1);
        const int blk = u >> logs;
        const int p0 = (blk << logB) + i0;
        long long t0 = clock64();
        float2 wb[LOGR];
        const int k1 = i0 << (kLog2TwMax - logB);
#pragma unroll
        for (int i = 0; i < LOGR; ++i) wb[i] = twiddle(twA, twB, k1 << i);
        int slot[R];
        const int s0 = swz(p0), ds = (1 << logs) + (1 << (logs - 4));
#pragma unroll
        for (int j = 0; j < R; ++j) slot[j] = s0 + j * ds;
        float2 v[R];
#pragma unroll
        for (int j = 0; j < R; ++j) v[j] = S[slot[j]];
        // force completion of the loads
        float sink = 0.f;
#pragma unroll
        for (int j = 0; j < R; ++j) sink += v[j].x;
#pragma unroll
        for (int i = 0; i < LOGR; ++i) sink += wb[i].x;
        if (sink == 1234.5f) out[100] = 1;
        long long t1 = clock64();
        Dft<R, -1>::run(v);
        sink = 0.f;
#pragma unroll
        for (int j = 0; j < R; ++j) sink += v[j].x + v[j].y;
        if (sink == 1234.5f) out[101] = 1;
        long long t2 = clock64();
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int q = qmap<R>(r);
            if (q != 0) v[r] = cmul(v[r], twiddle_power<LOGR>(wb, q));
        }
        sink = 0.f;
#pragma unroll
        for (int j = 0; j < R; ++j) sink += v[j].x + v[j].y;
        if (sink == 1234.5f) out[102] = 1;
        long long t3 = clock64();
#pragma unroll
        for (int r = 0; r < R; ++r) S[slot[brev<LOGR>(qmap<R>(r))]] = v[r];
        __syncwarp();
        long long t4 = clock64();
        acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2; acc[3] += t4 - t3; acc[4] += t4 - t0;
        __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0)
        for (int i = 0; i < 5; ++i) out[i] = acc[i] / reps;
}

int main() {
    long long* d; cudaMalloc(&d, 1024);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int threads : {32, 128, 256, 512}) {
        k<<<148, threads, 200 * 1024>>>(d, 13, 50);
        cudaDeviceSynchronize();
        long long h[5]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("threads %3d: load+twiddle-lookup %5lld  dft16 %5lld  twiddle-apply %5lld  store %5lld  total %5lld cycles (%s)\n",
               threads, h[0], h[1], h[2], h[3], h[4], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
