// Do tcgen05.mma instructions issued by DIFFERENT threads into the SAME tensor-memory accumulator lose updates?
// A = all ones (tensor memory), B = all ones (shared memory), accumulate always: after n issuers x iters MMAs of K = 8
// every element of D must be exactly n * iters * 8 (fp32 holds it exactly).  One CTA per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t b_desc(uint32_t a) {
    uint64_t d = 0;
    d |= (uint64_t)((a >> 4) & 0x3fff); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61;
    return d;
}
__global__ void __launch_bounds__(128, 1) k(int iters, int n_issuers, int stagger, int* bad, float* sample) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) unsigned long long bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 32 * 1024 / 4; i += 128) reinterpret_cast<float*>(raw + (base - smem_u32(raw)))[i] = 1.0f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(n_issuers) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    {   // zero D (columns 0..159), ones in A (columns 256..263) for this warp's lanes
        const uint32_t lanes = tmem + ((uint32_t)(32 * warp) << 16);
        const uint32_t one = __float_as_uint(1.0f);
        for (int j = 0; j < 20; ++j)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(lanes + 8 * j), "r"(0u) : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(lanes + 256), "r"(one) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // issuer 0: N = 160 into columns 0..159; issuers 1..: N = 80 into columns 80..159 (the overlap is the question)
    if (lane == 0 && warp < n_issuers) {
        const uint32_t N = warp == 0 ? 160 : 80;
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t bd = b_desc(base);
        if (stagger && warp) __nanosleep(37 * warp);
        for (int i = 0; i < iters; ++i)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(tmem + (warp == 0 ? 0u : 80u)), "r"(tmem + 256), "l"(bd), "r"(idesc) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        uint32_t done = 0;
        for (uint32_t spin = 0; !done && spin < (1u << 26); ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t lanes = tmem + ((uint32_t)(32 * warp) << 16);
    int wrong = 0;
    for (int j = 0; j < 20; ++j) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(lanes + 8 * j) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const float expect = (j < 10 ? 1.f : (float)n_issuers) * iters * 8.f;
        for (int i = 0; i < 8; ++i) wrong += __uint_as_float(r[i]) != expect;
        if (blockIdx.x == 0 && threadIdx.x == 0 && (j == 0 || j == 10)) sample[j / 10] = __uint_as_float(r[0]);
    }
    if (wrong) atomicAdd(bad, wrong);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
int main() {
    int* bad; float* sample; cudaMalloc(&bad, 4); cudaMalloc(&sample, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int n_issuers = 1; n_issuers <= 3; ++n_issuers)
        for (int stagger = 0; stagger < 2; ++stagger) {
            const int iters = 2048;
            int total_bad = 0; float h[2] = {0, 0};
            for (int rep = 0; rep < 20; ++rep) {
                cudaMemset(bad, 0, 4);
                k<<<148, 128, 64 * 1024>>>(iters, n_issuers, stagger, bad, sample);
                int b = 0; cudaMemcpy(&b, bad, 4, cudaMemcpyDeviceToHost); total_bad += b;
            }
            cudaMemcpy(h, sample, 8, cudaMemcpyDeviceToHost);
            printf("%d issuer(s)%s: D[head] = %.0f (expect %d), D[shared] = %.0f (expect %d), wrong elements over 20 x 148 CTAs: %d  (%s)\n", n_issuers,
                   stagger ? ", staggered" : "", h[0], iters * 8, h[1], n_issuers * iters * 8, total_bad, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
