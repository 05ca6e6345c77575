// Microbenchmark: tcgen05.mma issue-to-completion rate on sm_100a for the shapes the phase kernel could use:
// cta_group::1, M = 128, A from tensor memory, B from shared memory (K-major, 128-byte swizzle), kind::tf32 (K = 8) and
// kind::f16 with bf16 inputs (K = 16), N in {64, 80, 96, 128, 160, 256}.  One CTA per SM, one issuing thread, 4096 MMAs
// back to back into one accumulator (or into 2-4 independent ones in turn), one commit at the end; cycles = clock64 around issue + completion.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t b_desc(uint32_t a) {
    uint64_t d = 0;
    d |= (uint64_t)((a >> 4) & 0x3fff); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61;
    return d;
}
template <int KIND>   // 0: tf32, 1: bf16
__global__ void __launch_bounds__(128, 1) k(int N, int iters, int a_from_smem, int n_acc, int n_issuers, long long* cyc) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) unsigned long long bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float*>(raw + (base - smem_u32(raw)))[i] = 0.001f * (i & 63);
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
    // instruction descriptor: fp32 accumulate, A/B K-major, M = 128, N
    const uint32_t idesc = (1u << 4) | ((KIND ? 1u : 2u) << 7) | ((KIND ? 1u : 2u) << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (lane == 0 && warp < n_issuers) {
        const uint64_t bd = b_desc(base), ad = b_desc(base + 32 * 1024);
        const long long t0 = clock64();
        int acc = 0;
        for (int i = 0; i < iters; ++i) {
            const uint32_t tmem_d = tmem + 80 * (n_issuers > 1 ? warp : acc);   // independent accumulators: round robin, or one per issuer
            acc = acc + 1 == n_acc ? 0 : acc + 1;
            if (a_from_smem) {
                if (KIND == 0) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc), "r"(i) : "memory");
                else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc), "r"(i) : "memory");
            } else {
                if (KIND == 0) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                             ::"r"(tmem_d), "r"(tmem + 384), "l"(bd), "r"(idesc), "r"(i) : "memory");
                else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                             ::"r"(tmem_d), "r"(tmem + 384), "l"(bd), "r"(idesc), "r"(i) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = warp != 0;
        for (uint32_t spin = 0; !done && spin < (1u << 26); ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        const long long t1 = clock64();
        if (blockIdx.x == 0 && warp == 0) cyc[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
template <int KIND> void run(int N, int a_smem, int n_acc = 1, int n_issuers = 1) {
    long long* c; cudaMalloc(&c, 8);
    const int iters = 4096;
    cudaFuncSetAttribute(k<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int rep = 0; rep < 2; ++rep) k<KIND><<<148, 128, 100 * 1024>>>(N, iters, a_smem, n_acc, n_issuers, c);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / iters / n_issuers, macs = 128.0 * N * (KIND ? 16 : 8);
    printf("%s M=128 N=%3d K=%2d A from %s, %d accumulator(s) in turn: %6.1f cycles per MMA = %5.0f MAC/cycle/SM  (%s)\n", KIND ? "bf16" : "tf32", N, KIND ? 16 : 8,
           a_smem ? "smem" : "TMEM", n_issuers > 1 ? n_issuers : n_acc, per, macs / per, cudaGetErrorString(e));
    cudaFree(c);
}
int main() {
    for (int a = 0; a < 2; ++a) {
        for (int N : {64, 80, 96, 128, 160, 256}) run<0>(N, a);
        for (int N : {64, 80, 96, 128, 160, 256}) run<1>(N, a);
    }
    for (int n_acc : {2, 3, 4}) { run<0>(80, 0, n_acc); run<1>(80, 0, n_acc); }
    printf("several issuing threads (one warp each), one accumulator per issuer:\n");
    for (int n_is : {2, 3}) { run<0>(80, 0, 1, n_is); run<0>(160, 0, 1, n_is); }
    return 0;
}
