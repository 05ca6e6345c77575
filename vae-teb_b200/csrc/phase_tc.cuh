// Phase stage B, dense form, on the 5th-generation tensor cores (tcgen05 + TMEM).
//
//   out[row, n] = Re sum_t c_row(t) G[t, n],   c(t) = |z_i| exp(i p theta_i) conj(z_j)
//   (hdf5_dataset/kymatio_phase_scattering.py:211-218, :283 / :339, and _apply_phi_filter :233-273 as the
//    precomputed operator G)
//
// is the real GEMM  A' (rows x 2N) . B' (2N x n_out)  with A'[row, 2t] = Re c(t), A'[row, 2t+1] = -Im c(t) and
// B'[2t, n] = Re G[t, n], B'[2t+1, n] = Im G[t, n], in 3xTF32 (fp32-class accuracy): A' = Ah + Al, B' = Bh + Bl,
// out ~= Ah Bh + Al Bh + Ah Bl.
//
// One CTA = 128 rows x 80 columns, one CTA per SM, 25 warps:
//   * 16 PRODUCER warps compute the rows' products where they are consumed -- a thread owns one row (= one TMEM
//     lane) and four time samples of a slab of 16 -- split them into TF32 head and tail and write them straight into
//     TENSOR MEMORY with tcgen05.st: the A' operand never touches shared memory (tcgen05.mma with A from TMEM).
//     They also copy the slab of B' (pre-split on the host side of the plan, K-major, 128-byte swizzle) into shared
//     memory;
//   * ONE thread of the MMA warp issues tcgen05.mma.kind::tf32 (M = 128, N = 80, K = 8): 12 per slab, accumulating in
//     TMEM; tcgen05.commit releases the slab's stage and hands finished accumulators to
//   * 8 EPILOGUE warps, which drain an accumulator every kTcDrain slabs with tcgen05.ld and add it to running sums
//     in fp32 registers (a tensor-core accumulator that runs over all K = 2N = 9600 would carry its truncation bias,
//     see the mma.sync kernel), double-buffered so the drain overlaps the next group's MMAs.
// mbarriers: full[s] (producers -> MMA), empty[s] (MMA -> producers), acc_full[a] (MMA -> epilogue),
// acc_empty[a] (epilogue -> MMA).  TMEM (512 columns): accumulators at 0 and 128, A' stages from 256.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tebscat {

constexpr int kTcRows = 128;          // rows per CTA = TMEM lanes
constexpr int kTcCols = 80;           // output columns per CTA (UMMA N)
constexpr int kTcSlabT = 16;          // time samples per slab
constexpr int kTcK = 2 * kTcSlabT;    // K per slab (Re, -Im interleaved)
constexpr int kTcStages = 4;
constexpr int kTcDrain = 2;           // slabs per accumulator drain
constexpr int kTcEpiWarps = 8, kTcProdWarps = 16;
constexpr int kTcMmaWarp = kTcEpiWarps;                          // warp 8
constexpr int kTcThreads = 32 * (kTcEpiWarps + 1 + kTcProdWarps);   // 800
constexpr int kTcBTile = kTcCols * 128;                         // bytes of one B' tile (80 rows x 32 tf32)
constexpr int kTcStageBytes = 2 * kTcBTile;                      // head + tail
constexpr size_t kTcSmem = 1024 + (size_t)kTcStages * kTcStageBytes + 256;
constexpr uint32_t kTcAcc0 = 0, kTcAcc1 = 128, kTcA0 = 256;      // TMEM columns; A' stage s: head at kTcA0 + 64 s, tail + 32

struct PairTcParams {
    const float2* zp;
    const float2* zc;
    const float* Bs;            // [2 (head, tail)][n_cols_pad][k_pad] TF32 values in fp32 containers, k_pad = 32 n_slabs
    const int32_t* i_idx;
    const int32_t* j_idx;
    const float* powers;
    const int32_t* subset;
    float* out;
    long long rows;
    int32_t n_sel, F, N, n_out, n_cols_pad, n_slabs, k_pad;
};

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tc_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a protocol error traps instead of hanging the GPU
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && spin > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem descriptor]
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128-byte swizzle: rows of 128 bytes, groups of 8 rows 1024 bytes apart (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t tc_b_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset
    d |= (uint64_t)1 << 46;                    // descriptor version
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 80
constexpr uint32_t kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcCols >> 3) << 17) | ((uint32_t)(kTcRows >> 4) << 24);

__device__ __forceinline__ void tc_split(float v, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
    const float r = v - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}

__global__ void __launch_bounds__(kTcThreads, 1) phase_pair_tc_kernel(const PairTcParams p) {
    extern __shared__ uint8_t tc_raw[];
    // 1024-byte alignment of the swizzled tiles
    const uint32_t raw = tc_smem_u32(tc_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = tc_raw + (base - raw);
    const uint32_t bars = base + kTcStages * kTcStageBytes;          // full[4], empty[4], acc_full[2], acc_empty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kTcStages * kTcStageBytes + 128);
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (kTcStages + s); };
    auto acc_full = [&](int a) { return bars + 8u * (2 * kTcStages + a); };
    auto acc_empty = [&](int a) { return bars + 8u * (2 * kTcStages + 2 + a); };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kTcStages; ++s) {
            tc_mbar_init(full(s), kTcProdWarps);
            tc_mbar_init(empty(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            tc_mbar_init(acc_full(a), 1);
            tc_mbar_init(acc_empty(a), kTcEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kTcMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tc_smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const long long row0 = (long long)blockIdx.x * kTcRows;
    const int col0 = blockIdx.y * kTcCols;
    const int n_slabs = p.n_slabs;
    const int n_groups = (n_slabs + kTcDrain - 1) / kTcDrain;

    if (warp < kTcEpiWarps) {
        // ===== epilogue: lanes 32 q .. 32 q + 31, columns 40 h .. 40 h + 39 =====
        const int q = warp & 3, h = warp >> 2;
        float total[40];
#pragma unroll
        for (int i = 0; i < 40; ++i) total[i] = 0.f;
        for (int g = 0; g < n_groups; ++g) {
            const int a = g & 1;
            tc_mbar_wait(acc_full(a), (g >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (a ? kTcAcc1 : kTcAcc0) + 40 * h;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                float v[8];
                tc_ld8(taddr + 8 * j, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) total[8 * j + i] += v[i];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(acc_empty(a));
        }
        const long long row = row0 + 32 * q + lane;
        if (row < p.rows) {
#pragma unroll
            for (int i = 0; i < 40; ++i) {
                const int col = col0 + 40 * h + i;
                if (col < p.n_out) p.out[row * p.n_out + col] = total[i];
            }
        }
    } else if (warp == kTcMmaWarp) {
        // ===== MMA issuer: one thread =====
        if (lane == 0) {
            for (int i = 0; i < n_slabs; ++i) {
                const int s = i % kTcStages, g = i / kTcDrain, a = g & 1;
                const bool first = (i % kTcDrain) == 0;
                if (first) {
                    tc_mbar_wait(acc_empty(a), ((g >> 1) & 1) ^ 1);          // passes at once for the first two groups
                    tc_fence_after();
                }
                tc_mbar_wait(full(s), (i / kTcStages) & 1);
                tc_fence_after();
                const uint32_t d = tmem + (a ? kTcAcc1 : kTcAcc0);
                const uint32_t a_hi = tmem + kTcA0 + 64 * s, a_lo = a_hi + 32;
                const uint32_t b_hi = base + s * kTcStageBytes, b_lo = b_hi + kTcBTile;
#pragma unroll
                for (int k = 0; k < kTcK / 8; ++k) {
                    const uint64_t dh = tc_b_desc(b_hi + 32 * k), dl = tc_b_desc(b_lo + 32 * k);
                    tc_mma_ts(d, a_hi + 8 * k, dh, kTcIdesc, (first && k == 0) ? 0u : 1u);
                    tc_mma_ts(d, a_lo + 8 * k, dh, kTcIdesc, 1u);
                    tc_mma_ts(d, a_hi + 8 * k, dl, kTcIdesc, 1u);
                }
                tc_commit(empty(s));                                          // the stage is free once these MMAs retire
                if ((i % kTcDrain) == kTcDrain - 1 || i == n_slabs - 1) tc_commit(acc_full(a));
            }
        }
        __syncwarp();
    } else {
        // ===== producers: row 32 q + lane, samples 4 sub .. 4 sub + 3 of every slab; B' tiles =====
        const int pw_ = warp - (kTcMmaWarp + 1);
        const int q = warp & 3, sub = pw_ >> 2;
        const int ptid = pw_ * 32 + lane;
        const long long row = row0 + 32 * q + lane;
        long long zp_off = -1, zc_off = 0;
        float pw = 1.f;
        if (row < p.rows) {
            const long long b = row / p.n_sel;
            const int sidx = (int)(row - b * p.n_sel);
            const int pair = p.subset ? p.subset[sidx] : sidx;
            zp_off = (b * p.F + p.i_idx[pair]) * (long long)p.N;
            zc_off = (b * p.F + p.j_idx[pair]) * (long long)p.N;
            pw = p.powers[pair];
        }
        float2 rzp[4], rzc[4];
        auto fetch = [&](int i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int t = i * kTcSlabT + 4 * sub + j;
                const bool in = zp_off >= 0 && t < p.N;
                rzp[j] = in ? __ldg(p.zp + zp_off + t) : make_float2(0.f, 0.f);
                rzc[j] = in ? __ldg(p.zc + zc_off + t) : make_float2(0.f, 0.f);
            }
        };
        fetch(0);
        const float* Bh = p.Bs + (size_t)col0 * p.k_pad;
        const float* Bl = p.Bs + ((size_t)p.n_cols_pad + col0) * p.k_pad;
        for (int i = 0; i < n_slabs; ++i) {
            const int s = i % kTcStages;
            tc_mbar_wait(empty(s), ((i / kTcStages) & 1) ^ 1);
            tc_fence_after();
            // B' slab: 2 x 80 rows x 8 chunks of 16 bytes, swizzled
            float4 bv[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int idx = ptid + 512 * j;
                if (idx < 2 * kTcCols * 8) {
                    const int part = idx >= kTcCols * 8, rem = idx - part * kTcCols * 8;
                    const int n = rem >> 3, c = rem & 7;
                    bv[j] = __ldg(reinterpret_cast<const float4*>((part ? Bl : Bh) + (size_t)n * p.k_pad + i * kTcK + 4 * c));
                }
            }
            // A': the products of this thread's four samples, split, straight into tensor memory
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 c = accelerated_product(rzp[j], rzc[j], pw);
                tc_split(c.x, hi[2 * j], lo[2 * j]);
                tc_split(-c.y, hi[2 * j + 1], lo[2 * j + 1]);
            }
            if (i + 1 < n_slabs) fetch(i + 1);
            const uint32_t ta = tmem + ((uint32_t)(32 * q) << 16) + kTcA0 + 64 * s + 8 * sub;
            tc_st8(ta, hi);
            tc_st8(ta + 32, lo);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int idx = ptid + 512 * j;
                if (idx < 2 * kTcCols * 8) {
                    const int part = idx >= kTcCols * 8, rem = idx - part * kTcCols * 8;
                    const int n = rem >> 3, c = rem & 7;
                    uint8_t* dst = sm + s * kTcStageBytes + part * kTcBTile + (n >> 3) * 1024 + (n & 7) * 128 + ((c ^ (n & 7)) << 4);
                    *reinterpret_cast<float4*>(dst) = bv[j];
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes of B' -> the MMA's async proxy
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(full(s));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kTcMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

}  // namespace tebscat
