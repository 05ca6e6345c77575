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
// One CTA = 128 rows x 80 columns, one CTA per SM, 27 warps:
//   * 16 PRODUCER warps in four groups compute the rows' products where they are consumed -- group j owns the slabs
//     i = j (mod 4) and the stage j, a thread one row (= one TMEM lane) of such a slab of 16 samples -- split them
//     into TF32 head and tail and write them straight into TENSOR MEMORY with tcgen05.st: the A' operand never touches
//     shared memory (tcgen05.mma with A from TMEM).  The slab's inputs come through shared memory, every DISTINCT
//     128-byte line once (cp.async); its slab of B' (pre-split on the host side of the plan and stored as the
//     shared-memory image of the stage: K-major, 128-byte swizzle) with ONE bulk copy on the slab's barrier;
//   * THREE MMA-issuing threads (tcgen05.mma.kind::tf32, M = 128, N = 80, K = 8): one thread issues an MMA of this
//     size every ~98 cycles whatever N <= 160 is, threads issue independently (tools/micro/umma_rate.cu), and a
//     single issuer was the limit of the whole kernel (tools/tc_trace.py).  Issuers 0 / 1: Ah Bh of the drain
//     groups of their parity into their own head accumulator; issuer 2: Al Bh + Ah Bl of every slab into the
//     correction accumulator.  One writer per accumulator: the output does not depend on how the threads interleave;
//   * 8 EPILOGUE warps (two per TMEM lane quarter, 40 columns each), which drain a head accumulator every kTcDrain
//     slabs with tcgen05.ld, add it to running sums in fp32 registers (a tensor-core accumulator that runs over all
//     K = 2N = 9600 would carry its truncation bias, see the mma.sync kernel), double-buffered so
//     the drain overlaps the next group's MMAs.  The correction products are 2^-11 of the result, so their own
//     truncation error is irrelevant: their accumulator is read once at the end.
// mbarriers: full[s] (producers + bulk copy -> issuers), empty[s] (issuers -> producers), acc_full[a] (issuer a ->
// epilogue), acc_empty[a] (epilogue -> issuer a), tail_done (issuer 2 -> epilogue).  TMEM (512 columns, all in use):
// head accumulators at 0 and 160, the correction accumulator at 80, A' stages (head 32 + tail 32 columns) from 256.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tebscat {

constexpr int kTcRows = 128;          // rows per CTA = TMEM lanes
constexpr int kTcCols = 80;           // output columns per CTA (UMMA N)
constexpr int kTcSlabT = 16;          // time samples per slab
constexpr int kTcK = 2 * kTcSlabT;    // K per slab (Re, -Im interleaved)
constexpr int kTcStages = 4;           // one per producer group (a group that shared stages with others could run two barrier phases ahead)
constexpr int kTcDrain = 2;           // slabs per drain of the head-product accumulator: 8 tensor-core accumulations between
                                      // fp32 adds (the TMEM accumulator truncates like the register one: the error grows with
                                      // the number of accumulations into one running sum)
constexpr int kTcEpiWarps = 8, kTcProdWarps = 16, kTcGroups = 4;   // producer groups of four warps (one per TMEM lane quarter)
constexpr int kTcEpiCols = kTcCols / (kTcEpiWarps / 4);            // columns per epilogue warp
constexpr int kTcMmaWarp = kTcEpiWarps, kTcMmaWarps = 3;         // warps 8..10: one issuing thread each
constexpr int kTcThreads = 32 * (kTcEpiWarps + kTcMmaWarps + kTcProdWarps);   // 864: 72 registers per thread
constexpr int kTcBTile = kTcCols * 128;                         // bytes of one B' tile (80 rows x 32 tf32)
constexpr int kTcStageBytes = 2 * kTcBTile;                      // head + tail
constexpr int kTcInBytes = 2 * kTcRows * 128;                    // one group's input slab: 128 lines of (|z|, theta) + 128 of (re, im)
constexpr int kTcOffBars = kTcStages * kTcStageBytes + kTcGroups * kTcInBytes;
constexpr size_t kTcSmem = 1024 + (size_t)kTcOffBars + 256 + 4 * kTcRows * sizeof(int32_t) + 64;
// TMEM columns: head accumulator 0, the accumulator of the correction products (Al Bh + Ah Bl: 2^-11 of the head
// products, so its truncation error stays below 1e-7 of the result even over the whole contraction -- it is read
// once, at the end), head accumulator 1, and the A' stages (head at kTcA0 + 64 s, tail + 32)
constexpr uint32_t kTcAcc0 = 0, kTcAccS = 80, kTcAcc1 = 160, kTcA0 = 256;

#ifdef TEBSCAT_TC_TRACE
// timeline of one CTA (the last of the first wave: SM-resident alone with its wave), first kTcTraceSlabs slabs:
// [role 0..3 producer group g | 4 MMA | 5 epilogue][slab or group][event]
constexpr int kTcTraceSlabs = 64, kTcTraceEv = 6;
__device__ long long g_tc_trace[6 * kTcTraceSlabs * kTcTraceEv];
#define TC_TRACE(role, idx, ev) do { if (blockIdx.x == 1 && lane == 0 && (idx) < kTcTraceSlabs) \
    g_tc_trace[((role) * kTcTraceSlabs + (idx)) * kTcTraceEv + (ev)] = clock64(); } while (0)
#else
#define TC_TRACE(role, idx, ev) do {} while (0)
#endif

struct PairTcParams {
    const float2* zp;
    const float2* zc;
    const float* Bimg;          // [column tile][slab][kTcStageBytes]: the shared-memory image of the slab's B' stage (head, tail;
                                // TF32 values in fp32 containers, K-major, 128-byte swizzle), copied with one bulk copy
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
// one poll, no loop: the MMA issuers ask for the NEXT slab before they issue the current one, so that the ~250 cycles a
// poll takes (even of a barrier that is complete) pass behind the MMAs instead of in front of them
__device__ __forceinline__ uint32_t tc_mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done;
}
// the same for warps that wait long (epilogue, producers ahead of the tensor core): back off between polls so that
// the polling does not take issue slots from the producers
__device__ __forceinline__ void tc_mbar_wait_relaxed(uint32_t bar, uint32_t parity, unsigned ns) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"      // suspends up to the time hint
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(ns)
            : "memory");
        if (!done && spin > (1u << 22)) __trap();
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
__device__ __forceinline__ void tc_ld8_nowait(uint32_t taddr, float (&v)[8]) {     // the caller issues tcgen05.wait::ld
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "r"(taddr)
                 : "memory");
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

// TF32 head and tail of v: head = the top 19 bits of v (what the tensor core reads of an fp32 container anyway), tail =
// v - head (exact, |tail| < 2^-10 |v|), of which the tensor core again reads the top 19 bits: the dropped tail bits are
// below 2^-23 |v| and the product of the two tails (the term 3xTF32 leaves out) below 2^-21 |v b| -- two instructions
// per value on the ALU / FMA pipes (cvt.rna.tf32.f32 would occupy the 16-lane conversion unit that the sine and cosine
// of every product already use).  Rounding the head instead (one more integer add) was tried with the separate
// correction accumulator in place: same error (worst err / bound 0.69 both ways), 3 % slower.  B' is split with
// rounding on the host.
__device__ __forceinline__ void tc_split(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}
// the same for the two halves of a product (re, -im): one packed subtraction for both tails
__device__ __forceinline__ void tc_split2(float2 c, uint32_t& hi0, uint32_t& hi1, uint32_t& lo0, uint32_t& lo1) {
    hi0 = __float_as_uint(c.x) & 0xffffe000u;
    hi1 = (__float_as_uint(c.y) & 0xffffe000u) ^ 0x80000000u;               // head of -c.y
    const float2 t = csub(make_float2(c.x, -c.y), make_float2(__uint_as_float(hi0), __uint_as_float(hi1)));
    lo0 = __float_as_uint(t.x);
    lo1 = __float_as_uint(t.y);
}

__global__ void __launch_bounds__(kTcThreads, 1) phase_pair_tc_kernel(const PairTcParams p) {
    extern __shared__ uint8_t tc_raw[];
    // 1024-byte alignment of the swizzled tiles
    const uint32_t raw = tc_smem_u32(tc_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = tc_raw + (base - raw);
    const uint32_t bars = base + kTcOffBars;                         // full[4], empty[4], acc_full[2], acc_empty[2], tail_done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kTcOffBars + 128);
    int32_t* s_off = reinterpret_cast<int32_t*>(sm + kTcOffBars + 256);       // [2][128]: first sample of the rows' inputs (float2 units)
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (kTcStages + s); };
    auto acc_full = [&](int a) { return bars + 8u * (2 * kTcStages + a); };
    auto acc_empty = [&](int a) { return bars + 8u * (2 * kTcStages + 2 + a); };
    const uint32_t tail_done = bars + 8u * (2 * kTcStages + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kTcStages; ++s) {
            tc_mbar_init(full(s), kTcProdWarps / kTcGroups);
            tc_mbar_init(empty(s), 2);                                  // one issuer of each kind
        }
        for (int a = 0; a < 2; ++a) {
            tc_mbar_init(acc_full(a), 1);
            tc_mbar_init(acc_empty(a), kTcEpiWarps);
        }
        tc_mbar_init(tail_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // Input lines.  A row (pair) reads the (|z|, theta) line of its 'i' filter and the (re, im) line of its 'j' filter;
    // the 128 rows of a tile are consecutive pairs of one or two samples, so they share a handful of 'i' lines and at
    // most F 'j' lines per sample.  Each DISTINCT line is copied once per slab (s_line: its first sample in the
    // workspace, float2 units, or -1), and a row remembers the slot of its two lines (s_slot):
    //   side 0 ('i'): run-length slots -- the pair list is sorted by i, so equal keys are adjacent;
    //   side 1 ('j'): slot = (sample - first sample of the tile) * F + j while that fits 128 slots, else one per row.
    // At the headline configuration (741 pairs, F = 38) that is 4-10 + 38-76 lines instead of 256.
    int32_t* s_slot = s_off + 2 * kTcRows;                                      // [2][128]
    int32_t* s_cnt = s_off + 4 * kTcRows;                                       // [0..1] lines per side, [2..5] scratch
    const long long tile_row0 = (long long)blockIdx.x * kTcRows;
    const long long tile_last = (tile_row0 + kTcRows - 1 < p.rows ? tile_row0 + kTcRows - 1 : p.rows - 1);
    const long long b_first = tile_row0 / p.n_sel;
    const int nb_tile = (int)(tile_last / p.n_sel - b_first) + 1;
    const bool direct_j = nb_tile * p.F <= kTcRows;
    int32_t key_i = 0, key_j = 0, slot_j = 0;
    if (tid < kTcRows) {
        // rows beyond the batch (last CTA) read the lines of the tile's first row: their products are computed and never stored
        const long long row = tile_row0 + tid < p.rows ? tile_row0 + tid : tile_row0;
        const long long b = row / p.n_sel;
        const int sidx = (int)(row - b * p.n_sel);
        const int pair = p.subset ? p.subset[sidx] : sidx;
        key_i = (int32_t)(b * p.F + p.i_idx[pair]);                             // (key * N) < 2^31: the workspace chunk is bounded
        key_j = (int32_t)(b * p.F + p.j_idx[pair]);
        slot_j = direct_j ? (int32_t)((b - b_first) * p.F + p.j_idx[pair]) : tid;
        s_off[tid] = key_i;                                                     // neighbours compare keys
        s_off[kTcRows + tid] = -1;
        if (tid == 0) { s_cnt[6] = kTcRows; s_cnt[7] = 0; }
    }
    __syncthreads();
    if (tid < kTcRows) {
        const bool head = tid == 0 || s_off[tid - 1] != key_i;
        const unsigned m = __ballot_sync(0xffffffffu, head);
        if (lane == 0) s_cnt[2 + warp] = __popc(m);
        s_slot[kTcRows + tid] = slot_j;
        atomicMin(&s_cnt[6], slot_j);                                           // the used 'j' slots are one range
        atomicMax(&s_cnt[7], slot_j + 1);
        s_slot[tid] = __popc(m & (0xffffffffu >> (31 - lane))) - 1;             // rank inside the warp, completed below
    }
    __syncthreads();
    if (tid < kTcRows) {
        int before = 0;
        for (int w = 0; w < warp; ++w) before += s_cnt[2 + w];
        const int slot_i = s_slot[tid] + before;
        const bool head = tid == 0 || s_off[tid - 1] != key_i;
        s_slot[tid] = slot_i;
        if (tid == kTcRows - 1) {
            s_cnt[0] = slot_i + 1;
            s_cnt[1] = direct_j ? nb_tile * p.F : kTcRows;
        }
        __syncwarp();
        // (the keys in s_off[0..127] are overwritten by line offsets only after every thread has compared: next barrier)
        key_i = head ? key_i : -1;
    }
    __syncthreads();
    if (tid < kTcRows) {
        if (key_i >= 0) s_off[s_slot[tid]] = key_i * p.N;
        s_off[kTcRows + slot_j] = key_j * p.N;                                  // same value from every row that shares the line
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
        // ===== epilogue: warp (q, half) drains lanes 32 q .. 32 q + 31, columns 40 half .. 40 half + 39 (80 running sums
        // per thread do not fit the registers a CTA of this size leaves per thread: they were spilled) =====
        const int q = warp & 3, c0 = (warp >> 2) * kTcEpiCols;
        float total[kTcEpiCols];
#pragma unroll
        for (int i = 0; i < kTcEpiCols; ++i) total[i] = 0.f;
        for (int g = 0; g < n_groups; ++g) {
            const int a = g & 1;
            if (warp == 0) TC_TRACE(5, g, 0);
            tc_mbar_wait_relaxed(acc_full(a), (g >> 1) & 1, 4000);
            tc_fence_after();
            if (warp == 0) TC_TRACE(5, g, 1);
            const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (a ? kTcAcc1 : kTcAcc0) + c0;
            float v[kTcEpiCols];
#pragma unroll
            for (int j = 0; j < kTcEpiCols / 8; ++j) tc_ld8_nowait(taddr + 8 * j, *reinterpret_cast<float (*)[8]>(&v[8 * j]));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(acc_empty(a));                     // the accumulator is free before the adds
            if (warp == 0) TC_TRACE(5, g, 2);
#pragma unroll
            for (int i = 0; i < kTcEpiCols; ++i) total[i] += v[i];
        }
        // the correction products are complete when the third issuing thread says so
        tc_mbar_wait_relaxed(tail_done, 0, 1000);
        tc_fence_after();
        {
            const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + kTcAccS + c0;
            float v[kTcEpiCols];
#pragma unroll
            for (int j = 0; j < kTcEpiCols / 8; ++j) tc_ld8_nowait(taddr + 8 * j, *reinterpret_cast<float (*)[8]>(&v[8 * j]));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < kTcEpiCols; ++i) total[i] += v[i];
        }
        tc_fence_before();
        const long long row = row0 + 32 * q + lane;
        if (row < p.rows) {
#pragma unroll
            for (int i = 0; i < kTcEpiCols; ++i) {
                const int col = col0 + c0 + i;
                if (col < p.n_out) p.out[row * p.n_out + col] = total[i];
            }
        }
    } else if (warp < kTcMmaWarp + kTcMmaWarps) {
        // ===== MMA issuers: one thread of each of three warps.  TMEM columns [head a | corrections | head b].
        // Issuers 0 and 1 take the drain groups of their parity: head(A') head(B') into their own head accumulator.
        // Issuer 2 takes every slab: head(A') tail(B') + tail(A') head(B') into the correction accumulator.
        // Why three threads: an MMA of this size costs its issuing thread ~98 cycles whatever N <= 160 is, a wait on an
        // mbarrier that is already complete ~250, and threads issue independently (tools/micro/umma_rate.cu: one
        // thread 839, three threads 1833 of 1934 MAC/cycle/SM at N = 80); one thread issuing all twelve MMAs of a slab
        // was the limit of the whole kernel (tools/tc_trace.py).  Every accumulator has ONE writer, so the order of
        // its additions -- and the result, bit for bit -- does not depend on how the threads interleave. =====
        const int m = warp - kTcMmaWarp;
        if (lane == 0 && m < 2) {
            const int a = m;
            for (int g = a; g < n_groups; g += 2) {
                if (a == 0) TC_TRACE(4, g, 0);
                tc_mbar_wait(acc_empty(a), ((g >> 1) & 1) ^ 1);              // passes at once for the first group
                tc_fence_after();
                if (a == 0) TC_TRACE(4, g, 1);
                const int i_end = min((g + 1) * kTcDrain, n_slabs);
                uint32_t ready = 0;
                for (int i = g * kTcDrain; i < i_end; ++i) {
                    const int s = i % kTcStages;
                    if (!ready) tc_mbar_wait(full(s), (i / kTcStages) & 1);
                    tc_fence_after();
                    if (m == 0) TC_TRACE(4, g, 2 + 2 * (i - g * kTcDrain));
                    ready = i + 1 < i_end ? tc_mbar_test(full((i + 1) % kTcStages), ((i + 1) / kTcStages) & 1) : 0u;
                    const uint32_t d = tmem + (a ? kTcAcc1 : kTcAcc0);
                    const uint32_t a_hi = tmem + kTcA0 + 64 * s;
                    const uint32_t b_hi = base + s * kTcStageBytes;
#pragma unroll
                    for (int k = 0; k < kTcK / 8; ++k)                   // the first MMA of a drain group overwrites
                        tc_mma_ts(d, a_hi + 8 * k, tc_b_desc(b_hi + 32 * k), kTcIdesc, (i == g * kTcDrain && k == 0) ? 0u : 1u);
                    tc_commit(empty(s));                                      // the stage is free once the MMAs of both kinds retire
                    if (m == 0) TC_TRACE(4, g, 3 + 2 * (i - g * kTcDrain));
                }
                tc_commit(acc_full(a));
            }
        } else if (lane == 0 && m == 2) {
            uint32_t ready = 0;
            for (int i = 0; i < n_slabs; ++i) {
                const int s = i % kTcStages;
                if (!ready) tc_mbar_wait(full(s), (i / kTcStages) & 1);
                tc_fence_after();
                ready = i + 1 < n_slabs ? tc_mbar_test(full((i + 1) % kTcStages), ((i + 1) / kTcStages) & 1) : 0u;
                const uint32_t a_hi = tmem + kTcA0 + 64 * s, a_lo = a_hi + 32;
                const uint32_t b_hi = base + s * kTcStageBytes, b_lo = b_hi + kTcBTile;
#pragma unroll
                for (int k = 0; k < kTcK / 8; ++k) {
                    tc_mma_ts(tmem + kTcAccS, a_lo + 8 * k, tc_b_desc(b_hi + 32 * k), kTcIdesc, (i == 0 && k == 0) ? 0u : 1u);
                    tc_mma_ts(tmem + kTcAccS, a_hi + 8 * k, tc_b_desc(b_lo + 32 * k), kTcIdesc, 1u);
                }
                tc_commit(empty(s));
            }
            tc_commit(tail_done);
        }
        __syncwarp();
    } else {
        // ===== producers: four groups of four warps; group j owns the slabs i = j (mod 4), a thread the 16 samples of
        // row 32 q + lane in such a slab.  The slab's inputs -- one 128-byte line of (|z|, theta) and one of (re, im) per
        // row -- come through shared memory: the group copies them with COALESCED cp.async (eight lanes per line; a lane
        // reading its own row straight from global memory costs one L1 wavefront per lane, and the L1 data pipe was
        // the limit), swizzled so that the per-row 128-bit reads are conflict free.  While one group waits (its copies,
        // its stage, the tensor-memory stores), the other three compute. =====
        const int pw_ = warp - (kTcMmaWarp + kTcMmaWarps);
        const int q = warp & 3, grp = pw_ >> 2;
        const int gtid = q * 32 + lane;                         // 0..127 inside the group = the thread's row
        const long long row = row0 + gtid;
        float pw = 1.f;
        if (row < p.rows) {
            const int sidx = (int)(row % p.n_sel);
            pw = p.powers[p.subset ? p.subset[sidx] : sidx];
        }
        const uint32_t sin_base = base + kTcStages * kTcStageBytes + grp * kTcInBytes;
        const uint8_t* sin_ptr = sm + kTcStages * kTcStageBytes + grp * kTcInBytes;
        const bool wide = (p.N & 1) == 0;                       // rows start on 16-byte boundaries
        // thread (rb, c) of the group copies chunk c (16 bytes = 2 samples; 8 lanes cover a 128-byte line) of the line
        // slots rb, rb + 16, ... of both sides: the swizzle term (slot & 7) = (rb & 7) and the sample offset are the
        // thread's own constants
        const int cc = gtid & 7, rb = gtid >> 3;
        const uint32_t dst0 = sin_base + rb * 128 + ((cc ^ (rb & 7)) << 4);
        const int n_lines_i = s_cnt[0], n_lines_j = s_cnt[1];
        // 'j' side with the direct slot map: line slot -> (first sample of the tile * F + slot) * N, no table; only the
        // range of slots some row uses is copied
        const int j_lo = s_cnt[6] & ~15, j_hi = s_cnt[7];
        const long long j_base = b_first * p.F;
        auto copy_inputs = [&](int i) {
            if (wide) {
                const int t = i * kTcSlabT + 2 * cc;
                const int bytes_t = max(0, min(16, (p.N - t) * 8));
                for (int slot = rb; slot < n_lines_i; slot += 16) {
                    const int off = s_off[slot];
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + (slot - rb) * 128),
                                 "l"(p.zp + (bytes_t ? off + t : 0)), "r"(bytes_t) : "memory");
                }
                if (direct_j) {
                    for (int slot = j_lo + rb; slot < j_hi; slot += 16)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + kTcRows * 128 + (slot - rb) * 128),
                                     "l"(p.zc + (bytes_t ? (j_base + slot) * p.N + t : 0)), "r"(bytes_t) : "memory");
                } else {
                    for (int slot = rb; slot < n_lines_j; slot += 16) {
                        const int off = s_off[kTcRows + slot];
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + kTcRows * 128 + (slot - rb) * 128),
                                     "l"(p.zc + (bytes_t ? off + t : 0)), "r"(bytes_t) : "memory");
                    }
                }
            } else {
                // odd N: rows start on 8-byte boundaries only; sample by sample (thread (r8, c16): 16 lanes per line)
                const int c8 = gtid & 15, r16 = gtid >> 4;
                const int t = i * kTcSlabT + c8;
                const int bytes = t >= p.N ? 0 : 8;
#pragma unroll
                for (int side = 0; side < 2; ++side) {
                    const int n_lines = side ? n_lines_j : n_lines_i;
                    const float2* arr = side ? p.zc : p.zp;
                    for (int slot = r16; slot < n_lines; slot += 8) {
                        const int off = s_off[side * kTcRows + slot];
                        if (off < 0) continue;                          // a 'j' filter no row of this tile pairs with
                        const uint32_t dst = sin_base + side * (kTcRows * 128) + slot * 128 + (((c8 >> 1) ^ (slot & 7)) << 4) + (c8 & 1) * 8;
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(arr + (bytes ? off + t : 0)), "r"(bytes) : "memory");
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // this thread's row reads line slots li ('i' side) and lj ('j' side)
        const int li = s_slot[gtid], lj = s_slot[kTcRows + gtid];
        const uint8_t* line_i = sin_ptr + li * 128;
        const uint8_t* line_j = sin_ptr + kTcRows * 128 + lj * 128;
        const int swz_i = li & 7, swz_j = lj & 7;
        const uint8_t* Bimg = reinterpret_cast<const uint8_t*>(p.Bimg) + (size_t)blockIdx.y * p.n_slabs * kTcStageBytes;
        if (grp < n_slabs) copy_inputs(grp);
        for (int i = grp; i < n_slabs; i += kTcGroups) {
            const int s = i % kTcStages;
            if (q == 0) TC_TRACE(grp, i, 0);
            tc_mbar_wait_relaxed(empty(s), ((i / kTcStages) & 1) ^ 1, 1000);
            tc_fence_after();
            if (q == 0) TC_TRACE(grp, i, 1);
            // B' slab: its shared-memory image (head and tail, 2 x 80 rows of 128 bytes, swizzled) is one contiguous block
            // in global memory: ONE bulk copy by one thread, accounted on the slab's `full` barrier as transaction bytes
            // (the MMA issuers see the slab when the four warps have arrived AND the bytes have landed)
            if (gtid == 0) {
                asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(full(s)), "r"((uint32_t)kTcStageBytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(base + s * kTcStageBytes), "l"(Bimg + (size_t)i * kTcStageBytes), "r"((uint32_t)kTcStageBytes), "r"(full(s)) : "memory");
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");               // this slab's inputs
            asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");       // ... of every thread of the group
            if (q == 0) TC_TRACE(grp, i, 2);
            // A': the products of this thread's samples, split, straight into tensor memory
            const uint32_t ta = tmem + ((uint32_t)(32 * q) << 16) + kTcA0 + 64 * s;
#pragma unroll
            for (int ss = 0; ss < 4; ++ss) {
                float4 zp2[2], zc2[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    zp2[h] = *reinterpret_cast<const float4*>(line_i + (((2 * ss + h) ^ swz_i) << 4));
                    zc2[h] = *reinterpret_cast<const float4*>(line_j + (((2 * ss + h) ^ swz_j) << 4));
                }
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 a = zp2[j >> 1], b = zc2[j >> 1];
                    const float2 c = (j & 1) ? accelerated_product(make_float2(a.z, a.w), make_float2(b.z, b.w), pw)
                                             : accelerated_product(make_float2(a.x, a.y), make_float2(b.x, b.y), pw);
                    tc_split2(c, hi[2 * j], hi[2 * j + 1], lo[2 * j], lo[2 * j + 1]);
                }
                tc_st8(ta + 8 * ss, hi);
                tc_st8(ta + 32 + 8 * ss, lo);
            }
            // hand the slab to the tensor core first ...
            if (q == 0) TC_TRACE(grp, i, 3);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(full(s));
            if (q == 0) TC_TRACE(grp, i, 4);
            // ... then refill the inputs (off the producer -> MMA critical path): everyone has read them
            asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
            if (q == 0) TC_TRACE(grp, i, 5);
            if (i + kTcGroups < n_slabs) copy_inputs(i + kTcGroups);
            else asm volatile("cp.async.commit_group;" ::: "memory");
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kTcMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

}  // namespace tebscat
