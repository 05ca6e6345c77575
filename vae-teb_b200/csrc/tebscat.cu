// tebscat.cu -- C ABI (include/tebscat.h) + the sm_100a kernels of the scattering path.
//
// Product path only: there is no CPU fallback here.  Every entry point fails with
// TEBSCAT_ECUDA when no CUDA device is usable.
#include "../../include/tebscat.h"
#include "scat_core.cuh"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

using namespace tebscat;

// ---------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(TEBSCAT_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                  \
    } while (0)

// Every entry point runs on its plan's device and leaves the caller's current device as it found it (a
// single-process multi-GPU caller must not be moved to another device behind its back).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
        else if (err == cudaSuccess) prev = -1;              // already there: nothing to restore
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define ON_DEVICE(dev)                                                                                     \
    DeviceGuard device_guard_(dev);                                                                        \
    if (device_guard_.err != cudaSuccess)                                                                  \
        return fail(TEBSCAT_ECUDA, "cudaSetDevice(%d) failed: %s", (int)(dev), cudaGetErrorString(device_guard_.err))

// ---------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------
struct KParams {
    const float* arena;      // filters, fp32, bit-reversed bin order
    const float* win;        // optional analysis window of OP_LOAD ([N] or null)
    const float2* tw;        // kTwA coarse + kTwB fine twiddles
    const int4* warp_tab;    // [n_steps][kWarps] x 3 int4: the task of every warp in every step
    const int32_t* chan;     // channel table of the batched stores
    long long* prof;         // optional: clock64() of CTA 0 at every step boundary (first signal)
    float2* zc;              // phase stage A outputs (null for the scattering transform)
    float2* zp;
    long long x_stride;      // floats between the signals of consecutive jobs
    long long z_stride;      // float2 between the zc/zp blocks of consecutive jobs
    int32_t z_mode;
    int32_t n_steps;
    int32_t smem_complex;
    int32_t N, pad_left, log2_Np, n_paths, n_out;
    int32_t border;          // BORDER_* of OP_LOAD / OP_LOADPAIR
    // output epilogue (null mean = none)
    const float* ep_mean;
    const float* ep_std;
    const unsigned char* ep_mode;
    float ep_log_eps;
    int32_t ep_trim, ep_time_major;
    // phase stage B on the interpreter (pair_rows > 0): a job is n_paths consecutive (sample, pair) rows
    const float2* pair_zp;   // [Bc][F][N] (|z|, theta) of the 'i' channel
    const float2* pair_zc;   // [Bc][F][N] (re, im) of the 'j' channel
    const int32_t* pair_i;
    const int32_t* pair_j;
    const float* pair_pw;
    const int32_t* pair_subset;
    long long pair_rows;
    int32_t pair_n_sel, pair_F;
    // large-support level: job b works on the tile [b * g_slots, +g_slots) of a global complex buffer
    float2* gbuf;
    long long g_total;
    int32_t g_slots;
    // fused subtrees of the large-support level (OP_GMULFOLD, kernel variant GSRC): job b reads the spectrum at
    // gsrc + b * gsrc_stride (complex elements)
    const float2* gsrc;
    long long gsrc_stride;
    // schedules that PARK the signal's spectrum in global memory (plan_desc.scratch_complex > 0): CTA c owns
    // gscratch + c * gscratch_stride; it is the tile of OP_STOREC and the source of OP_GMULFOLD
    float2* gscratch;
    long long gscratch_stride;
};

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;

// One CTA per SM, persistent over the batch: signal b, b + gridDim.x, ...
// Every warp walks its own column of the task table.  The 12-int records travel global -> shared with
// cp.async, two steps ahead of their use, into a three-deep ring per warp: no register is held across a
// task body (the butterflies use all 128 -- a register-held prefetch gets spilled to local memory, which
// with the whole L1 carved out as shared memory is an L2 round trip at the start of every step), and the
// fields are read back one per lane and broadcast with shuffles when the step starts.
constexpr int kRing = 3;

__device__ __forceinline__ void fetch_record(int32_t* dst_smem, const int32_t* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src));
}
__device__ __forceinline__ void fetch_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void fetch_wait_all_but_2() { asm volatile("cp.async.wait_group 2;" ::: "memory"); }

template <bool PROF, bool GSRC = false>
__global__ void __launch_bounds__(kThreads, 1)
scat1d_kernel(const KParams p, const float* __restrict__ x, float* __restrict__ out, long long B) {
    extern __shared__ __align__(16) float2 smem[];
    __shared__ __align__(16) int32_t ring[kRing][kWarps][kTaskInts];
    float2* S = smem;
    float2* twA = smem + p.smem_complex;
    float2* twB = twA + kTwAP;
    for (int i = threadIdx.x; i < kTwAP + kTwBP; i += blockDim.x) twA[i] = p.tw[i];
    __syncthreads();

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int32_t* tab = reinterpret_cast<const int32_t*>(p.warp_tab) + kTaskInts * warp + lane;
    const int stride = kTaskInts * kWarps;
    const int n_steps = p.n_steps;
    // records of the first two steps
    if (lane < kTaskInts) fetch_record(&ring[0][warp][lane], tab);
    fetch_commit();
    if (lane < kTaskInts) fetch_record(&ring[1][warp][lane], tab + stride * (1 % n_steps));
    fetch_commit();
    int s_fetch = 2 % n_steps;       // step whose record is fetched next (wraps into the next signal)
    int slot = 0;                    // ring slot of the step about to run
    __shared__ SignalCtx c;          // per-signal context lives in shared memory: nothing to keep in
                                     // registers across the (partly out-of-line) task bodies
    for (long long b = blockIdx.x; b < B; b += gridDim.x) {
        if (tid == 0) {
            c.x = x + b * p.x_stride;
            c.win = p.win;
            c.out = out + b * (long long)p.n_paths * (p.ep_mean ? p.n_out - 2 * p.ep_trim : p.n_out);
            c.ep_mean = p.ep_mean;
            c.ep_std = p.ep_std;
            c.ep_mode = p.ep_mode;
            c.ep_log_eps = p.ep_log_eps;
            c.ep_trim = p.ep_trim;
            c.ep_time_major = p.ep_time_major;
            c.n_paths = p.n_paths;
            c.ch_limit = p.n_paths;
            c.gbuf = nullptr;
            c.g_valid = 0;
            if (p.gbuf) {
                const long long e0 = b * (long long)p.g_slots;
                c.gbuf = p.gbuf + e0;
                c.g_valid = (int)(p.g_total - e0 < p.g_slots ? p.g_total - e0 : p.g_slots);
            }
            if (p.pair_rows > 0) {
                const long long r0 = b * p.n_paths;
                c.ch_limit = (int)(p.pair_rows - r0 < p.n_paths ? p.pair_rows - r0 : p.n_paths);
                for (int q = 0; q < kMaxPairRows && q < p.n_paths; ++q) {
                    const long long r = r0 + q < p.pair_rows ? r0 + q : p.pair_rows - 1;    // a missing row repeats the last
                    const long long smp = r / p.pair_n_sel;
                    const int sel = (int)(r - smp * p.pair_n_sel);
                    const int pair = p.pair_subset ? p.pair_subset[sel] : sel;
                    c.pr_zp[q] = p.pair_zp + (smp * p.pair_F + p.pair_i[pair]) * (long long)p.N;
                    c.pr_zc[q] = p.pair_zc + (smp * p.pair_F + p.pair_j[pair]) * (long long)p.N;
                    c.pr_pw[q] = p.pair_pw[pair];
                }
            }
            c.chan = p.chan;
            c.zc = p.zc + b * p.z_stride;
            c.zp = p.zp + b * p.z_stride;
            c.z_mode = p.z_mode;
            c.N = p.N;
            c.pad_left = p.pad_left;
            c.log2_Np = p.log2_Np;
            c.border = p.border;
            c.n_out = p.n_out;
            if (GSRC) {
                if (p.gscratch) {
                    c.gbuf = p.gscratch + blockIdx.x * p.gscratch_stride;
                    c.g_valid = (int)p.gscratch_stride;
                    c.gsrc = c.gbuf;
                } else {
                    c.gsrc = p.gsrc + b * p.gsrc_stride;
                }
            }
        }
        __syncthreads();             // (the last step of the previous signal ended in a barrier too)
        const bool prof = PROF && blockIdx.x == 0 && b == blockIdx.x && tid == 0;
        for (int s = 0; s < n_steps; ++s) {
            {   // prefetch the record two steps ahead into the slot the previous step has just released
                const int fslot = slot == 0 ? kRing - 1 : slot - 1;
                if (lane < kTaskInts) fetch_record(&ring[fslot][warp][lane], tab + stride * s_fetch);
                fetch_commit();
                s_fetch = (s_fetch + 1 == n_steps) ? 0 : s_fetch + 1;
            }
            if (prof) p.prof[s] = clock64();
            fetch_wait_all_but_2();  // this step's record has landed (issued two steps ago)
            __syncwarp();
            // one field per lane, broadcast with shuffles: the compiler keeps shuffle-broadcast values in
            // UNIFORM registers, which the task bodies do not compete for
            const int cur = ring[slot][warp][lane < kTaskInts ? lane : 0];
            slot = slot + 1 == kRing ? 0 : slot + 1;
            const int op = __shfl_sync(0xffffffffu, cur, 0);
            const int f11 = __shfl_sync(0xffffffffu, cur, 11);
            const int warp_sync_only = f11 & 1;
            if ((op & 0xff) != OP_NOP) {
                Task t;
                t.op = op;
                t.t0 = __shfl_sync(0xffffffffu, cur, 1);
                t.nt = __shfl_sync(0xffffffffu, cur, 2);
                t.a = __shfl_sync(0xffffffffu, cur, 3);
                t.b = __shfl_sync(0xffffffffu, cur, 4);
                t.c = __shfl_sync(0xffffffffu, cur, 5);
                t.d = __shfl_sync(0xffffffffu, cur, 6);
                t.e = __shfl_sync(0xffffffffu, cur, 7);
                t.f = __shfl_sync(0xffffffffu, cur, 8);
                t.g = __shfl_sync(0xffffffffu, cur, 9);
                t.h = __shfl_sync(0xffffffffu, cur, 10);
                t.pad = f11;
#ifdef TEBSCAT_PROF_PHASES
                if (prof) p.prof[n_steps + 1 + 3 * s] = clock64();
#endif
#ifdef TEBSCAT_PROF_BFLY
                if (tid == 0 && blockIdx.x == 0) tebscat::g_bfly_dbg[7] = clock64();
#endif
                exec_task<GSRC>(S, twA, twB, p.arena, c, t, tid - t.t0);
#ifdef TEBSCAT_PROF_PHASES
                if (prof) p.prof[n_steps + 2 + 3 * s] = clock64();
#endif
            }
            // the plan marks the steps after which no warp touches a slot another warp has touched
            // since the last CTA barrier: there a warp-level fence is enough and the warps drift
            if (warp_sync_only) __syncwarp(); else __syncthreads();
#ifdef TEBSCAT_PROF_PHASES
            if (prof) p.prof[n_steps + 3 + 3 * s] = clock64();
#endif
        }
        if (prof) p.prof[n_steps] = clock64();
    }
}

// ---------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------
struct HostPipe {            // resources of the host-buffer entry point, created lazily
    // Three-stage pipeline over chunks of the batch: H2D on one stream, the kernel on a second, D2H on a third
    // (each direction has its own copy engine), kSlots chunks in flight, ordered by events.
    static constexpr int kSlots = 3;
    bool ready = false;
    int64_t chunk = 0;
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    cudaEvent_t loaded[kSlots] = {}, computed[kSlots] = {}, drained[kSlots] = {};
    float* d_x[kSlots] = {};
    float* d_S[kSlots] = {};
};

struct tebscat_plan {
    tebscat_plan_desc desc;
    int device = 0;
    int n_sms = 0;
    size_t smem_bytes = 0;
    float* d_arena = nullptr;
    float2* d_tw = nullptr;
    int32_t* d_warp_tab = nullptr;
    int32_t* d_chan = nullptr;
    float* d_win = nullptr;
    KParams kp;
    int64_t gsrc_extent = 0;     // > 0: the schedule reads a global source spectrum of that many complex bins (OP_GMULFOLD)
    // desc.scratch_complex > 0: the schedule parks U0 in a per-CTA scratch of that many complex elements.  One scratch
    // per stream the plan has been launched on (launches on one stream are ordered; different streams may overlap).
    int64_t scratch_complex = 0;
    mutable std::map<cudaStream_t, float2*> scratch;
    mutable std::mutex scratch_mu;
    float2* d_scratch_capture = nullptr;   // used while the launch stream is being captured into a CUDA graph (no allocation there)
    HostPipe pipe;
    std::mutex pipe_mu;
};

// the scratch of `p` for launches on `st` (allocated on first use; the plan's device is current)
static int plan_scratch(const tebscat_plan* p, cudaStream_t st, float2** out) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    CU(cudaStreamIsCapturing(st, &cap));
    if (cap != cudaStreamCaptureStatusNone) {          // replays of one graph are ordered among themselves
        *out = p->d_scratch_capture;
        return TEBSCAT_OK;
    }
    std::lock_guard<std::mutex> lock(p->scratch_mu);
    auto it = p->scratch.find(st);
    if (it == p->scratch.end()) {
        float2* d = nullptr;
        CU(cudaMalloc(&d, (size_t)p->n_sms * (size_t)p->scratch_complex * sizeof(float2)));
        it = p->scratch.emplace(st, d).first;
    }
    *out = it->second;
    return TEBSCAT_OK;
}

static int validate_schedule(const tebscat_plan_desc& d, const int32_t* tasks, const int32_t* steps,
                             size_t n_floats, size_t n_chan, int64_t* gsrc_extent = nullptr) {
    if (gsrc_extent) *gsrc_extent = 0;
    const int cap = (int)(((int64_t)d.smem_complex * 16) / 17);   // logical slots (1 pad slot per 16)
    for (int s = 0; s < d.n_steps; ++s) {
        const int b = steps[2 * s], e = steps[2 * s + 1];
        if (b < 0 || e < b || e > d.n_tasks) return fail(TEBSCAT_EINVAL, "step %d: bad task range [%d,%d)", s, b, e);
        // the tasks of a step own disjoint warps (a warp executes ONE task per step)
        unsigned busy = 0;
        for (int i = b; i < e; ++i) {
            const int32_t* t = tasks + kTaskInts * i;
            if ((t[0] & 0xff) == OP_NOP || t[1] < 0 || t[2] <= 0 || t[1] + t[2] > kThreads || ((t[1] | t[2]) & 31)) continue;
            const unsigned m = (t[2] >= 32 * 32 ? ~0u : ((1u << (t[2] / 32)) - 1u)) << (t[1] / 32);
            if (busy & m) return fail(TEBSCAT_EINVAL, "step %d: overlapping thread ranges", s);
            busy |= m;
        }
    }
    auto fits = [&](int64_t off, int64_t len) { return off >= 0 && (off & 15) == 0 && ((off + len + 15) & ~(int64_t)15) <= cap; };
    for (int i = 0; i < d.n_tasks; ++i) {
        const int32_t* t = tasks + kTaskInts * i;
        const int op = t[0] & 0xff;
        if (t[1] < 0 || t[2] <= 0 || t[1] + t[2] > d.n_threads || (t[1] & 31) || (t[2] & 31))
            return fail(TEBSCAT_EINVAL, "task %d: thread range [%d,+%d) outside the CTA", i, t[1], t[2]);
        switch (op) {
            case OP_LOAD:
                if (!fits(t[3], (int64_t)1 << d.log2_Np)) return fail(TEBSCAT_EINVAL, "task %d: LOAD out of range", i);
                break;
            case OP_FFT: {
                // a=region b=butterflies c=log2B d=log2R: the butterflies cover b*R slots in blocks of 2^c
                if ((t[7] & FFT_PACK) &&
                    (!(t[7] & FFT_FUSE_FWD) || t[9] < 0 || ((int64_t)t[9] << t[5]) > ((int64_t)t[4] << t[6]) ||
                     (t[9] > 0 && !fits(t[8], (int64_t)t[9] << t[5]))))
                    return fail(TEBSCAT_EINVAL, "task %d: bad packed pass (partner region %d, %d pairs)", i, t[8], t[9]);
                {   // chained passes (12 bits each in h and pad): blocks of at most 512 slots, never the fused pass
                    unsigned long long more = (unsigned long long)((unsigned)t[10] & 0xffffffu) |
                                              ((unsigned long long)(((unsigned)t[11] >> 4) & 0xfffu) << 24);
                    const int64_t slots = (int64_t)t[4] << t[6];
                    if (more && (t[5] > 9 || (t[7] & FFT_FUSE_FWD) || (slots & 15)))
                        return fail(TEBSCAT_EINVAL, "task %d: pass chain starts with a cross-warp pass", i);
                    for (; more & 0xfff; more >>= 12) {
                        const int lb = (int)(more & 15), lr = (int)((more >> 4) & 7), fl = (int)((more >> 7) & 15);
                        if (lb < 1 || lb > 9 || lr < 1 || lr > 4 || lr > lb || (fl & (FFT_FUSE_FWD | FFT_PACK)) ||
                            ((fl & FFT_MOD) && (lr != 4 || !(fl & FFT_INV))) || (slots & (((int64_t)1 << lb) - 1)))
                            return fail(TEBSCAT_EINVAL, "task %d: bad chained pass (B=2^%d R=2^%d)", i, lb, lr);
                    }
                }
                if (t[5] < 1 || t[5] > kLog2TwMax || t[6] < 1 || t[6] > 4 || t[6] > t[5] || t[4] < 1 ||
                    ((t[7] & (FFT_MOD | FFT_FUSE_FWD)) && (t[6] != 4 || !(t[7] & FFT_INV))) ||
                    (((int64_t)t[4] << t[6]) & (((int64_t)1 << t[5]) - 1)) || !fits(t[3], (int64_t)t[4] << t[6]))
                    return fail(TEBSCAT_EINVAL, "task %d: bad FFT pass (%d butterflies, B=2^%d R=2^%d at %d)", i, t[4], t[5], t[6], t[3]);
                break;
            }
            case OP_MULFOLD: {
                if (t[4] < 2 || t[4] > kLog2TwMax || t[5] < 0 || t[5] > t[4] || !fits(t[3], (int64_t)1 << t[4]) ||
                    !fits(t[6] & ~15, (t[6] & 15) + ((int64_t)1 << (t[4] - t[5]))) || t[7] < 0 ||
                    (t[7] & 3) || (t[5] == 0 && (t[6] & 3)) || (t[5] == 1 && (t[6] & 1)))
                    return fail(TEBSCAT_EINVAL, "task %d: bad MULFOLD", i);
                if (t[5] >= 2) {
                    const int logcw = t[10], n_chunks = 1 << (t[5] - logcw);
                    if (logcw < 2 || logcw > t[5] || n_chunks > 32 || (unsigned)t[8] == 0u ||
                        (n_chunks < 32 && ((unsigned)t[8] >> n_chunks) != 0u))
                        return fail(TEBSCAT_EINVAL, "task %d: bad MULFOLD chunk mask", i);
                }
                {   // filter extent: natural layout for k < 4, compacted to the active chunks for k >= 4
                    const size_t need = t[5] >= 2 ? (((size_t)1 << (t[4] - t[5])) << t[10]) * __builtin_popcount((unsigned)t[8])
                                                  : ((size_t)1 << t[4]);
                    if ((size_t)t[7] + need > n_floats) return fail(TEBSCAT_EINVAL, "task %d: MULFOLD filter outside the arena", i);
                }
                if (t[9] != 0 && (t[5] != 0 || t[9] < 0 || t[9] > 2))
                    return fail(TEBSCAT_EINVAL, "task %d: a first inverse pass can only be fused into a k=1 MULFOLD", i);
                break;
            }
            case OP_MULFOLD2: {
                // packed source: two destinations (d, g), no fused inverse pass
                const int64_t n_dst = (int64_t)1 << (t[4] - t[5]);
                if (t[4] < 2 || t[4] > kLog2TwMax || t[5] < 0 || t[5] > t[4] || !fits(t[3], (int64_t)1 << t[4]) ||
                    !fits(t[6] & ~15, (t[6] & 15) + n_dst) || !fits(t[9] & ~15, (t[9] & 15) + n_dst) || t[7] < 0 ||
                    (t[7] & 3) || (t[5] == 0 && ((t[6] | t[9]) & 3)) || (t[5] == 1 && ((t[6] | t[9]) & 1)))
                    return fail(TEBSCAT_EINVAL, "task %d: bad MULFOLD2", i);
                if (t[5] >= 2) {
                    const int logcw = t[10], n_chunks = 1 << (t[5] - logcw);
                    if (logcw < 2 || logcw > t[5] || n_chunks > 32 || (unsigned)t[8] == 0u ||
                        (n_chunks < 32 && ((unsigned)t[8] >> n_chunks) != 0u))
                        return fail(TEBSCAT_EINVAL, "task %d: bad MULFOLD2 chunk mask", i);
                }
                const size_t need = t[5] >= 2 ? (((size_t)1 << (t[4] - t[5])) << t[10]) * __builtin_popcount((unsigned)t[8])
                                              : ((size_t)1 << t[4]);
                if ((size_t)t[7] + need > n_floats) return fail(TEBSCAT_EINVAL, "task %d: MULFOLD2 filter outside the arena", i);
                break;
            }
            case OP_GMULFOLD: {
                // source in global memory (tebscat_scat1d_forward_gsrc): a = offset in complex bins, up to 2^17 bins
                const int log_dst = t[4] - t[5];
                if (t[3] < 0 || (t[3] & 3) || t[4] < 2 || t[4] > 17 || t[5] < 0 || t[5] > t[4] || log_dst > kLog2TwMax ||
                    !fits(t[6] & ~15, (t[6] & 15) + ((int64_t)1 << log_dst)) || t[7] < 0 || (t[7] & 3) ||
                    (t[5] == 0 && (t[6] & 3)) || (t[5] == 1 && (t[6] & 1)))
                    return fail(TEBSCAT_EINVAL, "task %d: bad GMULFOLD", i);
                if (t[5] >= 2) {
                    const int logcw = t[10], n_chunks = 1 << (t[5] - logcw);
                    if (logcw < 2 || logcw > t[5] || n_chunks > 32 || (unsigned)t[8] == 0u ||
                        (n_chunks < 32 && ((unsigned)t[8] >> n_chunks) != 0u))
                        return fail(TEBSCAT_EINVAL, "task %d: bad GMULFOLD chunk mask", i);
                }
                const size_t need = t[5] >= 2 ? (((size_t)1 << log_dst) << t[10]) * __builtin_popcount((unsigned)t[8])
                                              : ((size_t)1 << t[4]);
                if ((size_t)t[7] + need > n_floats) return fail(TEBSCAT_EINVAL, "task %d: GMULFOLD filter outside the arena", i);
                if (t[9] != 0 && (t[5] != 0 || t[9] < 0 || t[9] > 2))
                    return fail(TEBSCAT_EINVAL, "task %d: a first inverse pass can only be fused into a k=1 GMULFOLD", i);
                if (gsrc_extent && *gsrc_extent < (int64_t)t[3] + ((int64_t)1 << t[4])) *gsrc_extent = (int64_t)t[3] + ((int64_t)1 << t[4]);
                break;
            }
            case OP_GMULFOLD2: {
                // two filters (e, g) and two destinations (d, a) on one read of the global source
                const int log_dst = t[4] - t[5];
                const int64_t n_dst = (int64_t)1 << (log_dst < 0 ? 0 : log_dst);
                const int radix = t[10] >> 8, logcw = t[10] & 0xff;
                if (t[4] < 2 || t[4] > 17 || t[5] < 0 || t[5] > t[4] || log_dst > kLog2TwMax ||
                    !fits(t[6] & ~15, (t[6] & 15) + n_dst) || !fits(t[3] & ~15, (t[3] & 15) + n_dst) || t[7] < 0 || (t[7] & 3) ||
                    t[9] < 0 || (t[9] & 3) || (t[5] == 0 && ((t[6] | t[3]) & 3)) || (t[5] == 1 && ((t[6] | t[3]) & 1)))
                    return fail(TEBSCAT_EINVAL, "task %d: bad GMULFOLD2", i);
                if (t[5] >= 2) {
                    const int n_chunks = 1 << (t[5] - logcw);
                    if (radix != 0 || logcw < 2 || logcw > t[5] || n_chunks > 32 || (unsigned)t[8] == 0u ||
                        (n_chunks < 32 && ((unsigned)t[8] >> n_chunks) != 0u))
                        return fail(TEBSCAT_EINVAL, "task %d: bad GMULFOLD2 chunk mask", i);
                } else if (logcw != 0 || radix < 0 || radix > 2 || (radix && t[5] != 0)) {
                    return fail(TEBSCAT_EINVAL, "task %d: bad GMULFOLD2 fused pass", i);
                }
                const size_t need = t[5] >= 2 ? (((size_t)1 << log_dst) << logcw) * __builtin_popcount((unsigned)t[8])
                                              : ((size_t)1 << t[4]);
                if ((size_t)t[7] + need > n_floats || (size_t)t[9] + need > n_floats)
                    return fail(TEBSCAT_EINVAL, "task %d: GMULFOLD2 filter outside the arena", i);
                if (gsrc_extent && *gsrc_extent < ((int64_t)1 << t[4])) *gsrc_extent = (int64_t)1 << t[4];
                break;
            }
            case OP_STOREB:
                if (t[4] < 1 || t[6] != d.n_out || t[5] < 0 || t[8] < 0 || t[8] > kLog2TwMax || t[5] + t[6] > (1 << t[8]) ||
                    !fits(t[3], (int64_t)t[4] << t[8]) || t[7] < 0 || 2 * ((size_t)t[7] + (size_t)t[4]) > n_chan)
                    return fail(TEBSCAT_EINVAL, "task %d: bad STOREB", i);
                break;
            case OP_TINY:
                if (t[4] < 1 || t[5] < 1 || t[5] > 3 || !fits(t[3], (int64_t)t[4] << t[5]))
                    return fail(TEBSCAT_EINVAL, "task %d: bad TINY transform", i);
                break;
            case OP_STOREZ:
                if (t[4] < 0 || t[4] >= d.n_paths || t[6] != d.n_out || t[5] < 0 || !fits(t[3], (int64_t)t[5] + t[6]))
                    return fail(TEBSCAT_EINVAL, "task %d: bad STOREZ", i);
                break;
            case OP_LOADC:
            case OP_STOREC:
                if (t[4] < 1 || !fits(t[3], t[4])) return fail(TEBSCAT_EINVAL, "task %d: bad tile transfer", i);
                break;
            case OP_STOREU:
                if (t[5] < 0 || t[6] < 1 || t[7] < 0 || (int64_t)t[7] + t[6] > (int64_t)d.n_paths * d.n_out ||
                    !fits(t[3] & ~15, (t[3] & 15) + (int64_t)t[5] + t[6]))
                    return fail(TEBSCAT_EINVAL, "task %d: bad STOREU", i);
                break;
            case OP_LOADPAIR:
                if (t[4] < 0 || t[4] >= kMaxPairRows || t[4] >= d.n_paths || !fits(t[3], (int64_t)1 << d.log2_Np))
                    return fail(TEBSCAT_EINVAL, "task %d: bad LOADPAIR", i);
                break;
            case OP_NOP:
                break;
            default:
                return fail(TEBSCAT_EINVAL, "task %d: unknown opcode %d", i, op);
        }
    }
    return TEBSCAT_OK;
}

extern "C" int tebscat_abi_version(void) { return TEBSCAT_ABI_VERSION; }
extern "C" const char* tebscat_last_error(void) { return g_err; }
extern "C" int tebscat_last_launch_count(void) { return g_launches; }

extern "C" int tebscat_plan_create(const tebscat_plan_desc* desc, const float* arena, size_t n_floats,
                                   const int32_t* tasks, const int32_t* steps,
                                   const int32_t* chan, size_t n_chan, int device, tebscat_plan** out) {
    if (!desc || !arena || !tasks || !steps || !chan || !out) return fail(TEBSCAT_EINVAL, "null argument");
    if (desc->abi_version != TEBSCAT_ABI_VERSION)
        return fail(TEBSCAT_EINVAL, "ABI version %d != %d", desc->abi_version, TEBSCAT_ABI_VERSION);
    if (desc->log2_Np < 1 || desc->log2_Np > kLog2TwMax)
        return fail(TEBSCAT_EUNSUPPORTED, "padded length 2^%d exceeds the single-CTA limit 2^%d", desc->log2_Np, kLog2TwMax);
    if (desc->N < 2 || desc->pad_left < 0 || desc->pad_left >= desc->N ||
        (1 << desc->log2_Np) - desc->N - desc->pad_left >= desc->N || (1 << desc->log2_Np) < desc->N)
        return fail(TEBSCAT_EINVAL, "Indefinite padding size (larger than tensor).");
    if (desc->n_threads != kThreads) return fail(TEBSCAT_EUNSUPPORTED, "schedules must target 512-thread CTAs");
    if (desc->n_paths < 1 || desc->n_out < 1 || desc->n_tasks < 1 || desc->n_steps < 1 || desc->smem_complex < 1)
        return fail(TEBSCAT_EINVAL, "empty plan");
    // channel table: (real-part channel, imaginary-part channel or -1) per pool slot
    for (size_t i = 0; i < n_chan; ++i)
        if (chan[i] >= desc->n_paths || chan[i] < ((i & 1) ? -1 : 0))
            return fail(TEBSCAT_EINVAL, "channel table entry %zu out of range", i);
    int64_t gsrc_extent = 0;
    if (int rc = validate_schedule(*desc, tasks, steps, n_floats, n_chan, &gsrc_extent)) return rc;
    if (desc->scratch_complex < 0 || (desc->scratch_complex & 3) || (desc->scratch_complex && gsrc_extent > desc->scratch_complex))
        return fail(TEBSCAT_EINVAL, "bad scratch size %lld (the schedule reads %lld bins)", (long long)desc->scratch_complex,
                    (long long)gsrc_extent);

    int n_dev = 0;
    CU(cudaGetDeviceCount(&n_dev));
    if (device < 0 || device >= n_dev) return fail(TEBSCAT_EINVAL, "device %d not in [0,%d)", device, n_dev);
    ON_DEVICE(device);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));

    // (owned by a guard until the plan is complete: a failing CUDA call below must not leak it)
    std::unique_ptr<tebscat_plan, void (*)(tebscat_plan*)> guard(new tebscat_plan(), tebscat_plan_destroy);
    tebscat_plan* p = guard.get();
    p->desc = *desc;
    p->device = device;
    p->gsrc_extent = gsrc_extent;
    p->scratch_complex = desc->scratch_complex;
    p->n_sms = prop.multiProcessorCount;
    p->smem_bytes = ((size_t)desc->smem_complex + kTwAP + kTwBP) * sizeof(float2);
    if (p->smem_bytes > (size_t)prop.sharedMemPerBlockOptin) {
        const size_t wanted = p->smem_bytes;
        return fail(TEBSCAT_EUNSUPPORTED, "schedule needs %zu B of shared memory, device offers %zu",
                    wanted, (size_t)prop.sharedMemPerBlockOptin);
    }
    std::vector<float2> tw(kTwAP + kTwBP, make_float2(0.f, 0.f));
    const double w0 = -2.0 * M_PI / (double)(1 << kLog2TwMax);
    for (int a = 0; a < kTwA; ++a) tw[a + (a >> 4)] = make_float2((float)cos(w0 * 128.0 * a), (float)sin(w0 * 128.0 * a));
    for (int b = 0; b < kTwB; ++b) tw[kTwAP + b + (b >> 4)] = make_float2((float)cos(w0 * b), (float)sin(w0 * b));

    CU(cudaMalloc(&p->d_arena, n_floats * sizeof(float)));
    CU(cudaMalloc(&p->d_tw, tw.size() * sizeof(float2)));
    if (p->scratch_complex)
        CU(cudaMalloc(&p->d_scratch_capture, (size_t)p->n_sms * (size_t)p->scratch_complex * sizeof(float2)));
    // per-warp view of the schedule: record (step, warp) = the task whose thread range covers the warp
    std::vector<int32_t> wt((size_t)desc->n_steps * kWarps * kTaskInts, 0);
    for (int st = 0; st < desc->n_steps; ++st) {
        for (int ti = steps[2 * st]; ti < steps[2 * st + 1]; ++ti) {
            const int32_t* t = tasks + kTaskInts * ti;
            for (int w = t[1] / 32; w < (t[1] + t[2]) / 32; ++w) {
                int32_t* rec = wt.data() + ((size_t)st * kWarps + w) * kTaskInts;
                // (validate_schedule has rejected overlapping ranges; the guard owns the plan on every error path)
                if ((rec[0] & 0xff) != OP_NOP) return fail(TEBSCAT_EINVAL, "step %d: overlapping thread ranges", st);
                memcpy(rec, t, kTaskInts * sizeof(int32_t));
            }
        }
        // field 11 of a step's tasks: 1 = the barrier after the step may be a warp-level fence.
        // Every warp of the CTA (idle ones included) must take the same kind of barrier.
        int relaxed = steps[2 * st] < steps[2 * st + 1] ? 1 : 0;
        for (int ti = steps[2 * st]; ti < steps[2 * st + 1]; ++ti) relaxed &= (tasks[kTaskInts * ti + 11] & 1);
        if (st == desc->n_steps - 1) relaxed = 0;
        for (int w = 0; w < kWarps; ++w) {
            int32_t& f11 = wt[((size_t)st * kWarps + w) * kTaskInts + 11];
            f11 = (f11 & ~1) | relaxed;
        }
    }
    CU(cudaMalloc(&p->d_warp_tab, wt.size() * sizeof(int32_t)));
    CU(cudaMemcpy(p->d_warp_tab, wt.data(), wt.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&p->d_chan, (n_chan ? n_chan : 1) * sizeof(int32_t)));
    CU(cudaMemcpy(p->d_chan, chan, n_chan * sizeof(int32_t), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(p->d_arena, arena, n_floats * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(p->d_tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
    // the attribute belongs to the kernel, not to the plan: always allow the device maximum
    cudaFuncAttributes fa0, fa1, fa2;
    CU(cudaFuncGetAttributes(&fa0, scat1d_kernel<false>));
    CU(cudaFuncGetAttributes(&fa1, scat1d_kernel<true>));
    CU(cudaFuncGetAttributes(&fa2, (scat1d_kernel<false, true>)));
    CU(cudaFuncSetAttribute((scat1d_kernel<false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)(prop.sharedMemPerBlockOptin - fa2.sharedSizeBytes)));
    CU(cudaFuncSetAttribute((scat1d_kernel<true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)(prop.sharedMemPerBlockOptin - fa2.sharedSizeBytes)));
    CU(cudaFuncSetAttribute(scat1d_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)(prop.sharedMemPerBlockOptin - fa0.sharedSizeBytes)));
    CU(cudaFuncSetAttribute(scat1d_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)(prop.sharedMemPerBlockOptin - fa1.sharedSizeBytes)));
    if (p->smem_bytes + fa0.sharedSizeBytes > (size_t)prop.sharedMemPerBlockOptin ||
        p->smem_bytes + fa1.sharedSizeBytes > (size_t)prop.sharedMemPerBlockOptin ||
        p->smem_bytes + fa2.sharedSizeBytes > (size_t)prop.sharedMemPerBlockOptin)
        return fail(TEBSCAT_EUNSUPPORTED, "schedule needs more shared memory than the device offers");   // guard frees the plan

    KParams& k = p->kp;
    k.arena = p->d_arena;
    k.win = nullptr;
    k.tw = p->d_tw;
    k.warp_tab = reinterpret_cast<const int4*>(p->d_warp_tab);
    k.chan = p->d_chan;
    k.prof = nullptr;
    k.zc = nullptr;
    k.zp = nullptr;
    k.x_stride = desc->N;
    k.z_stride = 0;
    k.z_mode = 0;
    k.n_steps = desc->n_steps;
    k.smem_complex = desc->smem_complex;
    k.N = desc->N;
    k.pad_left = desc->pad_left;
    k.log2_Np = desc->log2_Np;
    k.border = desc->border_mode;
    k.n_paths = desc->n_paths;
    k.n_out = desc->n_out;
    k.ep_mean = nullptr;
    k.ep_std = nullptr;
    k.ep_mode = nullptr;
    k.ep_log_eps = 0.f;
    k.ep_trim = 0;
    k.ep_time_major = 0;
    k.pair_zp = nullptr;
    k.pair_zc = nullptr;
    k.pair_i = nullptr;
    k.pair_j = nullptr;
    k.pair_pw = nullptr;
    k.pair_subset = nullptr;
    k.pair_rows = 0;
    k.pair_n_sel = 0;
    k.pair_F = 0;
    k.gbuf = nullptr;
    k.g_total = 0;
    k.g_slots = 0;
    k.gsrc = nullptr;
    k.gsrc_stride = 0;
    k.gscratch = nullptr;
    k.gscratch_stride = 0;
    *out = guard.release();
    return TEBSCAT_OK;
}

extern "C" int tebscat_plan_get_desc(const tebscat_plan* p, tebscat_plan_desc* out) {
    if (!p || !out) return fail(TEBSCAT_EINVAL, "null argument");
    *out = p->desc;
    return TEBSCAT_OK;
}

// ---------------------------------------------------------------------------------
// plan files: everything tebscat_plan_create takes, in one self-describing file, so that a consumer without the
// Python scheduler (tebscat/schedule.py builds the schedule) can create plans at run time
// ---------------------------------------------------------------------------------
namespace {
struct PlanFileHeader {
    char magic[8];            // "TEBSCATP"
    int32_t abi_version;
    int32_t header_bytes;
    uint64_t n_floats, n_chan;
    uint64_t checksum;        // FNV-1a over desc + payload
    tebscat_plan_desc desc;
};
uint64_t fnv1a(uint64_t h, const void* data, size_t n) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}
}  // namespace

extern "C" int tebscat_plan_save(const char* path, const tebscat_plan_desc* desc, const float* arena, size_t n_floats,
                                 const int32_t* tasks, const int32_t* steps, const int32_t* chan, size_t n_chan) {
    if (!path || !desc || !arena || !tasks || !steps || !chan) return fail(TEBSCAT_EINVAL, "null argument");
    if (desc->abi_version != TEBSCAT_ABI_VERSION) return fail(TEBSCAT_EINVAL, "ABI version %d != %d", desc->abi_version, TEBSCAT_ABI_VERSION);
    if (desc->n_tasks < 1 || desc->n_steps < 1) return fail(TEBSCAT_EINVAL, "empty plan");
    if (int rc = validate_schedule(*desc, tasks, steps, n_floats, n_chan)) return rc;
    PlanFileHeader h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, "TEBSCATP", 8);
    h.abi_version = TEBSCAT_ABI_VERSION;
    h.header_bytes = (int32_t)sizeof(h);
    h.n_floats = n_floats;
    h.n_chan = n_chan;
    h.desc = *desc;
    uint64_t c = fnv1a(1469598103934665603ull, desc, sizeof(*desc));
    c = fnv1a(c, arena, n_floats * sizeof(float));
    c = fnv1a(c, tasks, (size_t)desc->n_tasks * kTaskInts * sizeof(int32_t));
    c = fnv1a(c, steps, (size_t)desc->n_steps * 2 * sizeof(int32_t));
    c = fnv1a(c, chan, n_chan * sizeof(int32_t));
    h.checksum = c;
    FILE* f = fopen(path, "wb");
    if (!f) return fail(TEBSCAT_EINVAL, "cannot open %s for writing", path);
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1 && fwrite(arena, sizeof(float), n_floats, f) == n_floats &&
              fwrite(tasks, sizeof(int32_t), (size_t)desc->n_tasks * kTaskInts, f) == (size_t)desc->n_tasks * kTaskInts &&
              fwrite(steps, sizeof(int32_t), (size_t)desc->n_steps * 2, f) == (size_t)desc->n_steps * 2 &&
              (n_chan == 0 || fwrite(chan, sizeof(int32_t), n_chan, f) == n_chan);
    ok = (fclose(f) == 0) && ok;
    return ok ? TEBSCAT_OK : fail(TEBSCAT_EINVAL, "short write to %s", path);
}

extern "C" int tebscat_plan_load(const char* path, int device, tebscat_plan** out) {
    if (!path || !out) return fail(TEBSCAT_EINVAL, "null argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(TEBSCAT_EINVAL, "cannot open %s", path);
    PlanFileHeader h;
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "TEBSCATP", 8) != 0 || h.header_bytes != (int32_t)sizeof(h)) {
        fclose(f);
        return fail(TEBSCAT_EINVAL, "%s is not a tebscat plan file", path);
    }
    if (h.abi_version != TEBSCAT_ABI_VERSION || h.desc.abi_version != TEBSCAT_ABI_VERSION) {
        fclose(f);
        return fail(TEBSCAT_EINVAL, "plan file of ABI version %d, library has %d", h.abi_version, TEBSCAT_ABI_VERSION);
    }
    if (h.desc.n_tasks < 1 || h.desc.n_steps < 1 || h.desc.n_tasks > (1 << 24) || h.desc.n_steps > (1 << 24) ||
        h.n_floats < 1 || h.n_floats > ((uint64_t)1 << 32) || h.n_chan > ((uint64_t)1 << 28)) {
        fclose(f);
        return fail(TEBSCAT_EINVAL, "plan file with implausible sizes");
    }
    std::vector<float> arena(h.n_floats);
    std::vector<int32_t> tasks((size_t)h.desc.n_tasks * kTaskInts), steps((size_t)h.desc.n_steps * 2), chan(h.n_chan ? h.n_chan : 1);
    bool ok = fread(arena.data(), sizeof(float), arena.size(), f) == arena.size() &&
              fread(tasks.data(), sizeof(int32_t), tasks.size(), f) == tasks.size() &&
              fread(steps.data(), sizeof(int32_t), steps.size(), f) == steps.size() &&
              (h.n_chan == 0 || fread(chan.data(), sizeof(int32_t), h.n_chan, f) == h.n_chan);
    char extra;
    ok = ok && fread(&extra, 1, 1, f) == 0;          // nothing may follow the payload
    fclose(f);
    if (!ok) return fail(TEBSCAT_EINVAL, "plan file %s is truncated or has trailing bytes", path);
    uint64_t c = fnv1a(1469598103934665603ull, &h.desc, sizeof(h.desc));
    c = fnv1a(c, arena.data(), arena.size() * sizeof(float));
    c = fnv1a(c, tasks.data(), tasks.size() * sizeof(int32_t));
    c = fnv1a(c, steps.data(), steps.size() * sizeof(int32_t));
    c = fnv1a(c, chan.data(), h.n_chan * sizeof(int32_t));
    if (c != h.checksum) return fail(TEBSCAT_EINVAL, "plan file %s fails its checksum", path);
    return tebscat_plan_create(&h.desc, arena.data(), arena.size(), tasks.data(), steps.data(), chan.data(), h.n_chan, device, out);
}

extern "C" void tebscat_plan_destroy(tebscat_plan* p) {
    if (!p) return;
    DeviceGuard device_guard_(p->device);
    {
        HostPipe& hp = p->pipe;
        if (hp.s_in) cudaStreamDestroy(hp.s_in);
        if (hp.s_run) cudaStreamDestroy(hp.s_run);
        if (hp.s_out) cudaStreamDestroy(hp.s_out);
        for (int i = 0; i < HostPipe::kSlots; ++i) {
            if (hp.loaded[i]) cudaEventDestroy(hp.loaded[i]);
            if (hp.computed[i]) cudaEventDestroy(hp.computed[i]);
            if (hp.drained[i]) cudaEventDestroy(hp.drained[i]);
            cudaFree(hp.d_x[i]);
            cudaFree(hp.d_S[i]);
        }
    }
    cudaFree(p->d_arena);
    cudaFree(p->d_tw);
    cudaFree(p->d_warp_tab);
    cudaFree(p->d_chan);
    cudaFree(p->d_win);
    for (auto& kv : p->scratch) cudaFree(kv.second);
    cudaFree(p->d_scratch_capture);
    delete p;
}

// Analysis window of the plan's loads: x[t] * w[t] before padding (the Tukey taper of
// KymatioPhaseScattering1D, hdf5_dataset/kymatio_phase_scattering.py:362-392, applied at :405-407).
// window_host: N floats, or NULL to remove it.  Not re-entrant with forward calls on the same plan.
extern "C" int tebscat_plan_set_window(tebscat_plan* p, const float* window_host) {
    if (!p) return fail(TEBSCAT_EINVAL, "null plan");
    ON_DEVICE(p->device);
    int rc = TEBSCAT_OK;
    do {
        if (!window_host) { p->kp.win = nullptr; break; }
        if (!p->d_win && cudaMalloc(&p->d_win, (size_t)p->desc.N * sizeof(float)) != cudaSuccess) {
            rc = fail(TEBSCAT_ECUDA, "cudaMalloc of the window failed");
            break;
        }
        if (cudaMemcpy(p->d_win, window_host, (size_t)p->desc.N * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
            rc = fail(TEBSCAT_ECUDA, "upload of the window failed");
            break;
        }
        p->kp.win = p->d_win;
    } while (0);
    return rc;
}

// launch of a plan's schedule with the parameters `kp` (the plan's own, possibly with an epilogue): plans that park
// U0 in global memory run on the kernel variant with the global-source op, on their per-stream scratch
static int launch_plan(const tebscat_plan* p, KParams kp, const float* x, int64_t B, float* S, cudaStream_t st) {
    if (p->gsrc_extent && !p->scratch_complex)
        return fail(TEBSCAT_EINVAL, "this plan reads a global source spectrum: use tebscat_scat1d_forward_gsrc");
    if (B == 0) return TEBSCAT_OK;
    const int grid = (int)(B < (int64_t)p->n_sms ? B : (int64_t)p->n_sms);
    if (p->scratch_complex) {
        if (int rc = plan_scratch(p, st, &kp.gscratch)) return rc;
        kp.gscratch_stride = p->scratch_complex;
        scat1d_kernel<false, true><<<grid, p->desc.n_threads, p->smem_bytes, st>>>(kp, x, S, (long long)B);
    } else {
        scat1d_kernel<false><<<grid, p->desc.n_threads, p->smem_bytes, st>>>(kp, x, S, (long long)B);
    }
    CU(cudaGetLastError());
    ++g_launches;
    return TEBSCAT_OK;
}

static int launch_scat1d(const tebscat_plan* p, const float* x, int64_t B, float* S, cudaStream_t st) {
    return launch_plan(p, p->kp, x, B, S, st);
}

extern "C" int tebscat_scat1d_forward(const tebscat_plan* p, const float* x_dev, int64_t B, float* S_dev,
                                      void* stream) {
    g_launches = 0;
    if (!p || B < 0 || (B > 0 && (!x_dev || !S_dev))) return fail(TEBSCAT_EINVAL, "null argument");
    ON_DEVICE(p->device);
    return launch_scat1d(p, x_dev, B, S_dev, (cudaStream_t)stream);
}

/* Fused subtrees of the large-support level: the schedule's OP_GMULFOLD tasks read job b's source spectrum
 * (complex64, bit-reversed bin order) at src_dev + b * src_stride complex elements; everything below runs out of
 * shared memory like the cascade itself and the leaves land in S_dev [B, n_paths, n_out] at their channels. */
extern "C" int tebscat_scat1d_forward_gsrc(const tebscat_plan* p, const float* src_dev, int64_t src_stride, int64_t B,
                                           float* S_dev, void* stream) {
    g_launches = 0;
    if (!p || B < 0 || (B > 0 && (!src_dev || !S_dev))) return fail(TEBSCAT_EINVAL, "null argument");
    if (!p->gsrc_extent || p->scratch_complex) return fail(TEBSCAT_EINVAL, "this plan has no global-source task");
    if (src_stride < p->gsrc_extent)
        return fail(TEBSCAT_EINVAL, "source stride %lld below the %lld bins the schedule reads", (long long)src_stride,
                    (long long)p->gsrc_extent);
    if ((reinterpret_cast<uintptr_t>(src_dev) & 31) || (src_stride & 3))
        return fail(TEBSCAT_EINVAL, "the source spectra must be 32-byte aligned");
    if (B == 0) return TEBSCAT_OK;
    ON_DEVICE(p->device);
    KParams kp = p->kp;
    kp.gsrc = reinterpret_cast<const float2*>(src_dev);
    kp.gsrc_stride = src_stride;
    const int grid = (int)(B < (int64_t)p->n_sms ? B : (int64_t)p->n_sms);
    scat1d_kernel<false, true><<<grid, p->desc.n_threads, p->smem_bytes, (cudaStream_t)stream>>>(kp, nullptr, S_dev, (long long)B);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(TEBSCAT_ECUDA, "launch failed: %s", cudaGetErrorString(e));
    ++g_launches;
    return TEBSCAT_OK;
}

extern "C" int tebscat_scat1d_forward_ex(const tebscat_plan* p, const float* x_dev, int64_t B, float* out_dev,
                                         const tebscat_epilogue* ep, void* stream) {
    g_launches = 0;
    if (!p || B < 0 || (B > 0 && (!x_dev || !out_dev))) return fail(TEBSCAT_EINVAL, "null argument");
    if (!ep) return tebscat_scat1d_forward(p, x_dev, B, out_dev, stream);
    if (!ep->mean_dev || !ep->std_dev || !ep->mode_dev) return fail(TEBSCAT_EINVAL, "epilogue: null statistics");
    if (ep->trim < 0 || 2 * ep->trim >= p->desc.n_out)
        return fail(TEBSCAT_EINVAL, "epilogue: trim %d leaves nothing of %d samples", ep->trim, p->desc.n_out);
    if (B == 0) return TEBSCAT_OK;
    ON_DEVICE(p->device);
    KParams kp = p->kp;
    kp.ep_mean = ep->mean_dev;
    kp.ep_std = ep->std_dev;
    kp.ep_mode = ep->mode_dev;
    kp.ep_log_eps = ep->log_eps;
    kp.ep_trim = ep->trim;
    kp.ep_time_major = ep->time_major ? 1 : 0;
    return launch_plan(p, kp, x_dev, B, out_dev, (cudaStream_t)stream);
}

extern "C" int tebscat_scat1d_profile_steps(const tebscat_plan* p, const float* x_dev, int64_t B, float* S_dev,
                                            long long* step_clocks_host, void* stream) {
    g_launches = 0;
    if (!p || !x_dev || !S_dev || !step_clocks_host || B < 1) return fail(TEBSCAT_EINVAL, "null argument");
    if (p->gsrc_extent && !p->scratch_complex)
        return fail(TEBSCAT_EUNSUPPORTED, "step profiling needs a schedule that starts from the signal");
    ON_DEVICE(p->device);
    long long* d_prof = nullptr;
#ifdef TEBSCAT_PROF_PHASES
    const size_t n = 4 * (size_t)p->desc.n_steps + 1;     // + (decode, exec, barrier) stamps of every step
#else
    const size_t n = (size_t)p->desc.n_steps + 1;
#endif
    CU(cudaMalloc(&d_prof, n * sizeof(long long)));
    CU(cudaMemsetAsync(d_prof, 0, n * sizeof(long long), (cudaStream_t)stream));
    KParams kp = p->kp;
    kp.prof = d_prof;
    const int grid = (int)(B < (int64_t)p->n_sms ? B : (int64_t)p->n_sms);
    if (p->scratch_complex) {
        if (int rc = plan_scratch(p, (cudaStream_t)stream, &kp.gscratch)) { cudaFree(d_prof); return rc; }
        kp.gscratch_stride = p->scratch_complex;
        scat1d_kernel<true, true><<<grid, p->desc.n_threads, p->smem_bytes, (cudaStream_t)stream>>>(kp, x_dev, S_dev, (long long)B);
    } else
    scat1d_kernel<true><<<grid, p->desc.n_threads, p->smem_bytes, (cudaStream_t)stream>>>(kp, x_dev, S_dev, (long long)B);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaMemcpy(step_clocks_host, d_prof, n * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(d_prof);
    if (e != cudaSuccess) return fail(TEBSCAT_ECUDA, "profile run failed: %s", cudaGetErrorString(e));
    ++g_launches;
    return TEBSCAT_OK;
}

#ifdef TEBSCAT_PROF_BFLY
extern "C" int tebscat_debug_bfly(long long* out8, int reset) {
    if (reset) { long long z[8] = {0}; cudaMemcpyToSymbol(tebscat::g_bfly_dbg, z, sizeof(z)); return 0; }
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out8, tebscat::g_bfly_dbg, 8 * sizeof(long long));
    return 0;
}
#endif

static int host_pipe_prepare(tebscat_plan* p) {
    HostPipe& hp = p->pipe;
    if (hp.ready) return TEBSCAT_OK;
    const size_t in_f = (size_t)p->desc.N, out_f = (size_t)p->desc.n_paths * p->desc.n_out;
    hp.chunk = (int64_t)p->n_sms * 8;              // 8 signals per SM per chunk (H: 22.7 MB in, 44.8 MB out)
    CU(cudaStreamCreateWithFlags(&hp.s_in, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&hp.s_run, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&hp.s_out, cudaStreamNonBlocking));
    for (int i = 0; i < HostPipe::kSlots; ++i) {
        CU(cudaEventCreateWithFlags(&hp.loaded[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&hp.computed[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&hp.drained[i], cudaEventDisableTiming));
        CU(cudaMalloc(&hp.d_x[i], hp.chunk * in_f * sizeof(float)));
        CU(cudaMalloc(&hp.d_S[i], hp.chunk * out_f * sizeof(float)));
    }
    hp.ready = true;
    return TEBSCAT_OK;
}

// `copies_only` != 0 runs the same pipeline without the kernel: the host <-> device copy ceiling of this rank for
// exactly the traffic of the real call (bench.py reports the end-to-end rate as a fraction of it).
static int forward_host_impl(tebscat_plan* p, const float* x_host, int64_t B, float* S_host, int copies_only) {
    g_launches = 0;
    if (!p || B < 0 || (B > 0 && (!x_host || !S_host))) return fail(TEBSCAT_EINVAL, "null argument");
    if (B == 0) return TEBSCAT_OK;
    std::lock_guard<std::mutex> lock(p->pipe_mu);
    ON_DEVICE(p->device);
    if (int rc = host_pipe_prepare(p)) return rc;
    HostPipe& hp = p->pipe;
    const size_t in_f = (size_t)p->desc.N, out_f = (size_t)p->desc.n_paths * p->desc.n_out;
    // Chunk sizes: the kernels of consecutive chunks run back to back, so what the copies add to the call is the H2D
    // of the FIRST chunk and the D2H of the LAST one.  Long batches therefore ramp up (1, 2, 4 signals per SM) to the
    // full chunk and down again at the end: 0.15 ms of exposed copies instead of 1.2 ms at the headline configuration.
    std::vector<int64_t> sizes;
    {
        std::vector<int64_t> ramp;
        int64_t ramp_total = 0;
        for (int64_t sz = p->n_sms; sz < hp.chunk; sz *= 2) { ramp.push_back(sz); ramp_total += sz; }
        if (B >= 2 * ramp_total + 2 * hp.chunk) {
            sizes = ramp;
            int64_t mid = B - 2 * ramp_total;
            if (mid % hp.chunk) { sizes.push_back(mid % hp.chunk); mid -= mid % hp.chunk; }
            for (; mid > 0; mid -= hp.chunk) sizes.push_back(hp.chunk);
            sizes.insert(sizes.end(), ramp.rbegin(), ramp.rend());
        } else {
            for (int64_t b0 = 0; b0 < B; b0 += hp.chunk) sizes.push_back(B - b0 < hp.chunk ? B - b0 : hp.chunk);
        }
    }
    int64_t b0 = 0;
    for (size_t n = 0; n < sizes.size(); b0 += sizes[n], ++n) {
        const int slot = (int)(n % HostPipe::kSlots);
        const int64_t nb = sizes[n];
        // the slot's input buffer is free once its previous kernel has run, its output buffer once drained
        if (n >= (size_t)HostPipe::kSlots) CU(cudaStreamWaitEvent(hp.s_in, hp.computed[slot], 0));
        CU(cudaMemcpyAsync(hp.d_x[slot], x_host + b0 * in_f, nb * in_f * sizeof(float), cudaMemcpyHostToDevice, hp.s_in));
        CU(cudaEventRecord(hp.loaded[slot], hp.s_in));
        CU(cudaStreamWaitEvent(hp.s_run, hp.loaded[slot], 0));
        if (n >= (size_t)HostPipe::kSlots) CU(cudaStreamWaitEvent(hp.s_run, hp.drained[slot], 0));
        if (!copies_only)
            if (int rc = launch_scat1d(p, hp.d_x[slot], nb, hp.d_S[slot], hp.s_run)) return rc;
        CU(cudaEventRecord(hp.computed[slot], hp.s_run));
        CU(cudaStreamWaitEvent(hp.s_out, hp.computed[slot], 0));
        CU(cudaMemcpyAsync(S_host + b0 * out_f, hp.d_S[slot], nb * out_f * sizeof(float), cudaMemcpyDeviceToHost, hp.s_out));
        CU(cudaEventRecord(hp.drained[slot], hp.s_out));
    }
    CU(cudaStreamSynchronize(hp.s_out));
    CU(cudaStreamSynchronize(hp.s_run));
    CU(cudaStreamSynchronize(hp.s_in));
    return TEBSCAT_OK;
}

extern "C" int tebscat_scat1d_forward_host(tebscat_plan* p, const float* x_host, int64_t B, float* S_host) {
    return forward_host_impl(p, x_host, B, S_host, 0);
}

// Diagnostic (bench.py, tools/host_copy_ceiling.py): the copies of tebscat_scat1d_forward_host without its kernel.
extern "C" int tebscat_scat1d_host_copies_only(tebscat_plan* p, const float* x_host, int64_t B, float* S_host) {
    return forward_host_impl(p, x_host, B, S_host, 1);
}

// =================================================================================
// Phase-harmonic correlation (hdf5_dataset/kymatio_phase_scattering.py:211-360)
// =================================================================================
//
// Stage A (per sample and channel) runs on the step interpreter above with a schedule whose
// leaves are STOREZ tasks: z_f = ifft(fft(pad(x)) psi1_f)[pad_left : pad_left+N] for every
// first-order filter f, written to an L2-resident workspace as (re, im) for the 'j' channel
// and as (|z|, theta) for the 'i' channel.
//
// Stage B is one kernel.  For every (sample, pair (i, j)) it forms the phase-accelerated
// product  c(t) = |z_i| exp(i p theta_i) conj(z_j)  (:211-218, :283 / :339) on the fly and
// contracts it with the precomputed operator of _apply_phi_filter (:233-273)
//     out[n] = Re sum_t c(t) G[t, n],
// G = reflect-pad -> FFT(Np) -> phi -> keep bins [0, Np/dec) -> iFFT(Np/dec) -> slice,
// built once on the host in float64.  The contraction is a GEMM (128 (sample, pair) rows x 80
// output samples per CTA, K = time in slabs of 16) on the tensor cores in 3xTF32; rows never
// materialise in memory -- the reference materialises (B, P, N) complex64 three times.

constexpr int kPR = 128;        // rows per CTA
constexpr int kPC = 80;         // output columns per CTA
constexpr int kPK = 16;         // time samples per slab
constexpr int kPThreads = 256;

struct PairParams {
    const float2* zp;           // [Bc][F][N] (|z|, theta) of the 'i' channel
    const float2* zc;           // [Bc][F][N] (re, im)     of the 'j' channel
    const float2* G;            // [N][n_cols_pad] (re, im) of the smoothing operator
    const int32_t* i_idx;
    const int32_t* j_idx;
    const float* powers;
    const int32_t* subset;      // selected pairs or null
    float* out;                 // [rows][n_out]
    long long rows;             // Bc * n_sel
    int32_t n_sel, F, N, n_out, n_cols_pad;
};

// (accelerated_product: scat_core.cuh -- shared with the interpreter's OP_LOADPAIR)

// ---- tensor-core contraction ------------------------------------------------------------
// out = A' B' with A' = [Re c, -Im c] (rows x 2N) and B' = [Re G; Im G] (2N x n_out) is a real
// GEMM.  It runs on the tensor cores as 3xTF32: every fp32 operand is split into a TF32 head and a
// TF32 tail, and hi*hi + lo*hi + hi*lo is accumulated in fp32 (relative error ~2^-21 per product,
// i.e. fp32-class accuracy; a plain TF32 product would be 1e-3).  mma.sync.m16n8k8 issues at
// 1 instruction/cycle/SM on B200 (tools/micro/mma_rate.cu): 991 MAC/cycle/SM, 330 after the
// three-way split, vs 128 FMA/cycle/SM on the FP32 pipe -- and the phase-acceleration math of
// the other resident CTA overlaps it on the FP32 pipe.
// Layout of the work inside a CTA (128 rows x 80 columns, 8 warps, K in slabs of 16 samples):
//  * warp w owns rows 16w..16w+15.  Its A' fragments never touch shared memory: with the K
//    index of an 8-wide MMA step ordered [Re c(t..t+3), -Im c(t..t+3)], lane (g, tq) needs exactly
//    c(row g, t+tq) and c(row g+8, t+tq) -- it computes those two products itself;
//  * B' is staged once per slab in "fragment-major" order, already split into TF32 head and
//    tail: one 128-bit load gives a lane (b0, b1) = (Re G, Im G)[t+tq][8n+g] as (hi, hi, lo, lo);
//  * the MMAs of a slab accumulate into a zeroed fragment that is then added to the running
//    sum with ordinary fp32 adds: the tensor core's accumulator truncates, and over the
//    K = 2N = 9600 of a whole row that bias would reach 5e-5.
constexpr int kBLane = 44;      // words per lane slot of the B' slab (10 tiles x 4 + 4 pad: conflict free)
constexpr size_t kPairSmem = (size_t)(kPK / 4) * 32 * kBLane * sizeof(uint32_t) + kPR * (2 * sizeof(int32_t) + sizeof(float));

__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
    const float r = v - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kPThreads, 2) phase_pair_kernel(const PairParams p) {
    extern __shared__ __align__(16) uint32_t psm[];
    uint32_t* sB = psm;                                 // [ks][lane][kBLane]
    int32_t* s_zp = reinterpret_cast<int32_t*>(sB + (kPK / 4) * 32 * kBLane);
    int32_t* s_zc = s_zp + kPR;
    float* s_pw = reinterpret_cast<float*>(s_zc + kPR);

    const int tid = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * kPR;
    const int col0 = blockIdx.y * kPC;
    if (tid < kPR) {
        const long long r = row0 + tid;
        int zp_off = -1, zc_off = -1;
        float pw = 1.f;
        if (r < p.rows) {
            const int b = (int)(r / p.n_sel), s = (int)(r - (long long)b * p.n_sel);
            const int pair = p.subset ? p.subset[s] : s;
            zp_off = (b * p.F + p.i_idx[pair]) * p.N;
            zc_off = (b * p.F + p.j_idx[pair]) * p.N;
            pw = p.powers[pair];
        }
        s_zp[tid] = zp_off;
        s_zc[tid] = zc_off;
        s_pw[tid] = pw;
    }
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
    // the two rows whose products this lane generates
    const int zp0 = s_zp[warp * 16 + g], zc0 = s_zc[warp * 16 + g];
    const int zp1 = s_zp[warp * 16 + g + 8], zc1 = s_zc[warp * 16 + g + 8];
    const float pw0 = s_pw[warp * 16 + g], pw1 = s_pw[warp * 16 + g + 8];

    float total[kPC / 8][4];
#pragma unroll
    for (int n = 0; n < kPC / 8; ++n)
#pragma unroll
        for (int c = 0; c < 4; ++c) total[n][c] = 0.f;

    // raw inputs of the products, fetched one MMA step ahead (software pipeline: the loads of
    // step k+1 are in flight while the 30 MMAs of step k issue)
    float2 r_zp0, r_zc0, r_zp1, r_zc1;
    auto fetch = [&](int t) {
        const bool in = t < p.N;
        r_zp0 = (in && zp0 >= 0) ? __ldg(p.zp + zp0 + t) : make_float2(0.f, 0.f);
        r_zc0 = (in && zp0 >= 0) ? __ldg(p.zc + zc0 + t) : make_float2(0.f, 0.f);
        r_zp1 = (in && zp1 >= 0) ? __ldg(p.zp + zp1 + t) : make_float2(0.f, 0.f);
        r_zc1 = (in && zp1 >= 0) ? __ldg(p.zc + zc1 + t) : make_float2(0.f, 0.f);
    };
    fetch(tq);

    for (int t0 = 0; t0 < p.N; t0 += kPK) {
        // stage B': G[t0 + tl][col0 + col] -> slot (ks = tl/4, tq = tl%4, g = col%8), tile n = col/8
#pragma unroll
        for (int pass = 0; pass < kPK * kPC / kPThreads; ++pass) {
            const int e = tid + pass * kPThreads;
            const int tl = e / kPC, col = e - tl * kPC;
            const int t = t0 + tl;
            const float2 gv = (t < p.N) ? __ldg(p.G + (long long)t * p.n_cols_pad + col0 + col) : make_float2(0.f, 0.f);
            uint4 w;
            split_tf32(gv.x, w.x, w.z);
            split_tf32(gv.y, w.y, w.w);
            *reinterpret_cast<uint4*>(sB + ((tl >> 2) * 32 + (col & 7) * 4 + (tl & 3)) * kBLane + (col >> 3) * 4) = w;
        }
        __syncthreads();
        float acc[kPC / 8][4];
#pragma unroll
        for (int n = 0; n < kPC / 8; ++n)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[n][c] = 0.f;
#pragma unroll
        for (int ks = 0; ks < kPK / 4; ++ks) {
            // zero magnitude (padding rows / samples beyond N) gives a zero product
            const float2 c0 = accelerated_product(r_zp0, r_zc0, pw0);
            const float2 c1 = accelerated_product(r_zp1, r_zc1, pw1);
            fetch(t0 + 4 * (ks + 1) + tq);              // next step (possibly of the next slab)
            // A' fragment: a0 = (row g, Re), a1 = (row g+8, Re), a2 = (row g, -Im), a3 = (row g+8, -Im)
            uint32_t a_hi[4], a_lo[4];
            split_tf32(c0.x, a_hi[0], a_lo[0]);
            split_tf32(c1.x, a_hi[1], a_lo[1]);
            split_tf32(-c0.y, a_hi[2], a_lo[2]);
            split_tf32(-c1.y, a_hi[3], a_lo[3]);
            const uint4* bq = reinterpret_cast<const uint4*>(sB + (ks * 32 + lane) * kBLane);
#pragma unroll
            for (int n = 0; n < kPC / 8; ++n) {
                const uint4 b = bq[n];                  // (hi Re, hi Im, lo Re, lo Im) of G[t][8n + g]
                mma_tf32(acc[n], a_lo, b.x, b.y);
                mma_tf32(acc[n], a_hi, b.z, b.w);
                mma_tf32(acc[n], a_hi, b.x, b.y);
            }
        }
#pragma unroll
        for (int n = 0; n < kPC / 8; ++n)
#pragma unroll
            for (int c = 0; c < 4; ++c) total[n][c] += acc[n][c];
        __syncthreads();
    }
    // accumulator fragment: rows g, g+8 of the warp's 16; columns 8n + 2*tq, +1
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const long long row = row0 + warp * 16 + g + 8 * half;
        if (row >= p.rows) continue;
#pragma unroll
        for (int n = 0; n < kPC / 8; ++n) {
            const int col = col0 + 8 * n + 2 * tq;
            if (col < p.n_out) p.out[row * p.n_out + col] = total[n][2 * half];
            if (col + 1 < p.n_out) p.out[row * p.n_out + col + 1] = total[n][2 * half + 1];
        }
    }
}

#include "phase_tc.cuh"

// cross_phase_low_pass=False (:356-360): the full-rate real part of the product
__global__ void phase_product_kernel(const PairParams p) {
    const long long total = p.rows * p.N;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / p.N;
        const int t = (int)(e - r * p.N);
        const int b = (int)(r / p.n_sel), s = (int)(r - (long long)b * p.n_sel);
        const int pair = p.subset ? p.subset[s] : s;
        const float2 c = accelerated_product(__ldg(p.zp + (long long)(b * p.F + p.i_idx[pair]) * p.N + t),
                                             __ldg(p.zc + (long long)(b * p.F + p.j_idx[pair]) * p.N + t), p.powers[pair]);
        p.out[e] = c.x;
    }
}

// samples per workspace chunk: two stage-A jobs per SM and launch (148 SMs), 0.86 GB of
// analytic signals at N=4800, F=38
constexpr int64_t kPhaseChunk = 296;

struct tebscat_phase_plan {
    tebscat_phase_desc desc;
    tebscat_plan* stage_a = nullptr;
    int device = 0;
    float2* d_G = nullptr;
    float* d_Bs = nullptr;          // tcgen05 form: shared-memory images of the slabs of B' (TF32 head / tail, phase_tc.cuh)
    int32_t n_slabs = 0, k_pad = 0;
    int32_t* d_i = nullptr;
    int32_t* d_j = nullptr;
    float* d_pw = nullptr;
    int32_t* d_subset = nullptr;
    int32_t* d_subset2 = nullptr;
    float2* d_zc = nullptr;
    float2* d_zp = nullptr;
    float2* d_zc2 = nullptr;       // second cartesian workspace of the dataset entry point
    int64_t ws_samples = 0;
    int64_t ws2_samples = 0;
    tebscat_plan* pair_plan = nullptr;   // optional: stage B as FFTs on the interpreter (owned)
    std::mutex mu;
    // The workspaces (analytic signals, subset tables) belong to the plan: a call on another stream than the previous
    // one waits for that call's work first (calls on one stream are ordered by the stream itself).
    cudaEvent_t last_done = nullptr;
    cudaStream_t last_stream = nullptr;
    bool used = false;
    // optional stage timing (tebscat_phase_plan_profile): events around stage A and stage B of every chunk
    bool prof = false;
    std::vector<cudaEvent_t> prof_ev;    // triples (before A, between, after B)
};

static int phase_plan_create_impl(const tebscat_phase_desc* d, tebscat_plan* stage_a, int device, const float* G_host,
                                  const int32_t* i_idx, const int32_t* j_idx, const float* powers, tebscat_phase_plan** out);

extern "C" int tebscat_phase_plan_create(const tebscat_phase_desc* d, tebscat_plan* stage_a, const float* G_host,
                                         const int32_t* i_idx, const int32_t* j_idx, const float* powers,
                                         tebscat_phase_plan** out) {
    if (!stage_a) return fail(TEBSCAT_EINVAL, "null argument");
    if (stage_a->gsrc_extent) return fail(TEBSCAT_EINVAL, "a stage-A schedule cannot read a global source");
    return phase_plan_create_impl(d, stage_a, stage_a->device, G_host, i_idx, j_idx, powers, out);
}

// A phase plan WITHOUT a stage-A plan: for padded lengths above 2^13 the analytic signals come from the ops of the
// large-support level (tebscat_large_*, driven by tebscat/phase.py) and stage B runs on the workspaces the caller
// hands to tebscat_phase_pairs.
extern "C" int tebscat_phase_plan_create_pairs_only(const tebscat_phase_desc* d, int device, const float* G_host,
                                                    const int32_t* i_idx, const int32_t* j_idx, const float* powers,
                                                    tebscat_phase_plan** out) {
    if (!G_host) return fail(TEBSCAT_EINVAL, "a pairs-only plan needs the dense operator");
    int n_dev = 0;
    CU(cudaGetDeviceCount(&n_dev));
    if (device < 0 || device >= n_dev) return fail(TEBSCAT_EINVAL, "device %d not in [0,%d)", device, n_dev);
    return phase_plan_create_impl(d, nullptr, device, G_host, i_idx, j_idx, powers, out);
}

static int phase_plan_create_impl(const tebscat_phase_desc* d, tebscat_plan* stage_a, int device, const float* G_host,
                                  const int32_t* i_idx, const int32_t* j_idx, const float* powers, tebscat_phase_plan** out) {
    // G_host may be NULL: the plan then has no dense form of stage B and needs an attached pair plan
    // (configurations without decimation, where the dense operator would be N x N)
    if (!d || !i_idx || !j_idx || !powers || !out) return fail(TEBSCAT_EINVAL, "null argument");
    if (d->abi_version != TEBSCAT_ABI_VERSION) return fail(TEBSCAT_EINVAL, "ABI version mismatch");
    if (d->N < 2 || d->n_filters < 1) return fail(TEBSCAT_EINVAL, "bad phase description");
    if (stage_a && (d->N != stage_a->desc.N || d->n_filters != stage_a->desc.n_paths || stage_a->desc.n_out != d->N))
        return fail(TEBSCAT_EINVAL, "stage-A plan does not match the phase description");
    if (d->n_pairs < 1 || d->n_out < 1 || (G_host && (d->n_cols_pad < d->n_out || d->n_cols_pad % kPC != 0)))
        return fail(TEBSCAT_EINVAL, "bad phase description");
    for (int k = 0; k < d->n_pairs; ++k)
        if (i_idx[k] < 0 || i_idx[k] >= d->n_filters || j_idx[k] < 0 || j_idx[k] >= d->n_filters)
            return fail(TEBSCAT_EINVAL, "pair %d references a filter outside [0,%d)", k, d->n_filters);
    ON_DEVICE(device);
    CU(cudaFuncSetAttribute(phase_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPairSmem));
    std::unique_ptr<tebscat_phase_plan, void (*)(tebscat_phase_plan*)> guard(new tebscat_phase_plan(), [](tebscat_phase_plan* q) {
        q->stage_a = nullptr;                    // on failure the caller keeps the stage-A plan
        tebscat_phase_plan_destroy(q);
    });
    tebscat_phase_plan* p = guard.get();
    p->desc = *d;
    p->stage_a = stage_a;
    p->device = device;
    const size_t g_elems = (size_t)d->N * d->n_cols_pad;
    if (G_host) {
        CU(cudaMalloc(&p->d_G, g_elems * sizeof(float2)));
        CU(cudaMemcpy(p->d_G, G_host, g_elems * sizeof(float2), cudaMemcpyHostToDevice));
    }
    if (G_host) {   // B'[n][2t + c] = G[t][n][c], split into TF32 head and tail (round to nearest, ties away, like cvt.rna.tf32.f32)
        p->n_slabs = (d->N + kTcSlabT - 1) / kTcSlabT;
        p->k_pad = p->n_slabs * kTcK;
        const size_t per = (size_t)d->n_cols_pad * p->k_pad;
        std::vector<float> bs(2 * per, 0.f);
        auto rna = [](float v) {
            uint32_t u;
            memcpy(&u, &v, 4);
            u = (u + 0x1000u) & 0xffffe000u;
            float r;
            memcpy(&r, &u, 4);
            return r;
        };
        for (int t = 0; t < d->N; ++t)
            for (int n = 0; n < d->n_cols_pad; ++n)
                for (int c = 0; c < 2; ++c) {
                    const float v = G_host[((size_t)t * d->n_cols_pad + n) * 2 + c];
                    const float hi = rna(v);
                    bs[(size_t)n * p->k_pad + 2 * t + c] = hi;
                    bs[per + (size_t)n * p->k_pad + 2 * t + c] = rna(v - hi);
                }
        // the kernel copies a slab of B' into shared memory with ONE bulk copy: lay the slabs out as their shared-memory
        // image -- [column tile][slab][head, tail][80 rows x 128 bytes, 16-byte chunk c of row n at c ^ (n & 7)]
        const int n_tiles = d->n_cols_pad / kTcCols, stage_f = kTcStageBytes / 4, tile_f = kTcBTile / 4;
        std::vector<float> img((size_t)n_tiles * p->n_slabs * stage_f, 0.f);
        for (int ct = 0; ct < n_tiles; ++ct)
            for (int i = 0; i < p->n_slabs; ++i) {
                float* dst = img.data() + ((size_t)ct * p->n_slabs + i) * stage_f;
                for (int part = 0; part < 2; ++part)
                    for (int n = 0; n < kTcCols; ++n)
                        for (int c = 0; c < 8; ++c)
                            for (int e = 0; e < 4; ++e)
                                dst[part * tile_f + (n >> 3) * 256 + (n & 7) * 32 + ((c ^ (n & 7)) << 2) + e] =
                                    bs[part * per + (size_t)(ct * kTcCols + n) * p->k_pad + i * kTcK + 4 * c + e];
            }
        CU(cudaMalloc(&p->d_Bs, img.size() * sizeof(float)));
        CU(cudaMemcpy(p->d_Bs, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice));
        CU(cudaFuncSetAttribute(phase_pair_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
    }
    CU(cudaMalloc(&p->d_i, d->n_pairs * sizeof(int32_t)));
    CU(cudaMalloc(&p->d_j, d->n_pairs * sizeof(int32_t)));
    CU(cudaMalloc(&p->d_pw, d->n_pairs * sizeof(float)));
    CU(cudaMalloc(&p->d_subset, d->n_pairs * sizeof(int32_t)));
    CU(cudaMalloc(&p->d_subset2, d->n_pairs * sizeof(int32_t)));
    CU(cudaMemcpy(p->d_i, i_idx, d->n_pairs * sizeof(int32_t), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(p->d_j, j_idx, d->n_pairs * sizeof(int32_t), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(p->d_pw, powers, d->n_pairs * sizeof(float), cudaMemcpyHostToDevice));
    *out = guard.release();
    return TEBSCAT_OK;
}

extern "C" void tebscat_phase_plan_destroy(tebscat_phase_plan* p) {
    if (!p) return;
    DeviceGuard device_guard_(p->device);
    cudaFree(p->d_G);
    cudaFree(p->d_Bs);
    cudaFree(p->d_i);
    cudaFree(p->d_j);
    cudaFree(p->d_pw);
    cudaFree(p->d_subset);
    cudaFree(p->d_subset2);
    cudaFree(p->d_zc);
    cudaFree(p->d_zp);
    cudaFree(p->d_zc2);
    for (cudaEvent_t e : p->prof_ev) cudaEventDestroy(e);
    if (p->last_done) cudaEventDestroy(p->last_done);
    tebscat_plan_destroy(p->stage_a);
    tebscat_plan_destroy(p->pair_plan);
    delete p;
}

// the window acts where the samples enter: stage A's loads (both channels go through the same stage-A plan)
extern "C" int tebscat_phase_plan_set_window(tebscat_phase_plan* p, const float* window_host) {
    if (!p) return fail(TEBSCAT_EINVAL, "null plan");
    if (!p->stage_a) return fail(TEBSCAT_EUNSUPPORTED, "this phase plan has no stage A (its loads belong to a tebscat_large context)");
    std::lock_guard<std::mutex> lock(p->mu);
    return tebscat_plan_set_window(p->stage_a, window_host);
}

// The dense form of stage B: tcgen05 + TMEM (phase_tc.cuh) unless TEBSCAT_PHASE_MMA=sync asks for the mma.sync kernel.
static bool use_tcgen05_pairs() {
    const char* e = getenv("TEBSCAT_PHASE_MMA");          // read per call: the tests switch it
    return !(e && strcmp(e, "sync") == 0);
}
static void launch_pair_gemm(const tebscat_phase_plan* p, const PairParams& pp, cudaStream_t st) {
    const tebscat_phase_desc& d = p->desc;
    // (the tcgen05 kernel keeps 32-bit sample offsets into the workspace chunk)
    const long long ws_elems = (pp.rows / pp.n_sel + 1) * (long long)pp.F * pp.N;
    if (use_tcgen05_pairs() && ws_elems < (1LL << 31)) {
        PairTcParams q;
        q.zp = pp.zp; q.zc = pp.zc; q.Bimg = p->d_Bs; q.i_idx = pp.i_idx; q.j_idx = pp.j_idx; q.powers = pp.powers;
        q.subset = pp.subset; q.out = pp.out; q.rows = pp.rows; q.n_sel = pp.n_sel; q.F = pp.F; q.N = pp.N;
        q.n_out = pp.n_out; q.n_cols_pad = pp.n_cols_pad; q.n_slabs = p->n_slabs; q.k_pad = p->k_pad;
        dim3 grid((unsigned)((pp.rows + kTcRows - 1) / kTcRows), (unsigned)(d.n_cols_pad / kTcCols));
        phase_pair_tc_kernel<<<grid, kTcThreads, kTcSmem, st>>>(q);
    } else {
        dim3 grid((unsigned)((pp.rows + kPR - 1) / kPR), (unsigned)(d.n_cols_pad / kPC));
        phase_pair_kernel<<<grid, kPThreads, kPairSmem, st>>>(pp);
    }
}

// Stage B as transforms on the step interpreter: `pair_plan` is a schedule whose jobs are n_paths (up to 8)
// consecutive (sample, pair) rows -- LOADPAIR, forward transform, phi on the kept bins, reduced inverse
// transform, unpad (the literal cascade of _apply_phi_filter :233-273).  5 N log N instead of the dense
// operator's 4 N n_out flops per row: it wins when the output is long (production config: n_out = 360).
// Only for power-of-two decimation; the dense tensor-core kernel stays the general path.
extern "C" int tebscat_phase_plan_attach_pair_plan(tebscat_phase_plan* p, tebscat_plan* pair_plan) {
    if (!p || !pair_plan) return fail(TEBSCAT_EINVAL, "null argument");
    if (pair_plan->device != p->device || pair_plan->desc.N != p->desc.N || pair_plan->desc.n_out != p->desc.n_out ||
        pair_plan->desc.n_paths < 1 || pair_plan->desc.n_paths > kMaxPairRows)
        return fail(TEBSCAT_EINVAL, "pair plan does not match the phase description");
    if (pair_plan->gsrc_extent) return fail(TEBSCAT_EINVAL, "a pair-stage schedule cannot read a global source");
    std::lock_guard<std::mutex> lock(p->mu);
    tebscat_plan_destroy(p->pair_plan);
    p->pair_plan = pair_plan;
    return TEBSCAT_OK;
}

static int launch_pairs_fft(const tebscat_phase_plan* p, const float2* zp, const float2* zc, const int32_t* subset_dev,
                            int n_sel, int64_t nb, float* out, cudaStream_t st) {
    const tebscat_plan* a = p->pair_plan;
    KParams kp = a->kp;
    kp.pair_zp = zp;
    kp.pair_zc = zc;
    kp.pair_i = p->d_i;
    kp.pair_j = p->d_j;
    kp.pair_pw = p->d_pw;
    kp.pair_subset = subset_dev;
    kp.pair_rows = (long long)nb * n_sel;
    kp.pair_n_sel = n_sel;
    kp.pair_F = p->desc.n_filters;
    const long long jobs = (kp.pair_rows + a->desc.n_paths - 1) / a->desc.n_paths;
    const int grid = (int)(jobs < (long long)a->n_sms ? jobs : (long long)a->n_sms);
    scat1d_kernel<false><<<grid, a->desc.n_threads, a->smem_bytes, st>>>(kp, nullptr, out, jobs);
    CU(cudaGetLastError());
    ++g_launches;
    return TEBSCAT_OK;
}

// order a call on `st` behind the plan's previous call if that ran on another stream; `phase_call_done` marks the end
static int phase_call_begin(tebscat_phase_plan* p, cudaStream_t st) {
    if (!p->last_done) CU(cudaEventCreateWithFlags(&p->last_done, cudaEventDisableTiming));
    if (p->used && p->last_stream != st) CU(cudaStreamWaitEvent(st, p->last_done, 0));
    return TEBSCAT_OK;
}
struct PhaseCallScope {                       // declared after the plan's lock: records the end of the call on every exit path
    tebscat_phase_plan* p;
    cudaStream_t st;
    ~PhaseCallScope() {
        if (p->last_done && cudaEventRecord(p->last_done, st) == cudaSuccess) {
            p->last_stream = st;
            p->used = true;
        }
    }
};

static int launch_stage_a(const tebscat_plan* a, const float* x, long long x_stride, int64_t jobs,
                          float2* zc, float2* zp, int z_mode, cudaStream_t st) {
    KParams kp = a->kp;
    kp.x_stride = x_stride;
    kp.zc = zc;
    kp.zp = zp;
    kp.z_stride = (long long)a->desc.n_paths * a->desc.N;
    kp.z_mode = z_mode;
    const int grid = (int)(jobs < (int64_t)a->n_sms ? jobs : (int64_t)a->n_sms);
    scat1d_kernel<false><<<grid, a->desc.n_threads, a->smem_bytes, st>>>(kp, x, nullptr, (long long)jobs);
    CU(cudaGetLastError());
    ++g_launches;
    return TEBSCAT_OK;
}

extern "C" int tebscat_phase_forward(tebscat_phase_plan* p, const float* x_dev, int64_t B, int n_channels,
                                     int ch_i, int ch_j, const int32_t* pair_subset_host, int n_subset,
                                     int apply_low_pass, float* out_dev, void* stream) {
    g_launches = 0;
    if (!p || B < 0 || (B > 0 && (!x_dev || !out_dev))) return fail(TEBSCAT_EINVAL, "null argument");
    if (n_channels < 1 || ch_i < 0 || ch_i >= n_channels || ch_j < 0 || ch_j >= n_channels)
        return fail(TEBSCAT_EINVAL, "channel index outside [0,%d)", n_channels);
    if (pair_subset_host && (n_subset < 1 || n_subset > p->desc.n_pairs)) return fail(TEBSCAT_EINVAL, "bad pair subset size");
    if (pair_subset_host)
        for (int k = 0; k < n_subset; ++k)
            if (pair_subset_host[k] < 0 || pair_subset_host[k] >= p->desc.n_pairs)
                return fail(TEBSCAT_EINVAL, "pair subset entry %d outside [0,%d)", k, p->desc.n_pairs);
    if (B == 0) return TEBSCAT_OK;
    if (!p->stage_a) return fail(TEBSCAT_EUNSUPPORTED, "this phase plan has no stage A: use tebscat_phase_pairs on analytic signals");
    std::lock_guard<std::mutex> lock(p->mu);
    cudaStream_t st = (cudaStream_t)stream;
    ON_DEVICE(p->device);
    if (int rc = phase_call_begin(p, st)) return rc;
    PhaseCallScope call_scope{p, st};
    const tebscat_phase_desc& d = p->desc;
    const int n_sel = pair_subset_host ? n_subset : d.n_pairs;
    const size_t per_sample = (size_t)d.n_filters * d.N;
    // the dense form of stage B runs over three stage-A chunks at once: 888 samples x 741 pairs are 34.7 waves of row tiles
    // on 148 SMs (99 % of the last wave used) where 296 samples are 11.6 (the twelfth wave 58 % full)
    const int64_t span = (apply_low_pass && !p->pair_plan && p->d_G) ? 3 * kPhaseChunk : kPhaseChunk;
    const int64_t chunk = B < span ? B : span;
    if (p->ws_samples < chunk) {
        CU(cudaStreamSynchronize(st));
        cudaFree(p->d_zc);
        cudaFree(p->d_zp);
        p->d_zc = p->d_zp = nullptr;
        CU(cudaMalloc(&p->d_zc, (size_t)chunk * per_sample * sizeof(float2)));
        CU(cudaMalloc(&p->d_zp, (size_t)chunk * per_sample * sizeof(float2)));
        p->ws_samples = chunk;
    }
    if (pair_subset_host) CU(cudaMemcpyAsync(p->d_subset, pair_subset_host, n_sel * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    const long long x_stride = (long long)n_channels * d.N;
    const size_t out_per_sample = (size_t)n_sel * (apply_low_pass ? d.n_out : d.N);
    auto mark = [&]() {                          // stage timing, when asked for
        if (!p->prof) return;
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        cudaEventRecord(e, st);
        p->prof_ev.push_back(e);
    };
    for (int64_t b0 = 0; b0 < B; b0 += chunk) {
        const int64_t nb = (B - b0 < chunk) ? (B - b0) : chunk;
        const float* xb = x_dev + b0 * x_stride;
        mark();
        for (int64_t a0 = 0; a0 < nb; a0 += kPhaseChunk) {            // stage A: two jobs per SM and launch
            const int64_t na = (nb - a0 < kPhaseChunk) ? (nb - a0) : kPhaseChunk;
            const float* xa = xb + a0 * x_stride;
            float2* zc = p->d_zc + (size_t)a0 * per_sample;
            float2* zp = p->d_zp + (size_t)a0 * per_sample;
            if (ch_i == ch_j) {
                if (int rc = launch_stage_a(p->stage_a, xa + (size_t)ch_i * d.N, x_stride, na, zc, zp, Z_CART | Z_POLAR, st)) return rc;
            } else {
                if (int rc = launch_stage_a(p->stage_a, xa + (size_t)ch_i * d.N, x_stride, na, zc, zp, Z_POLAR, st)) return rc;
                if (int rc = launch_stage_a(p->stage_a, xa + (size_t)ch_j * d.N, x_stride, na, zc, zp, Z_CART, st)) return rc;
            }
        }
        mark();
        PairParams pp;
        pp.zp = p->d_zp;
        pp.zc = p->d_zc;
        pp.G = p->d_G;
        pp.i_idx = p->d_i;
        pp.j_idx = p->d_j;
        pp.powers = p->d_pw;
        pp.subset = pair_subset_host ? p->d_subset : nullptr;
        pp.out = out_dev + (size_t)b0 * out_per_sample;
        pp.rows = (long long)nb * n_sel;
        pp.n_sel = n_sel;
        pp.F = d.n_filters;
        pp.N = d.N;
        pp.n_out = d.n_out;
        pp.n_cols_pad = d.n_cols_pad;
        if (apply_low_pass && p->pair_plan) {
            if (int rc = launch_pairs_fft(p, p->d_zp, p->d_zc, pp.subset, n_sel, nb, pp.out, st)) return rc;
            mark();
            continue;
        } else if (apply_low_pass && !p->d_G) {
            return fail(TEBSCAT_EUNSUPPORTED, "this phase plan has no dense operator and no pair plan attached");
        } else if (apply_low_pass) {
            launch_pair_gemm(p, pp, st);
        } else {
            phase_product_kernel<<<1184, 256, 0, st>>>(pp);
        }
        CU(cudaGetLastError());
        ++g_launches;
        mark();
    }
    return TEBSCAT_OK;
}

// Stage timing of tebscat_phase_forward (bench.py's roofline of the phase metric): while enabled, every chunk records
// CUDA events before stage A, between the stages and after stage B on the caller's stream.
extern "C" int tebscat_phase_plan_profile(tebscat_phase_plan* p, int enable) {
    if (!p) return fail(TEBSCAT_EINVAL, "null plan");
    std::lock_guard<std::mutex> lock(p->mu);
    for (cudaEvent_t e : p->prof_ev) cudaEventDestroy(e);
    p->prof_ev.clear();
    p->prof = enable != 0;
    return TEBSCAT_OK;
}
// Sum of the recorded stage times in milliseconds since profiling was enabled (waits for the last event); clears them.
extern "C" int tebscat_phase_plan_profile_read(tebscat_phase_plan* p, double* stage_a_ms, double* stage_b_ms, int* n_chunks) {
    if (!p || !stage_a_ms || !stage_b_ms || !n_chunks) return fail(TEBSCAT_EINVAL, "null argument");
    std::lock_guard<std::mutex> lock(p->mu);
    *stage_a_ms = *stage_b_ms = 0.0;
    *n_chunks = (int)(p->prof_ev.size() / 3);
    if (!p->prof_ev.empty()) CU(cudaEventSynchronize(p->prof_ev.back()));
    for (size_t k = 0; k + 2 < p->prof_ev.size(); k += 3) {
        float a = 0.f, b = 0.f;
        CU(cudaEventElapsedTime(&a, p->prof_ev[k], p->prof_ev[k + 1]));
        CU(cudaEventElapsedTime(&b, p->prof_ev[k + 1], p->prof_ev[k + 2]));
        *stage_a_ms += a;
        *stage_b_ms += b;
    }
    for (cudaEvent_t e : p->prof_ev) cudaEventDestroy(e);
    p->prof_ev.clear();
    return TEBSCAT_OK;
}

static int launch_pairs(const tebscat_phase_plan* p, const float2* zp, const float2* zc, const int32_t* subset_dev,
                        int n_sel, int64_t nb, float* out, cudaStream_t st) {
    if (p->pair_plan) return launch_pairs_fft(p, zp, zc, subset_dev, n_sel, nb, out, st);
    if (!p->d_G) return fail(TEBSCAT_EUNSUPPORTED, "this phase plan has no dense operator and no pair plan attached");
    const tebscat_phase_desc& d = p->desc;
    PairParams pp;
    pp.zp = zp;
    pp.zc = zc;
    pp.G = p->d_G;
    pp.i_idx = p->d_i;
    pp.j_idx = p->d_j;
    pp.powers = p->d_pw;
    pp.subset = subset_dev;
    pp.out = out;
    pp.rows = (long long)nb * n_sel;
    pp.n_sel = n_sel;
    pp.F = d.n_filters;
    pp.N = d.N;
    pp.n_out = d.n_out;
    pp.n_cols_pad = d.n_cols_pad;
    launch_pair_gemm(p, pp, st);
    CU(cudaGetLastError());
    ++g_launches;
    return TEBSCAT_OK;
}

// Stage B alone, on analytic signals the caller has produced (padded lengths above 2^13: stage A runs on the ops of the
// large-support level): zp_dev [nb][F][N] (|z|, theta) of the 'i' channel, zc_dev [nb][F][N] (re, im) of the 'j' channel.
extern "C" int tebscat_phase_pairs(tebscat_phase_plan* p, const float* zp_dev, const float* zc_dev, int64_t nb,
                                   const int32_t* pair_subset_host, int n_subset, int apply_low_pass, float* out_dev, void* stream) {
    g_launches = 0;
    if (!p || nb < 0 || (nb > 0 && (!zp_dev || !zc_dev || !out_dev))) return fail(TEBSCAT_EINVAL, "null argument");
    if (pair_subset_host && (n_subset < 1 || n_subset > p->desc.n_pairs)) return fail(TEBSCAT_EINVAL, "bad pair subset size");
    if (pair_subset_host)
        for (int k = 0; k < n_subset; ++k)
            if (pair_subset_host[k] < 0 || pair_subset_host[k] >= p->desc.n_pairs)
                return fail(TEBSCAT_EINVAL, "pair subset entry %d outside [0,%d)", k, p->desc.n_pairs);
    if (nb == 0) return TEBSCAT_OK;
    if ((long long)nb * p->desc.n_filters * p->desc.N >= (1LL << 31))
        return fail(TEBSCAT_EINVAL, "workspace of %lld samples: chunk the batch (32-bit sample offsets)", (long long)nb);
    std::lock_guard<std::mutex> lock(p->mu);
    cudaStream_t st = (cudaStream_t)stream;
    ON_DEVICE(p->device);
    if (int rc = phase_call_begin(p, st)) return rc;
    PhaseCallScope call_scope{p, st};
    const int n_sel = pair_subset_host ? n_subset : p->desc.n_pairs;
    if (pair_subset_host) CU(cudaMemcpyAsync(p->d_subset, pair_subset_host, n_sel * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    const int32_t* sub = pair_subset_host ? p->d_subset : nullptr;
    if (apply_low_pass)
        return launch_pairs(p, reinterpret_cast<const float2*>(zp_dev), reinterpret_cast<const float2*>(zc_dev), sub, n_sel, nb, out_dev, st);
    PairParams pp;
    pp.zp = reinterpret_cast<const float2*>(zp_dev);
    pp.zc = reinterpret_cast<const float2*>(zc_dev);
    pp.G = p->d_G;
    pp.i_idx = p->d_i;
    pp.j_idx = p->d_j;
    pp.powers = p->d_pw;
    pp.subset = sub;
    pp.out = out_dev;
    pp.rows = (long long)nb * n_sel;
    pp.n_sel = n_sel;
    pp.F = p->desc.n_filters;
    pp.N = p->desc.N;
    pp.n_out = p->desc.n_out;
    pp.n_cols_pad = p->desc.n_cols_pad;
    phase_product_kernel<<<1184, 256, 0, st>>>(pp);
    CU(cudaGetLastError());
    ++g_launches;
    return TEBSCAT_OK;
}

// Single-pass dataset entry: within-channel correlations of channel ch_i for the
// pairs `within_subset` AND cross-channel correlations ch_i x ch_j for `cross_subset`, with the
// analytic signals of ch_i computed once.  Replaces the two st_model(...) calls and the masking
// of hdf5_dataset/create_hdf5_dataset.py:421-441.
extern "C" int tebscat_phase_forward_dual(tebscat_phase_plan* p, const float* x_dev, int64_t B, int n_channels,
                                          int ch_i, int ch_j, const int32_t* within_subset_host, int n_within,
                                          const int32_t* cross_subset_host, int n_cross,
                                          float* out_within_dev, float* out_cross_dev, void* stream) {
    g_launches = 0;
    if (!p || B < 0 || (B > 0 && (!x_dev || !out_within_dev || !out_cross_dev))) return fail(TEBSCAT_EINVAL, "null argument");
    if (n_channels < 2 || ch_i < 0 || ch_i >= n_channels || ch_j < 0 || ch_j >= n_channels || ch_i == ch_j)
        return fail(TEBSCAT_EINVAL, "the dataset entry point needs two distinct channels in [0,%d)", n_channels);
    if (!p->stage_a) return fail(TEBSCAT_EUNSUPPORTED, "this phase plan has no stage A: use tebscat_phase_pairs on analytic signals");
    if (!within_subset_host || !cross_subset_host || n_within < 1 || n_cross < 1 ||
        n_within + n_cross > 2 * p->desc.n_pairs)
        return fail(TEBSCAT_EINVAL, "bad pair subsets");
    for (int k = 0; k < n_within; ++k)
        if (within_subset_host[k] < 0 || within_subset_host[k] >= p->desc.n_pairs) return fail(TEBSCAT_EINVAL, "within subset entry out of range");
    for (int k = 0; k < n_cross; ++k)
        if (cross_subset_host[k] < 0 || cross_subset_host[k] >= p->desc.n_pairs) return fail(TEBSCAT_EINVAL, "cross subset entry out of range");
    if (n_within > p->desc.n_pairs || n_cross > p->desc.n_pairs) return fail(TEBSCAT_EINVAL, "subset larger than the pair list");
    if (B == 0) return TEBSCAT_OK;
    std::lock_guard<std::mutex> lock(p->mu);
    cudaStream_t st = (cudaStream_t)stream;
    ON_DEVICE(p->device);
    if (int rc = phase_call_begin(p, st)) return rc;
    PhaseCallScope call_scope{p, st};
    const tebscat_phase_desc& d = p->desc;
    const size_t per_sample = (size_t)d.n_filters * d.N;
    const int64_t chunk = B < kPhaseChunk ? B : kPhaseChunk;
    if (p->ws_samples < chunk || p->ws2_samples < chunk) {
        CU(cudaStreamSynchronize(st));
        if (p->ws_samples < chunk) {
            cudaFree(p->d_zc);
            cudaFree(p->d_zp);
            p->d_zc = p->d_zp = nullptr;
            CU(cudaMalloc(&p->d_zc, (size_t)chunk * per_sample * sizeof(float2)));
            CU(cudaMalloc(&p->d_zp, (size_t)chunk * per_sample * sizeof(float2)));
            p->ws_samples = chunk;
        }
        if (p->ws2_samples < chunk) {
            cudaFree(p->d_zc2);
            p->d_zc2 = nullptr;
            CU(cudaMalloc(&p->d_zc2, (size_t)chunk * per_sample * sizeof(float2)));
            p->ws2_samples = chunk;
        }
    }
    // both subsets live in the plan's subset buffer (n_pairs entries each half)
    static_assert(sizeof(int32_t) == 4, "");
    int32_t* sub_w = p->d_subset;
    int32_t* sub_c = p->d_subset2;
    CU(cudaMemcpyAsync(sub_w, within_subset_host, n_within * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(sub_c, cross_subset_host, n_cross * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    const long long x_stride = (long long)n_channels * d.N;
    for (int64_t b0 = 0; b0 < B; b0 += chunk) {
        const int64_t nb = (B - b0 < chunk) ? (B - b0) : chunk;
        const float* xb = x_dev + b0 * x_stride;
        if (int rc = launch_stage_a(p->stage_a, xb + (size_t)ch_i * d.N, x_stride, nb, p->d_zc, p->d_zp, Z_CART | Z_POLAR, st)) return rc;
        if (int rc = launch_stage_a(p->stage_a, xb + (size_t)ch_j * d.N, x_stride, nb, p->d_zc2, p->d_zp, Z_CART, st)) return rc;
        if (int rc = launch_pairs(p, p->d_zp, p->d_zc, sub_w, n_within, nb, out_within_dev + (size_t)b0 * n_within * d.n_out, st)) return rc;
        if (int rc = launch_pairs(p, p->d_zp, p->d_zc2, sub_c, n_cross, nb, out_cross_dev + (size_t)b0 * n_cross * d.n_out, st)) return rc;
    }
    return TEBSCAT_OK;
}


// =================================================================================
// Large-support level: padded lengths above 2^13
// =================================================================================
// Spectra of more than 8192 samples live in global memory (L2/HBM).  The cascade keeps the reference's op order
// (core/scattering1d.py:269-370); every op is one launch over the whole batch:
//   pad+load, transform (tile jobs of the step interpreter; above 8192 samples one global radix pass plus
//   tile jobs), filter multiply + periodisation, modulus, unpad + store.
// Spectra are in bit-reversed order like everywhere else, so the periodisation is a sum of adjacent elements.

constexpr int kLargeMaxLog2 = 17;

struct tebscat_large {
    int device = 0;
    int n_sms = 0;
    // [log2 length][forward / inverse / inverse-modulus-forward][small / large job]: owned.  Large jobs (three
    // 8192-blocks) amortise the per-step latency of the interpreter, small ones keep every SM busy on short buffers.
    tebscat_plan* tile[kLog2TwMax + 1][3][2] = {};
    int tile_slots[kLog2TwMax + 1][3][2] = {};      // complex elements one job of the plan covers
    float2* d_tw[kLargeMaxLog2 + 1] = {};           // W_L^m, m < L, for L = 2^14 .. 2^17
    float* d_win = nullptr;                         // optional analysis window of pad_load / pad_adjoint
    int win_n = 0;
};

extern "C" int tebscat_large_create(int device, tebscat_large** out) {
    if (!out) return fail(TEBSCAT_EINVAL, "null argument");
    int n_dev = 0;
    CU(cudaGetDeviceCount(&n_dev));
    if (device < 0 || device >= n_dev) return fail(TEBSCAT_EINVAL, "device %d not in [0,%d)", device, n_dev);
    ON_DEVICE(device);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    std::unique_ptr<tebscat_large, void (*)(tebscat_large*)> guard(new tebscat_large(), tebscat_large_destroy);
    tebscat_large* g = guard.get();
    g->device = device;
    g->n_sms = prop.multiProcessorCount;
    for (int n = kLog2TwMax + 1; n <= kLargeMaxLog2; ++n) {
        const size_t L = (size_t)1 << n;
        std::vector<float2> tw(L);
        const double w0 = -2.0 * M_PI / (double)L;
        for (size_t m = 0; m < L; ++m) tw[m] = make_float2((float)cos(w0 * (double)m), (float)sin(w0 * (double)m));
        CU(cudaMalloc(&g->d_tw[n], L * sizeof(float2)));
        CU(cudaMemcpy(g->d_tw[n], tw.data(), L * sizeof(float2), cudaMemcpyHostToDevice));
    }
    *out = guard.release();
    return TEBSCAT_OK;
}

extern "C" void tebscat_large_destroy(tebscat_large* g) {
    if (!g) return;
    DeviceGuard device_guard_(g->device);
    for (int n = 0; n <= kLog2TwMax; ++n)
        for (int d = 0; d < 3; ++d)
            for (int z = 0; z < 2; ++z) tebscat_plan_destroy(g->tile[n][d][z]);
    for (int n = 0; n <= kLargeMaxLog2; ++n) cudaFree(g->d_tw[n]);
    cudaFree(g->d_win);
    delete g;
}

// analysis window of pad_load / pad_adjoint (see tebscat_plan_set_window): n floats, or NULL to remove it
extern "C" int tebscat_large_set_window(tebscat_large* g, const float* window_host, int n) {
    if (!g || (window_host && n < 1)) return fail(TEBSCAT_EINVAL, "bad window");
    ON_DEVICE(g->device);
    cudaFree(g->d_win);
    g->d_win = nullptr;
    g->win_n = 0;
    if (!window_host) return TEBSCAT_OK;
    CU(cudaMalloc(&g->d_win, (size_t)n * sizeof(float)));
    CU(cudaMemcpy(g->d_win, window_host, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    g->win_n = n;
    return TEBSCAT_OK;
}

/* Hand a tile plan (schedule.build_tile_plan) for transforms of 2^log2_len samples to the context (it takes ownership). */
extern "C" int tebscat_large_set_tile_plan(tebscat_large* g, int log2_len, int kind, int slots, tebscat_plan* plan) {
    if (!g || !plan || log2_len < 1 || log2_len > kLog2TwMax || kind < 0 || kind > 2 || slots < (1 << log2_len) ||
        (slots & ((1 << log2_len) - 1)))
        return fail(TEBSCAT_EINVAL, "bad tile plan");
    if (plan->device != g->device) return fail(TEBSCAT_EINVAL, "tile plan lives on another device");
    if (plan->gsrc_extent) return fail(TEBSCAT_EINVAL, "a tile schedule cannot read a global source");
    const int z = slots > 8192 ? 1 : 0;
    tebscat_plan_destroy(g->tile[log2_len][kind][z]);
    g->tile[log2_len][kind][z] = plan;
    g->tile_slots[log2_len][kind][z] = slots;
    return TEBSCAT_OK;
}

// reflect padding (torch_backend.py:50-78) + real -> complex, natural order
__global__ void g_pad_load_kernel(const float* __restrict__ x, float2* __restrict__ u, long long B, int N, int pad_left, int log2_Np,
                                  const float* __restrict__ win, int border) {
    const long long total = B << log2_Np;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long b = e >> log2_Np;
        int r = (int)(e - (b << log2_Np)) - pad_left;
        bool inside = true;
        if (border == BORDER_REFLECT) {                      // torch_backend.py:50-78; _pad_signal 'reflect' (pad < N)
            if (r < 0) r = -r;
            if (r >= N) r = 2 * (N - 1) - r;
        } else if (border == BORDER_CIRCULAR) {              // kymatio_phase_scattering.py:162-173 (pad <= N)
            if (r < 0) r += N;
            if (r >= N) r -= N;
        } else {                                             // 'constant': zeros
            inside = r >= 0 && r < N;
        }
        float v = inside ? __ldg(x + b * N + r) : 0.f;
        if (win && inside) v *= __ldg(win + r);
        u[e] = make_float2(v, 0.f);
    }
}

// One radix-R pass (R = 2^logr <= 16) across the R blocks of 8192 samples of every length-2^n transform:
// forward = first decimation-in-frequency pass (natural -> R blocks, twiddled), inverse = last decimation-in-time pass.
// The R values a thread owns go through the register DFTs of the step interpreter (scat_core.cuh: Dft<R, SGN>, the
// same index maps as fft_butterfly); the only table reads are the R - 1 twiddles W_L^(i0 q).
template <int LOGR>
__global__ void g_radix_kernel(float2* __restrict__ buf, long long n_transforms, int n, int inverse, const float2* __restrict__ tw) {
    constexpr int R = 1 << LOGR;
    const int logm = n - LOGR;                               // block length 2^logm (= 8192)
    const long long total = n_transforms << logm;
    const long long L = (long long)1 << n;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long s = e >> logm;
        const int i0 = (int)(e - (s << logm));
        float2* base = buf + s * L + i0;
        float2 v[R];
        if (!inverse) {
#pragma unroll
            for (int j = 0; j < R; ++j) v[j] = base[(long long)j << logm];
            Dft<R, -1>::run(v);                              // register r holds X_q, q = qmap<R>(r)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int q = qmap<R>(r);
                float2 y = v[r];
                if (q != 0) y = cmul(y, __ldg(tw + (long long)i0 * q));                      // times W_L^(i0 q)
                base[(long long)brev<LOGR>(q) << logm] = y;
            }
        } else {
#pragma unroll
            for (int q = 0; q < R; ++q) {
                float2 z = base[(long long)brev<LOGR>(q) << logm];
                if (q != 0) z = cmulc(z, __ldg(tw + (long long)i0 * q));                     // times conj(W_L^(i0 q))
                v[q] = z;
            }
            Dft<R, +1>::run(v);                              // register r holds x_j, j = qmap<R>(r)
#pragma unroll
            for (int r = 0; r < R; ++r) base[(long long)qmap<R>(r) << logm] = v[r];
        }
    }
}

static int launch_tile_jobs(const tebscat_large* g, float2* buf, long long total_elems, int log2_len, int kind, cudaStream_t st) {
    // large jobs only when there are enough of them for two waves over the SMs
    int z = (g->tile[log2_len][kind][1] && total_elems / g->tile_slots[log2_len][kind][1] >= 2LL * g->n_sms) ? 1 : 0;
    if (!g->tile[log2_len][kind][z]) z ^= 1;
    const tebscat_plan* a = g->tile[log2_len][kind][z];
    if (!a) return fail(TEBSCAT_EINVAL, "no tile plan for transforms of 2^%d samples", log2_len);
    KParams kp = a->kp;
    kp.gbuf = buf;
    kp.g_total = total_elems;
    kp.g_slots = g->tile_slots[log2_len][kind][z];
    const long long jobs = (total_elems + kp.g_slots - 1) / kp.g_slots;
    const int grid = (int)(jobs < (long long)a->n_sms ? jobs : (long long)a->n_sms);
    scat1d_kernel<false><<<grid, a->desc.n_threads, a->smem_bytes, st>>>(kp, nullptr, nullptr, jobs);
    CU(cudaGetLastError());
    ++g_launches;
    return TEBSCAT_OK;
}

/* In-place transforms of `n_transforms` contiguous sequences of 2^log2_len complex samples (2 <= len <= 2^17).
 * Forward: natural -> bit-reversed order; inverse: bit-reversed -> natural, unnormalised. */
extern "C" int tebscat_large_fft(tebscat_large* g, float* buf_dev, int64_t n_transforms, int log2_len, int inverse, void* stream) {
    if (!g || !buf_dev || n_transforms < 1 || log2_len < 1 || log2_len > kLargeMaxLog2) return fail(TEBSCAT_EINVAL, "bad transform request");
    ON_DEVICE(g->device);
    cudaStream_t st = (cudaStream_t)stream;
    float2* buf = reinterpret_cast<float2*>(buf_dev);
    const long long total = (long long)n_transforms << log2_len;
    if (log2_len <= kLog2TwMax) return launch_tile_jobs(g, buf, total, log2_len, inverse, st);
    const int logr = log2_len - kLog2TwMax;
    const int blocks = g->n_sms * 8;
    auto radix = [&]() {
        switch (logr) {
            case 1: g_radix_kernel<1><<<blocks, 256, 0, st>>>(buf, n_transforms, log2_len, inverse, g->d_tw[log2_len]); break;
            case 2: g_radix_kernel<2><<<blocks, 256, 0, st>>>(buf, n_transforms, log2_len, inverse, g->d_tw[log2_len]); break;
            case 3: g_radix_kernel<3><<<blocks, 256, 0, st>>>(buf, n_transforms, log2_len, inverse, g->d_tw[log2_len]); break;
            default: g_radix_kernel<4><<<blocks, 256, 0, st>>>(buf, n_transforms, log2_len, inverse, g->d_tw[log2_len]); break;
        }
        ++g_launches;
    };
    if (!inverse) {
        radix();
        CU(cudaGetLastError());
        return launch_tile_jobs(g, buf, total, kLog2TwMax, 0, st);
    }
    if (int rc = launch_tile_jobs(g, buf, total, kLog2TwMax, 1, st)) return rc;
    radix();
    CU(cudaGetLastError());
    return TEBSCAT_OK;
}

// Middle of a long "iFFT -> modulus -> FFT" (core/scattering1d.py:312-318): last inverse radix pass, modulus and
// first forward radix pass on the R values each thread owns -- one trip through memory instead of three.
template <int LOGR>
__global__ void g_radix_pair_kernel(float2* __restrict__ buf, long long n_transforms, int n, const float2* __restrict__ tw) {
    constexpr int R = 1 << LOGR;
    const int logm = n - LOGR;
    const long long total = n_transforms << logm;
    const long long L = (long long)1 << n;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long s = e >> logm;
        const int i0 = (int)(e - (s << logm));
        float2* base = buf + s * L + i0;
        float2 v[R], f[R];
#pragma unroll
        for (int q = 0; q < R; ++q) {
            float2 z = base[(long long)brev<LOGR>(q) << logm];
            if (q != 0) z = cmulc(z, __ldg(tw + (long long)i0 * q));
            v[q] = z;
        }
        Dft<R, +1>::run(v);                                  // last inverse pass: register r holds sample qmap<R>(r)
#pragma unroll
        for (int r = 0; r < R; ++r) f[qmap<R>(r)] = make_float2(sqrtf(fmaf(v[r].x, v[r].x, v[r].y * v[r].y)), 0.f);
        Dft<R, -1>::run(f);                                  // first forward pass on the moduli
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int q = qmap<R>(r);
            float2 y = f[r];
            if (q != 0) y = cmul(y, __ldg(tw + (long long)i0 * q));
            base[(long long)brev<LOGR>(q) << logm] = y;
        }
    }
}

/* ifft -> modulus -> fft (core/scattering1d.py:312-318) in place on `n_transforms` bit-reversed spectra of 2^log2_len
 * samples: the result is the bit-reversed spectrum of |ifft(.)| (unnormalised inverse). */
extern "C" int tebscat_large_pair(tebscat_large* g, float* buf_dev, int64_t n_transforms, int log2_len, void* stream) {
    if (!g || !buf_dev || n_transforms < 1 || log2_len < 1 || log2_len > kLargeMaxLog2) return fail(TEBSCAT_EINVAL, "bad transform request");
    ON_DEVICE(g->device);
    cudaStream_t st = (cudaStream_t)stream;
    float2* buf = reinterpret_cast<float2*>(buf_dev);
    const long long total = (long long)n_transforms << log2_len;
    if (log2_len <= kLog2TwMax) return launch_tile_jobs(g, buf, total, log2_len, 2, st);
    if (int rc = launch_tile_jobs(g, buf, total, kLog2TwMax, 1, st)) return rc;
    const int blocks = g->n_sms * 8;
    switch (log2_len - kLog2TwMax) {
        case 1: g_radix_pair_kernel<1><<<blocks, 256, 0, st>>>(buf, n_transforms, log2_len, g->d_tw[log2_len]); break;
        case 2: g_radix_pair_kernel<2><<<blocks, 256, 0, st>>>(buf, n_transforms, log2_len, g->d_tw[log2_len]); break;
        case 3: g_radix_pair_kernel<3><<<blocks, 256, 0, st>>>(buf, n_transforms, log2_len, g->d_tw[log2_len]); break;
        default: g_radix_pair_kernel<4><<<blocks, 256, 0, st>>>(buf, n_transforms, log2_len, g->d_tw[log2_len]); break;
    }
    CU(cudaGetLastError());
    ++g_launches;
    return launch_tile_jobs(g, buf, total, kLog2TwMax, 0, st);
}

extern "C" int tebscat_large_pad_load_mode(tebscat_large* g, const float* x_dev, int64_t B, int N, int pad_left, int log2_Np,
                                           int border_mode, float* u_dev, void* stream) {
    if (!g || !x_dev || !u_dev || B < 1 || N < 2 || pad_left < 0 || pad_left >= N || ((1LL << log2_Np) - N - pad_left) >= N ||
        log2_Np < 1 || log2_Np > kLargeMaxLog2)
        return fail(TEBSCAT_EINVAL, "Indefinite padding size (larger than tensor).");
    if (border_mode != BORDER_REFLECT && border_mode != BORDER_CONSTANT && border_mode != BORDER_CIRCULAR)
        return fail(TEBSCAT_EINVAL, "unknown border mode %d", border_mode);
    ON_DEVICE(g->device);
    if (g->d_win && g->win_n != N) return fail(TEBSCAT_EINVAL, "window of %d samples on signals of %d", g->win_n, N);
    g_pad_load_kernel<<<g->n_sms * 8, 256, 0, (cudaStream_t)stream>>>(x_dev, reinterpret_cast<float2*>(u_dev), B, N, pad_left, log2_Np,
                                                                       g->d_win, border_mode);
    CU(cudaGetLastError());
    ++g_launches;
    return TEBSCAT_OK;
}

extern "C" int tebscat_large_pad_load(tebscat_large* g, const float* x_dev, int64_t B, int N, int pad_left, int log2_Np,
                                      float* u_dev, void* stream) {
    return tebscat_large_pad_load_mode(g, x_dev, B, N, pad_left, log2_Np, BORDER_REFLECT, u_dev, stream);
}

// dst[b, m] = 2^-sexp * sum_{t<k} src[b, m k + t] * f[m k + t]  (bit-reversed order; 4-bin chunks of the k-block that
// the host found negligible for every output are skipped: `mask` over up to 32 chunks of 2^logcw bins)
__global__ void g_mulfold_kernel(const float2* __restrict__ src, const float* __restrict__ f, float2* __restrict__ dst,
                                 long long B, int log_src, int logk, unsigned mask, int logcw, float scale) {
    const int log_dst = log_src - logk;
    const long long total = B << log_dst;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long b = e >> log_dst;
        const long long m = e - (b << log_dst);
        const float2* s = src + (b << log_src) + (m << logk);
        const float* ff = f + (m << logk);
        float ax = 0.f, ay = 0.f;
        if (logk < 2) {
            for (int t = 0; t < (1 << logk); ++t) {
                const float2 z = s[t];
                const float w = __ldg(ff + t);
                ax = fmaf(z.x, w, ax);
                ay = fmaf(z.y, w, ay);
            }
        } else {
            unsigned rest = mask;
            while (rest) {
                const int i_chunk = (__ffs(rest) - 1) << logcw;
                rest &= rest - 1;
                for (int t = i_chunk; t < i_chunk + (1 << logcw); t += 4) {
                    const float4 w = __ldg(reinterpret_cast<const float4*>(ff + t));
                    const float2 z0 = s[t], z1 = s[t + 1], z2 = s[t + 2], z3 = s[t + 3];
                    ax = fmaf(z0.x, w.x, ax); ay = fmaf(z0.y, w.x, ay);
                    ax = fmaf(z1.x, w.y, ax); ay = fmaf(z1.y, w.y, ay);
                    ax = fmaf(z2.x, w.z, ax); ay = fmaf(z2.y, w.z, ay);
                    ax = fmaf(z3.x, w.w, ax); ay = fmaf(z3.y, w.w, ay);
                }
            }
        }
        dst[e] = make_float2(ax * scale, ay * scale);
    }
}

/* cdgmm + subsample_fourier (kymatio/backend/torch_backend.py:147-219, torch_backend.py:18-48) on global spectra. */
extern "C" int tebscat_large_mulfold(tebscat_large* g, const float* src_dev, const float* filt_dev, float* dst_dev, int64_t B,
                                     int log_src, int logk, uint32_t chunk_mask, int log_chunk, int scale_exp, void* stream) {
    if (!g || !src_dev || !filt_dev || !dst_dev || B < 1 || log_src < 1 || log_src > kLargeMaxLog2 || logk < 0 || logk > log_src ||
        (logk >= 2 && (chunk_mask == 0 || log_chunk < 2 || log_chunk > logk || (logk - log_chunk) > 5)))
        return fail(TEBSCAT_EINVAL, "bad filter multiply request");
    ON_DEVICE(g->device);
    g_mulfold_kernel<<<g->n_sms * 8, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(src_dev), filt_dev,
                                                                      reinterpret_cast<float2*>(dst_dev), B, log_src, logk, chunk_mask,
                                                                      log_chunk, ldexpf(1.0f, -scale_exp));
    CU(cudaGetLastError());
    ++g_launches;
    return TEBSCAT_OK;
}

__global__ void g_modulus_kernel(float2* __restrict__ buf, long long total) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const float2 z = buf[e];
        buf[e] = make_float2(sqrtf(fmaf(z.x, z.x, z.y * z.y)), 0.f);
    }
}

/* modulus (kymatio/backend/torch_backend.py:137-141), in place, imaginary part zeroed */
extern "C" int tebscat_large_modulus(tebscat_large* g, float* buf_dev, int64_t n_complex, void* stream) {
    if (!g || !buf_dev || n_complex < 1) return fail(TEBSCAT_EINVAL, "null argument");
    ON_DEVICE(g->device);
    g_modulus_kernel<<<g->n_sms * 8, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float2*>(buf_dev), n_complex);
    CU(cudaGetLastError());
    ++g_launches;
    return TEBSCAT_OK;
}

__global__ void g_store_kernel(const float2* __restrict__ buf, float* __restrict__ out, long long B, int log_len, int i0, int n_out,
                               int n_paths, int channel) {
    const long long total = B * n_out;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long b = e / n_out;
        const int n = (int)(e - b * n_out);
        out[(b * n_paths + channel) * n_out + n] = buf[(b << log_len) + i0 + n].x;
    }
}

/* unpad + concatenate (torch_backend.py:80-102, kymatio/backend/torch_backend.py:143-145): real part of samples
 * [i0, i0 + n_out) of every length-2^log_len signal -> channel `channel` of out [B, n_paths, n_out] */
extern "C" int tebscat_large_store(tebscat_large* g, const float* buf_dev, int64_t B, int log_len, int i0, int n_out, int n_paths,
                                   int channel, float* out_dev, void* stream) {
    if (!g || !buf_dev || !out_dev || B < 1 || i0 < 0 || n_out < 1 || i0 + n_out > (1 << log_len) || channel < 0 || channel >= n_paths)
        return fail(TEBSCAT_EINVAL, "bad store request");
    ON_DEVICE(g->device);
    g_store_kernel<<<g->n_sms * 4, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(buf_dev), out_dev, B, log_len, i0,
                                                                    n_out, n_paths, channel);
    CU(cudaGetLastError());
    ++g_launches;
    return TEBSCAT_OK;
}

// ---- fused leaves ---------------------------------------------------------------------------------------------------
// A leaf of the cascade (core/scattering1d.py:287-292 / :320-327 / :358-364) is "phi multiply + periodise down to
// 2^lf bins -> inverse transform of 2^lf samples -> unpad -> channel": three launches and two round trips of a tiny
// buffer in the op-by-op form.  Here one block does it per signal: the periodised spectrum (<= 1024 bins) lives in
// shared memory, the transform is a radix-2 decimation-in-time pass sequence (bit-reversed in, natural out, twiddles
// from sincospif), and the adjoint kernel runs the same steps transposed (decimation in frequency, natural in,
// bit-reversed out) for the backward pass.
constexpr int kLeafMaxLog2 = 10;

__global__ void __launch_bounds__(256) g_leaf_kernel(const float2* __restrict__ src, const float* __restrict__ f, float* __restrict__ out,
                                                     long long B, int log_src, int logk, unsigned mask, int logcw, float scale, int lf,
                                                     int i0, int n_out, int n_paths, int channel) {
    __shared__ float2 buf[1 << kLeafMaxLog2];
    const int M = 1 << lf;
    for (long long b = blockIdx.x; b < B; b += gridDim.x) {
        const float2* sb = src + (b << log_src);
        for (int m = threadIdx.x; m < M; m += blockDim.x) {
            const float2* s = sb + ((long long)m << logk);
            const float* ff = f + ((long long)m << logk);
            float ax = 0.f, ay = 0.f;
            if (logk < 2) {
                for (int t = 0; t < (1 << logk); ++t) {
                    const float2 z = s[t];
                    const float w = __ldg(ff + t);
                    ax = fmaf(z.x, w, ax);
                    ay = fmaf(z.y, w, ay);
                }
            } else {
                unsigned rest = mask;
                while (rest) {
                    const int i_chunk = (__ffs(rest) - 1) << logcw;
                    rest &= rest - 1;
                    for (int t = i_chunk; t < i_chunk + (1 << logcw); t += 4) {
                        const float4 w = __ldg(reinterpret_cast<const float4*>(ff + t));
                        const float2 z0 = s[t], z1 = s[t + 1], z2 = s[t + 2], z3 = s[t + 3];
                        ax = fmaf(z0.x, w.x, ax); ay = fmaf(z0.y, w.x, ay);
                        ax = fmaf(z1.x, w.y, ax); ay = fmaf(z1.y, w.y, ay);
                        ax = fmaf(z2.x, w.z, ax); ay = fmaf(z2.y, w.z, ay);
                        ax = fmaf(z3.x, w.w, ax); ay = fmaf(z3.y, w.w, ay);
                    }
                }
            }
            buf[m] = make_float2(ax * scale, ay * scale);
        }
        __syncthreads();
        for (int st = 0; st < lf; ++st) {                      // decimation in time: bit-reversed -> natural, e^{+i...}
            const int half = 1 << st;
            for (int j = threadIdx.x; j < (M >> 1); j += blockDim.x) {
                const int pos = j & (half - 1), lo = ((j >> st) << (st + 1)) + pos, hi = lo + half;
                float sn, cs;
                sincospif((float)pos / (float)half, &sn, &cs);
                const float2 a = buf[lo], z = buf[hi];
                const float2 w = make_float2(fmaf(z.x, cs, -z.y * sn), fmaf(z.x, sn, z.y * cs));
                buf[lo] = make_float2(a.x + w.x, a.y + w.y);
                buf[hi] = make_float2(a.x - w.x, a.y - w.y);
            }
            __syncthreads();
        }
        for (int n = threadIdx.x; n < n_out; n += blockDim.x) out[(b * n_paths + channel) * n_out + n] = buf[i0 + n].x;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) g_leaf_adjoint_kernel(const float* __restrict__ gout, const float* __restrict__ f,
                                                             float2* __restrict__ gsrc, long long B, int log_src, int logk, unsigned mask,
                                                             int logcw, float scale, int lf, int i0, int n_out, int n_paths, int channel,
                                                             int accumulate) {
    __shared__ float2 buf[1 << kLeafMaxLog2];
    const int M = 1 << lf;
    for (long long b = blockIdx.x; b < B; b += gridDim.x) {
        for (int n = threadIdx.x; n < M; n += blockDim.x) {
            const int k = n - i0;
            buf[n] = make_float2((k >= 0 && k < n_out) ? __ldg(gout + (b * n_paths + channel) * n_out + k) : 0.f, 0.f);
        }
        __syncthreads();
        for (int st = lf - 1; st >= 0; --st) {                 // the transposed passes: natural -> bit-reversed, e^{-i...}
            const int half = 1 << st;
            for (int j = threadIdx.x; j < (M >> 1); j += blockDim.x) {
                const int pos = j & (half - 1), lo = ((j >> st) << (st + 1)) + pos, hi = lo + half;
                float sn, cs;
                sincospif((float)pos / (float)half, &sn, &cs);
                const float2 a = buf[lo], z = buf[hi];
                const float2 d = make_float2(a.x - z.x, a.y - z.y);
                buf[lo] = make_float2(a.x + z.x, a.y + z.y);
                buf[hi] = make_float2(fmaf(d.x, cs, d.y * sn), fmaf(d.y, cs, -d.x * sn));
            }
            __syncthreads();
        }
        float2* gb = gsrc + (b << log_src);
        for (int p = threadIdx.x; p < (1 << log_src); p += blockDim.x) {
            const int t = p & ((1 << logk) - 1);
            const bool live = logk < 2 || ((mask >> (t >> logcw)) & 1u);
            float2 v = make_float2(0.f, 0.f);
            if (live) {
                const float w = __ldg(f + p) * scale;
                const float2 z = buf[p >> logk];
                v = make_float2(z.x * w, z.y * w);
            }
            if (!accumulate) gb[p] = v;
            else if (live) {
                const float2 o = gb[p];
                gb[p] = make_float2(o.x + v.x, o.y + v.y);
            }
        }
        __syncthreads();
    }
}

static bool leaf_args_ok(int log_src, int logk, uint32_t chunk_mask, int log_chunk, int lf, int i0, int n_out, int n_paths, int channel) {
    return log_src >= 1 && log_src <= kLargeMaxLog2 && logk >= 0 && logk <= log_src && lf == log_src - logk && lf >= 1 &&
           lf <= kLeafMaxLog2 && !(logk >= 2 && (chunk_mask == 0 || log_chunk < 2 || log_chunk > logk || (logk - log_chunk) > 5)) &&
           i0 >= 0 && n_out >= 1 && i0 + n_out <= (1 << lf) && channel >= 0 && channel < n_paths;
}

extern "C" int tebscat_large_leaf(tebscat_large* g, const float* src_dev, const float* filt_dev, int64_t B, int log_src, int logk,
                                  uint32_t chunk_mask, int log_chunk, int scale_exp, int i0, int n_out, int n_paths, int channel,
                                  float* out_dev, void* stream) {
    const int lf = log_src - logk;
    if (!g || !src_dev || !filt_dev || !out_dev || B < 1 || !leaf_args_ok(log_src, logk, chunk_mask, log_chunk, lf, i0, n_out, n_paths, channel))
        return fail(TEBSCAT_EINVAL, "bad leaf request");
    ON_DEVICE(g->device);
    const int grid = (int)(B < (int64_t)g->n_sms * 8 ? B : (int64_t)g->n_sms * 8);
    g_leaf_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(src_dev), filt_dev, out_dev, B, log_src, logk,
                                                          chunk_mask, log_chunk, ldexpf(1.0f, -scale_exp), lf, i0, n_out, n_paths, channel);
    CU(cudaGetLastError());
    ++g_launches;
    return TEBSCAT_OK;
}

extern "C" int tebscat_large_leaf_adjoint(tebscat_large* g, const float* gout_dev, const float* filt_dev, int64_t B, int log_src, int logk,
                                          uint32_t chunk_mask, int log_chunk, int scale_exp, int i0, int n_out, int n_paths, int channel,
                                          float* gsrc_dev, int accumulate, void* stream) {
    const int lf = log_src - logk;
    if (!g || !gout_dev || !filt_dev || !gsrc_dev || B < 1 || !leaf_args_ok(log_src, logk, chunk_mask, log_chunk, lf, i0, n_out, n_paths, channel))
        return fail(TEBSCAT_EINVAL, "bad leaf request");
    ON_DEVICE(g->device);
    const int grid = (int)(B < (int64_t)g->n_sms * 8 ? B : (int64_t)g->n_sms * 8);
    g_leaf_adjoint_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gout_dev, filt_dev, reinterpret_cast<float2*>(gsrc_dev), B, log_src, logk,
                                                                  chunk_mask, log_chunk, ldexpf(1.0f, -scale_exp), lf, i0, n_out, n_paths,
                                                                  channel, accumulate);
    CU(cudaGetLastError());
    ++g_launches;
    return TEBSCAT_OK;
}

// ---- backward pass: adjoints of the ops above ----------------------------------------------------------
// The reference is differentiable through torch autograd with ModulusStable (kymatio/backend/torch_backend.py:5-96;
// test_differentiability_scattering, tests/scattering1d/test_torch_scattering1d.py:292-315).  Here the gradient is the
// transposed cascade on the same global buffers: recompute a first-order node, walk its subtree backwards.

// dst = (|src|, 0) out of place: the pre-modulus signal is kept for the backward of the modulus
__global__ void g_modulus_to_kernel(const float2* __restrict__ src, float2* __restrict__ dst, long long total) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const float2 z = src[e];
        dst[e] = make_float2(sqrtf(fmaf(z.x, z.x, z.y * z.y)), 0.f);
    }
}

extern "C" int tebscat_large_modulus_to(tebscat_large* g, const float* src_dev, float* dst_dev, int64_t n_complex, void* stream) {
    if (!g || !src_dev || !dst_dev || n_complex < 1) return fail(TEBSCAT_EINVAL, "null argument");
    ON_DEVICE(g->device);
    g_modulus_to_kernel<<<g->n_sms * 8, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(src_dev),
                                                                         reinterpret_cast<float2*>(dst_dev), n_complex);
    CU(cudaGetLastError());
    ++g_launches;
    return TEBSCAT_OK;
}

// ModulusStable.backward (kymatio/backend/torch_backend.py:59-96): grad_u = grad_|u| * u / |u|, 0 where |u| = 0.
// `grad` holds the gradient with respect to the (real) modulus in its real part; it is overwritten by grad_u.
__global__ void g_modulus_backward_kernel(const float2* __restrict__ u, float2* __restrict__ grad, long long total) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const float2 z = u[e];
        const float m = sqrtf(fmaf(z.x, z.x, z.y * z.y));
        const float q = m == 0.f ? 0.f : grad[e].x / m;
        grad[e] = make_float2(z.x * q, z.y * q);
    }
}

extern "C" int tebscat_large_modulus_backward(tebscat_large* g, const float* u_dev, float* grad_dev, int64_t n_complex, void* stream) {
    if (!g || !u_dev || !grad_dev || n_complex < 1) return fail(TEBSCAT_EINVAL, "null argument");
    ON_DEVICE(g->device);
    g_modulus_backward_kernel<<<g->n_sms * 8, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(u_dev),
                                                                               reinterpret_cast<float2*>(grad_dev), n_complex);
    CU(cudaGetLastError());
    ++g_launches;
    return TEBSCAT_OK;
}

// adjoint of g_mulfold_kernel: gsrc[b, p] (+)= 2^-sexp * f[p] * gdst[b, p >> logk]; the chunks the forward skips get 0
__global__ void g_unfold_kernel(const float2* __restrict__ gdst, const float* __restrict__ f, float2* __restrict__ gsrc,
                                long long B, int log_src, int logk, unsigned mask, int logcw, float scale, int accumulate) {
    const long long total = B << log_src;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long b = e >> log_src;
        const int p = (int)(e - (b << log_src));
        const int t = p & ((1 << logk) - 1);
        const bool live = logk < 2 || ((mask >> (t >> logcw)) & 1u);
        float2 v = make_float2(0.f, 0.f);
        if (live) {
            const float w = __ldg(f + p) * scale;
            const float2 z = gdst[(b << (log_src - logk)) + (p >> logk)];
            v = make_float2(z.x * w, z.y * w);
        }
        if (accumulate) {
            if (live) {
                const float2 o = gsrc[e];
                gsrc[e] = make_float2(o.x + v.x, o.y + v.y);
            }
        } else {
            gsrc[e] = v;
        }
    }
}

extern "C" int tebscat_large_unfold(tebscat_large* g, const float* gdst_dev, const float* filt_dev, float* gsrc_dev, int64_t B,
                                    int log_src, int logk, uint32_t chunk_mask, int log_chunk, int scale_exp, int accumulate,
                                    void* stream) {
    if (!g || !gdst_dev || !filt_dev || !gsrc_dev || B < 1 || log_src < 1 || log_src > kLargeMaxLog2 || logk < 0 || logk > log_src ||
        (logk >= 2 && (chunk_mask == 0 || log_chunk < 2 || log_chunk > logk || (logk - log_chunk) > 5)))
        return fail(TEBSCAT_EINVAL, "bad filter multiply request");
    ON_DEVICE(g->device);
    g_unfold_kernel<<<g->n_sms * 8, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(gdst_dev), filt_dev,
                                                                     reinterpret_cast<float2*>(gsrc_dev), B, log_src, logk, chunk_mask,
                                                                     log_chunk, ldexpf(1.0f, -scale_exp), accumulate);
    CU(cudaGetLastError());
    ++g_launches;
    return TEBSCAT_OK;
}

// adjoint of g_store_kernel: buf[b, :] = 0 except Re buf[b, i0 + n] = gout[b, channel, n]
__global__ void g_unstore_kernel(const float* __restrict__ gout, float2* __restrict__ buf, long long B, int log_len, int i0, int n_out,
                                 int n_paths, int channel) {
    const long long total = B << log_len;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long b = e >> log_len;
        const int n = (int)(e - (b << log_len)) - i0;
        const float v = (n >= 0 && n < n_out) ? __ldg(gout + (b * n_paths + channel) * n_out + n) : 0.f;
        buf[e] = make_float2(v, 0.f);
    }
}

extern "C" int tebscat_large_unstore(tebscat_large* g, const float* gout_dev, int64_t B, int log_len, int i0, int n_out, int n_paths,
                                     int channel, float* buf_dev, void* stream) {
    if (!g || !buf_dev || !gout_dev || B < 1 || i0 < 0 || n_out < 1 || i0 + n_out > (1 << log_len) || channel < 0 || channel >= n_paths)
        return fail(TEBSCAT_EINVAL, "bad store request");
    ON_DEVICE(g->device);
    g_unstore_kernel<<<g->n_sms * 4, 256, 0, (cudaStream_t)stream>>>(gout_dev, reinterpret_cast<float2*>(buf_dev), B, log_len, i0,
                                                                      n_out, n_paths, channel);
    CU(cudaGetLastError());
    ++g_launches;
    return TEBSCAT_OK;
}

// Phase stage A on the large-support level (kymatio_phase_scattering.py:220-231): the unpadded analytic signal of
// filter f, buf[b, i0 : i0 + N], into the workspaces of the pair stage -- cartesian and / or polar like STOREZ
__global__ void g_storez_kernel(const float2* __restrict__ buf, long long B, int log_len, int i0, int N, int F, int f, int mode,
                                float2* __restrict__ zc, float2* __restrict__ zp) {
    const long long total = B * N;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long b = e / N;
        const int t = (int)(e - b * N);
        const float2 z = buf[(b << log_len) + i0 + t];
        const long long o = (b * F + f) * (long long)N + t;
        if (mode & Z_CART) zc[o] = z;
        if (mode & Z_POLAR) zp[o] = make_float2(sqrtf(fmaf(z.x, z.x, z.y * z.y)), atan2f(z.y, z.x));
    }
}

extern "C" int tebscat_large_storez(tebscat_large* g, const float* buf_dev, int64_t B, int log_len, int i0, int N, int F, int f,
                                    int mode, float* zc_dev, float* zp_dev, void* stream) {
    if (!g || !buf_dev || B < 1 || log_len < 1 || log_len > kLargeMaxLog2 || i0 < 0 || N < 1 || i0 + N > (1 << log_len) ||
        F < 1 || f < 0 || f >= F || !(mode & (Z_CART | Z_POLAR)) || ((mode & Z_CART) && !zc_dev) || ((mode & Z_POLAR) && !zp_dev))
        return fail(TEBSCAT_EINVAL, "bad analytic-signal store request");
    ON_DEVICE(g->device);
    g_storez_kernel<<<g->n_sms * 4, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(buf_dev), B, log_len, i0, N, F, f, mode,
                                                                     reinterpret_cast<float2*>(zc_dev), reinterpret_cast<float2*>(zp_dev));
    CU(cudaGetLastError());
    ++g_launches;
    return TEBSCAT_OK;
}

// adjoint of the un-averaged store (OP_STOREU / core/scattering1d.py:329-330, :366-367): the output gradient of one path,
// grow[b, offset : offset + len], enters the real part of samples [i0, i0 + len) of the length-2^log_len signal --
// either as the whole signal (everything else zero) or added to what the buffer already holds there
__global__ void g_unstore_row_kernel(const float* __restrict__ grow, float2* __restrict__ buf, long long B, long long row_stride,
                                     long long offset, int log_len, int i0, int len, int accumulate) {
    if (accumulate) {
        const long long total = B * len;
        for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
            const long long b = e / len;
            const int n = (int)(e - b * len);
            buf[(b << log_len) + i0 + n].x += __ldg(grow + b * row_stride + offset + n);
        }
        return;
    }
    const long long total = B << log_len;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long b = e >> log_len;
        const int n = (int)(e - (b << log_len)) - i0;
        const float v = (n >= 0 && n < len) ? __ldg(grow + b * row_stride + offset + n) : 0.f;
        buf[e] = make_float2(v, 0.f);
    }
}

extern "C" int tebscat_large_unstore_row(tebscat_large* g, const float* grow_dev, int64_t B, int64_t row_stride, int64_t offset,
                                         int log_len, int i0, int len, int accumulate, float* buf_dev, void* stream) {
    if (!g || !buf_dev || !grow_dev || B < 1 || i0 < 0 || len < 1 || log_len < 0 || log_len > kLargeMaxLog2 ||
        i0 + len > (1 << log_len) || offset < 0 || offset + len > row_stride)
        return fail(TEBSCAT_EINVAL, "bad un-averaged store request");
    ON_DEVICE(g->device);
    g_unstore_row_kernel<<<g->n_sms * 4, 256, 0, (cudaStream_t)stream>>>(grow_dev, reinterpret_cast<float2*>(buf_dev), B, row_stride,
                                                                          offset, log_len, i0, len, accumulate);
    CU(cudaGetLastError());
    ++g_launches;
    return TEBSCAT_OK;
}

// adjoint of g_pad_load_kernel: gx[b, r] = Re gu[b, pad_left + r] + its left and right mirror images (pad < N: one fold)
__global__ void g_pad_adjoint_kernel(const float2* __restrict__ gu, float* __restrict__ gx, long long B, int N, int pad_left, int log2_Np,
                                     const float* __restrict__ win) {
    const long long total = B * N;
    const int pad_right = (1 << log2_Np) - N - pad_left;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long b = e / N;
        const int r = (int)(e - b * N);
        const float2* row = gu + (b << log2_Np);
        float v = row[pad_left + r].x;
        if (r >= 1 && r <= pad_left) v += row[pad_left - r].x;
        if (r <= N - 2 && r >= N - 1 - pad_right) v += row[pad_left + 2 * (N - 1) - r].x;
        if (win) v *= __ldg(win + r);
        gx[e] = v;
    }
}

extern "C" int tebscat_large_pad_adjoint(tebscat_large* g, const float* gu_dev, int64_t B, int N, int pad_left, int log2_Np,
                                         float* gx_dev, void* stream) {
    if (!g || !gx_dev || !gu_dev || B < 1 || N < 2 || pad_left < 0 || pad_left >= N || ((1LL << log2_Np) - N - pad_left) >= N)
        return fail(TEBSCAT_EINVAL, "Indefinite padding size (larger than tensor).");
    ON_DEVICE(g->device);
    if (g->d_win && g->win_n != N) return fail(TEBSCAT_EINVAL, "window of %d samples on signals of %d", g->win_n, N);
    g_pad_adjoint_kernel<<<g->n_sms * 8, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(gu_dev), gx_dev, B, N, pad_left,
                                                                          log2_Np, g->d_win);
    CU(cudaGetLastError());
    ++g_launches;
    return TEBSCAT_OK;
}

#ifdef TEBSCAT_TC_TRACE
extern "C" int tebscat_debug_tc_trace(long long* out, int n) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, tebscat::g_tc_trace, (size_t)n * sizeof(long long));
    return 0;
}
#endif
