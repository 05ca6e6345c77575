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
#include <cstring>
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

// ---------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------
struct KParams {
    const float* arena;      // filters, fp32, bit-reversed bin order
    const float2* tw;        // kTwA coarse + kTwB fine twiddles
    const int4* warp_tab;    // [n_steps][kWarps] x 3 int4: the task of every warp in every step
    const int32_t* chan;     // channel table of the batched stores
    long long* prof;         // optional: clock64() of CTA 0 at every step boundary (first signal)
    int32_t n_steps;
    int32_t smem_complex;
    int32_t N, pad_left, log2_Np, n_paths, n_out;
};

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;

// One CTA per SM, persistent over the batch: signal b, b + gridDim.x, ...
// Every warp walks its own column of the task table.  A 12-int record is held ONE INT PER
// LANE (lane l keeps field l), so the two records prefetched ahead of the executing one
// cost two registers, and the fields are broadcast with shuffles when the step starts.
// Descriptor latency (an L2 hit) therefore never sits between two barriers.
__global__ void __launch_bounds__(kThreads, 1)
scat1d_kernel(const KParams p, const float* __restrict__ x, float* __restrict__ out, long long B) {
    extern __shared__ __align__(16) float2 smem[];
    float2* S = smem;
    float2* twA = smem + p.smem_complex;
    float2* twB = twA + kTwA;
    for (int i = threadIdx.x; i < kTwA + kTwB; i += blockDim.x) twA[i] = p.tw[i];
    __syncthreads();

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int32_t* tab = reinterpret_cast<const int32_t*>(p.warp_tab) + kTaskInts * warp + (lane < kTaskInts ? lane : 0);
    const int stride = kTaskInts * kWarps;
    const int n_steps = p.n_steps;
    int cur = __ldg(tab);
    int nxt = __ldg(tab + stride * (1 % n_steps));
    int s_fetch = 2 % n_steps;
    for (long long b = blockIdx.x; b < B; b += gridDim.x) {
        SignalCtx c;
        c.x = x + b * p.N;
        c.out = out + b * (long long)p.n_paths * p.n_out;
        c.chan = p.chan;
        c.N = p.N;
        c.pad_left = p.pad_left;
        c.log2_Np = p.log2_Np;
        c.n_out = p.n_out;
        const bool prof = p.prof != nullptr && blockIdx.x == 0 && b == blockIdx.x && tid == 0;
        for (int s = 0; s < n_steps; ++s) {
            const int fut = __ldg(tab + stride * s_fetch);             // wraps into the next signal
            s_fetch = (s_fetch + 1 == n_steps) ? 0 : s_fetch + 1;
            if (prof) p.prof[s] = clock64();
            const int op = __shfl_sync(0xffffffffu, cur, 0);
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
                t.g = 0; t.h = 0; t.pad = 0;
                exec_task(S, twA, twB, p.arena, c, t, tid - t.t0);
            }
            __syncthreads();
            cur = nxt;
            nxt = fut;
        }
        if (prof) p.prof[n_steps] = clock64();
    }
}

// ---------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------
struct HostPipe {            // resources of the host-buffer entry point, created lazily
    bool ready = false;
    int64_t chunk = 0;
    cudaStream_t stream[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    float* d_x[2] = {nullptr, nullptr};
    float* d_S[2] = {nullptr, nullptr};
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
    KParams kp;
    HostPipe pipe;
    std::mutex pipe_mu;
};

static int validate_schedule(const tebscat_plan_desc& d, const int32_t* tasks, const int32_t* steps,
                             size_t n_floats, size_t n_chan) {
    const int cap = (int)(((int64_t)d.smem_complex * 16) / 17);   // logical slots (1 pad slot per 16)
    for (int s = 0; s < d.n_steps; ++s) {
        const int b = steps[2 * s], e = steps[2 * s + 1];
        if (b < 0 || e < b || e > d.n_tasks) return fail(TEBSCAT_EINVAL, "step %d: bad task range [%d,%d)", s, b, e);
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
                if (t[5] < 1 || t[5] > kLog2TwMax || t[6] < 1 || t[6] > 4 || t[6] > t[5] || t[4] < 1 ||
                    (((int64_t)t[4] << t[6]) & (((int64_t)1 << t[5]) - 1)) || !fits(t[3], (int64_t)t[4] << t[6]))
                    return fail(TEBSCAT_EINVAL, "task %d: bad FFT pass (%d butterflies, B=2^%d R=2^%d at %d)", i, t[4], t[5], t[6], t[3]);
                break;
            }
            case OP_MULFOLD: {
                if (t[4] < 2 || t[4] > kLog2TwMax || t[5] < 0 || t[5] > t[4] || t[5] > 6 || !fits(t[3], (int64_t)1 << t[4]) ||
                    !fits(t[6] & ~15, (t[6] & 15) + ((int64_t)1 << (t[4] - t[5]))) || t[7] < 0 ||
                    (size_t)t[7] + ((size_t)1 << t[4]) > n_floats || (t[7] & 3) || (t[6] & 3 && t[5] < 2))
                    return fail(TEBSCAT_EINVAL, "task %d: bad MULFOLD", i);
                if (t[5] >= 2 && ((unsigned)t[8] == 0u || ((unsigned)t[8] >> (1 << (t[5] - 2))) != 0u))
                    return fail(TEBSCAT_EINVAL, "task %d: bad MULFOLD chunk mask", i);
                break;
            }
            case OP_STOREB:
                if (t[4] < 1 || t[6] != d.n_out || t[5] < 0 || t[8] < 0 || t[8] > kLog2TwMax || t[5] + t[6] > (1 << t[8]) ||
                    !fits(t[3], (int64_t)t[4] << t[8]) || t[7] < 0 || (size_t)t[7] + (size_t)t[4] > n_chan)
                    return fail(TEBSCAT_EINVAL, "task %d: bad STOREB", i);
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
    for (size_t i = 0; i < n_chan; ++i)
        if (chan[i] < 0 || chan[i] >= desc->n_paths) return fail(TEBSCAT_EINVAL, "channel table entry %zu out of range", i);
    if (int rc = validate_schedule(*desc, tasks, steps, n_floats, n_chan)) return rc;

    int n_dev = 0;
    CU(cudaGetDeviceCount(&n_dev));
    if (device < 0 || device >= n_dev) return fail(TEBSCAT_EINVAL, "device %d not in [0,%d)", device, n_dev);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));

    tebscat_plan* p = new tebscat_plan();
    p->desc = *desc;
    p->device = device;
    p->n_sms = prop.multiProcessorCount;
    p->smem_bytes = ((size_t)desc->smem_complex + kTwA + kTwB) * sizeof(float2);
    if (p->smem_bytes > (size_t)prop.sharedMemPerBlockOptin) {
        delete p;
        return fail(TEBSCAT_EUNSUPPORTED, "schedule needs %zu B of shared memory, device offers %zu",
                    p->smem_bytes, (size_t)prop.sharedMemPerBlockOptin);
    }
    std::vector<float2> tw(kTwA + kTwB);
    const double w0 = -2.0 * M_PI / (double)(1 << kLog2TwMax);
    for (int a = 0; a < kTwA; ++a) tw[a] = make_float2((float)cos(w0 * 128.0 * a), (float)sin(w0 * 128.0 * a));
    for (int b = 0; b < kTwB; ++b) tw[kTwA + b] = make_float2((float)cos(w0 * b), (float)sin(w0 * b));

    CU(cudaMalloc(&p->d_arena, n_floats * sizeof(float)));
    CU(cudaMalloc(&p->d_tw, tw.size() * sizeof(float2)));
    // per-warp view of the schedule: record (step, warp) = the task whose thread range covers the warp
    std::vector<int32_t> wt((size_t)desc->n_steps * kWarps * kTaskInts, 0);
    for (int st = 0; st < desc->n_steps; ++st) {
        for (int ti = steps[2 * st]; ti < steps[2 * st + 1]; ++ti) {
            const int32_t* t = tasks + kTaskInts * ti;
            for (int w = t[1] / 32; w < (t[1] + t[2]) / 32; ++w) {
                int32_t* rec = wt.data() + ((size_t)st * kWarps + w) * kTaskInts;
                if ((rec[0] & 0xff) != OP_NOP) { delete p; return fail(TEBSCAT_EINVAL, "step %d: overlapping thread ranges", st); }
                memcpy(rec, t, kTaskInts * sizeof(int32_t));
            }
        }
    }
    CU(cudaMalloc(&p->d_warp_tab, wt.size() * sizeof(int32_t)));
    CU(cudaMemcpy(p->d_warp_tab, wt.data(), wt.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&p->d_chan, (n_chan ? n_chan : 1) * sizeof(int32_t)));
    CU(cudaMemcpy(p->d_chan, chan, n_chan * sizeof(int32_t), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(p->d_arena, arena, n_floats * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(p->d_tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
    // the attribute belongs to the kernel, not to the plan: always allow the device maximum
    CU(cudaFuncSetAttribute(scat1d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)prop.sharedMemPerBlockOptin));

    KParams& k = p->kp;
    k.arena = p->d_arena;
    k.tw = p->d_tw;
    k.warp_tab = reinterpret_cast<const int4*>(p->d_warp_tab);
    k.chan = p->d_chan;
    k.prof = nullptr;
    k.n_steps = desc->n_steps;
    k.smem_complex = desc->smem_complex;
    k.N = desc->N;
    k.pad_left = desc->pad_left;
    k.log2_Np = desc->log2_Np;
    k.n_paths = desc->n_paths;
    k.n_out = desc->n_out;
    *out = p;
    return TEBSCAT_OK;
}

extern "C" void tebscat_plan_destroy(tebscat_plan* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->pipe.ready) {
        for (int i = 0; i < 2; ++i) {
            cudaStreamDestroy(p->pipe.stream[i]);
            cudaEventDestroy(p->pipe.done[i]);
            cudaFree(p->pipe.d_x[i]);
            cudaFree(p->pipe.d_S[i]);
        }
    }
    cudaFree(p->d_arena);
    cudaFree(p->d_tw);
    cudaFree(p->d_warp_tab);
    cudaFree(p->d_chan);
    delete p;
}

static int launch_scat1d(const tebscat_plan* p, const float* x, int64_t B, float* S, cudaStream_t st) {
    if (B == 0) return TEBSCAT_OK;
    const int grid = (int)(B < (int64_t)p->n_sms ? B : (int64_t)p->n_sms);
    scat1d_kernel<<<grid, p->desc.n_threads, p->smem_bytes, st>>>(p->kp, x, S, (long long)B);
    CU(cudaGetLastError());
    ++g_launches;
    return TEBSCAT_OK;
}

extern "C" int tebscat_scat1d_forward(const tebscat_plan* p, const float* x_dev, int64_t B, float* S_dev,
                                      void* stream) {
    g_launches = 0;
    if (!p || B < 0 || (B > 0 && (!x_dev || !S_dev))) return fail(TEBSCAT_EINVAL, "null argument");
    int cur = -1;
    CU(cudaGetDevice(&cur));
    if (cur != p->device) CU(cudaSetDevice(p->device));
    int rc = launch_scat1d(p, x_dev, B, S_dev, (cudaStream_t)stream);
    if (cur != p->device && cur >= 0) cudaSetDevice(cur);
    return rc;
}

extern "C" int tebscat_scat1d_profile_steps(const tebscat_plan* p, const float* x_dev, int64_t B, float* S_dev,
                                            long long* step_clocks_host, void* stream) {
    g_launches = 0;
    if (!p || !x_dev || !S_dev || !step_clocks_host || B < 1) return fail(TEBSCAT_EINVAL, "null argument");
    CU(cudaSetDevice(p->device));
    long long* d_prof = nullptr;
    const size_t n = (size_t)p->desc.n_steps + 1;
    CU(cudaMalloc(&d_prof, n * sizeof(long long)));
    CU(cudaMemsetAsync(d_prof, 0, n * sizeof(long long), (cudaStream_t)stream));
    KParams kp = p->kp;
    kp.prof = d_prof;
    const int grid = (int)(B < (int64_t)p->n_sms ? B : (int64_t)p->n_sms);
    scat1d_kernel<<<grid, p->desc.n_threads, p->smem_bytes, (cudaStream_t)stream>>>(kp, x_dev, S_dev, (long long)B);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaMemcpy(step_clocks_host, d_prof, n * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(d_prof);
    if (e != cudaSuccess) return fail(TEBSCAT_ECUDA, "profile run failed: %s", cudaGetErrorString(e));
    ++g_launches;
    return TEBSCAT_OK;
}

extern "C" int tebscat_scat1d_forward_host(tebscat_plan* p, const float* x_host, int64_t B, float* S_host) {
    g_launches = 0;
    if (!p || B < 0 || (B > 0 && (!x_host || !S_host))) return fail(TEBSCAT_EINVAL, "null argument");
    if (B == 0) return TEBSCAT_OK;
    std::lock_guard<std::mutex> lock(p->pipe_mu);
    CU(cudaSetDevice(p->device));
    HostPipe& hp = p->pipe;
    const size_t in_f = (size_t)p->desc.N, out_f = (size_t)p->desc.n_paths * p->desc.n_out;
    if (!hp.ready) {
        hp.chunk = (int64_t)p->n_sms * 8;          // 8 signals per SM per chunk
        for (int i = 0; i < 2; ++i) {
            CU(cudaStreamCreateWithFlags(&hp.stream[i], cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&hp.done[i], cudaEventDisableTiming));
            CU(cudaMalloc(&hp.d_x[i], hp.chunk * in_f * sizeof(float)));
            CU(cudaMalloc(&hp.d_S[i], hp.chunk * out_f * sizeof(float)));
        }
        hp.ready = true;
    }
    int slot = 0;
    for (int64_t b0 = 0; b0 < B; b0 += hp.chunk, slot ^= 1) {
        const int64_t nb = (B - b0 < hp.chunk) ? (B - b0) : hp.chunk;
        cudaStream_t st = hp.stream[slot];
        // the slot's previous D2H is ordered before this H2D by stream order
        CU(cudaMemcpyAsync(hp.d_x[slot], x_host + b0 * in_f, nb * in_f * sizeof(float), cudaMemcpyHostToDevice, st));
        if (int rc = launch_scat1d(p, hp.d_x[slot], nb, hp.d_S[slot], st)) return rc;
        CU(cudaMemcpyAsync(S_host + b0 * out_f, hp.d_S[slot], nb * out_f * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(hp.stream[0]));
    CU(cudaStreamSynchronize(hp.stream[1]));
    return TEBSCAT_OK;
}
