// Host emulator of the step interpreter (TEST INFRASTRUCTURE ONLY).
//
// Compiles vae-teb_b200/csrc/scat_core.cuh as plain C++ and executes the threads of
// every step one after another.  It lets the CPU-only test-suite check the
// schedule builder and the task semantics (index maps, twiddles, swizzle) against
// the oracle without a GPU.  It is never loaded by the product path.
#define TEBSCAT_HOST_EMU 1
#include "../../vae-teb_b200/csrc/scat_core.cuh"

#include <cstdlib>
#include <cstring>
#include <vector>

using namespace tebscat;

static int emu_run(int N, int log2_Np, int pad_left, int n_paths, int n_out,
                   int smem_complex, int n_tasks, int n_steps,
                   const float* arena, const int32_t* tasks, const int32_t* steps,
                   const int32_t* chan, const float* x, long long B, float* out,
                   float* zc, float* zp, int z_mode,
                   const float* ep_mean, const float* ep_std, const unsigned char* ep_mode,
                   float ep_log_eps, int ep_trim, int ep_time_major, int border,
                   const float* gsrc, long long gsrc_stride, int scratch_complex = 0) {
    std::vector<float2> scratch((size_t)scratch_complex, make_float2(NAN, NAN));   // plans that park U0 in global memory
    std::vector<float2> S((size_t)smem_complex);
    std::vector<float2> tw(kTwAP + kTwBP, make_float2(0.f, 0.f));
    const double w0 = -2.0 * M_PI / (double)(1 << kLog2TwMax);
    for (int a = 0; a < kTwA; ++a) tw[a + (a >> 4)] = make_float2((float)cos(w0 * 128.0 * a), (float)sin(w0 * 128.0 * a));
    for (int b = 0; b < kTwB; ++b) tw[kTwAP + b + (b >> 4)] = make_float2((float)cos(w0 * b), (float)sin(w0 * b));
    for (long long b = 0; b < B; ++b) {
        // poison shared memory so that reads of never-written slots show up as NaNs
        for (auto& z : S) z = make_float2(NAN, NAN);
        SignalCtx c;
        c.x = x ? x + b * N : nullptr;
        c.gsrc = gsrc ? reinterpret_cast<const float2*>(gsrc) + b * gsrc_stride : nullptr;
        c.out = out + b * (long long)n_paths * (ep_mean ? n_out - 2 * ep_trim : n_out);
        c.ep_mean = ep_mean; c.ep_std = ep_std; c.ep_mode = ep_mode; c.ep_log_eps = ep_log_eps;
        c.ep_trim = ep_trim; c.ep_time_major = ep_time_major; c.n_paths = n_paths;
        c.ch_limit = n_paths;
        c.gbuf = nullptr; c.g_valid = 0;
        if (z_mode == 8) {                          // tile jobs: `zc` is the global complex buffer, n_out = slots per job
            c.gbuf = reinterpret_cast<float2*>(zc) + b * (long long)n_out;
            c.g_valid = n_out;
        }
        // phase stage B (OP_LOADPAIR): job b covers rows b*n_paths + q; zc/zp are then [rows][N] INPUTS
        if (z_mode == 4) {
            for (int q = 0; q < kMaxPairRows && q < n_paths; ++q) {
                c.pr_zp[q] = reinterpret_cast<const float2*>(zp) + (b * n_paths + q) * (long long)N;
                c.pr_zc[q] = reinterpret_cast<const float2*>(zc) + (b * n_paths + q) * (long long)N;
                c.pr_pw[q] = x[b * n_paths + q];           // the rows' powers travel in `x`
            }
        }
        if (scratch_complex > 0) {
            c.gbuf = scratch.data();
            c.g_valid = scratch_complex;
            c.gsrc = scratch.data();
        }
        c.chan = chan;
        c.zc = reinterpret_cast<float2*>(zc) + b * (long long)n_paths * n_out;
        c.zp = reinterpret_cast<float2*>(zp) + b * (long long)n_paths * n_out;
        c.z_mode = z_mode;
        c.N = N; c.pad_left = pad_left; c.log2_Np = log2_Np; c.n_out = n_out; c.border = border; c.win = nullptr;
        for (int s = 0; s < n_steps; ++s) {
            for (int ti = steps[2 * s]; ti < steps[2 * s + 1]; ++ti) {
                Task t;
                memcpy(&t, tasks + kTaskInts * ti, sizeof(Task));
                // multi-pass FFT tasks: pass by pass over all lanes (the kernel separates them by warp fences)
                for (int k = 0; k < task_passes(t); ++k)
                    for (int lt = 0; lt < t.nt; ++lt) exec_task<true>(S.data(), tw.data(), tw.data() + kTwAP, arena, c, t, lt, k);
            }
        }
    }
    return 0;
}

extern "C" int emu_scat1d_forward(int N, int log2_Np, int pad_left, int n_paths, int n_out,
                                  int smem_complex, int n_tasks, int n_steps,
                                  const float* arena, const int32_t* tasks, const int32_t* steps,
                                  const int32_t* chan, const float* x, long long B, float* out,
                                  float* zc, float* zp, int z_mode,
                                  const float* ep_mean, const float* ep_std, const unsigned char* ep_mode,
                                  float ep_log_eps, int ep_trim, int ep_time_major, int border, int scratch_complex) {
    return emu_run(N, log2_Np, pad_left, n_paths, n_out, smem_complex, n_tasks, n_steps, arena, tasks, steps, chan, x, B, out,
                   zc, zp, z_mode, ep_mean, ep_std, ep_mode, ep_log_eps, ep_trim, ep_time_major, border, nullptr, 0,
                   scratch_complex);
}

// fused subtrees of the large-support level (tebscat_scat1d_forward_gsrc): job b reads the complex spectrum at
// gsrc + b * gsrc_stride; only the plan's channels of `out` are written
extern "C" int emu_scat1d_forward_gsrc(int n_paths, int n_out, int smem_complex, int n_tasks, int n_steps,
                                       const float* arena, const int32_t* tasks, const int32_t* steps, const int32_t* chan,
                                       const float* gsrc, long long gsrc_stride, long long B, float* out) {
    return emu_run(1 << kLog2TwMax, kLog2TwMax, 0, n_paths, n_out, smem_complex, n_tasks, n_steps, arena, tasks, steps, chan,
                   nullptr, B, out, nullptr, nullptr, 0, nullptr, nullptr, nullptr, 0.f, 0, 0, 0, gsrc, gsrc_stride);
}
