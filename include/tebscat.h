/*
 * tebscat -- C ABI of the B200-native wavelet-scattering hot path of VAE-TEB.
 *
 * This is the drop-in boundary.  Nothing like it exists in
 * the reference, whose plugin API is the op-by-op kymatio backend class
 *   kymatio/kymatio/scattering1d/backend/torch_backend.py:17-174
 * (pad, rfft, cdgmm, subsample_fourier, ifft, modulus, irfft, unpad, concatenate).
 * An op-by-op backend cannot express a fused cascade, so the boundary sits one
 * level higher: one call replaces the whole of
 *   kymatio/kymatio/scattering1d/core/scattering1d.py:197-399      (tebscat_scat1d_*)
 *   hdf5_dataset/kymatio_phase_scattering.py:220-360               (tebscat_phase_*)
 * The Python frontends in vae-teb_b200/tebscat/ bind these entry points with
 * ctypes and keep the reference's Scattering1D / KymatioPhaseScattering1D surface.
 *
 * Conventions
 *  - plain pointers and sizes only; all device pointers are fp32, contiguous,
 *    on the plan's device; the caller owns every buffer it passes in;
 *  - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream,
 *    the same stream convention as the reference's skcuda backend,
 *    kymatio/kymatio/scattering1d/backend/torch_skcuda_backend.py:87,164);
 *  - every function returns 0 on success or a TEBSCAT_E* code; the message is
 *    available from tebscat_last_error() (thread-local); nothing throws;
 *  - a plan is immutable after creation: forward calls are re-entrant from several
 *    host threads on different streams.  One plan per device.
 */
#ifndef TEBSCAT_H_
#define TEBSCAT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TEBSCAT_ABI_VERSION 2

enum {
    TEBSCAT_OK = 0,
    TEBSCAT_EINVAL = 1,      /* bad argument / inconsistent plan description   */
    TEBSCAT_ECUDA = 2,       /* a CUDA runtime call failed                     */
    TEBSCAT_EUNSUPPORTED = 3 /* valid request this build cannot serve          */
};

/* Geometry of one Scattering1D instance plus the sizes of its flat tables.
 * Replaces the attributes set by ScatteringBase1D.build
 * (kymatio/kymatio/scattering1d/frontend/base_frontend.py:27-77). */
typedef struct tebscat_plan_desc {
    int32_t abi_version;   /* TEBSCAT_ABI_VERSION                                      */
    int32_t N;             /* input length                                             */
    int32_t log2_Np;       /* J_pad                                                    */
    int32_t pad_left;      /* reflect padding on the left                              */
    int32_t n_paths;       /* C: number of output channels (meta()['key'] order)       */
    int32_t n_out;         /* ind_end[log2T] - ind_start[log2T]                        */
    int32_t n_threads;     /* CTA size the schedule was built for                      */
    int32_t smem_complex;  /* complex64 slots of shared memory the schedule addresses  */
    int32_t n_tasks;       /* entries of `tasks` (12 x int32 each)                     */
    int32_t n_steps;       /* entries of `steps` (2 x int32 each: [task_begin, task_end)) */
    int32_t border_mode;   /* padding rule of the plan's loads: 0 reflect (the scattering transform, always),
                              1 constant (zeros), 2 circular -- KymatioPhaseScattering1D(border_mode=...),
                              hdf5_dataset/kymatio_phase_scattering.py:162-173                          */
    int32_t scratch_complex; /* 0, or the complex64 elements of global scratch each CTA needs: schedules that park the
                              signal's spectrum in L2 instead of shared memory (tebscat/schedule.py, u0_scratch);
                              the library allocates it (n_SM x scratch_complex x 8 bytes per stream in use)  */
    int32_t reserved[4];
} tebscat_plan_desc;

typedef struct tebscat_plan tebscat_plan;

/* Upload the filter arena (fp32, bit-reversed bin order), the step
 * schedule and the channel table of its batched stores to `device`.  Replaces ScatteringTorch1D.register_filters /
 * load_filters (kymatio/kymatio/scattering1d/frontend/torch_frontend.py:75-116). */
int tebscat_plan_create(const tebscat_plan_desc* desc,
                        const float* filter_arena_host, size_t n_floats,
                        const int32_t* tasks_host, const int32_t* steps_host,
                        const int32_t* channel_table_host, size_t n_channel_entries,
                        int device, tebscat_plan** out);

void tebscat_plan_destroy(tebscat_plan* plan);

/* Plan files.  The schedule of a configuration is built by tebscat/schedule.py (host Python, like the reference's own
 * filter factory); _save writes everything tebscat_plan_create takes into one self-describing, checksummed file
 * (`python -m tebscat.export_plan` does it from the command line) and _load creates the plan from it, so that a consumer
 * without Python (C, C++, or any FFI) can run the transform: tebscat_plan_load + tebscat_scat1d_forward.  The file is
 * validated like a direct call (schedule checks, ABI version, sizes, checksum): a malformed file is TEBSCAT_EINVAL. */
int tebscat_plan_save(const char* path, const tebscat_plan_desc* desc, const float* filter_arena_host, size_t n_floats,
                      const int32_t* tasks_host, const int32_t* steps_host,
                      const int32_t* channel_table_host, size_t n_channel_entries);
int tebscat_plan_load(const char* path, int device, tebscat_plan** out);
/* The description a plan was created with (input length N, output geometry n_paths x n_out, ...). */
int tebscat_plan_get_desc(const tebscat_plan* plan, tebscat_plan_desc* out);

/* Analysis window applied to the samples as the plan's loads read them: x[t] * window[t] before padding.
 * window_host: N floats (copied), or NULL to remove it.  Replaces the eager `x = x * window` of
 * KymatioPhaseScattering1D.forward (hdf5_dataset/kymatio_phase_scattering.py:405-407; the Tukey taper of
 * _create_tukey_window, :362-392).  Call it between forward calls, not concurrently with them. */
int tebscat_plan_set_window(tebscat_plan* plan, const float* window_host);

/* S[b, c, :] for b < B.  x_dev: [B, N]; S_dev: [B, n_paths, n_out]; both device
 * fp32 contiguous.  Replaces core.scattering1d
 * (kymatio/kymatio/scattering1d/core/scattering1d.py:197-399) as called from
 * ScatteringTorch1D.scattering (frontend/torch_frontend.py:219-227). */
int tebscat_scat1d_forward(const tebscat_plan* plan, const float* x_dev, int64_t B,
                           float* S_dev, void* stream);

/* Output epilogue fused into the transform's stores: what CombinedHDF5Dataset.__getitem__ does
 * to a stored `fhr_st` record before the model sees it --
 *   trim `trim` decimated samples at both ends                  (hdf5_dataset/hdf5_dataset.py:733-741),
 *   normalize_tensor_data: log(clamp(x, 0) + log_eps) or asinh(x) on the flagged channels, then
 *   (x - mean[c]) / (std[c] + 1e-8)                             (hdf5_dataset/hdf5_dataset.py:96-135),
 *   (channels, time) -> (time, channels)                        (hdf5_dataset/hdf5_dataset.py:758-759).
 * All pointers are DEVICE pointers to n_paths entries; mode: 0 none, 1 log, 2 asinh. */
typedef struct tebscat_epilogue {
    const float* mean_dev;
    const float* std_dev;
    const unsigned char* mode_dev;
    float log_eps;
    int32_t trim;
    int32_t time_major;    /* 1: out_dev is [B, n_out - 2 trim, n_paths]; 0: [B, n_paths, n_out - 2 trim] */
} tebscat_epilogue;

/* tebscat_scat1d_forward with the epilogue above (epilogue == NULL: identical to tebscat_scat1d_forward). */
int tebscat_scat1d_forward_ex(const tebscat_plan* plan, const float* x_dev, int64_t B, float* out_dev,
                              const tebscat_epilogue* epilogue, void* stream);

/* Fused subtrees of the large-support level (padded lengths 2^14 .. 2^17).  A plan whose schedule
 * starts from global-source filter multiplies (OP_GMULFOLD, tebscat/schedule.py:build_hybrid_plans) runs every part of
 * the cascade below a spectrum that is too long for one SM -- the signal's spectrum U0, or the spectrum of a
 * first-order modulus of more than 8192 samples -- as ONE launch: psi multiply + periodisation straight from global
 * memory, then iFFT -> modulus -> FFT -> second order -> phi leaves out of shared memory
 * (kymatio/kymatio/scattering1d/core/scattering1d.py:300-364).  src_dev: [B][src_stride] complex64 (interleaved
 * floats), bit-reversed bin order, 32-byte aligned; S_dev: [B, n_paths, n_out], only the plan's channels are written.
 * Plans with such tasks are refused by the other forward entry points, and vice versa. */
int tebscat_scat1d_forward_gsrc(const tebscat_plan* plan, const float* src_dev, int64_t src_stride_complex, int64_t B,
                                float* S_dev, void* stream);

/* Same transform with HOST buffers: pinned staging, chunked H2D / compute / D2H
 * overlap on the plan's own streams; returns when S_host is complete.  This is
 * the call the dataset builder makes per record
 * (hdf5_dataset/create_hdf5_dataset.py:418-441: .to(device) ... .cpu().numpy()). */
int tebscat_scat1d_forward_host(tebscat_plan* plan, const float* x_host, int64_t B,
                                float* S_host);

/* Diagnostic: the host <-> device copies of tebscat_scat1d_forward_host (same chunks, streams and buffers) without
 * its kernel -- the copy ceiling the end-to-end rate is reported against (bench.py `e2e.host_copy_ceiling`). */
int tebscat_scat1d_host_copies_only(tebscat_plan* plan, const float* x_host, int64_t B, float* S_host);

/* Diagnostic: run the transform once and return clock64() of CTA 0 at every step
 * boundary of its first signal (n_steps + 1 values).  Used to calibrate the host
 * scheduler's cost model; not on the product path. */
int tebscat_scat1d_profile_steps(const tebscat_plan* plan, const float* x_dev, int64_t B,
                                 float* S_dev, long long* step_clocks_host, void* stream);

/* ---- phase-harmonic correlation (hdf5_dataset/kymatio_phase_scattering.py) ---------- */

/* Geometry of KymatioPhaseScattering1D's phase path (:100-160, :233-273). */
typedef struct tebscat_phase_desc {
    int32_t abi_version;
    int32_t N;             /* input length                                                     */
    int32_t n_filters;     /* F: first-order filters (rows of psi1_filters, :123-124)          */
    int32_t n_pairs;       /* P: pairs (i, j) with xi_j >= xi_i (:141-152)                      */
    int32_t n_out;         /* N // dec: length of the decimated output (:258-268)              */
    int32_t n_cols_pad;    /* columns of the smoothing operator, n_out rounded up to 80        */
    int32_t reserved[10];
} tebscat_phase_desc;

typedef struct tebscat_phase_plan tebscat_phase_plan;

/* `stage_a` is a plan whose schedule ends in STOREZ tasks (n_paths = F, n_out = N); its
 * ownership passes to the phase plan.  G_host: [N][n_cols_pad][2] fp32, the operator
 * reflect-pad -> FFT -> phi -> truncate -> iFFT -> slice of _apply_phi_filter (:233-273).
 * i_idx/j_idx/powers replace the buffers of _build_coupling_indices (:134-160). */
int tebscat_phase_plan_create(const tebscat_phase_desc* desc, tebscat_plan* stage_a,
                              const float* G_host, const int32_t* i_idx, const int32_t* j_idx,
                              const float* powers, tebscat_phase_plan** out);

void tebscat_phase_plan_destroy(tebscat_phase_plan* plan);

/* The same window on the loads of stage A (both channels). */
int tebscat_phase_plan_set_window(tebscat_phase_plan* plan, const float* window_host);

/* out[b, s, :] = Re smooth( |z_i| exp(i p theta_i) conj(z_j) ) for the selected pairs s,
 * z_i from channel ch_i and z_j from channel ch_j of x_dev [B, n_channels, N]
 * (ch_i == ch_j: _compute_phase_correlation :275-301; else
 * _compute_cross_channel_phase_correlation :303-360).  pair_subset_host (nullable) selects
 * pairs like `same_pairs_only` / the dataset masks (create_hdf5_dataset.py:440-441).
 * apply_low_pass == 0 returns the full-rate real part, out [B, n_sel, N] (:356-360).
 * Calls on one phase plan are serialised (it owns an L2-sized workspace): the enqueue under a lock, and a call on another
 * stream than the previous one waits (on the device) for that call to finish. */
int tebscat_phase_forward(tebscat_phase_plan* plan, const float* x_dev, int64_t B, int n_channels,
                          int ch_i, int ch_j, const int32_t* pair_subset_host, int n_subset,
                          int apply_low_pass, float* out_dev, void* stream);

/* Diagnostic (bench.py): time the two stages of tebscat_phase_forward with CUDA events on the caller's stream.
 * _profile(plan, 1) starts recording, _profile_read returns the summed stage times (ms) and clears the record. */
int tebscat_phase_plan_profile(tebscat_phase_plan* plan, int enable);
int tebscat_phase_plan_profile_read(tebscat_phase_plan* plan, double* stage_a_ms, double* stage_b_ms, int* n_chunks);

/* Padded lengths above 2^13: no stage-A plan exists (the analytic signals come from tebscat_large_* ops, see
 * tebscat_large_storez); the plan holds the pair tables and the dense operator only, and stage B runs on workspaces the
 * caller owns: zp_dev [nb][F][N] (|z|, theta) of the 'i' channel, zc_dev [nb][F][N] (re, im) of the 'j' channel
 * (kymatio_phase_scattering.py:275-360).  nb * F * N < 2^31. */
int tebscat_phase_plan_create_pairs_only(const tebscat_phase_desc* desc, int device, const float* G_host,
                                         const int32_t* i_idx, const int32_t* j_idx, const float* powers,
                                         tebscat_phase_plan** out);
int tebscat_phase_pairs(tebscat_phase_plan* plan, const float* zp_dev, const float* zc_dev, int64_t nb,
                        const int32_t* pair_subset_host, int n_subset, int apply_low_pass, float* out_dev, void* stream);

/* Optional: run stage B (the low-pass of every (sample, pair) product, _apply_phi_filter :233-273) as
 * transforms on the step interpreter instead of the dense operator.  `pair_plan` is a plan created with
 * tebscat_plan_create whose schedule starts with LOADPAIR tasks (tebscat/phase.py builds it); the phase
 * plan takes ownership.  Requires a power-of-two decimation factor. */
int tebscat_phase_plan_attach_pair_plan(tebscat_phase_plan* plan, tebscat_plan* pair_plan);

/* Single-pass dataset entry: within-channel correlations of channel ch_i for the
 * pairs `within_subset` and cross-channel correlations ch_i x ch_j for `cross_subset`, sharing the
 * analytic signals of ch_i.  Replaces the two st_model(...) calls and the 903 -> 44 / 130 masking
 * of hdf5_dataset/create_hdf5_dataset.py:421-441.  out_within [B, n_within, n_out],
 * out_cross [B, n_cross, n_out]. */
int tebscat_phase_forward_dual(tebscat_phase_plan* plan, const float* x_dev, int64_t B, int n_channels,
                               int ch_i, int ch_j, const int32_t* within_subset_host, int n_within,
                               const int32_t* cross_subset_host, int n_cross,
                               float* out_within_dev, float* out_cross_dev, void* stream);

/* Number of kernels the last forward call on this thread launched. */
int tebscat_last_launch_count(void);

/* ---- large-support level: padded lengths 2^14 .. 2^17 ----------------------
 * At these lengths a spectrum no longer fits one SM.  The cascade keeps the reference's op order
 * (core/scattering1d.py:269-370) with every op one launch over the batch on GLOBAL buffers of complex64
 * (bit-reversed spectra); the transforms run as tile jobs of the step interpreter.  tebscat/large.py drives it. */
typedef struct tebscat_large tebscat_large;
int tebscat_large_create(int device, tebscat_large** out);
void tebscat_large_destroy(tebscat_large* ctx);
/* analysis window of pad_load / pad_adjoint (see tebscat_plan_set_window): n floats (copied), or NULL to remove it */
int tebscat_large_set_window(tebscat_large* ctx, const float* window_host, int n);
/* tile plan (tebscat.schedule.build_tile_plan) for in-place transforms of 2^log2_len <= 8192 samples;
 * kind: 0 forward, 1 inverse, 2 inverse -> modulus -> forward; one job covers slots_per_job consecutive complex
 * elements (a multiple of the length); ownership moves */
int tebscat_large_set_tile_plan(tebscat_large* ctx, int log2_len, int kind, int slots_per_job, tebscat_plan* plan);
/* pad (torch_backend.py:50-78) + real -> complex: x_dev [B, N] -> u_dev [B, 2^log2_Np] complex64 */
int tebscat_large_pad_load(tebscat_large* ctx, const float* x_dev, int64_t B, int N, int pad_left, int log2_Np,
                           float* u_dev, void* stream);
/* the same with the phase module's padding rules (kymatio_phase_scattering.py:162-173): border_mode 0 reflect, 1 constant
 * (zeros), 2 circular -- stage A of the phase path at padded lengths above 2^13 */
int tebscat_large_pad_load_mode(tebscat_large* ctx, const float* x_dev, int64_t B, int N, int pad_left, int log2_Np,
                                int border_mode, float* u_dev, void* stream);
/* fft / ifft (torch_backend.py:106-128, unnormalised), in place, forward natural -> bit-reversed, inverse back */
int tebscat_large_fft(tebscat_large* ctx, float* buf_dev, int64_t n_transforms, int log2_len, int inverse, void* stream);
/* ifft -> modulus -> fft (core/scattering1d.py:312-318) in place on bit-reversed spectra, unnormalised */
int tebscat_large_pair(tebscat_large* ctx, float* buf_dev, int64_t n_transforms, int log2_len, void* stream);
/* cdgmm + subsample_fourier (kymatio/backend/torch_backend.py:147-219, torch_backend.py:18-48), times 2^-scale_exp */
int tebscat_large_mulfold(tebscat_large* ctx, const float* src_dev, const float* filt_dev, float* dst_dev, int64_t B,
                          int log_src, int logk, uint32_t chunk_mask, int log_chunk, int scale_exp, void* stream);
/* modulus (kymatio/backend/torch_backend.py:137-141), in place */
int tebscat_large_modulus(tebscat_large* ctx, float* buf_dev, int64_t n_complex, void* stream);
/* unpad + concatenate (torch_backend.py:80-102, kymatio/backend/torch_backend.py:143-145) */
int tebscat_large_store(tebscat_large* ctx, const float* buf_dev, int64_t B, int log_len, int i0, int n_out, int n_paths,
                        int channel, float* out_dev, void* stream);

/* phase stage A on this level (hdf5_dataset/kymatio_phase_scattering.py:220-231): buf_dev[b, i0 : i0 + N] of filter f ->
 * zc_dev [B][F][N] (re, im) if mode & 1, zp_dev [B][F][N] (|z|, atan2) if mode & 2 */
int tebscat_large_storez(tebscat_large* ctx, const float* buf_dev, int64_t B, int log_len, int i0, int N, int F, int f,
                         int mode, float* zc_dev, float* zp_dev, void* stream);

/* A whole leaf in one launch: phi multiply + periodise (as tebscat_large_mulfold) down to 2^lf = 2^(log_src - logk)
 * <= 1024 bins -> inverse transform -> samples [i0, i0 + n_out) -> channel (core/scattering1d.py:287-292, :320-327,
 * :358-364), and its adjoint for the backward pass (gsrc (+)= ...). */
int tebscat_large_leaf(tebscat_large* ctx, const float* src_dev, const float* filt_dev, int64_t B, int log_src, int logk,
                       uint32_t chunk_mask, int log_chunk, int scale_exp, int i0, int n_out, int n_paths, int channel,
                       float* out_dev, void* stream);
int tebscat_large_leaf_adjoint(tebscat_large* ctx, const float* gout_dev, const float* filt_dev, int64_t B, int log_src, int logk,
                               uint32_t chunk_mask, int log_chunk, int scale_exp, int i0, int n_out, int n_paths, int channel,
                               float* gsrc_dev, int accumulate, void* stream);

/* ---- backward pass: the reference is differentiable through torch autograd with ModulusStable
 * (kymatio/backend/torch_backend.py:5-96).  Adjoints of the ops above; tebscat/large.py walks the cascade backwards. */
/* dst = (|src|, 0), out of place (the pre-modulus signal stays available for the backward of the modulus) */
int tebscat_large_modulus_to(tebscat_large* ctx, const float* src_dev, float* dst_dev, int64_t n_complex, void* stream);
/* ModulusStable.backward (:59-96): grad <- Re(grad) * u / |u|, 0 where |u| = 0 */
int tebscat_large_modulus_backward(tebscat_large* ctx, const float* u_dev, float* grad_dev, int64_t n_complex, void* stream);
/* adjoint of tebscat_large_mulfold: gsrc[b, p] (+)= 2^-scale_exp f[p] gdst[b, p >> logk] (bit-reversed order) */
int tebscat_large_unfold(tebscat_large* ctx, const float* gdst_dev, const float* filt_dev, float* gsrc_dev, int64_t B,
                         int log_src, int logk, uint32_t chunk_mask, int log_chunk, int scale_exp, int accumulate, void* stream);
/* adjoint of tebscat_large_store: zero signal with gout[b, channel, :] in the real part of samples [i0, i0 + n_out) */
int tebscat_large_unstore(tebscat_large* ctx, const float* gout_dev, int64_t B, int log_len, int i0, int n_out, int n_paths,
                          int channel, float* buf_dev, void* stream);
/* adjoint of the un-averaged store (average=False: the unpadded modulus itself is the output, core/scattering1d.py:329-330,
 * :366-367): grow[b, offset : offset + len] (rows of row_stride floats) -> real part of samples [i0, i0 + len) of the
 * length-2^log_len buffer; accumulate == 0 writes the whole buffer (zero elsewhere), != 0 adds inside the window */
int tebscat_large_unstore_row(tebscat_large* ctx, const float* grow_dev, int64_t B, int64_t row_stride, int64_t offset,
                              int log_len, int i0, int len, int accumulate, float* buf_dev, void* stream);
/* adjoint of tebscat_large_pad_load: real part of the padded gradient folded back onto the N samples */
int tebscat_large_pad_adjoint(tebscat_large* ctx, const float* gu_dev, int64_t B, int N, int pad_left, int log2_Np,
                              float* gx_dev, void* stream);

/* Measured FP32 FMA peak of `device` in TFLOP/s (bench.py's FP32 roofline denominator;
 * MEASURED_PEAKS.json has no FP32 figure). */
int tebscat_bench_fp32_peak(int device, double* tflops_out);

const char* tebscat_last_error(void);
int tebscat_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TEBSCAT_H_ */
