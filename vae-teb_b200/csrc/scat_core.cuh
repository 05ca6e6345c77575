// scat_core.cuh -- the per-CTA "step interpreter" of the fused scattering cascade.
//
// One CTA owns one signal at a time.  The whole cascade of
//   kymatio/kymatio/scattering1d/core/scattering1d.py:269-370
// (reflect pad, FFT, psi multiply + Fourier periodisation, reduced-length iFFT,
// modulus, FFT, second-order bank, phi low-pass, unpad) runs out of shared memory
// as a host-built schedule of STEPS; each step is a set of independent TASKS
// assigned to disjoint thread ranges and ends in one __syncthreads().
//
// Spectral data lives in BIT-REVERSED bin order: forward transforms are
// decimation-in-frequency (natural in -> bit-reversed out), inverse transforms
// decimation-in-time (bit-reversed in -> natural out), both in place.  In that
// order the Fourier periodisation of torch_backend.py:18-48,
//   out[m] = mean_i in[i * L/k + m],
// is a sum over k ADJACENT slots, so "filter multiply + periodise" is one
// streaming pass and no transposition ever happens.  Filters are stored
// pre-permuted (plan.py).
//
// The same source compiles as plain C++ (TEBSCAT_HOST_EMU) where a test harness
// executes the threads of a step one after another -- valid because within a
// step no task reads a slot another thread writes.  That build is test
// infrastructure only; the product path is the CUDA kernel in tebscat.cu.
#pragma once

#include <stdint.h>
#include <math.h>

#ifdef TEBSCAT_HOST_EMU
#define TEB_D inline
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
#define TEB_LDG(p) (*(p))
#define TEB_UNROLL
#define TEB_UNROLL2
#define TEB_UNROLL4
#define TEB_FFS(x) __builtin_ffs((int)(x))
#define TEB_CLZ(x) __builtin_clz((unsigned)(x))
#define __popc(x) __builtin_popcount(x)
#else
#include <cuda_runtime.h>
#define TEB_D __device__ __forceinline__
#define TEB_LDG(p) __ldg(p)
#define TEB_UNROLL _Pragma("unroll")
#define TEB_UNROLL2 _Pragma("unroll 2")
#define TEB_UNROLL4 _Pragma("unroll 4")
#define TEB_FFS(x) __ffs((int)(x))
#define TEB_CLZ(x) __clz((int)(x))
#endif

namespace tebscat {

// ---- task encoding (12 x int32); keep in sync with tebscat/schedule.py ----------
enum : int32_t {
    OP_NOP = 0,
    OP_LOAD = 1,     // a=dst                              reflect-padded signal -> (x, 0)
    OP_FFT = 2,      // a=region b=butterflies (all blocks) c=log2B d=log2R e=flags(FFT_INV|FFT_MOD)
    OP_MULFOLD = 3,  // a=src b=log2Lsrc c=log2k d=dst e=filter offset (floats) f=chunk mask g=fused first inverse radix; scale 2^-sexp
    OP_STOREB = 4,   // a=pool base b=slots c=first index d=count e=channel-table offset f=log2 slot length
    OP_STOREZ = 5,   // a=src b=row (filter index) c=first index d=count: complex crop -> global (phase stage A)
    OP_TINY = 6,     // a=region b=transforms c=log2L (1..3) e=flags: whole transforms of 2, 4 or 8 samples
    OP_LOADPAIR = 8, // phase stage B on the interpreter: a=dst b=row of the job; c(t) = |z_i| e^{i p theta_i} conj(z_j)
                     // of the row's pair, reflect-padded (kymatio_phase_scattering.py:211-218, :283/:339, :162-209)
    OP_LOADC = 10,   // large-support level: a=dst b=slots: complex tile of the job's global buffer -> shared memory
    OP_STOREC = 11,  //                      a=src b=slots: and back
    OP_STOREU = 9,   // average=False: a=src c=first index d=count e=offset in the output row: the modulus itself
    OP_MULFOLD2 = 7, // like MULFOLD with k >= 1 on a PACKED source (spectrum of u_a + i u_b): a=src b=log2Lsrc c=log2k
                     // d=dst of the a-child e=filter offset f=chunk mask g=dst of the b-child h=log2 chunk width
    OP_GMULFOLD2 = 13, // two filters on ONE read of the global source (the partners of a packed pair): a=dst of filter B
                     // b=log2Lsrc c=log2k d=dst of filter A e=filter A f=chunk mask (shared) g=filter B
                     // h=log2 chunk width | fused first inverse radix << 8; the source starts at SignalCtx::gsrc
    OP_GMULFOLD = 12 // MULFOLD whose SOURCE is a spectrum in global memory (SignalCtx::gsrc + a, up to 2^17 bins, bit-reversed
                     // order, unswizzled): the subtrees of <= 8192 samples under a longer parent run on the fused cascade
                     // (DESIGN 6.1).  Fields as MULFOLD.  Only the kernel variant built with GSRC executes it.
};
enum : int32_t { Z_CART = 1, Z_POLAR = 2 };
// FFT_PACK (with FFT_INV | FFT_FUSE_FWD): the moduli of two transforms -- block i at `a` and block i
// at `f` (its partner, for i < g) -- enter ONE forward transform as real and imaginary part.
enum : int32_t { FFT_INV = 1, FFT_MOD = 2, FFT_FUSE_FWD = 4, FFT_PACK = 8 };
constexpr int kTaskInts = 12;

struct Task {
    int32_t op;    // opcode | (sexp << 8)
    int32_t t0;    // first thread of the range
    int32_t nt;    // threads in the range
    int32_t a, b, c, d, e, f, g, h, pad;
};

constexpr int kMaxPairRows = 8;    // (sample, pair) rows of one phase stage-B job

struct SignalCtx {
    const float* x;          // this signal's N input samples
    const float* win;        // optional analysis window applied to the samples as they are loaded ([N] or null): the
                             // Tukey taper of KymatioPhaseScattering1D(tukey_alpha=...) (kymatio_phase_scattering.py:362-392, :405-407)
    float* out;              // this signal's [n_paths, n_out] block
    const int32_t* chan;     // channel table of the batched stores
    float2* zc;              // phase stage A: analytic signals of this job, cartesian [F][N]
    float2* zp;              // phase stage A: analytic signals of this job, polar (|z|, theta) [F][N]
    int32_t z_mode;          // Z_CART | Z_POLAR
    int32_t N, pad_left, log2_Np, n_out;
    // optional output epilogue (hdf5_dataset/hdf5_dataset.py:18-137, :733-741, :758-759): per-channel
    // log / asinh transform, (x - mean) / (std + 1e-8), trim of `trim` samples at both ends, and the
    // (channels, time) -> (time, channels) transposition the model consumes
    const float* ep_mean;    // [C] or null: no epilogue
    const float* ep_std;     // [C]
    const unsigned char* ep_mode;   // [C]: EP_NONE / EP_LOG / EP_ASINH
    float ep_log_eps;
    int32_t ep_trim, ep_time_major, n_paths;
    // phase stage B (OP_LOADPAIR): the rows (sample, pair) of this job; ch_limit = rows that exist
    const float2* pr_zp[kMaxPairRows];  // (|z_i|, theta_i)[N] of the row's 'i' filter
    const float2* pr_zc[kMaxPairRows];  // (re, im)[N] of the row's 'j' filter
    float pr_pw[kMaxPairRows];
    int32_t ch_limit;        // channels >= ch_limit are not stored (n_paths for the scattering transform)
    // large-support level (OP_LOADC / OP_STOREC): this job's tile of a global complex buffer
    float2* gbuf;
    int32_t g_valid;         // complex elements of the tile that exist (the last job may be partial)
    // padding rule of OP_LOAD / OP_LOADPAIR: the scattering transform always reflects (torch_backend.py:50-78);
    // the phase module's _pad_signal (kymatio_phase_scattering.py:162-173) also offers 'constant' and 'circular'
    int32_t border;
    // OP_GMULFOLD: this job's source spectrum in global memory (null unless the kernel variant with GSRC runs)
    const float2* gsrc;
};
enum : int32_t { EP_NONE = 0, EP_LOG = 1, EP_ASINH = 2 };
enum : int32_t { BORDER_REFLECT = 0, BORDER_CONSTANT = 1, BORDER_CIRCULAR = 2 };

constexpr int kLog2TwMax = 13;                 // twiddle tables cover lengths up to 8192
constexpr int kTwA = 1 << (kLog2TwMax - 7);    // coarse table entries: W^(128 a)
constexpr int kTwB = 128;                      // fine table entries:   W^b
// both tables are stored with one pad entry per 16 so that lanes reading entries a
// power-of-two apart (the usual case: k = i0 * 2^m) hit distinct banks
constexpr int kTwAP = kTwA + kTwA / 16;
constexpr int kTwBP = kTwB + kTwB / 16;

// modulus (kymatio/backend/torch_backend.py:57): MUFU.SQRT on the device (<= 2 ulp, no
// slow path), sqrtf in the host emulator.
#ifdef TEBSCAT_HOST_EMU
TEB_D float teb_sqrt(float v) { return sqrtf(v); }
#else
TEB_D float teb_sqrt(float v) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
#endif

// ---- complex helpers -----------------------------------------------------------
// On the device every complex operation is written with the PACKED fp32 forms of sm_100
// (add/mul/fma.rn.f32x2 on a 64-bit register pair): measured on B200 (tools/micro/fp2_rate.cu) the scalar
// forms are limited by register-operand bandwidth -- FADD 93, three-register FFMA 62 lane-ops/cycle/SM --
// while FADD2/FMUL2 reach 124 and FFMA2 82, in half the issue slots.  ptxas folds the mov.b64 packs below
// into operand modifiers (broadcast .F32, swap .LO_HI, per-half negation), so a complex add is ONE
// instruction and a complex multiply TWO.  The formulas -- which product is rounded, which is fused -- are
// the same in the scalar host versions, so the host emulator stays bit-comparable.
#ifdef TEBSCAT_HOST_EMU
TEB_D float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
TEB_D float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
TEB_D float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x));
}
TEB_D float2 cmulc(float2 a, float2 b) {   // a * conj(b)
    return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -(a.x * b.y)));
}
// multiply by the constant (c + SGN*i*s)
template <int SGN> TEB_D float2 cmulk(float2 a, float c, float s) {
    return SGN < 0 ? make_float2(fmaf(a.x, c, a.y * s), fmaf(a.y, c, -(a.x * s)))
                   : make_float2(fmaf(a.x, c, -(a.y * s)), fmaf(a.y, c, a.x * s));
}
TEB_D float2 cmul_r(float2 z, float g) { return make_float2(z.x * g, z.y * g); }                 // z * real
TEB_D float2 cfma_r(float2 z, float g, float2 acc) { return make_float2(fmaf(z.x, g, acc.x), fmaf(z.y, g, acc.y)); }
#else
typedef unsigned long long teb_u64;
TEB_D teb_u64 pk(float x, float y) { teb_u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
TEB_D teb_u64 pk(float2 a) { return pk(a.x, a.y); }
TEB_D float2 unpk(teb_u64 v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
TEB_D teb_u64 add2(teb_u64 a, teb_u64 b) { teb_u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
TEB_D teb_u64 mul2(teb_u64 a, teb_u64 b) { teb_u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
TEB_D teb_u64 fma2(teb_u64 a, teb_u64 b, teb_u64 c) {
    teb_u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
TEB_D float2 cadd(float2 a, float2 b) { return unpk(add2(pk(a), pk(b))); }
TEB_D float2 csub(float2 a, float2 b) { return unpk(add2(pk(a), pk(-b.x, -b.y))); }
TEB_D float2 cmul(float2 a, float2 b) {
    const float2 t = unpk(mul2(pk(a.y, a.y), pk(b.y, b.x)));             // (a.y b.y, a.y b.x)
    return unpk(fma2(pk(a.x, a.x), pk(b), pk(-t.x, t.y)));
}
TEB_D float2 cmulc(float2 a, float2 b) {   // a * conj(b)
    const float2 t = unpk(mul2(pk(b.y, b.y), pk(a.y, a.x)));             // (a.y b.y, a.x b.y)
    return unpk(fma2(pk(b.x, b.x), pk(a), pk(t.x, -t.y)));
}
template <int SGN> TEB_D float2 cmulk(float2 a, float c, float s) {
    const float2 t = unpk(mul2(pk(s, s), pk(a.y, a.x)));                 // (a.y s, a.x s)
    return SGN < 0 ? unpk(fma2(pk(c, c), pk(a), pk(t.x, -t.y))) : unpk(fma2(pk(c, c), pk(a), pk(-t.x, t.y)));
}
TEB_D float2 cmul_r(float2 z, float g) { return unpk(mul2(pk(z), pk(g, g))); }
TEB_D float2 cfma_r(float2 z, float g, float2 acc) { return unpk(fma2(pk(z), pk(g, g), pk(acc))); }
#endif
// multiply by exp(SGN * i*pi/2)
template <int SGN> TEB_D float2 rot90(float2 a) {
    return SGN < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}

// shared-memory slot of logical complex index i: one pad slot after every 16.  It keeps
// every access pattern of the passes below (unit stride across lanes, 16 contiguous per
// thread, stride 2^m) free of bank conflicts for 8-byte accesses, and it is AFFINE for
// strides that are multiples of 16: pad(p + j*s) = pad(p) + j*(s + s/16), so a butterfly
// addresses its R elements with one pad() and constant increments.  Buffers start on
// multiples of 16 slots.
TEB_D int swz(int i) { return i + (i >> 4); }

// ---- register DFTs --------------------------------------------------------------
// dft<R, SGN>(v): v[] <- DFT_R of v[] with kernel exp(SGN*2*pi*i*j*q/R); on return
// register r holds frequency qmap<R>(r).
template <int R> TEB_D constexpr int qmap(int r) {
    return R == 16 ? ((r >> 2) + 4 * (r & 3)) : R == 8 ? ((r >> 2) + 2 * (r & 3)) : r;
}
template <int LOGR> TEB_D constexpr int brev(int q) {
    int r = 0;
    for (int b = 0; b < LOGR; ++b) r |= ((q >> b) & 1) << (LOGR - 1 - b);
    return r;
}

template <int SGN> TEB_D void dft2(float2& a, float2& b) {
    float2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}
template <int SGN> TEB_D void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = rot90<SGN>(csub(a1, a3));
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = cadd(t1, t3);
    a3 = csub(t1, t3);
}

template <int R, int SGN> struct Dft;
template <int SGN> struct Dft<2, SGN> {
    static TEB_D void run(float2 (&v)[2]) { dft2<SGN>(v[0], v[1]); }
};
template <int SGN> struct Dft<4, SGN> {
    static TEB_D void run(float2 (&v)[4]) { dft4<SGN>(v[0], v[1], v[2], v[3]); }
};
template <int SGN> struct Dft<8, SGN> {
    static TEB_D void run(float2 (&v)[8]) {
        const float h = 0.70710678118654752f;
        TEB_UNROLL for (int b = 0; b < 4; ++b) dft2<SGN>(v[b], v[4 + b]);
        v[5] = cmulk<SGN>(v[5], h, h);        // w8^1
        v[6] = rot90<SGN>(v[6]);              // w8^2
        v[7] = cmulk<SGN>(v[7], -h, h);       // w8^3
        dft4<SGN>(v[0], v[1], v[2], v[3]);
        dft4<SGN>(v[4], v[5], v[6], v[7]);
    }
};
template <int SGN> struct Dft<16, SGN> {
    static TEB_D void run(float2 (&v)[16]) {
        const float h = 0.70710678118654752f, c1 = 0.92387953251128674f, s1 = 0.38268343236508977f;
        TEB_UNROLL for (int b = 0; b < 4; ++b) dft4<SGN>(v[b], v[4 + b], v[8 + b], v[12 + b]);
        // v[4*q1 + b] *= w16^(b*q1)
        v[5] = cmulk<SGN>(v[5], c1, s1);      // w16^1
        v[6] = cmulk<SGN>(v[6], h, h);        // w16^2
        v[7] = cmulk<SGN>(v[7], s1, c1);      // w16^3
        v[9] = cmulk<SGN>(v[9], h, h);        // w16^2
        v[10] = rot90<SGN>(v[10]);            // w16^4
        v[11] = cmulk<SGN>(v[11], -h, h);     // w16^6
        v[13] = cmulk<SGN>(v[13], s1, c1);    // w16^3
        v[14] = cmulk<SGN>(v[14], -h, h);     // w16^6
        v[15] = cmulk<SGN>(v[15], -c1, -s1);  // w16^9
        TEB_UNROLL for (int q1 = 0; q1 < 4; ++q1)
            dft4<SGN>(v[4 * q1], v[4 * q1 + 1], v[4 * q1 + 2], v[4 * q1 + 3]);
    }
};

// W_8192^k = exp(-2*pi*i*k/8192) from the two shared-memory tables
TEB_D float2 twiddle(const float2* twA, const float2* twB, int k) {
    const int a = k >> 7, b = k & 127;
    return cmul(twA[a + (a >> 4)], twB[b + (b >> 4)]);
}

// w^q for q = 1..R-1 from the base powers wb[i] = w^(2^i): at most popcount(q)-1 products
template <int LOGR> TEB_D float2 twiddle_power(const float2 (&wb)[LOGR], int q) {
    float2 w = make_float2(1.f, 0.f);
    bool first = true;
    TEB_UNROLL for (int i = 0; i < LOGR; ++i) {
        if ((q >> i) & 1) {
            w = first ? wb[i] : cmul(w, wb[i]);
            first = false;
        }
    }
    return w;
}

#ifdef TEBSCAT_PROF_BFLY
__device__ long long g_bfly_dbg[8];
#endif

// Shared-memory accesses of the butterflies go through BYTE offsets (slot index * 8, computed once
// per butterfly and reused by the stores): the address of an access is then "constant base + register"
// and costs no instruction of its own.
TEB_D float2 sld(const float2* S, int byte_off) {
    return *reinterpret_cast<const float2*>(reinterpret_cast<const char*>(S) + byte_off);
}
TEB_D void sst(float2* S, int byte_off, float2 v) {
    *reinterpret_cast<float2*>(reinterpret_cast<char*>(S) + byte_off) = v;
}

// One radix-2^LOGR pass over butterfly u of a length-2^logL transform stored at `base`.
//   forward (INV=0): decimation in frequency, block size 2^logB, natural -> bit-reversed
//   inverse (INV=1): decimation in time, the exact adjoint of the forward pass
// Unit-stride passes (logs == 0) have i0 == 0, so their twiddles are exactly 1: no special case.
template <int LOGR, bool INV, bool MOD, bool FUSE = false>
TEB_D void fft_butterfly(float2* S, const float2* twA, const float2* twB, int base, int logB, int u,
                         int partner = 0, int n_paired = 0) {
    constexpr int R = 1 << LOGR;
#ifdef TEBSCAT_PROF_BFLY
    const bool dbg = threadIdx.x == 0 && blockIdx.x == 0 && !INV && LOGR == 4;
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    if (dbg) { c0 = clock64(); g_bfly_dbg[6] = c0; }
#endif
    const int logs = logB - LOGR;                  // log2 of the sub-block stride
    const int i0 = u & ((1 << logs) - 1);
    const int blk = u >> logs;
    const int p0 = base + (blk << logB) + i0;
    float2 v[R];
    float2 wb[LOGR];
    {
        const int k1 = i0 << (kLog2TwMax - logB);
        TEB_UNROLL for (int i = 0; i < LOGR; ++i) wb[i] = twiddle(twA, twB, k1 << i);
    }
    // element j of the butterfly lives at slot swz(p0) + swz(j << logs) (p0's low four bits are below
    // the stride, so the pad terms add without carry); for strides >= 16 that is affine in j
    int slot[R];
    {
        const int s0 = swz(p0) << 3;
        if (logs >= 4) {
            const int ds = ((1 << logs) + (1 << (logs - 4))) << 3;
            TEB_UNROLL for (int j = 0; j < R; ++j) slot[j] = s0 + j * ds;
        } else {
            TEB_UNROLL for (int j = 0; j < R; ++j) slot[j] = s0 + (swz(j << logs) << 3);
        }
    }
#define TEB_SLOT(j) slot[j]
    if (!INV) {
        TEB_UNROLL for (int j = 0; j < R; ++j) v[j] = sld(S, TEB_SLOT(j));
#ifdef TEBSCAT_PROF_BFLY
        if (dbg) { float sink = 0.f; TEB_UNROLL for (int j = 0; j < R; ++j) sink += v[j].x; if (sink == 1234.5f) g_bfly_dbg[3] = 1; c1 = clock64(); }
#endif
        Dft<R, -1>::run(v);
#ifdef TEBSCAT_PROF_BFLY
        if (dbg) { float sink = 0.f; TEB_UNROLL for (int j = 0; j < R; ++j) sink += v[j].x + v[j].y; if (sink == 1234.5f) g_bfly_dbg[3] = 1; c2 = clock64(); }
#endif
        TEB_UNROLL for (int r = 0; r < R; ++r) {
            const int q = qmap<R>(r);
            float2 y = v[r];
            if (q != 0) y = cmul(y, twiddle_power<LOGR>(wb, q));
            sst(S, TEB_SLOT(brev<LOGR>(q)), y);
        }
#ifdef TEBSCAT_PROF_BFLY
        if (dbg) { c3 = clock64(); g_bfly_dbg[0] += c1 - c0; g_bfly_dbg[1] += c2 - c1; g_bfly_dbg[2] += c3 - c2; g_bfly_dbg[3] += 1; }
#endif
    } else {
        TEB_UNROLL for (int q = 0; q < R; ++q) {
            float2 y = sld(S, TEB_SLOT(brev<LOGR>(q)));
            if (q != 0) y = cmulc(y, twiddle_power<LOGR>(wb, q));
            v[q] = y;
        }
        Dft<R, +1>::run(v);
        if (!FUSE) {
            TEB_UNROLL for (int r = 0; r < R; ++r) {
                float2 y = v[r];
                if (MOD) y = make_float2(teb_sqrt(fmaf(y.x, y.x, y.y * y.y)), 0.f);
                sst(S, TEB_SLOT(qmap<R>(r)), y);
            }
        } else {
            // The last inverse pass and the first forward pass of "ifft -> modulus -> fft" touch
            // the same R elements of the same thread: take the modulus in registers and go
            // straight into the forward butterfly (one shared-memory round trip and one barrier less).
            float2 f[R];
            TEB_UNROLL for (int r = 0; r < R; ++r)
                f[qmap<R>(r)] = make_float2(teb_sqrt(fmaf(v[r].x, v[r].x, v[r].y * v[r].y)), 0.f);
            if (blk < n_paired) {
                // Packed pair: the partner transform's last inverse pass; its modulus becomes the
                // IMAGINARY part, so one forward transform serves both real signals
                // (FFT(u_a + i u_b) = U_a + i U_b; the consumers separate or keep them packed).
                const int delta = ((partner - base) + ((partner - base) >> 4)) << 3;    // both multiples of 16
                TEB_UNROLL for (int q = 0; q < R; ++q) {
                    float2 y = sld(S, TEB_SLOT(brev<LOGR>(q)) + delta);
                    if (q != 0) y = cmulc(y, twiddle_power<LOGR>(wb, q));
                    v[q] = y;
                }
                Dft<R, +1>::run(v);
                TEB_UNROLL for (int r = 0; r < R; ++r)
                    f[qmap<R>(r)].y = teb_sqrt(fmaf(v[r].x, v[r].x, v[r].y * v[r].y));
            }
            Dft<R, -1>::run(f);
            TEB_UNROLL for (int r = 0; r < R; ++r) {
                const int q = qmap<R>(r);
                float2 y = f[r];
                if (q != 0) y = cmul(y, twiddle_power<LOGR>(wb, q));
                sst(S, TEB_SLOT(brev<LOGR>(q)), y);
            }
        }
    }
}

#undef TEB_SLOT

// Unit-stride pass of a small radix (the remainder pass: last DIF / first DIT pass, no
// twiddles): 16 contiguous slots per thread and trip -- 16/R butterflies whose loads are
// all in flight together -- instead of one tiny butterfly per trip.
template <int LOGR, bool INV>
TEB_D void fft_unit_stride_group(float2* S, int first_slot) {
    constexpr int R = 1 << LOGR;
    const int q0 = swz(first_slot);                          // 16 slots of one group: contiguous
    float2 v[16];
    TEB_UNROLL for (int j = 0; j < 16; ++j) v[j] = S[q0 + j];
    TEB_UNROLL for (int b = 0; b < 16 / R; ++b) {
        float2 w[R];
        if (!INV) {
            TEB_UNROLL for (int j = 0; j < R; ++j) w[j] = v[b * R + j];
            Dft<R, -1>::run(w);
            TEB_UNROLL for (int r = 0; r < R; ++r) v[b * R + brev<LOGR>(qmap<R>(r))] = w[r];
        } else {
            TEB_UNROLL for (int q = 0; q < R; ++q) w[q] = v[b * R + brev<LOGR>(q)];
            Dft<R, +1>::run(w);
            TEB_UNROLL for (int r = 0; r < R; ++r) v[b * R + qmap<R>(r)] = w[r];
        }
    }
    TEB_UNROLL for (int j = 0; j < 16; ++j) S[q0 + j] = v[j];
}

// One pass of an FFT task over `slots` slots at t.a: radix 2^LOGR, blocks of 2^logB.
template <int LOGR>
TEB_D void fft_pass(float2* S, const float2* twA, const float2* twB, const Task& t, int lt, int logB, int flags,
                    int slots) {
    const int n_bfly = slots >> LOGR;
#ifdef TEBSCAT_PROF_BFLY
    if (threadIdx.x == 0 && blockIdx.x == 0 && LOGR == 4) g_bfly_dbg[5] = clock64();
#endif
    const bool inv = (flags & FFT_INV) != 0, mod = (flags & FFT_MOD) != 0, fuse = (flags & FFT_FUSE_FWD) != 0;
    if (LOGR <= 3 && logB == LOGR && !mod && ((slots & 15) == 0)) {
        const int n_groups = slots >> 4;
        for (int g = lt; g < n_groups; g += t.nt) {
            if (!inv) fft_unit_stride_group<(LOGR <= 3 ? LOGR : 1), false>(S, t.a + (g << 4));
            else fft_unit_stride_group<(LOGR <= 3 ? LOGR : 1), true>(S, t.a + (g << 4));
        }
        return;
    }
    // the modulus / fused passes are always radix 16 (the last inverse pass of any transform
    // of 16 samples or more): only that instantiation carries them, which keeps the kernel small
    if (!inv) {
        for (int u = lt; u < n_bfly; u += t.nt) fft_butterfly<LOGR, false, false>(S, twA, twB, t.a, logB, u);
    } else if (LOGR == 4 && fuse) {
        const int n_paired = (flags & FFT_PACK) ? t.g : 0;
        for (int u = lt; u < n_bfly; u += t.nt) fft_butterfly<4, true, true, true>(S, twA, twB, t.a, logB, u, t.f, n_paired);
    } else if (LOGR == 4 && mod) {
        for (int u = lt; u < n_bfly; u += t.nt) fft_butterfly<4, true, true>(S, twA, twB, t.a, logB, u);
    } else {
        for (int u = lt; u < n_bfly; u += t.nt) fft_butterfly<LOGR, true, false>(S, twA, twB, t.a, logB, u);
    }
}

// An FFT task is a SEQUENCE of up to four passes over the same slots.  Passes after the first are
// packed 12 bits each -- log2B | log2R << 4 | flags << 7 -- into h (two) and pad bits 4..15 (one).
// The host only chains passes whose blocks are at most 512 slots: 32 consecutive work items (radix-16
// butterflies, or 16-slot groups of the unit-stride pass) then cover the same 512 slots in every pass, so
// a warp only ever reads what it wrote itself and a warp-level fence between the passes is enough --
// no CTA barrier, no trip through the dispatcher.
TEB_D unsigned long long fft_more_passes(const Task& t) {
    return (unsigned long long)((unsigned)t.h & 0xffffffu) | ((unsigned long long)(((unsigned)t.pad >> 4) & 0xfffu) << 24);
}
TEB_D int task_passes(const Task& t) {
    if ((t.op & 0xff) != OP_FFT) return 1;
    unsigned long long more = fft_more_passes(t);
    int n = 1;
    while (more & 0xfff) { ++n; more >>= 12; }
    return n;
}
TEB_D void fft_task_pass(float2* S, const float2* twA, const float2* twB, const Task& t, int lt, int logB, int logR,
                         int flags) {
    const int slots = t.b << t.d;
    // (an if-chain, radix 16 first: no jump table through the constant cache on the hot path)
    if (logR == 4) fft_pass<4>(S, twA, twB, t, lt, logB, flags, slots);
    else if (logR == 3) fft_pass<3>(S, twA, twB, t, lt, logB, flags, slots);
    else if (logR == 2) fft_pass<2>(S, twA, twB, t, lt, logB, flags, slots);
    else fft_pass<1>(S, twA, twB, t, lt, logB, flags, slots);
}
#ifndef TEBSCAT_HOST_EMU
TEB_D void fft_task(float2* S, const float2* twA, const float2* twB, const Task& t, int lt) {
#ifdef TEBSCAT_PROF_BFLY
    if (threadIdx.x == 0 && blockIdx.x == 0) g_bfly_dbg[4] = clock64();
#endif
    unsigned long long more = fft_more_passes(t);
    int logB = t.c, logR = t.d, flags = t.e;
    for (;;) {
        fft_task_pass(S, twA, twB, t, lt, logB, logR, flags);
        if (!(more & 0xfff)) break;
        __syncwarp();
        logB = (int)(more & 15);
        logR = (int)((more >> 4) & 7);
        flags = (int)((more >> 7) & 15);
        more >>= 12;
    }
}
#endif

// Transforms of 2, 4 or 8 samples (output-rate lengths of short signals): one thread per
// transform, direct O(L^2) DFT with the eighth roots of unity as constants.  Same conventions
// as the passes above: spectra in bit-reversed order, inverse unnormalised, optional modulus
// and "inverse -> modulus -> forward" in one go.
TEB_D void tiny_task(float2* S, const Task& t, int lt) {
    const int n = t.c, L = 1 << n;
    const bool inv = (t.e & FFT_INV) != 0, mod = (t.e & FFT_MOD) != 0, fuse = (t.e & FFT_FUSE_FWD) != 0;
    const float h = 0.70710678118654752f;
    const float wr[8] = {1.f, h, 0.f, -h, -1.f, -h, 0.f, h};          // exp(-2 pi i m / 8)
    const float wi[8] = {0.f, -h, -1.f, -h, 0.f, h, 1.f, h};
    for (int u = lt; u < t.b; u += t.nt) {
        const int base = t.a + (u << n);
        float2 x[8], y[8];
        for (int p = 0; p < L; ++p) x[p] = S[swz(base + p)];
        auto rev = [&](int k) { int r = 0; for (int b = 0; b < n; ++b) r |= ((k >> b) & 1) << (n - 1 - b); return r; };
        if (inv) {
            for (int tt = 0; tt < L; ++tt) {                        // x_t = sum_k X_k conj(W)^(k t)
                float ax = 0.f, ay = 0.f;
                for (int k = 0; k < L; ++k) {
                    const int m = ((k * tt) << (3 - n)) & 7;
                    const float2 X = x[rev(k)];
                    ax += X.x * wr[m] + X.y * wi[m];
                    ay += X.y * wr[m] - X.x * wi[m];
                }
                y[tt] = (mod || fuse) ? make_float2(teb_sqrt(fmaf(ax, ax, ay * ay)), 0.f) : make_float2(ax, ay);
            }
            if (fuse) {
                for (int p = 0; p < L; ++p) x[p] = y[p];
            } else {
                for (int p = 0; p < L; ++p) S[swz(base + p)] = y[p];
                continue;
            }
        }
        for (int k = 0; k < L; ++k) {                               // X_k = sum_t x_t W^(k t)
            float ax = 0.f, ay = 0.f;
            for (int tt = 0; tt < L; ++tt) {
                const int m = ((k * tt) << (3 - n)) & 7;
                ax += x[tt].x * wr[m] - x[tt].y * wi[m];
                ay += x[tt].y * wr[m] + x[tt].x * wi[m];
            }
            y[rev(k)] = make_float2(ax, ay);
        }
        for (int p = 0; p < L; ++p) S[swz(base + p)] = y[p];
    }
}

// dst[m] = 2^-sexp * sum_{i<k} src[m*k + i] * filt[m*k + i]      (bit-reversed bin order)
// Filters are real fp32 in global memory (L2 resident).  Every work item covers four
// consecutive source slots with one 128-bit filter load; for k >= 4 only the 4-slot chunks
// named in the task's mask are visited (the host clears chunks where the filter is below
// 1e-9 of its peak for every output bin -- e.g. 2 of 16 chunks for the phi low-pass).
TEB_D void mulfold_task(float2* S, const float* __restrict__ arena, const Task& t, int lt) {
    const int logk = t.c;
    const float scale = ldexpf(1.0f, -(t.op >> 8));
    const float* f = arena + t.e;
    const float2 zero = make_float2(0.f, 0.f);
    if (logk >= 2) {
        // Filter layout for k >= 4: COMPACTED by the host to the active chunks only,
        // f[(m * nch + c) * 4 + r], so that consecutive outputs read consecutive memory
        // (the natural layout would stride by k floats: one L2 request per lane).
        const int n_dst = 1 << (t.b - logk);
        const unsigned mask = (unsigned)t.f;
        const int logcw = t.h;                         // chunk width 2^logcw bins (4, or k/32 for k > 128)
        const int nch = __popc(mask) << (logcw - 2);   // active 4-bin groups per output
        if (nch <= 2 && logcw == 2) {
            // at most two active chunks (every phi low-pass leaf): all filter loads of a trip --
            // 2 chunks x 4 outputs -- are issued before the first one is consumed (one L2 round trip)
            const int i0 = (TEB_FFS(mask) - 1) << 2;
            const unsigned m2 = mask & (mask - 1);
            const int i1 = m2 ? ((TEB_FFS(m2) - 1) << 2) : i0;
            for (int m0 = lt; m0 < n_dst; m0 += 4 * t.nt) {
                float4 g0[4], g1[4];
                TEB_UNROLL for (int j = 0; j < 4; ++j) {
                    const int m = m0 + j * t.nt;
                    const bool ok = m < n_dst;
                    const float4* fm = reinterpret_cast<const float4*>(f) + m * nch;
                    g0[j] = ok ? TEB_LDG(fm) : float4{0.f, 0.f, 0.f, 0.f};
                    g1[j] = (ok && m2) ? TEB_LDG(fm + 1) : float4{0.f, 0.f, 0.f, 0.f};
                }
                TEB_UNROLL for (int j = 0; j < 4; ++j) {
                    const int m = m0 + j * t.nt;
                    if (m < n_dst) {
                        const int q = swz(t.a + (m << logk) + i0);
                        float2 acc = cmul_r(S[q], g0[j].x);
                        acc = cfma_r(S[q + 1], g0[j].y, acc);
                        acc = cfma_r(S[q + 2], g0[j].z, acc);
                        acc = cfma_r(S[q + 3], g0[j].w, acc);
                        if (m2) {
                            const int r = swz(t.a + (m << logk) + i1);
                            acc = cfma_r(S[r], g1[j].x, acc);
                            acc = cfma_r(S[r + 1], g1[j].y, acc);
                            acc = cfma_r(S[r + 2], g1[j].z, acc);
                            acc = cfma_r(S[r + 3], g1[j].w, acc);
                        }
                        S[swz(t.d + m)] = cmul_r(acc, scale);
                    }
                }
            }
            return;
        }
        // four outputs per thread and trip: the chunk loop is uniform over the task, so the
        // four 128-bit filter loads of one chunk are in flight together
        for (int m0 = lt; m0 < n_dst; m0 += 4 * t.nt) {
            float2 acc[4] = {zero, zero, zero, zero};
            unsigned rest = mask;
            int c = 0;
            while (rest) {
                const int i_chunk = (TEB_FFS(rest) - 1) << logcw;
                rest &= rest - 1;
                for (int sub = 0; sub < (1 << (logcw - 2)); ++sub, ++c) {
                    const int i = i_chunk + (sub << 2);
                    float4 g[4];
                    TEB_UNROLL for (int j = 0; j < 4; ++j) {
                        const int m = m0 + j * t.nt;
                        g[j] = (m < n_dst) ? TEB_LDG(reinterpret_cast<const float4*>(f) + m * nch + c)
                                           : float4{0.f, 0.f, 0.f, 0.f};
                    }
                    TEB_UNROLL for (int j = 0; j < 4; ++j) {
                        const int m = m0 + j * t.nt;
                        if (m < n_dst) {
                            const int q = swz(t.a + (m << logk) + i);      // 4 slots of one 16-group: contiguous
                            acc[j] = cfma_r(S[q], g[j].x, acc[j]);
                            acc[j] = cfma_r(S[q + 1], g[j].y, acc[j]);
                            acc[j] = cfma_r(S[q + 2], g[j].z, acc[j]);
                            acc[j] = cfma_r(S[q + 3], g[j].w, acc[j]);
                        }
                    }
                }
            }
            TEB_UNROLL for (int j = 0; j < 4; ++j) {
                const int m = m0 + j * t.nt;
                if (m < n_dst) S[swz(t.d + m)] = cmul_r(acc[j], scale);
            }
        }
    } else {
        const int n_items = 1 << (t.b - 2);                    // 4 source slots per item
        TEB_UNROLL4 for (int it = lt; it < n_items; it += t.nt) {
            const float4 g = TEB_LDG(reinterpret_cast<const float4*>(f + 4 * it));
            const int q = swz(t.a + 4 * it);
            const float2 z0 = S[q], z1 = S[q + 1], z2 = S[q + 2], z3 = S[q + 3];
            if (logk == 0) {
                const int o = swz(t.d + 4 * it);
                float2 w0 = cmul_r(z0, g.x * scale);
                float2 w1 = cmul_r(z1, g.y * scale);
                float2 w2 = cmul_r(z2, g.z * scale);
                float2 w3 = cmul_r(z3, g.w * scale);
                // optionally the first (unit-stride, twiddle-free) inverse pass of the transform
                // that follows, on the four slots this thread owns: one round trip less
                if (t.g == 1) {
                    dft2<+1>(w0, w1);
                    dft2<+1>(w2, w3);
                } else if (t.g == 2) {
                    float2 a0 = w0, a1 = w2, a2 = w1, a3 = w3;        // inputs in bit-reversed order
                    dft4<+1>(a0, a1, a2, a3);
                    w0 = a0; w1 = a1; w2 = a2; w3 = a3;
                }
                S[o] = w0;
                S[o + 1] = w1;
                S[o + 2] = w2;
                S[o + 3] = w3;
            } else {
                const int o = swz(t.d + 2 * it);
                S[o] = cmul_r(cfma_r(z0, g.x, cmul_r(z1, g.y)), scale);
                S[o + 1] = cmul_r(cfma_r(z2, g.z, cmul_r(z3, g.w)), scale);
            }
        }
    }
}

// Four consecutive complex bins of a global spectrum (32-byte aligned: two 128-bit loads).  Plain coherent loads, not
// the read-only path: a schedule that parks U0 in its CTA's scratch rewrites that memory for every signal.
TEB_D void gload4(const float2* g, float2 (&z)[4]) {
#ifdef TEBSCAT_HOST_EMU
    z[0] = g[0]; z[1] = g[1]; z[2] = g[2]; z[3] = g[3];
#else
    const float4 lo = *reinterpret_cast<const float4*>(g), hi = *(reinterpret_cast<const float4*>(g) + 1);
    z[0] = make_float2(lo.x, lo.y); z[1] = make_float2(lo.z, lo.w);
    z[2] = make_float2(hi.x, hi.y); z[3] = make_float2(hi.z, hi.w);
#endif
}

// MULFOLD with the source spectrum in GLOBAL memory (large-support level): the parent -- the signal's
// spectrum U0, or the spectrum of a first-order modulus of more than 8192 samples -- was produced by the level's
// own kernels; everything below it that fits one SM runs here.  Same arithmetic as mulfold_task (same products, same
// order); every trip has its filter AND source loads in flight before the first use (one trip to L2 / HBM).
TEB_D void gmulfold_task(float2* S, const float* __restrict__ arena, const SignalCtx& c, const Task& t, int lt) {
    const int logk = t.c;
    const float scale = ldexpf(1.0f, -(t.op >> 8));
    const float* f = arena + t.e;
    const float2* G = c.gsrc + t.a;
    if (logk >= 2) {
        const int n_dst = 1 << (t.b - logk);
        const unsigned mask = (unsigned)t.f;
        const int logcw = t.h;
        const int nch = __popc(mask) << (logcw - 2);
        const float2 zero = make_float2(0.f, 0.f);
        for (int m0 = lt; m0 < n_dst; m0 += 4 * t.nt) {
            float2 acc[4] = {zero, zero, zero, zero};
            unsigned rest = mask;
            int cc = 0;
            while (rest) {
                const int i_chunk = (TEB_FFS(rest) - 1) << logcw;
                rest &= rest - 1;
                for (int sub = 0; sub < (1 << (logcw - 2)); ++sub, ++cc) {
                    const int i = i_chunk + (sub << 2);
                    float4 g[4];
                    float2 z[4][4];
                    TEB_UNROLL for (int j = 0; j < 4; ++j) {
                        const int m = m0 + j * t.nt;
                        if (m < n_dst) {
                            g[j] = TEB_LDG(reinterpret_cast<const float4*>(f) + m * nch + cc);
                            gload4(G + ((long long)m << logk) + i, z[j]);
                        } else {
                            g[j] = float4{0.f, 0.f, 0.f, 0.f};
                            z[j][0] = z[j][1] = z[j][2] = z[j][3] = zero;
                        }
                    }
                    TEB_UNROLL for (int j = 0; j < 4; ++j) {
                        acc[j] = cfma_r(z[j][0], g[j].x, acc[j]);
                        acc[j] = cfma_r(z[j][1], g[j].y, acc[j]);
                        acc[j] = cfma_r(z[j][2], g[j].z, acc[j]);
                        acc[j] = cfma_r(z[j][3], g[j].w, acc[j]);
                    }
                }
            }
            TEB_UNROLL for (int j = 0; j < 4; ++j) {
                const int m = m0 + j * t.nt;
                if (m < n_dst) S[swz(t.d + m)] = cmul_r(acc[j], scale);
            }
        }
    } else {
        const int n_items = 1 << (t.b - 2);                    // 4 source bins per item
        for (int it0 = lt; it0 < n_items; it0 += 4 * t.nt) {
            float4 gg[4];
            float2 zz[4][4];
            TEB_UNROLL for (int j = 0; j < 4; ++j) {           // four items in flight per thread
                const int it = it0 + j * t.nt;
                if (it < n_items) {
                    gg[j] = TEB_LDG(reinterpret_cast<const float4*>(f + 4 * it));
                    gload4(G + 4 * it, zz[j]);
                } else {
                    gg[j] = float4{0.f, 0.f, 0.f, 0.f};
                    zz[j][0] = zz[j][1] = zz[j][2] = zz[j][3] = make_float2(0.f, 0.f);
                }
            }
            TEB_UNROLL for (int j = 0; j < 4; ++j) {
                const int it = it0 + j * t.nt;
                if (it >= n_items) continue;
                const float4 g = gg[j];
                const float2 z0 = zz[j][0], z1 = zz[j][1], z2 = zz[j][2], z3 = zz[j][3];
                if (logk == 0) {
                    const int o = swz(t.d + 4 * it);
                    float2 w0 = cmul_r(z0, g.x * scale);
                    float2 w1 = cmul_r(z1, g.y * scale);
                    float2 w2 = cmul_r(z2, g.z * scale);
                    float2 w3 = cmul_r(z3, g.w * scale);
                    if (t.g == 1) {                            // first (unit-stride) inverse pass, as in mulfold_task
                        dft2<+1>(w0, w1);
                        dft2<+1>(w2, w3);
                    } else if (t.g == 2) {
                        float2 a0 = w0, a1 = w2, a2 = w1, a3 = w3;
                        dft4<+1>(a0, a1, a2, a3);
                        w0 = a0; w1 = a1; w2 = a2; w3 = a3;
                    }
                    S[o] = w0;
                    S[o + 1] = w1;
                    S[o + 2] = w2;
                    S[o + 3] = w3;
                } else {
                    const int o = swz(t.d + 2 * it);
                    S[o] = cmul_r(cfma_r(z0, g.x, cmul_r(z1, g.y)), scale);
                    S[o + 1] = cmul_r(cfma_r(z2, g.z, cmul_r(z3, g.w)), scale);
                }
            }
        }
    }
}

// Two filters of one scale on ONE read of the global source (OP_GMULFOLD2): the source bins cross the SM's L2 port
// once per pair instead of once per filter.  Arithmetic per filter as in gmulfold_task / mulfold_task.
TEB_D void gmulfold2_task(float2* S, const float* __restrict__ arena, const SignalCtx& c, const Task& t, int lt) {
    const int logk = t.c;
    const float scale = ldexpf(1.0f, -(t.op >> 8));
    const float* fa = arena + t.e;
    const float* fb = arena + t.g;
    const float2* G = c.gsrc;
    const int da = t.d, db = t.a;
    if (logk >= 2) {
        const int n_dst = 1 << (t.b - logk);
        const unsigned mask = (unsigned)t.f;
        const int logcw = t.h & 0xff;
        const int nch = __popc(mask) << (logcw - 2);
        const float2 zero = make_float2(0.f, 0.f);
        if (nch == 1 && logcw == 2) {
            // one active 4-bin chunk (the usual case for folds by 4 and 8): FOUR outputs per thread and trip, their
            // sixteen 128-bit loads (two filters + eight source bins each) in flight together -- one trip to L2
            const int i = (TEB_FFS(mask) - 1) << 2;
            for (int m0 = lt; m0 < n_dst; m0 += 4 * t.nt) {
                float4 ga[4], gb[4];
                float2 z[4][4];
                TEB_UNROLL for (int j = 0; j < 4; ++j) {
                    const int m = m0 + j * t.nt;
                    if (m < n_dst) {
                        ga[j] = TEB_LDG(reinterpret_cast<const float4*>(fa) + m);
                        gb[j] = TEB_LDG(reinterpret_cast<const float4*>(fb) + m);
                        gload4(G + ((long long)m << logk) + i, z[j]);
                    } else {
                        ga[j] = gb[j] = float4{0.f, 0.f, 0.f, 0.f};
                        z[j][0] = z[j][1] = z[j][2] = z[j][3] = zero;
                    }
                }
                TEB_UNROLL for (int j = 0; j < 4; ++j) {
                    const int m = m0 + j * t.nt;
                    if (m >= n_dst) continue;
                    float2 a = cfma_r(z[j][0], ga[j].x, zero), b = cfma_r(z[j][0], gb[j].x, zero);
                    a = cfma_r(z[j][1], ga[j].y, a); b = cfma_r(z[j][1], gb[j].y, b);
                    a = cfma_r(z[j][2], ga[j].z, a); b = cfma_r(z[j][2], gb[j].z, b);
                    a = cfma_r(z[j][3], ga[j].w, a); b = cfma_r(z[j][3], gb[j].w, b);
                    S[swz(da + m)] = cmul_r(a, scale);
                    S[swz(db + m)] = cmul_r(b, scale);
                }
            }
            return;
        }
        if (nch == 2 && logcw == 2) {
            // two active chunks: two outputs per trip, the loads of BOTH chunks in flight together
            const int i0 = (TEB_FFS(mask) - 1) << 2;
            const int i1 = (TEB_FFS(mask & (mask - 1)) - 1) << 2;
            for (int m0 = lt; m0 < n_dst; m0 += 2 * t.nt) {
                float4 ga[2][2], gb[2][2];
                float2 z[2][2][4];
                TEB_UNROLL for (int j = 0; j < 2; ++j) {
                    const int m = m0 + j * t.nt;
                    if (m < n_dst) {
                        ga[j][0] = TEB_LDG(reinterpret_cast<const float4*>(fa) + 2 * m);
                        ga[j][1] = TEB_LDG(reinterpret_cast<const float4*>(fa) + 2 * m + 1);
                        gb[j][0] = TEB_LDG(reinterpret_cast<const float4*>(fb) + 2 * m);
                        gb[j][1] = TEB_LDG(reinterpret_cast<const float4*>(fb) + 2 * m + 1);
                        gload4(G + ((long long)m << logk) + i0, z[j][0]);
                        gload4(G + ((long long)m << logk) + i1, z[j][1]);
                    } else {
                        ga[j][0] = ga[j][1] = gb[j][0] = gb[j][1] = float4{0.f, 0.f, 0.f, 0.f};
                        TEB_UNROLL for (int q = 0; q < 4; ++q) z[j][0][q] = z[j][1][q] = zero;
                    }
                }
                TEB_UNROLL for (int j = 0; j < 2; ++j) {
                    const int m = m0 + j * t.nt;
                    if (m >= n_dst) continue;
                    float2 a = zero, b = zero;
                    TEB_UNROLL for (int cidx = 0; cidx < 2; ++cidx) {
                        a = cfma_r(z[j][cidx][0], ga[j][cidx].x, a); b = cfma_r(z[j][cidx][0], gb[j][cidx].x, b);
                        a = cfma_r(z[j][cidx][1], ga[j][cidx].y, a); b = cfma_r(z[j][cidx][1], gb[j][cidx].y, b);
                        a = cfma_r(z[j][cidx][2], ga[j][cidx].z, a); b = cfma_r(z[j][cidx][2], gb[j][cidx].z, b);
                        a = cfma_r(z[j][cidx][3], ga[j][cidx].w, a); b = cfma_r(z[j][cidx][3], gb[j][cidx].w, b);
                    }
                    S[swz(da + m)] = cmul_r(a, scale);
                    S[swz(db + m)] = cmul_r(b, scale);
                }
            }
            return;
        }
        for (int m0 = lt; m0 < n_dst; m0 += 2 * t.nt) {
            float2 accA[2] = {zero, zero}, accB[2] = {zero, zero};
            unsigned rest = mask;
            int cc = 0;
            while (rest) {
                const int i_chunk = (TEB_FFS(rest) - 1) << logcw;
                rest &= rest - 1;
                for (int sub = 0; sub < (1 << (logcw - 2)); ++sub, ++cc) {
                    const int i = i_chunk + (sub << 2);
                    float4 ga[2], gb[2];
                    float2 z[2][4];
                    TEB_UNROLL for (int j = 0; j < 2; ++j) {
                        const int m = m0 + j * t.nt;
                        if (m < n_dst) {
                            ga[j] = TEB_LDG(reinterpret_cast<const float4*>(fa) + m * nch + cc);
                            gb[j] = TEB_LDG(reinterpret_cast<const float4*>(fb) + m * nch + cc);
                            gload4(G + ((long long)m << logk) + i, z[j]);
                        } else {
                            ga[j] = gb[j] = float4{0.f, 0.f, 0.f, 0.f};
                            z[j][0] = z[j][1] = z[j][2] = z[j][3] = zero;
                        }
                    }
                    TEB_UNROLL for (int j = 0; j < 2; ++j) {
                        accA[j] = cfma_r(z[j][0], ga[j].x, accA[j]); accB[j] = cfma_r(z[j][0], gb[j].x, accB[j]);
                        accA[j] = cfma_r(z[j][1], ga[j].y, accA[j]); accB[j] = cfma_r(z[j][1], gb[j].y, accB[j]);
                        accA[j] = cfma_r(z[j][2], ga[j].z, accA[j]); accB[j] = cfma_r(z[j][2], gb[j].z, accB[j]);
                        accA[j] = cfma_r(z[j][3], ga[j].w, accA[j]); accB[j] = cfma_r(z[j][3], gb[j].w, accB[j]);
                    }
                }
            }
            TEB_UNROLL for (int j = 0; j < 2; ++j) {
                const int m = m0 + j * t.nt;
                if (m < n_dst) {
                    S[swz(da + m)] = cmul_r(accA[j], scale);
                    S[swz(db + m)] = cmul_r(accB[j], scale);
                }
            }
        }
    } else {
        const int radix = t.h >> 8;
        const int n_items = 1 << (t.b - 2);                    // 4 source bins per item
        for (int it0 = lt; it0 < n_items; it0 += 4 * t.nt) {   // four items (sixteen 128-bit loads) in flight per thread
            float4 gga[4], ggb[4];
            float2 zz[4][4];
            TEB_UNROLL for (int j = 0; j < 4; ++j) {
                const int it = it0 + j * t.nt;
                if (it < n_items) {
                    gga[j] = TEB_LDG(reinterpret_cast<const float4*>(fa + 4 * it));
                    ggb[j] = TEB_LDG(reinterpret_cast<const float4*>(fb + 4 * it));
                    gload4(G + 4 * it, zz[j]);
                } else {
                    gga[j] = ggb[j] = float4{0.f, 0.f, 0.f, 0.f};
                    zz[j][0] = zz[j][1] = zz[j][2] = zz[j][3] = make_float2(0.f, 0.f);
                }
            }
            TEB_UNROLL for (int j = 0; j < 4; ++j) {
                const int it = it0 + j * t.nt;
                if (it >= n_items) continue;
                TEB_UNROLL for (int which = 0; which < 2; ++which) {
                    const float4 g = which ? ggb[j] : gga[j];
                    const int dst = which ? db : da;
                    const float2 z0 = zz[j][0], z1 = zz[j][1], z2 = zz[j][2], z3 = zz[j][3];
                    if (logk == 0) {
                        const int o = swz(dst + 4 * it);
                        float2 w0 = cmul_r(z0, g.x * scale);
                        float2 w1 = cmul_r(z1, g.y * scale);
                        float2 w2 = cmul_r(z2, g.z * scale);
                        float2 w3 = cmul_r(z3, g.w * scale);
                        if (radix == 1) {
                            dft2<+1>(w0, w1);
                            dft2<+1>(w2, w3);
                        } else if (radix == 2) {
                            float2 a0 = w0, a1 = w2, a2 = w1, a3 = w3;
                            dft4<+1>(a0, a1, a2, a3);
                            w0 = a0; w1 = a1; w2 = a2; w3 = a3;
                        }
                        S[o] = w0;
                        S[o + 1] = w1;
                        S[o + 2] = w2;
                        S[o + 3] = w3;
                    } else {
                        const int o = swz(dst + 2 * it);
                        S[o] = cmul_r(cfma_r(z0, g.x, cmul_r(z1, g.y)), scale);
                        S[o + 1] = cmul_r(cfma_r(z2, g.z, cmul_r(z3, g.w)), scale);
                    }
                }
            }
        }
    }
}

// Slot of bin -k in a bit-reversed spectrum when p is the slot of bin k: the highest set bit of p
// stays, every bit below it flips (the mirror image inside p's dyadic block).
TEB_D int mirror_slot(int p) { return p ? (p ^ ((1 << (31 - TEB_CLZ(p))) - 1)) : 0; }

// MULFOLD on a PACKED source Z = FFT(u_a + i u_b) of two real signals (see FFT_PACK):
//   U_a[k] = (Z[k] + conj Z[-k]) / 2,   U_b[k] = (Z[k] - conj Z[-k]) / (2i),
// so with A = sum f Z[p] and Bc = sum f Z[mirror(p)] over the k slots of an output bin
//   dst_a[m] = (A + conj Bc) / 2,       dst_b[m] = -i (A - conj Bc) / 2
// (the 1/2 is folded into the task's power-of-two scale).  One filter load serves both children.
// the source slots of one 4-slot group p..p+3 and of their mirrors (bins -k), as (z[4], y[4])
TEB_D void load_group_and_mirror(const float2* S, int base, int p, float2 (&z)[4], float2 (&y)[4]) {
    const int q = swz(base + p);
    z[0] = S[q]; z[1] = S[q + 1]; z[2] = S[q + 2]; z[3] = S[q + 3];
    if (p >= 4) {                                   // mirrors of p..p+3: pm, pm-1, pm-2, pm-3 (one aligned group)
        const int r = swz(base + mirror_slot(p) - 3);
        y[3] = S[r]; y[2] = S[r + 1]; y[1] = S[r + 2]; y[0] = S[r + 3];
    } else {                                        // slots 0,1,2,3 <-> 0,1,3,2
        y[0] = z[0]; y[1] = z[1]; y[2] = z[3]; y[3] = z[2];
    }
}

// outputs of one packed-source bin: A = sum f Z[p], Bc = sum f Z[mirror(p)]
//   dst_a = (A + conj Bc) * scale,   dst_b = -i (A - conj Bc) * scale
TEB_D void mulfold2_store(float2* S, int oa, int ob, float2 A, float2 Bc, float scale) {
    const float2 cb = make_float2(Bc.x, -Bc.y);
    S[oa] = cmul_r(cadd(A, cb), scale);
    S[ob] = cmul_r(rot90<-1>(csub(A, cb)), scale);
}

#ifndef TEBSCAT_MF2_ITEMS
#define TEBSCAT_MF2_ITEMS 4
#endif
constexpr int kMf2Items = TEBSCAT_MF2_ITEMS;       // 4-slot items one thread takes per trip of a k < 4 MULFOLD2 (one trip to L2)

TEB_D void mulfold2_task(float2* S, const float* __restrict__ arena, const Task& t, int lt) {
    const int logk = t.c;
    const float scale = ldexpf(1.0f, -(t.op >> 8));
    const float* f = arena + t.e;
    const float2 zero = make_float2(0.f, 0.f);
    if (logk >= 2) {
        const int n_dst = 1 << (t.b - logk);
        const unsigned mask = (unsigned)t.f;
        const int logcw = t.h;
        const int nch = __popc(mask) << (logcw - 2);
        if (nch <= 2 && logcw == 2) {
            // at most two active chunks (the usual case for the second-order bank): the filter loads of
            // four outputs are all in flight before the first is used -- one L2 round trip per trip
            const int i0 = (TEB_FFS(mask) - 1) << 2;
            const unsigned m2 = mask & (mask - 1);
            const int i1 = m2 ? ((TEB_FFS(m2) - 1) << 2) : i0;
            for (int m0 = lt; m0 < n_dst; m0 += 4 * t.nt) {
                float4 g0[4], g1[4];
                TEB_UNROLL for (int j = 0; j < 4; ++j) {
                    const int m = m0 + j * t.nt;
                    const bool ok = m < n_dst;
                    const float4* fm = reinterpret_cast<const float4*>(f) + m * nch;
                    g0[j] = ok ? TEB_LDG(fm) : float4{0.f, 0.f, 0.f, 0.f};
                    g1[j] = (ok && m2) ? TEB_LDG(fm + 1) : float4{0.f, 0.f, 0.f, 0.f};
                }
                TEB_UNROLL for (int j = 0; j < 4; ++j) {
                    const int m = m0 + j * t.nt;
                    if (m < n_dst) {
                        float2 z[4], y[4];
                        load_group_and_mirror(S, t.a, (m << logk) + i0, z, y);
                        float2 A = cmul_r(z[0], g0[j].x), Bc = cmul_r(y[0], g0[j].x);
                        A = cfma_r(z[1], g0[j].y, A); Bc = cfma_r(y[1], g0[j].y, Bc);
                        A = cfma_r(z[2], g0[j].z, A); Bc = cfma_r(y[2], g0[j].z, Bc);
                        A = cfma_r(z[3], g0[j].w, A); Bc = cfma_r(y[3], g0[j].w, Bc);
                        if (m2) {
                            load_group_and_mirror(S, t.a, (m << logk) + i1, z, y);
                            A = cfma_r(z[0], g1[j].x, A); Bc = cfma_r(y[0], g1[j].x, Bc);
                            A = cfma_r(z[1], g1[j].y, A); Bc = cfma_r(y[1], g1[j].y, Bc);
                            A = cfma_r(z[2], g1[j].z, A); Bc = cfma_r(y[2], g1[j].z, Bc);
                            A = cfma_r(z[3], g1[j].w, A); Bc = cfma_r(y[3], g1[j].w, Bc);
                        }
                        mulfold2_store(S, swz(t.d + m), swz(t.g + m), A, Bc, scale);
                    }
                }
            }
            return;
        }
        for (int m0 = lt; m0 < n_dst; m0 += 2 * t.nt) {
            float2 A[2] = {zero, zero}, Bc[2] = {zero, zero};
            unsigned rest = mask;
            int c = 0;
            while (rest) {
                const int i_chunk = (TEB_FFS(rest) - 1) << logcw;
                rest &= rest - 1;
                for (int sub = 0; sub < (1 << (logcw - 2)); ++sub, ++c) {
                    const int i = i_chunk + (sub << 2);
                    float4 g[2];
                    TEB_UNROLL for (int j = 0; j < 2; ++j) {
                        const int m = m0 + j * t.nt;
                        g[j] = (m < n_dst) ? TEB_LDG(reinterpret_cast<const float4*>(f) + m * nch + c)
                                           : float4{0.f, 0.f, 0.f, 0.f};
                    }
                    TEB_UNROLL for (int j = 0; j < 2; ++j) {
                        const int m = m0 + j * t.nt;
                        if (m < n_dst) {
                            float2 z[4], y[4];
                            load_group_and_mirror(S, t.a, (m << logk) + i, z, y);
                            A[j] = cfma_r(z[0], g[j].x, A[j]); Bc[j] = cfma_r(y[0], g[j].x, Bc[j]);
                            A[j] = cfma_r(z[1], g[j].y, A[j]); Bc[j] = cfma_r(y[1], g[j].y, Bc[j]);
                            A[j] = cfma_r(z[2], g[j].z, A[j]); Bc[j] = cfma_r(y[2], g[j].z, Bc[j]);
                            A[j] = cfma_r(z[3], g[j].w, A[j]); Bc[j] = cfma_r(y[3], g[j].w, Bc[j]);
                        }
                    }
                }
            }
            TEB_UNROLL for (int j = 0; j < 2; ++j) {
                const int m = m0 + j * t.nt;
                if (m < n_dst) mulfold2_store(S, swz(t.d + m), swz(t.g + m), A[j], Bc[j], scale);
            }
        }
    } else {
        const int n_items = 1 << (t.b - 2);                    // 4 source slots per item
        for (int it0 = lt; it0 < n_items; it0 += kMf2Items * t.nt) {
            float4 gg[kMf2Items];
            TEB_UNROLL for (int j = 0; j < kMf2Items; ++j) {   // the filter loads of a trip are in flight together
                const int it = it0 + j * t.nt;
                gg[j] = (it < n_items) ? TEB_LDG(reinterpret_cast<const float4*>(f + 4 * it)) : float4{0.f, 0.f, 0.f, 0.f};
            }
            TEB_UNROLL for (int j = 0; j < kMf2Items; ++j) {
                const int it = it0 + j * t.nt;
                if (it >= n_items) continue;
                const float4 g = gg[j];
                float2 z[4], y[4];
                load_group_and_mirror(S, t.a, 4 * it, z, y);
                if (logk == 0) {
                    const int oa = swz(t.d + 4 * it), ob = swz(t.g + 4 * it);
                    mulfold2_store(S, oa, ob, cmul_r(z[0], g.x), cmul_r(y[0], g.x), scale);
                    mulfold2_store(S, oa + 1, ob + 1, cmul_r(z[1], g.y), cmul_r(y[1], g.y), scale);
                    mulfold2_store(S, oa + 2, ob + 2, cmul_r(z[2], g.z), cmul_r(y[2], g.z), scale);
                    mulfold2_store(S, oa + 3, ob + 3, cmul_r(z[3], g.w), cmul_r(y[3], g.w), scale);
                } else {
                    const int oa = swz(t.d + 2 * it), ob = swz(t.g + 2 * it);
                    mulfold2_store(S, oa, ob, cfma_r(z[0], g.x, cmul_r(z[1], g.y)), cfma_r(y[0], g.x, cmul_r(y[1], g.y)), scale);
                    mulfold2_store(S, oa + 1, ob + 1, cfma_r(z[2], g.z, cmul_r(z[3], g.w)),
                                   cfma_r(y[2], g.z, cmul_r(y[3], g.w)), scale);
                }
            }
        }
    }
}

// reflect padding of torch_backend.py:50-78 (F.pad(..., mode='reflect'); pad < N); for the phase module's
// stage A also F.pad(..., 'constant', 0) and F.pad(..., 'circular') (kymatio_phase_scattering.py:162-173)
TEB_D void load_task(float2* S, const SignalCtx& c, const Task& t, int lt) {
    const int Np = 1 << c.log2_Np;
    const int N = c.N, pad_left = c.pad_left, border = c.border;
    const float* x = c.x;
    const float* win = c.win;
    for (int i0 = lt; i0 < Np; i0 += 8 * t.nt) {
        float v[8];
        TEB_UNROLL for (int j = 0; j < 8; ++j) {
            const int i = i0 + j * t.nt;
            int r = i - pad_left;
            bool inside = i < Np;
            if (border == BORDER_REFLECT) {
                if (r < 0) r = -r;
                if (r >= N) r = 2 * (N - 1) - r;
            } else if (border == BORDER_CIRCULAR) {
                if (r < 0) r += N;
                if (r >= N) r -= N;
            } else {
                inside = inside && r >= 0 && r < N;
            }
            v[j] = inside ? TEB_LDG(x + r) : 0.f;
            if (win && inside) v[j] *= TEB_LDG(win + r);
        }
        TEB_UNROLL for (int j = 0; j < 8; ++j) {
            const int i = i0 + j * t.nt;
            if (i < Np) S[swz(t.a + i)] = make_float2(v[j], 0.f);
        }
    }
}

// Phase stage B: the phase-accelerated product of one (sample, pair) row, computed where it is consumed.
//   theta * p in fp32 like the reference (:215), an exact-enough two-constant reduction to [-pi, pi],
//   c = |z_i| (cos + i sin)(p theta_i) * conj(z_j)                         (:216-218, :283 / :339)
// and the reflect padding of _pad_signal (:162-209; pad < N): every sample is written to its own slot and
// to the slots of its mirror images, so the transcendental work is done once per sample.
TEB_D float2 accelerated_product(float2 pz, float2 zj, float power) {
    const float ph = pz.y * power;
    const float k = rintf(ph * 0.15915494309189535f);
    float r = fmaf(-k, 6.2831854820251465f, ph);
    r = fmaf(-k, -1.7484555314695172e-7f, r);
    float sn, cs;
#ifdef TEBSCAT_HOST_EMU
    sn = sinf(r); cs = cosf(r);
#else
    __sincosf(r, &sn, &cs);
#endif
    return cmulc(cmul_r(make_float2(cs, sn), pz.x), zj);     // |z| (cos + i sin) * conj(zj): three packed instructions
}

TEB_D void loadpair_task(float2* S, const SignalCtx& c, const Task& t, int lt) {
    // (the context lives in shared memory like S: read its fields once, the stores below would force reloads)
    const int N = c.N, pad_left = c.pad_left, border = c.border;
    const int Np = 1 << c.log2_Np;
    const int pad_right = Np - N - pad_left;
    const float2* zp = c.pr_zp[t.b];
    const float2* zc = c.pr_zc[t.b];
    const float pw = c.pr_pw[t.b];
    for (int t0 = lt; t0 < N; t0 += 4 * t.nt) {
        float2 a[4], b[4];
        TEB_UNROLL for (int j = 0; j < 4; ++j) {
            const int tt = t0 + j * t.nt;
            a[j] = tt < N ? TEB_LDG(zp + tt) : make_float2(0.f, 0.f);
            b[j] = tt < N ? TEB_LDG(zc + tt) : make_float2(0.f, 0.f);
        }
        TEB_UNROLL for (int j = 0; j < 4; ++j) {
            const int tt = t0 + j * t.nt;
            if (tt >= N) continue;
            const float2 v = accelerated_product(a[j], b[j], pw);
            S[swz(t.a + pad_left + tt)] = v;
            if (border == BORDER_REFLECT) {
                if (tt >= 1 && tt <= pad_left) S[swz(t.a + pad_left - tt)] = v;                       // left mirror
                if (tt <= N - 2 && tt >= N - 1 - pad_right) S[swz(t.a + pad_left + 2 * (N - 1) - tt)] = v;   // right mirror
            } else if (border == BORDER_CIRCULAR) {                         // padded[i] = c[(i - pad_left) mod N], pad <= N
                if (tt >= N - pad_left) S[swz(t.a + pad_left + tt - N)] = v;
                if (tt < pad_right) S[swz(t.a + pad_left + tt + N)] = v;
            }
        }
    }
    if (border == BORDER_CONSTANT) {                                        // zeros on both sides
        const float2 z = make_float2(0.f, 0.f);
        for (int i = lt; i < pad_left; i += t.nt) S[swz(t.a + i)] = z;
        for (int i = lt; i < pad_right; i += t.nt) S[swz(t.a + pad_left + N + i)] = z;
    }
}

// unpad (torch_backend.py:80-102) + concatenate (kymatio/backend/torch_backend.py:143-145)
// for a pool of `b` finished low-pass outputs: slot s goes to channels chan[2 (e + s)] (real part)
// and chan[2 (e + s) + 1] (imaginary part of a packed pair; -1 = none).
// normalize_tensor_data (hdf5_dataset/hdf5_dataset.py:96-135) for one coefficient of channel ch
TEB_D float epilogue_value(const SignalCtx& c, int ch, float v) {
    const int mode = c.ep_mode[ch];
    if (mode == EP_LOG) v = logf(fmaxf(v, 0.0f) + c.ep_log_eps);          // :107  log(clamp(x, min=0) + eps)
    else if (mode == EP_ASINH) v = asinhf(v);                                // :118
    return (v - c.ep_mean[ch]) / (c.ep_std[ch] + 1e-8f);                     // :133-135
}
TEB_D void store_coefficient(const SignalCtx& c, int ch, int n, float v) {
    if (ch >= c.ch_limit) return;
    if (!c.ep_mean) {
        c.out[(int64_t)ch * c.n_out + n] = v;
        return;
    }
    const int n_keep = c.n_out - 2 * c.ep_trim, m = n - c.ep_trim;           // :733-741
    if (m < 0 || m >= n_keep) return;
    v = epilogue_value(c, ch, v);
    if (c.ep_time_major) c.out[(int64_t)m * c.n_paths + ch] = v;             // :758-759
    else c.out[(int64_t)ch * n_keep + m] = v;
}

TEB_D void storeb_task(const float2* S, const SignalCtx& c, const Task& t, int lt) {
    const int total = t.b * t.d;
    for (int i = lt; i < total; i += t.nt) {
        const int slot = i / t.d, n = i - slot * t.d;
        // two channels per pool slot: the real part and, for a packed pair, the imaginary part
        const int ch_re = TEB_LDG(c.chan + 2 * (t.e + slot)), ch_im = TEB_LDG(c.chan + 2 * (t.e + slot) + 1);
        const float2 y = S[swz(t.a + (slot << t.f) + t.c + n)];
        store_coefficient(c, ch_re, n, y.x);
        if (ch_im >= 0) store_coefficient(c, ch_im, n, y.y);
    }
}

// Large-support level: transforms of more than 8192 samples live in global memory; their
// 8192-sample blocks -- and every shorter transform of that level -- pass through shared memory as tiles.
TEB_D void loadc_task(float2* S, const SignalCtx& c, const Task& t, int lt) {
    for (int i0 = lt; i0 < t.b; i0 += 4 * t.nt) {
        float2 v[4];
        TEB_UNROLL for (int j = 0; j < 4; ++j) {
            const int i = i0 + j * t.nt;
            v[j] = (i < t.b && i < c.g_valid) ? c.gbuf[i] : make_float2(0.f, 0.f);
        }
        TEB_UNROLL for (int j = 0; j < 4; ++j) {
            const int i = i0 + j * t.nt;
            if (i < t.b) S[swz(t.a + i)] = v[j];
        }
    }
}
TEB_D void storec_task(const float2* S, const SignalCtx& c, const Task& t, int lt) {
    for (int i = lt; i < t.b && i < c.g_valid; i += t.nt) c.gbuf[i] = S[swz(t.a + i)];
}

// average=False (core/scattering1d.py:329-330, :366-367): the unpadded modulus U1 / U2 at its own rate
TEB_D void storeu_task(const float2* S, const SignalCtx& c, const Task& t, int lt) {
    for (int n = lt; n < t.d; n += t.nt) c.out[(int64_t)t.e + n] = S[swz(t.a + t.c + n)].x;
}

// Phase stage A (hdf5_dataset/kymatio_phase_scattering.py:220-231): the unpadded analytic
// signal z = ifft(fft(pad(x)) psi1_f)[pad_left : pad_left + N], stored for the pair stage as
// (re, im) and/or as (|z|, atan2(im, re)) -- the polar form of _accelerate_phase (:214-215).
TEB_D void storez_task(const float2* S, const SignalCtx& c, const Task& t, int lt) {
    const int64_t row = (int64_t)t.b * t.d;
    for (int i = lt; i < t.d; i += t.nt) {
        const float2 z = S[swz(t.a + t.c + i)];
        if (c.z_mode & Z_CART) c.zc[row + i] = z;
        if (c.z_mode & Z_POLAR) c.zp[row + i] = make_float2(sqrtf(fmaf(z.x, z.x, z.y * z.y)), atan2f(z.y, z.x));
    }
}

// `pass` selects the pass of a multi-pass FFT task; the kernel runs them back to back (fft_task), the host
// emulator -- which executes the lanes of a task one after another -- runs pass by pass.
// GSRC: the kernel variant of the large-support level's fused subtrees (OP_GMULFOLD); the other variants -- the
// headline kernel among them -- do not contain the op.
template <bool GSRC = false>
TEB_D void exec_task(float2* S, const float2* twA, const float2* twB, const float* __restrict__ arena,
                     const SignalCtx& c, const Task& t, int lt, int pass = -1) {
    switch (t.op & 0xff) {
        case OP_LOAD: load_task(S, c, t, lt); break;
        case OP_FFT: {
#ifndef TEBSCAT_HOST_EMU
            fft_task(S, twA, twB, t, lt);
#else
            unsigned long long more = fft_more_passes(t);
            int logB = t.c, logR = t.d, flags = t.e;
            for (int k = 0; k < pass; ++k) {
                logB = (int)(more & 15); logR = (int)((more >> 4) & 7); flags = (int)((more >> 7) & 15);
                more >>= 12;
            }
            fft_task_pass(S, twA, twB, t, lt, logB, logR, flags);
#endif
            break;
        }
        case OP_MULFOLD: mulfold_task(S, arena, t, lt); break;
        case OP_MULFOLD2: mulfold2_task(S, arena, t, lt); break;
        case OP_GMULFOLD: if (GSRC) gmulfold_task(S, arena, c, t, lt); break;
        case OP_GMULFOLD2: if (GSRC) gmulfold2_task(S, arena, c, t, lt); break;
        case OP_LOADPAIR: loadpair_task(S, c, t, lt); break;
        case OP_STOREU: storeu_task(S, c, t, lt); break;
        case OP_LOADC: loadc_task(S, c, t, lt); break;
        case OP_STOREC: storec_task(S, c, t, lt); break;
        case OP_STOREB: storeb_task(S, c, t, lt); break;
        case OP_STOREZ: storez_task(S, c, t, lt); break;
        case OP_TINY: tiny_task(S, t, lt); break;
        default: break;
    }
}

}  // namespace tebscat
