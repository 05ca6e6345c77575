"""``KymatioPhaseScattering1D`` -- the reference's phase-scattering module on the fused CUDA path.

Drop-in for ``hdf5_dataset/kymatio_phase_scattering.py`` (constructor :60-98, ``forward``
:394-473, ``meta`` :497-499, coefficient-selection helpers :501-760, buffers
``psi1_filters, phi_filter, center_freqs, i_idx, j_idx, powers, autoc_idx`` :124-160),
as called from ``hdf5_dataset/create_hdf5_dataset.py:360-365, 418-441``.

Scattering coefficients come from :class:`tebscat.Scattering1D`; the phase-harmonic
correlations from ``tebscat_phase_forward`` (include/tebscat.h): stage A (analytic signals)
on the step interpreter, stage B (phase-accelerated products contracted with the low-pass /
truncation operator) as one fused kernel.  Host code here only builds the plan.
"""
import ctypes
import math
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from . import filterbank as fbk
from . import schedule as sch
from .torch_frontend import Scattering1D, _DevicePlan

COL_TILE = 80            # kPC of the pair kernel


BORDER_MODES = {'reflect': 0, 'constant': 1, 'circular': 2}        # _pad_signal (:162-173)


def smoothing_operator(phi0_f32, N, J_pad, pad_left, dec, border_mode='reflect'):
    """The linear map of ``_apply_phi_filter`` (:233-273) with decimation, as a dense matrix.

    c (N complex) -> pad (Np; reflect, zeros or circular) -> FFT -> * phi -> bins [0, Np//dec) -> iFFT of that
    length -> samples [pad_left//dec, pad_left//dec + N//dec).  Returns G (N, n_out) complex128
    with out = c @ G."""
    Np = 1 << J_pad
    M = max(Np // dec, 1)                                                    # :243-246
    start = pad_left // dec                                                  # :258
    stop = min(start + N // dec, M)                                          # :259-266
    n_abs = np.arange(start, stop)
    k = np.arange(M)
    phi = phi0_f32.astype(np.float64)[:M]
    E1 = (phi / M)[:, None] * np.exp(2j * np.pi * np.outer(k, n_abs) / M)    # (M, n_out)
    G = np.zeros((N, n_abs.size), np.complex128)
    tp = np.arange(Np)
    src = tp - pad_left
    live = np.ones(Np, bool)
    if border_mode == 'reflect':                                             # single fold: pad < N
        src = np.where(src < 0, -src, src)
        src = np.where(src >= N, 2 * (N - 1) - src, src)
    elif border_mode == 'circular':                                          # pad <= N
        src = np.mod(src, N)
    else:                                                                    # 'constant': padded samples are zero
        live = (src >= 0) & (src < N)
    for lo in range(0, Np, 1024):                                            # bounded temporaries
        hi = min(lo + 1024, Np)
        sel = live[lo:hi]
        E2 = np.exp(-2j * np.pi * np.outer(tp[lo:hi][sel], k) / Np)          # (chunk, M)
        np.add.at(G, src[lo:hi][sel], E2 @ E1)
    return G


def tukey_window(n, alpha):
    """Tukey (tapered cosine) window with the reference's conventions (kymatio_phase_scattering.py:362-392): alpha
    outside (0, 1] -> rectangular; alpha >= 1 -> symmetric Hann of n points; otherwise each edge of
    L = int(alpha (n - 1) / 2) samples is one half of a symmetric Hann window of 2 L points.  float64."""
    w = np.ones(int(n), np.float64)
    if alpha is None or not 0 < alpha <= 1:
        return w

    def hann(m):                                        # 0.5 - 0.5 cos(2 pi k / (m - 1)), k < m
        return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(m) / max(m - 1, 1))

    if alpha >= 1.0:
        return hann(int(n))
    edge = int(alpha * (n - 1) / 2.0)
    if edge > 0:
        h = hann(2 * edge)
        w[:edge], w[n - edge:] = h[:edge], h[edge:]
    return w


PAIR_FFT_MIN_OUT = 160      # outputs per row from which the transform form of stage B beats the dense operator


class PhasePlan:
    """Host description of the phase path of one (J, Q, T, N) configuration."""

    def __init__(self, J, Q, T, N, n_out_scattering, border_mode='reflect'):
        Q1 = fbk._as_Q1(Q)
        self.J, self.Q, self.T, self.N = J, Q1, T, N
        self.border_mode, self.border = border_mode, BORDER_MODES[border_mode]
        self.geo = fbk.build_geometry(N, J, Q1, T, clamp_to_signal=True)     # :100-113
        # padded lengths above 2^13: stage A on the ops of the large-support level, stage B in its dense form
        self.large = self.geo.J_pad > sch.LOG2_NP_MAX
        if self.geo.J_pad > 17:
            raise NotImplementedError('padded length 2**%d exceeds the large-support level (max 2**17)' % self.geo.J_pad)
        bank = fbk.build_filter_bank(self.geo.J_pad, J, Q1, T)               # :117-120
        self.bank = bank
        self.center_freqs = np.array([p.xi for p in bank.psi1], dtype=np.float32)       # :128
        F = len(bank.psi1)
        pairs = [(i, j) for i in range(F) for j in range(F)
                 if self.center_freqs[j] >= self.center_freqs[i]]                        # :141-146
        self.i_idx = np.array([p[0] for p in pairs], dtype=np.int64)
        self.j_idx = np.array([p[1] for p in pairs], dtype=np.int64)
        xi_i, xi_j = self.center_freqs[self.i_idx], self.center_freqs[self.j_idx]
        self.powers = np.where(xi_i > np.float32(1e-8), xi_j / xi_i, np.float32(1.0)).astype(np.float32)
        self.autoc_idx = np.array([k for k, (i, j) in enumerate(pairs) if i == j], dtype=np.int64)
        # decimation (:285-291 / :342-348)
        if n_out_scattering > 0 and N > n_out_scattering:
            self.dec = max(1, min(N, N // n_out_scattering))
        else:
            self.dec = 1
        phi0 = bank.phi.levels[0].astype(np.float32)
        if self.dec > 1:
            Np, start = 1 << self.geo.J_pad, self.geo.pad_left // self.dec
            if min(start + N // self.dec, max(Np // self.dec, 1)) - start <= 0:
                self.dec = 1                          # zero-length decimated output: the reference falls back (:296-299)
        self.pair_plan = None
        if self.large:
            if self.dec == 1:
                raise NotImplementedError('phase path without decimation at a padded length of 2**%d (the transform '
                                          'form of stage B needs one spectrum per SM: <= 2**13)' % self.geo.J_pad)
            arena = sch._Arena()
            self.psi1_off = [arena.add(p.levels[0]) for p in bank.psi1]
            self.arena = arena.finish()
            self.tile_lengths = [sch.LOG2_NP_MAX]             # the long transforms run as 8192-sample tiles + one global pass
            self.stage_a = None
        else:
            self._build_stage_a()
        if self.dec == 1:
            # no decimation (target length >= N, e.g. T = 1 or oversampling >= log2 T): the full-length low-pass
            # ifft(fft(pad(c)) phi)[pad_left : pad_left + N] (:268-273).  The dense operator would be N x N, so
            # stage B exists in the transform form only: one (sample, pair) row per job.
            self.n_out, self.n_cols_pad, self.G = N, 0, None
            self._build_pair_plan(phi0, rows_per_job=1)
            return
        G = smoothing_operator(phi0, N, self.geo.J_pad, self.geo.pad_left, self.dec, border_mode)
        self.n_out = G.shape[1]
        self.n_cols_pad = -(-self.n_out // COL_TILE) * COL_TILE
        Gp = np.zeros((N, self.n_cols_pad, 2), np.float32)
        Gp[:, :self.n_out, 0] = G.real
        Gp[:, :self.n_out, 1] = G.imag
        self.G = Gp
        if not self.large and self.dec & (self.dec - 1) == 0 and (1 << self.geo.J_pad) // self.dec >= 16:
            self._build_pair_plan(phi0)

    def _build_pair_plan(self, phi0, rows_per_job=None):
        """Stage B as transforms (power-of-two decimation): per (sample, pair) row the literal cascade of
        _apply_phi_filter (:233-273) on the step interpreter --
            LOADPAIR (product + reflect pad) -> FFT(Np) -> phi on bins [0, Np/dec) -> iFFT(Np/dec) -> unpad.
        Keeping bins [0, M) and multiplying by phi IS a 'filter multiply + periodise by dec' with the filter
        phi * 1[k < M] (the periodisation sums bins m + i M, of which only i = 0 survives), times dec to undo
        the periodisation's mean: the transform's leaf machinery does the rest.  Several rows share a job:
        as many as fit run side by side, and while the short tail of one wave (reduced iFFT, store) runs, the
        next wave's LOADPAIR already uses the buffers the first has released -- the list scheduler overlaps
        them."""
        if rows_per_job is None:
            rows_per_job = int(os.environ.get('TEBSCAT_PAIR_ROWS', '8'))
        n = self.geo.J_pad
        Np, dec = 1 << n, self.dec
        M = Np // dec
        lf = int(math.log2(M))
        start = self.geo.pad_left // dec
        arena = sch._Arena()
        f = np.zeros(Np, np.float64)
        f[:M] = phi0.astype(np.float64)[:M] * dec                       # :239-250 (float32 phi, exact power-of-two scale)
        off = arena.add(f)
        chains = []
        for q in range(rows_per_job):
            xq = sch.Buf(Np, 'C[%d]' % q)
            st = [[sch.TaskSpec(sch.OP_LOADPAIR, -(-self.N // 4), 1500.0, 120.0, a=(xq, 0), b=q, trip=900.0)]]
            st += sch._merge_local_passes(sch._fft_stages((xq, 0), n, 1, 'fwd'))
            c = sch.Chain(xq.name, st, owns=[xq], depth=1)
            chains.append(c)
            leaf = sch._mulfold(arena, (xq, 0), n, n - lf, sch.LEAF, off, (q, -1))
            chains.append(sch.Chain('Y[%d]' % q, [[leaf]], after=[c], reads=[xq], depth=2))
        steps, high, chan, stats = sch.schedule_chains(chains, sch.smem_capacity(), lf, start, self.n_out,
                                                       pool_slots=max(rows_per_job << lf, 256))
        tasks, ranges = sch.emit(steps)
        logical = sch._round16(high)

        class _P:
            pass
        a = _P()
        a.N, a.geo, a.n_paths, a.n_out = self.N, self.geo, rows_per_job, self.n_out
        a.border = self.border
        a.n_threads, a.smem_complex = sch.N_THREADS, logical + logical // 16
        a.tasks, a.steps, a.arena = tasks, ranges, arena.finish()
        a.chan = np.asarray(chan, np.int32)
        a.stats = stats
        self.pair_plan = a

    def _build_stage_a(self):
        """Schedule of stage A: root transform, then per filter psi multiply -> iFFT -> STOREZ."""
        n = self.geo.J_pad
        arena = sch._Arena()
        u0 = sch.Buf(1 << n, 'U0')
        root = sch.Chain('root', [[sch.TaskSpec(sch.OP_LOAD, 1 << n, 300.0, 16.0, a=(u0, 0))]] +
                         sch._fft_stages((u0, 0), n, 1, 'fwd'), owns=[u0], depth=0)
        chains = [root]
        for f, p in enumerate(self.bank.psi1):
            off = arena.add(p.levels[0])
            x = sch.Buf(1 << n, 'Z[%d]' % f)
            mf = [sch._mulfold(arena, (u0, 0), n, 0, (x, 0), off)]
            st = [mf] + sch._fuse_first_inverse_pass(mf, sch._fft_stages((x, 0), n, 1, 'inv'), n)
            st.append([sch.TaskSpec(sch.OP_STOREZ, self.N, 200.0, 40.0, a=(x, 0), b=f, c=self.geo.pad_left, d=self.N)])
            chains.append(sch.Chain(x.name, st, after=[root], reads=[u0], owns=[x], frees_own_at_end=True, depth=1))
        steps, high, chan, stats = sch.schedule_chains(chains, sch.smem_capacity(), n, 0, self.N)
        tasks, ranges = sch.emit(steps)
        logical = sch._round16(high)

        class _A:
            pass
        a = _A()
        a.N, a.geo, a.n_paths, a.n_out = self.N, self.geo, len(self.bank.psi1), self.N
        a.border = self.border
        a.n_threads, a.smem_complex = sch.N_THREADS, logical + logical // 16
        a.tasks, a.steps, a.arena = tasks, ranges, arena.finish()
        a.chan = np.zeros(1, np.int32)
        a.stats = stats
        self.stage_a = a


class _DevicePhasePlan:
    def __init__(self, plan: PhasePlan, device_index: int):
        lib = _lib.load()
        self.plan = plan
        self.large = None
        stage_a = _DevicePlan(plan.stage_a, device_index) if not plan.large else None
        d = _lib.PhaseDesc()
        d.abi_version = _lib.ABI_VERSION
        d.N, d.n_filters, d.n_pairs = plan.N, len(plan.bank.psi1), len(plan.i_idx)
        d.n_out, d.n_cols_pad = plan.n_out, plan.n_cols_pad
        G = np.ascontiguousarray(plan.G, np.float32) if plan.G is not None else None
        ii = np.ascontiguousarray(plan.i_idx, np.int32)
        jj = np.ascontiguousarray(plan.j_idx, np.int32)
        pw = np.ascontiguousarray(plan.powers, np.float32)
        handle = ctypes.c_void_p()
        i32p, fp = ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_float)
        if plan.large:
            rc = lib.tebscat_phase_plan_create_pairs_only(ctypes.byref(d), int(device_index), G.ctypes.data_as(fp),
                                                          ii.ctypes.data_as(i32p), jj.ctypes.data_as(i32p),
                                                          pw.ctypes.data_as(fp), ctypes.byref(handle))
            _lib.check(rc)
            from .large import LargeDevicePlan
            self.large = LargeDevicePlan(plan, device_index)   # context + tile plans + the psi1 arena on the device
        else:
            rc = lib.tebscat_phase_plan_create(ctypes.byref(d), stage_a.handle, G.ctypes.data_as(fp) if G is not None else None,
                                               ii.ctypes.data_as(i32p), jj.ctypes.data_as(i32p),
                                               pw.ctypes.data_as(fp), ctypes.byref(handle))
            _lib.check(rc)
            stage_a.handle = None                       # ownership moved into the phase plan
        self.handle = handle
        self._lib = lib
        self._ws = None
        # stage B as transforms where that is the cheaper form (long outputs); TEBSCAT_PHASE_FFT=0/1 overrides
        mode = os.environ.get('TEBSCAT_PHASE_FFT', 'auto')
        use_fft = plan.pair_plan is not None and (mode == '1' or (mode == 'auto' and plan.n_out >= PAIR_FFT_MIN_OUT)
                                                  or plan.G is None)
        self.uses_fft_pairs = bool(use_fft)
        if use_fft:
            pp = _DevicePlan(plan.pair_plan, device_index)
            _lib.check(lib.tebscat_phase_plan_attach_pair_plan(handle, pp.handle))
            pp.handle = None

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                self._lib.tebscat_phase_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- padded lengths above 2^13 ---------------------------------------------------------------------------------
    LARGE_CHUNK_BYTES = 1 << 30                  # workspace of analytic signals per chunk (two arrays of this size)

    def forward_large(self, x3, ch_i, ch_j, sub_ptr, n_sel, low_pass, out):
        """Stage A on the large-support level -- pad + load, FFT, and per filter psi multiply -> iFFT -> crop
        (kymatio_phase_scattering.py:220-231) as one launch each over a chunk of samples -- into workspaces of
        (|z|, theta) / (re, im), then stage B (tebscat_phase_pairs) on them."""
        lib, p, g = self._lib, self.plan, self.large.handle
        B, C, N = x3.shape
        dev = x3.device
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        F, n, Np = len(p.bank.psi1), p.geo.J_pad, 1 << p.geo.J_pad
        chunk = int(max(1, min(B, self.LARGE_CHUNK_BYTES // (F * N * 8), ((1 << 31) - 1) // (F * N))))
        if self._ws is None or self._ws[0] < chunk or self._ws[1].device != dev:
            self._ws = (chunk, torch.empty(chunk * F * N * 2, dtype=torch.float32, device=dev),      # zc
                        torch.empty(chunk * F * N * 2, dtype=torch.float32, device=dev),              # zp
                        torch.empty(chunk * Np * 2, dtype=torch.float32, device=dev),                 # spectrum of the channel
                        torch.empty(chunk * Np * 2, dtype=torch.float32, device=dev),                 # one filtered signal
                        torch.empty(chunk * N, dtype=torch.float32, device=dev))                      # the channel, contiguous
        _, zc, zp, U0, W, xc = self._ws
        vp = ctypes.c_void_p
        fa = self.large.arena.data_ptr()
        width = p.n_out if low_pass else N

        def stage_a(xb, ch, mode):
            nb = xb.shape[0]
            xcv = xc[:nb * N].view(nb, N)
            xcv.copy_(xb[:, ch, :])
            _lib.check(lib.tebscat_large_pad_load_mode(g, vp(xcv.data_ptr()), nb, N, p.geo.pad_left, n, p.border,
                                                       vp(U0.data_ptr()), st))
            _lib.check(lib.tebscat_large_fft(g, vp(U0.data_ptr()), nb, n, 0, st))
            for f, off in enumerate(p.psi1_off):
                _lib.check(lib.tebscat_large_mulfold(g, vp(U0.data_ptr()), vp(fa + 4 * off), vp(W.data_ptr()), nb, n, 0, 0, 0, n, st))
                _lib.check(lib.tebscat_large_fft(g, vp(W.data_ptr()), nb, n, 1, st))
                _lib.check(lib.tebscat_large_storez(g, vp(W.data_ptr()), nb, n, p.geo.pad_left, N, F, f, mode,
                                                    vp(zc.data_ptr()), vp(zp.data_ptr()), st))

        for b0 in range(0, B, chunk):
            xb = x3[b0:b0 + chunk]
            nb = xb.shape[0]
            if ch_i == ch_j:
                stage_a(xb, ch_i, 3)
            else:
                stage_a(xb, ch_i, 2)
                stage_a(xb, ch_j, 1)
            _lib.check(lib.tebscat_phase_pairs(self.handle, vp(zp.data_ptr()), vp(zc.data_ptr()), nb, sub_ptr, n_sel,
                                               1 if low_pass else 0, vp(out[b0:b0 + nb].data_ptr()), st))
        return out


class KymatioPhaseScattering1D(nn.Module):
    def __init__(self, J, Q, T, shape, device=None, oversampling=0, max_order=2,
                 border_mode='reflect', tukey_alpha=None):
        super().__init__()
        self.J = J
        if isinstance(Q, tuple):                       # :71-76
            self.Q_scattering = Q
            self.Q = Q[0]
        else:
            self.Q_scattering = Q
            self.Q = Q
        self.T = T
        self.oversampling = oversampling
        self.max_order = max_order
        self.border_mode = border_mode
        self.tukey_alpha = tukey_alpha
        self.device = device if device is not None else torch.device('cuda' if torch.cuda.is_available() else 'cpu')
        self.device = torch.device(self.device)
        self.eps = 1e-14
        self.N = int(shape) if isinstance(shape, (int, float)) else int(shape[0])
        if border_mode not in ('reflect', 'constant', 'circular'):
            raise ValueError(f"Unsupported border_mode: {border_mode}")

        self.scattering = Scattering1D(J=J, shape=shape, Q=self.Q_scattering, max_order=max_order, average=True,
                                       oversampling=oversampling, vectorize=True, out_type='array', T=T).to(self.device)
        k_out = max(int(math.floor(math.log2(T))) - oversampling, 0)       # core/scattering1d.py:260-261; target_length (:445)
        n_out_scat = self.scattering.ind_end[k_out] - self.scattering.ind_start[k_out]
        self._plan = PhasePlan(J, self.Q, T, self.N, n_out_scat, border_mode)
        g = self._plan.geo
        self.J_pad, self.pad_left, self.pad_right = g.J_pad, g.pad_left, g.pad_right
        self.ind_start, self.ind_end = g.ind_start, g.ind_end
        self.N_padded = 2 ** self.J_pad
        bank = self._plan.bank
        filters = np.stack([p.levels[0] for p in bank.psi1], axis=0)                        # :123-125
        self.register_buffer('psi1_filters', torch.from_numpy(filters).to(torch.complex64).to(self.device))
        self.register_buffer('phi_filter', torch.from_numpy(bank.phi.levels[0]).to(torch.complex64).to(self.device))
        self.register_buffer('center_freqs', torch.from_numpy(self._plan.center_freqs).to(self.device))
        self.register_buffer('i_idx', torch.from_numpy(self._plan.i_idx).to(self.device))
        self.register_buffer('j_idx', torch.from_numpy(self._plan.j_idx).to(self.device))
        self.register_buffer('powers', torch.from_numpy(self._plan.powers).to(self.device))
        self.register_buffer('autoc_idx', torch.from_numpy(self._plan.autoc_idx).to(self.device))
        self._dev_plans = {}
        self._windowed = set()

    # ---- plumbing -------------------------------------------------------------------------
    def _dev_plan(self, index):
        if index not in self._dev_plans:
            self._dev_plans[index] = _DevicePhasePlan(self._plan, index)
        return self._dev_plans[index]

    def _phase(self, x3, ch_i, ch_j, subset=None, low_pass=True):
        """x3: (B, C, N) float32 CUDA contiguous -> (B, n_sel, n_out) (or (B, n_sel, N))."""
        if x3.device.type != 'cuda':
            raise TypeError('Input must be on GPU.')
        if x3.dtype is not torch.float32:
            raise TypeError('Input and filter must be of the same dtype.')
        x3 = x3.contiguous()
        B, C, N = x3.shape
        if N != self.N:
            raise ValueError('Input length {} does not match shape={}'.format(N, self.N))
        index = x3.device.index if x3.device.index is not None else torch.cuda.current_device()
        plan = self._dev_plan(index)
        if subset is not None:
            sub = np.ascontiguousarray(np.asarray(subset, dtype=np.int64).astype(np.int32))
            n_sel, sub_ptr = int(sub.size), sub.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
        else:
            n_sel, sub_ptr = len(self._plan.i_idx), None
        width = self._plan.n_out if low_pass else N
        out = torch.empty((B, n_sel, width), dtype=torch.float32, device=x3.device)
        if n_sel == 0 or B == 0:
            return out
        if self._plan.large:
            return plan.forward_large(x3, int(ch_i), int(ch_j), sub_ptr, n_sel, low_pass, out)
        rc = _lib.load().tebscat_phase_forward(plan.handle, x3.data_ptr(), B, C, int(ch_i), int(ch_j), sub_ptr, n_sel,
                                               1 if low_pass else 0, out.data_ptr(),
                                               torch.cuda.current_stream(x3.device).cuda_stream)
        _lib.check(rc)
        return out

    # ---- Tukey window (:362-392, applied at :405-407) ---------------------------------------
    def _window(self, n):
        """The reference multiplies the input by a Tukey taper before anything else.  Here the taper is a table the
        kernels' loads apply (OP_LOAD / pad_load: tebscat_plan_set_window): no elementwise pass over the batch."""
        if self.tukey_alpha is None:
            return None
        return tukey_window(n, self.tukey_alpha)

    def _install_window(self, index):
        """Hand the taper to the device plans of the scattering transform and of stage A (once per device)."""
        if self.tukey_alpha is None or index in self._windowed:
            return
        w = np.ascontiguousarray(self._window(self.N), np.float32)
        self.scattering.set_window(w)
        plan = self._dev_plan(index)
        wp = w.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        if self._plan.large:
            _lib.check(_lib.load().tebscat_large_set_window(plan.large.handle, wp, int(self.N)))
        else:
            _lib.check(_lib.load().tebscat_phase_plan_set_window(plan.handle, wp))
        self._windowed.add(index)

    # ---- forward (:394-473) -------------------------------------------------------------------
    def forward(self, x, compute_phase=True, compute_cross_phase=False, cross_phase_same_pairs_only=False,
                cross_phase_low_pass=True, scattering_channel=0, phase_channels=None,
                phase_pairs=None):
        """Same contract as the reference.  ``phase_pairs`` (extension) restricts the phase
        output to the given pair indices, e.g. the masks of get_optimal_coefficients_for_fhr."""
        x = x.to(self.device)
        if self.tukey_alpha is not None and x.device.type == 'cuda':
            if x.shape[-1] != self.N:
                raise ValueError('Input length {} does not match shape={}'.format(x.shape[-1], self.N))
            self._install_window(x.device.index if x.device.index is not None else torch.cuda.current_device())
        ch = None
        if x.dim() == 3:
            B, n_channels, N = x.shape
            if scattering_channel >= n_channels:
                raise ValueError(f"scattering_channel {scattering_channel} >= {n_channels}")
            scattering_input = x[:, scattering_channel, :].contiguous()
            if compute_cross_phase:                                           # :477-484
                if phase_channels is None:
                    if n_channels < 2:
                        raise ValueError("Cross-channel correlation requires at least 2 channels")
                    phase_channels = [0, 1]
                if len(phase_channels) != 2 or any(c >= n_channels for c in phase_channels):
                    raise ValueError("Invalid phase_channels for cross-channel correlation")
                ch = (phase_channels[0], phase_channels[1])
            elif phase_channels is not None:                                  # :486-492
                if len(phase_channels) != 1:
                    raise ValueError("Single-channel phase correlation requires exactly 1 channel")
                if phase_channels[0] >= n_channels:
                    raise ValueError(f"phase_channel {phase_channels[0]} >= {n_channels}")
                ch = (phase_channels[0], phase_channels[0])
            else:
                ch = (scattering_channel, scattering_channel)
            x3 = x
        elif x.dim() == 2:
            if scattering_channel != 0:
                raise ValueError("scattering_channel must be 0 for single-channel input")
            if compute_cross_phase:
                raise ValueError("Cross-channel correlation requires multi-channel input")
            scattering_input = x
            x3 = x.unsqueeze(1)
            ch = (0, 0) if compute_phase else None
        else:
            raise ValueError(f"Input must be 2D or 3D, got shape {x.shape}")

        scattering_coeffs, _ = self.scattering(scattering_input)
        if scattering_coeffs.shape[-1] == 0:
            raise ValueError(f"Scattering output has zero temporal dimension: {scattering_coeffs.shape}")
        results = {'scattering': scattering_coeffs}
        if (compute_phase or compute_cross_phase) and ch is not None:
            subset = None if phase_pairs is None else np.asarray(
                phase_pairs.cpu() if torch.is_tensor(phase_pairs) else phase_pairs).reshape(-1)
            if subset is not None and subset.dtype == np.bool_:
                subset = np.nonzero(subset)[0]
            if compute_cross_phase:
                if cross_phase_same_pairs_only:                               # :325-328
                    base = self._plan.autoc_idx
                    subset = base if subset is None else np.intersect1d(base, subset)
                results['cross_phase_corr'] = self._phase(x3, ch[0], ch[1], subset, low_pass=cross_phase_low_pass)
            elif compute_phase:
                results['phase_corr'] = self._phase(x3, ch[0], ch[0], subset, low_pass=True)
            results['autoc_idx'] = self.autoc_idx
        return results

    def forward_dataset(self, x, phase_pairs, cross_pairs, scattering_channel=0, phase_channels=(0, 1)):
        """Single-pass dataset entry: what ``create_hdf5_dataset.py:418-441`` gets from
        two ``st_model(...)`` calls plus masking, in one call --

            {'scattering': S(x[:, scattering_channel]),
             'phase_corr': within-channel correlations of phase_channels[0] for `phase_pairs`,
             'cross_phase_corr': phase_channels[0] x phase_channels[1] for `cross_pairs`,
             'autoc_idx': ...}

        with the scattering transform and the analytic signals of phase_channels[0] computed once
        and only the selected pairs contracted.  `phase_pairs` / `cross_pairs`: boolean masks over
        the P pairs (e.g. get_optimal_coefficients_for_fhr()['recommendations']) or index arrays."""
        x = x.to(self.device)
        if self.tukey_alpha is not None and x.device.type == 'cuda' and x.shape[-1] == self.N:
            self._install_window(x.device.index if x.device.index is not None else torch.cuda.current_device())
        if x.dim() != 3 or x.shape[1] < 2:
            raise ValueError("Cross-channel correlation requires at least 2 channels")
        B, C, N = x.shape
        if N != self.N:
            raise ValueError('Input length {} does not match shape={}'.format(N, self.N))
        ch_i, ch_j = int(phase_channels[0]), int(phase_channels[1])
        if len(phase_channels) != 2 or max(ch_i, ch_j) >= C or ch_i == ch_j:
            raise ValueError("Invalid phase_channels for cross-channel correlation")
        if scattering_channel >= C:
            raise ValueError(f"scattering_channel {scattering_channel} >= {C}")

        def as_index(sel):
            a = np.asarray(sel.cpu() if torch.is_tensor(sel) else sel).reshape(-1)
            return (np.nonzero(a)[0] if a.dtype == np.bool_ else a).astype(np.int32)

        sub_w, sub_c = np.ascontiguousarray(as_index(phase_pairs)), np.ascontiguousarray(as_index(cross_pairs))
        if sub_w.size == 0 or sub_c.size == 0:
            raise ValueError('forward_dataset needs at least one within-channel and one cross-channel pair')
        if x.dtype is not torch.float32:
            raise TypeError('Input and filter must be of the same dtype.')
        x = x.contiguous()
        S, _ = self.scattering(x[:, scattering_channel, :].contiguous())
        index = x.device.index if x.device.index is not None else torch.cuda.current_device()
        plan = self._dev_plan(index)
        n_out = self._plan.n_out
        if self._plan.large:                          # above 2^13 the two correlations are two passes of the same driver
            return {'scattering': S, 'phase_corr': self._phase(x, ch_i, ch_i, sub_w), 'cross_phase_corr': self._phase(x, ch_i, ch_j, sub_c),
                    'autoc_idx': self.autoc_idx}
        within = torch.empty((B, sub_w.size, n_out), dtype=torch.float32, device=x.device)
        cross = torch.empty((B, sub_c.size, n_out), dtype=torch.float32, device=x.device)
        i32p = ctypes.POINTER(ctypes.c_int32)
        rc = _lib.load().tebscat_phase_forward_dual(
            plan.handle, x.data_ptr(), B, C, ch_i, ch_j, sub_w.ctypes.data_as(i32p), int(sub_w.size),
            sub_c.ctypes.data_as(i32p), int(sub_c.size), within.data_ptr(), cross.data_ptr(),
            torch.cuda.current_stream(x.device).cuda_stream)
        _lib.check(rc)
        return {'scattering': S, 'phase_corr': within, 'cross_phase_corr': cross, 'autoc_idx': self.autoc_idx}

    def meta(self):
        return self.scattering.meta()

    # ---- coefficient selection (:501-760): index logic on center_freqs / powers -------------------
    def select_fhr_phase_coefficients(self, min_freq=0.006, max_harmonic_power=8, include_autocorr=True,
                                      harmonic_ratios=[2, 3]):
        cf, pw, ii, jj = self.center_freqs, self.powers, self.i_idx, self.j_idx
        in_band = cf >= min_freq
        both = in_band[ii] & in_band[jj]
        masks = {}
        if include_autocorr:
            auto = torch.zeros(len(ii), dtype=torch.bool, device=cf.device)
            auto[self.autoc_idx] = True
            masks['autocorr'] = both & auto
        for ratio in harmonic_ratios:
            masks[f'harmonic_{ratio}'] = both & (torch.abs(pw - ratio) < 0.1) & (pw <= max_harmonic_power)
        optimal = torch.zeros(len(ii), dtype=torch.bool, device=cf.device)
        for m in masks.values():
            optimal |= m
        some = bool(optimal.any())
        metadata = {
            'total_pairs': len(ii),
            'selected_pairs': optimal.sum().item(),
            'frequency_range': (cf.min().item(), cf.max().item()),
            'selected_freq_range': (cf[ii[optimal]].min().item() if some else 0, cf[jj[optimal]].max().item() if some else 0),
            'power_range': (pw[optimal].min().item() if some else 0, pw[optimal].max().item() if some else 0),
        }
        return {'masks': masks, 'optimal_mask': optimal, 'metadata': metadata,
                'i_idx_selected': ii[optimal], 'j_idx_selected': jj[optimal], 'powers_selected': pw[optimal],
                'freqs_i_selected': cf[ii[optimal]], 'freqs_j_selected': cf[jj[optimal]]}

    def select_fhr_up_cross_coefficients(self, up_max_freq=0.02, fhr_min_freq=0.04, fhr_max_freq=0.5,
                                         max_harmonic_power=32):
        cf, pw, ii, jj = self.center_freqs, self.powers, self.i_idx, self.j_idx
        up_band = cf < up_max_freq
        fhr_band = (cf >= fhr_min_freq) & (cf <= fhr_max_freq)
        cross = up_band[ii] & fhr_band[jj] & (pw >= 1) & (pw <= max_harmonic_power)
        some = bool(cross.any())
        metadata = {
            'total_pairs': len(ii), 'cross_selected_pairs': cross.sum().item(),
            'up_freq_range': (0.0, up_max_freq), 'fhr_freq_range': (fhr_min_freq, fhr_max_freq),
            'up_filters_available': up_band.sum().item(), 'fhr_filters_available': fhr_band.sum().item(),
            'power_range': (pw[cross].min().item() if some else 0, pw[cross].max().item() if some else 0),
        }
        return {'cross_mask': cross, 'up_band_mask': up_band, 'fhr_band_mask': fhr_band, 'metadata': metadata,
                'i_idx_selected': ii[cross], 'j_idx_selected': jj[cross], 'powers_selected': pw[cross],
                'up_freqs_selected': cf[ii[cross]], 'fhr_freqs_selected': cf[jj[cross]]}

    def get_optimal_coefficients_for_fhr(self, j_config=11, q_config=4, t_config=16):
        min_freq = 0.006 if j_config >= 11 else 0.003                          # :714-717
        phase_selection = self.select_fhr_phase_coefficients(min_freq=min_freq, max_harmonic_power=8,
                                                             include_autocorr=True, harmonic_ratios=[2, 3])
        cross_selection = self.select_fhr_up_cross_coefficients(up_max_freq=0.02, fhr_min_freq=0.04,
                                                                fhr_max_freq=0.5, max_harmonic_power=32)
        n_phase = phase_selection['optimal_mask'].sum().item()
        n_cross = cross_selection['cross_mask'].sum().item()
        analysis = {
            'current_config': {'J': j_config, 'Q': q_config, 'T': t_config},
            'total_scattering_coeffs': j_config * q_config + 1,
            'selected_phase_coeffs': n_phase,
            'selected_cross_coeffs': n_cross,
            'efficiency_gain': {
                'phase_reduction': f"{100 * (1 - n_phase / len(self.i_idx)):.1f}%",
                'focus_improvement': f"Focused on {phase_selection['metadata']['selected_pairs']} most relevant pairs",
            },
        }
        return {'phase_selection': phase_selection, 'cross_selection': cross_selection, 'config_analysis': analysis,
                'recommendations': {'use_phase_mask': phase_selection['optimal_mask'],
                                    'use_cross_mask': cross_selection['cross_mask'],
                                    'total_selected_features': analysis['total_scattering_coeffs'] + n_phase + n_cross}}

    def verify_phase_correlation_properties(self, x, tol=1e-6):
        """:762-811 -- autocorrelations non-negative, xi_j >= xi_i, powers >= 1."""
        results = {'passed': True, 'details': {}}
        x_test = x[:1] if x.dim() == 2 else x[:1, :1]
        x3 = x_test.to(self.device).reshape(1, 1, -1)
        try:
            corr = self._phase(x3, 0, 0, self._plan.autoc_idx, low_pass=True)
            worst = corr.min(dim=-1).values[0]
            for k in torch.nonzero(worst < -tol).flatten().tolist():
                results['passed'] = False
                results['details'][f'autocorr_{k}_negative'] = worst[k].item()
        except Exception as e:                                                   # mirror the reference's catch-all
            results['passed'] = False
            results['details']['phase_computation_error'] = str(e)
        cf, ii, jj = self.center_freqs, self.i_idx, self.j_idx
        bad = torch.nonzero(cf[jj] < cf[ii] - tol).flatten().tolist()
        for k in bad:
            results['passed'] = False
            results['details'][f'frequency_ordering_violation_{k}'] = (cf[ii[k]].item(), cf[jj[k]].item())
        if torch.any(self.powers < 1.0 - tol):
            results['passed'] = False
            results['details']['invalid_powers'] = self.powers[self.powers < 1.0 - tol].tolist()
        return results


__all__ = ['KymatioPhaseScattering1D', 'PhasePlan', 'smoothing_operator', 'tukey_window']
