"""Large-support level: padded lengths 2^14 .. 2^17.

At these lengths a spectrum (128 KB .. 1 MB) no longer fits one SM, so the fused single-kernel cascade of
``schedule.py`` does not apply.  This driver keeps the reference's op order
(``kymatio/scattering1d/core/scattering1d.py:269-370``) with every op ONE launch over the whole batch on global
(L2/HBM-resident) complex buffers -- pad + load, transform, filter multiply + periodisation, modulus, unpad + store
-- all hand-written CUDA behind the C ABI (``tebscat_large_*``).  Transforms run as tile jobs of the step
interpreter (``schedule.build_tile_plan``); above 8192 samples one global radix pass precedes / follows them.
Spectra are kept in bit-reversed order, so the periodisation is a sum of adjacent elements here too.

There is no CPU or torch fallback in here: torch only owns the buffers.
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import _lib
from . import filterbank as fbk
from . import schedule as sch

LOG2_MAX = 17
LEAF_FUSED_MAX_LOG2 = 10          # leaves of up to 1024 output-rate samples run as one launch (tebscat_large_leaf)
FUSED_LEAVES = os.environ.get('TEBSCAT_FUSED_LEAVES', '1') != '0'
# Above a padded length of 2^13 the subtrees that fit one SM run on the fused kernel with a global source
# (schedule.build_hybrid_plans, tebscat_scat1d_forward_gsrc); TEBSCAT_HYBRID=0 keeps every op on this level (A/B)
HYBRID = os.environ.get('TEBSCAT_HYBRID', '1') != '0'


class _TilePlanOwner:
    """Creates the device plan of a tile schedule and hands it to the large-support context."""

    def __init__(self, ctx, n, kind, device_index, tile_slots):
        from .torch_frontend import _DevicePlan
        plan = sch.build_tile_plan(n, kind == 1, kind='pair' if kind == 2 else None, tile_slots=tile_slots)
        dp = _DevicePlan(plan, device_index)
        _lib.check(_lib.load().tebscat_large_set_tile_plan(ctx, n, kind, plan.slots, dp.handle))
        dp.handle = None


class LargePlan:
    """Host description: geometry, filter arena (bit-reversed, fp32 like register_filters) and the op list."""

    def __init__(self, J, N, Q, T, max_order=2, oversampling=0):
        Q1 = fbk._as_Q1(Q)
        self.J, self.N, self.Q, self.T, self.max_order = J, N, Q1, T, max_order
        self.geo = geo = fbk.build_geometry(N, J, Q1, T)
        n = geo.J_pad
        if n > LOG2_MAX:
            raise NotImplementedError('padded length 2**%d exceeds the large-support level (max 2**%d)' % (n, LOG2_MAX))
        bank = fbk.build_filter_bank(n, J, Q1, T)
        log2_T = int(math.floor(math.log2(T)))
        os_ = int(oversampling)
        kf = max(log2_T - os_, 0)
        self.lf = lf = n - kf
        if lf < 1:
            raise NotImplementedError('output-rate length below 2 samples is not supported')
        self.i0, i1 = geo.ind_start[kf], geo.ind_end[kf]
        self.n_out = i1 - self.i0
        arena = sch._Arena()
        phi_off = [arena.add(a) for a in bank.phi.levels]
        psi1_off = [arena.add(p.levels[0]) for p in bank.psi1]
        psi2_off = [[arena.add(a) for a in p.levels] for p in bank.psi2]

        def mf(off, log_src, logk):
            """(filter offset, log_src, logk, chunk mask, log2 chunk width, scale exponent)"""
            if logk >= 2:
                mask, logcw = arena.chunk_mask(off, logk), sch._Arena.chunk_log2(logk)
            else:
                mask, logcw = 0, 0
            return (off, log_src, logk, mask & 0xffffffff, logcw, logk + (log_src - logk))   # mean over k, 1/L of the iFFT

        keys = [()] + [(i,) for i in range(len(bank.psi1))]
        if max_order == 2:
            for n1, p1 in enumerate(bank.psi1):
                for n2, p2 in enumerate(bank.psi2):
                    if p2.j > p1.j:
                        keys.append((n1, n2))
        self.keys = keys
        channel = {k: c for c, k in enumerate(keys)}
        # the cascade as a list of paths: (channel, [first-order op], [second-order ops])
        self.s0 = mf(phi_off[0], n, n - lf)                                       # core :285-292
        self.first = []
        for n1, p1 in enumerate(bank.psi1):
            k1 = max(min(p1.j - os_, log2_T - os_), 0)                            # :304
            if not p1.xi < 0.5 / (2 ** k1):
                raise AssertionError('psi1 aliasing assertion of the reference violated')
            l1 = n - k1
            entry = dict(n1=n1, ch=channel[(n1,)], l1=l1, mul=mf(psi1_off[n1], n, k1), leaf=mf(phi_off[k1], l1, l1 - lf), kids=[])
            if max_order == 2:
                for n2, p2 in enumerate(bank.psi2):
                    if p2.j > p1.j:
                        k2 = max(min(p2.j - k1 - os_, log2_T - k1 - os_), 0)      # :344-345
                        l2 = l1 - k2
                        entry['kids'].append(dict(n2=n2, ch=channel[(n1, n2)], l2=l2, mul=mf(psi2_off[n2][k1], l1, k2),
                                                  leaf=mf(phi_off[k1 + k2], l2, l2 - lf)))
            self.first.append(entry)
        self.arena = arena.finish()
        self.n_paths = len(keys)
        lens = {n, lf} | {e['l1'] for e in self.first} | {k['l2'] for e in self.first for k in e['kids']}
        self.tile_lengths = sorted(min(v, sch.LOG2_NP_MAX) for v in lens)
        # longest second-order transform: Np/2 without oversampling, but a child is not subsampled at all when
        # oversampling >= its j2 - k1 (core :344-345), and then it is as long as its parent
        self.max_l2 = max([k['l2'] for e in self.first for k in e['kids']], default=1)
        self._channel = channel
        self._os = os_
        self._hybrid = False                      # built on first use (seconds of host scheduling)

    def hybrid_plans(self):
        """Schedules of the subtrees that fit one SM (padded lengths above 2^13 only), or None."""
        if self._hybrid is False:
            self._hybrid = None
            if HYBRID and self.geo.J_pad > sch.LOG2_NP_MAX:
                self._hybrid = sch.build_hybrid_plans(self.J, self.N, self.Q, self.T, self.max_order, self._os)
        return self._hybrid


class LargeDevicePlan:
    def __init__(self, plan: LargePlan, device_index: int):
        lib = _lib.load()
        self._lib = lib
        self.plan = plan
        self.device_index = device_index
        handle = ctypes.c_void_p()
        _lib.check(lib.tebscat_large_create(int(device_index), ctypes.byref(handle)))
        self.handle = handle
        for nlen in sorted(set(plan.tile_lengths)):
            for kind in (0, 1, 2):                   # forward, inverse, inverse -> modulus -> forward
                for slots in sch.TILE_SLOTS:
                    _TilePlanOwner(handle, nlen, kind, device_index, slots)
        self.arena = torch.from_numpy(np.ascontiguousarray(plan.arena, np.float32)).to(torch.device('cuda', device_index))
        self._ws, self._bws, self._graphs = {}, {}, {}
        # fused subtrees: device plans of the schedules with a global source
        self._first, self._first_n1, self._kids = None, frozenset(), {}
        hyb = plan.hybrid_plans() if hasattr(plan, 'hybrid_plans') else None    # (bare contexts: transforms only)
        if hyb:
            from .torch_frontend import _DevicePlan
            if hyb['first'] is not None:
                self._first, self._first_n1 = _DevicePlan(hyb['first'], device_index), frozenset(hyb['first_n1'])
            ch = plan._channel
            for kp, members, done in hyb['kids']:
                dp = _DevicePlan(kp, device_index)
                for n1 in members:                 # same children, channels a constant shift apart
                    shift = ch[(n1, done[0])] - ch[(kp.head, done[0])]
                    assert all(ch[(n1, n2)] - ch[(kp.head, n2)] == shift for n2 in done)
                    self._kids[n1] = (dp, shift, frozenset(done))

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                self._lib.tebscat_large_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- workspaces and CUDA graphs ------------------------------------------------------------------------------
    # One grow-only workspace per direction, sized for the largest batch seen (a smaller batch uses a prefix): a
    # loader whose last batch is short, or a batch chunked with a remainder, does not reallocate anything.  Graphs
    # are cached per (direction, batch size) in a small LRU; growing a workspace drops the graphs that captured it.
    GRAPH_SLOTS = 6

    def _workspace(self, B, dev):
        if B > self._ws.get('cap', 0) or self._ws.get('dev') != dev.index:
            Np = 1 << self.plan.geo.J_pad
            self._ws = dict(cap=B, dev=dev.index, bufs=(
                torch.empty(B * Np * 2, dtype=torch.float32, device=dev),                       # U0
                torch.empty(B * Np * 2, dtype=torch.float32, device=dev),                       # first-order work
                torch.empty(B * (2 << self.plan.max_l2), dtype=torch.float32, device=dev),      # second-order work
                torch.empty(B * (2 << self.plan.lf), dtype=torch.float32, device=dev)))         # leaf
            self._graphs = {k: v for k, v in self._graphs.items() if k[0] != 'fwd'}
        return self._ws['bufs']

    def invalidate_graphs(self):
        """Forget every captured graph (a plan-level setting the kernels read through the context has changed)."""
        self._graphs = {}

    def _graphed(self, kind, run, args, ins, outs, direct=True):
        """Run `run(*args)` through a CUDA graph.  `ins` / `outs`: the caller-owned tensors among `args`.

        A forward is a few thousand short launches (a backward ~1200 at the headline configuration): below ~1000
        signals they, not the arithmetic, set the time, so the op list of a batch size is captured once -- on static
        buffers of that batch size -- and replayed between one copy in and one copy out (0.4 % of a forward at a
        padded length of 2^14).  Capturing on the caller's own buffers instead (keyed by their addresses, measured:
        +5 % on a loop that reuses its buffers) was dropped: a training loop's gradient buffers alternate between a
        few addresses, every new combination costs a capture, and short runs never amortise them (`direct` is kept
        in the signature for the callers and ignored)."""
        import os
        if os.environ.get('TEBSCAT_LARGE_GRAPH', '1') == '0':
            return run(*args)
        dev = args[0].device
        key = (kind, args[0].shape[0], dev.index)
        entry = self._graphs.get(key)
        if entry is None:
            static = tuple(torch.empty_like(t) for t in args)
            for t, sbuf in zip(args, static):
                if any(t is i for i in ins):
                    sbuf.copy_(t)
            entry = self._capture(run, static, dev)
            entry['static'] = static
            self._remember(key, entry)
        else:
            self._graphs[key] = self._graphs.pop(key)          # most recently used last
        for t, sbuf in zip(args, entry['static']):
            if any(t is i for i in ins):
                sbuf.copy_(t)
        entry['graph'].replay()
        for t, sbuf in zip(args, entry['static']):
            if any(t is o for o in outs):
                t.copy_(sbuf)
        return args[-1]

    def _capture(self, run, args, dev):
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            run(*args)                                               # warm-up outside the capture (workspace allocation)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, capture_error_mode='thread_local'):      # other threads (e.g. DDP's reducer) keep working
            run(*args)
        return dict(graph=graph, static=None)

    def _remember(self, key, entry):
        self._graphs[key] = entry
        while len(self._graphs) > self.GRAPH_SLOTS:
            self._graphs.pop(next(iter(self._graphs)))

    def forward(self, x2, out, direct=True):
        """x2: (B, N) float32 CUDA contiguous; out: (B, C, n_out) float32 CUDA."""
        return self._graphed('fwd', self._run, (x2, out), ins=(x2,), outs=(out,), direct=direct)

    def bytes_per_signal(self, backward=False):
        """Workspace bytes one signal occupies on this level (what has to stay in L2 for the ops of a chunk to find
        their operands there instead of in HBM)."""
        Np = 1 << self.plan.geo.J_pad
        if backward:
            return 8 * Np * 8 + (2 << self.plan.lf) * 4
        return 2 * Np * 8 + (2 << self.plan.max_l2) * 4 + (2 << self.plan.lf) * 4

    def _run(self, x2, out):
        p, lib, g = self.plan, self._lib, self.handle
        B = x2.shape[0]
        dev = x2.device
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        U0, W1, W2, WL = self._workspace(B, dev)
        n, lf = p.geo.J_pad, p.lf
        fa = self.arena.data_ptr()

        def mulfold(src, spec, dst):
            off, log_src, logk, mask, logcw, sexp = spec
            _lib.check(lib.tebscat_large_mulfold(g, ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(fa + 4 * off),
                                                 ctypes.c_void_p(dst.data_ptr()), B, log_src, logk, mask, logcw, sexp, st))

        def fft(buf, log_len, inverse):
            _lib.check(lib.tebscat_large_fft(g, ctypes.c_void_p(buf.data_ptr()), B, log_len, 1 if inverse else 0, st))

        def pair(buf, log_len):                               # ifft -> modulus -> fft, one trip per level
            _lib.check(lib.tebscat_large_pair(g, ctypes.c_void_p(buf.data_ptr()), B, log_len, st))

        def leaf(src, spec, ch):                              # phi multiply + periodise -> iFFT -> unpad -> channel (:287-292)
            if lf <= LEAF_FUSED_MAX_LOG2 and FUSED_LEAVES:       # one launch, the 2^lf bins never leave shared memory
                off, log_src, logk, mask, logcw, sexp = spec
                _lib.check(lib.tebscat_large_leaf(g, ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(fa + 4 * off), B, log_src, logk,
                                                  mask, logcw, sexp, p.i0, p.n_out, p.n_paths, ch, ctypes.c_void_p(out.data_ptr()), st))
                return
            mulfold(src, spec, WL)
            fft(WL, lf, True)
            _lib.check(lib.tebscat_large_store(g, ctypes.c_void_p(WL.data_ptr()), B, lf, p.i0, p.n_out, p.n_paths, ch,
                                               ctypes.c_void_p(out.data_ptr()), st))

        def fused(dev_plan, src, log_len, chan_shift=0):
            """everything below the spectra `src` (B x 2^log_len) that fits one SM: one launch of the fused kernel"""
            _lib.check(lib.tebscat_scat1d_forward_gsrc(dev_plan.handle, ctypes.c_void_p(src.data_ptr()), 1 << log_len, B,
                                                       ctypes.c_void_p(out.data_ptr() + 4 * chan_shift * p.n_out), st))

        _lib.check(lib.tebscat_large_pad_load(g, ctypes.c_void_p(x2.data_ptr()), B, p.N, p.geo.pad_left, n,
                                              ctypes.c_void_p(U0.data_ptr()), st))                    # :278
        fft(U0, n, False)                                                                             # :280
        leaf(U0, p.s0, 0)
        if self._first is not None:                # first-order filters of <= 8192 samples with their whole subtrees
            fused(self._first, U0, n)
        for e in p.first:
            if e['n1'] in self._first_n1:
                continue
            mulfold(U0, e['mul'], W1)                                                                 # :307-310
            pair(W1, e['l1'])                                                                         # :312-318
            leaf(W1, e['leaf'], e['ch'])                                                              # :320-327
            kid_plan = self._kids.get(e['n1'])
            for k in e['kids']:
                if kid_plan is not None and k['n2'] in kid_plan[2]:
                    continue
                mulfold(W1, k['mul'], W2)                                                             # :347-348
                pair(W2, k['l2'])                                                                     # :350-355
                leaf(W2, k['leaf'], k['ch'])                                                          # :358-364
            if kid_plan is not None:               # the children of <= 8192 samples of a longer parent
                fused(kid_plan[0], W1, e['l1'], kid_plan[1])
        return out

    # ---- backward pass --------------------------------------------------------------------------------
    def _bwd_workspace(self, B, dev):
        if B > self._bws.get('cap', 0) or self._bws.get('dev') != dev.index:
            Np = 1 << self.plan.geo.J_pad
            self._bws = dict(cap=B, dev=dev.index, bufs=tuple(
                torch.empty(B * Np * 2, dtype=torch.float32, device=dev) for _ in range(8)) +
                (torch.empty(B * (2 << self.plan.lf), dtype=torch.float32, device=dev),))
            self._graphs = {k: v for k, v in self._graphs.items() if k[0] != 'bwd'}
        return self._bws['bufs']

    def backward(self, x2, gout, gx, direct=True):
        """gx = (dS/dx)^T gout, through the graph cache of `_graphed`."""
        return self._graphed('bwd', self._run_backward, (x2, gout, gx), ins=(x2, gout), outs=(gx,), direct=direct)

    def _run_backward(self, x2, gout, gx):
        """gx = (dS/dx)^T gout for x2 (B, N), gout (B, C, n_out), gx (B, N), all float32 CUDA contiguous.

        The reference differentiates the cascade with torch autograd (ModulusStable,
        kymatio/backend/torch_backend.py:5-96).  Here the transposed cascade runs on the same global buffers: the
        spectrum U0 is recomputed once, every first-order node is recomputed when its turn comes (the pre-modulus
        signals u1, u2 are what the modulus' backward needs) and its subtree is walked backwards --
            store^T (zero signal carrying the output gradient) -> iFFT^T = FFT -> (phi multiply + periodise)^T
            -> FFT^T = iFFT, real part -> modulus backward -> iFFT^T = FFT -> (psi multiply + periodise)^T, accumulated.
        With spectra in bit-reversed order the adjoint of the forward transform (natural -> bit-reversed) IS the
        unnormalised inverse transform (bit-reversed -> natural) and vice versa, so no new transform is needed."""
        p, lib, g = self.plan, self._lib, self.handle
        B, dev = x2.shape[0], x2.device
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        U0, gU0, W1, H1, gH1, W2, gW2, gA, WL = self._bwd_workspace(B, dev)
        n, lf = p.geo.J_pad, p.lf
        fa = self.arena.data_ptr()
        vp = ctypes.c_void_p

        def mulfold(src, spec, dst):
            off, log_src, logk, mask, logcw, sexp = spec
            _lib.check(lib.tebscat_large_mulfold(g, vp(src.data_ptr()), vp(fa + 4 * off), vp(dst.data_ptr()), B, log_src, logk,
                                                 mask, logcw, sexp, st))

        def unfold(gdst, spec, gsrc, accumulate):
            off, log_src, logk, mask, logcw, sexp = spec
            _lib.check(lib.tebscat_large_unfold(g, vp(gdst.data_ptr()), vp(fa + 4 * off), vp(gsrc.data_ptr()), B, log_src, logk,
                                                mask, logcw, sexp, 1 if accumulate else 0, st))

        def fft(buf, log_len, inverse):
            _lib.check(lib.tebscat_large_fft(g, vp(buf.data_ptr()), B, log_len, 1 if inverse else 0, st))

        def leaf_adjoint(spec, ch, gsrc, accumulate):
            """(phi multiply + periodise -> iFFT -> unpad -> channel)^T, core :287-292 / :320-327 / :358-364"""
            if lf <= LEAF_FUSED_MAX_LOG2 and FUSED_LEAVES:
                off, log_src, logk, mask, logcw, sexp = spec
                _lib.check(lib.tebscat_large_leaf_adjoint(g, vp(gout.data_ptr()), vp(fa + 4 * off), B, log_src, logk, mask, logcw, sexp,
                                                          p.i0, p.n_out, p.n_paths, ch, vp(gsrc.data_ptr()), 1 if accumulate else 0, st))
                return
            _lib.check(lib.tebscat_large_unstore(g, vp(gout.data_ptr()), B, lf, p.i0, p.n_out, p.n_paths, ch, vp(WL.data_ptr()), st))
            fft(WL, lf, False)
            unfold(WL, spec, gsrc, accumulate)

        def modulus_node(y, log_len, mod_out):
            """y: periodised spectrum -> u = ifft(y) (kept in y) -> mod_out = fft(|u|)"""
            fft(y, log_len, True)
            _lib.check(lib.tebscat_large_modulus_to(g, vp(y.data_ptr()), vp(mod_out.data_ptr()), B << log_len, st))
            fft(mod_out, log_len, False)

        def modulus_adjoint(gspec, u, log_len):
            """gspec: gradient w.r.t. fft(|u|) (bit-reversed) -> gradient w.r.t. the periodised spectrum ifft's input"""
            fft(gspec, log_len, True)                          # FFT^T; the modulus is real: only the real part matters
            _lib.check(lib.tebscat_large_modulus_backward(g, vp(u.data_ptr()), vp(gspec.data_ptr()), B << log_len, st))
            fft(gspec, log_len, False)                         # iFFT^T (its 1/L lives in the multiply's scale)

        _lib.check(lib.tebscat_large_pad_load(g, vp(x2.data_ptr()), B, p.N, p.geo.pad_left, n, vp(U0.data_ptr()), st))
        fft(U0, n, False)
        leaf_adjoint(p.s0, 0, gU0, False)                      # initialises gU0
        for e in p.first:
            l1 = e['l1']
            mulfold(U0, e['mul'], W1)
            modulus_node(W1, l1, H1)                     # W1 = u1, H1 = fft(|u1|)
            leaf_adjoint(e['leaf'], e['ch'], gH1, False)       # initialises gH1
            for k in e['kids']:
                l2 = k['l2']
                mulfold(H1, k['mul'], W2)
                fft(W2, l2, True)                              # W2 = u2
                leaf_adjoint(k['leaf'], k['ch'], gW2, False)   # gradient w.r.t. fft(|u2|)
                modulus_adjoint(gW2, W2, l2)
                unfold(gW2, k['mul'], gH1, True)
            modulus_adjoint(gH1, W1, l1)
            unfold(gH1, e['mul'], gU0, True)
        fft(gU0, n, True)                                      # FFT^T of the real padded signal
        _lib.check(lib.tebscat_large_pad_adjoint(g, vp(gU0.data_ptr()), B, p.N, p.geo.pad_left, n, vp(gx.data_ptr()), st))
        return gx

    # ---- backward of average=False (the un-averaged moduli are the outputs) ----------------------------------------
    def backward_unaveraged(self, x2, grow, gx, segments, direct=True):
        """gx = (d row / dx)^T grow for the average=False transform: `row` holds the unpadded moduli U1 / U2 back to back
        (`segments` = [(key, offset, length)] of schedule.build_plan_unaveraged; order 0 -- the input itself -- is the
        frontend's business).  Same graph cache as the other directions."""
        self._useg = {tuple(k): (int(off), int(ln)) for k, off, ln in segments}
        self._urow = int(grow.shape[-1])
        return self._graphed('bwdu', self._run_backward_unaveraged, (x2, grow, gx), ins=(x2, grow), outs=(gx,), direct=direct)

    def _run_backward_unaveraged(self, x2, grow, gx):
        """The transposed cascade of core/scattering1d.py:300-367 with average=False: a path's output is |u| itself
        (unpadded, at its own rate), so its cotangent enters the modulus' backward directly in the time domain --
        for a first-order path on top of what its second-order children send back through fft(|u1|)."""
        p, lib, g = self.plan, self._lib, self.handle
        B, dev = x2.shape[0], x2.device
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        U0, gU0, W1, H1, gH1, W2, gW2, gA, WL = self._bwd_workspace(B, dev)
        n = p.geo.J_pad
        fa = self.arena.data_ptr()
        vp = ctypes.c_void_p
        seg, row = self._useg, self._urow

        def mulfold(src, spec, dst):
            off, log_src, logk, mask, logcw, sexp = spec
            _lib.check(lib.tebscat_large_mulfold(g, vp(src.data_ptr()), vp(fa + 4 * off), vp(dst.data_ptr()), B, log_src, logk,
                                                 mask, logcw, sexp, st))

        def unfold(gdst, spec, gsrc, accumulate):
            off, log_src, logk, mask, logcw, sexp = spec
            _lib.check(lib.tebscat_large_unfold(g, vp(gdst.data_ptr()), vp(fa + 4 * off), vp(gsrc.data_ptr()), B, log_src, logk,
                                                mask, logcw, sexp, 1 if accumulate else 0, st))

        def fft(buf, log_len, inverse):
            _lib.check(lib.tebscat_large_fft(g, vp(buf.data_ptr()), B, log_len, 1 if inverse else 0, st))

        def cotangent(key, log_len, buf, accumulate):
            """+= / = the path's output gradient in the real part of samples [ind_start, ind_start + len)"""
            off, ln = seg[key]
            _lib.check(lib.tebscat_large_unstore_row(g, vp(grow.data_ptr()), B, row, off, log_len, p.geo.ind_start[n - log_len], ln,
                                                     1 if accumulate else 0, vp(buf.data_ptr()), st))

        def through_modulus(grad, u, log_len):
            """time-domain gradient w.r.t. |u| -> gradient w.r.t. the periodised spectrum the inverse transform read"""
            _lib.check(lib.tebscat_large_modulus_backward(g, vp(u.data_ptr()), vp(grad.data_ptr()), B << log_len, st))
            fft(grad, log_len, False)                          # iFFT^T (its 1/L lives in the multiply's scale)

        _lib.check(lib.tebscat_large_pad_load(g, vp(x2.data_ptr()), B, p.N, p.geo.pad_left, n, vp(U0.data_ptr()), st))
        fft(U0, n, False)
        for idx, e in enumerate(p.first):
            l1, key1 = e['l1'], p.keys[e['ch']]
            mulfold(U0, e['mul'], W1)
            fft(W1, l1, True)                                  # W1 = u1
            if e['kids']:
                _lib.check(lib.tebscat_large_modulus_to(g, vp(W1.data_ptr()), vp(H1.data_ptr()), B << l1, st))
                fft(H1, l1, False)                             # H1 = fft(|u1|)
                for c, k in enumerate(e['kids']):
                    l2 = k['l2']
                    mulfold(H1, k['mul'], W2)
                    fft(W2, l2, True)                          # W2 = u2
                    cotangent(p.keys[k['ch']], l2, gW2, False)
                    through_modulus(gW2, W2, l2)
                    unfold(gW2, k['mul'], gH1, c > 0)
                fft(gH1, l1, True)                             # FFT^T: what the children send back to |u1|
                cotangent(key1, l1, gH1, True)
            else:
                cotangent(key1, l1, gH1, False)
            through_modulus(gH1, W1, l1)
            unfold(gH1, e['mul'], gU0, idx > 0)
        fft(gU0, n, True)                                      # FFT^T of the real padded signal
        _lib.check(lib.tebscat_large_pad_adjoint(g, vp(gU0.data_ptr()), B, p.N, p.geo.pad_left, n, vp(gx.data_ptr()), st))
        return gx
