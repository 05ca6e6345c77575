"""``Scattering1D`` -- the reference's torch frontend surface on the fused CUDA path.

Same constructor, attributes, ``forward``/``scattering`` (returning ``[S, P]``),
``meta()`` and ``output_size()`` as ``ScatteringTorch1D``
(kymatio/scattering1d/frontend/torch_frontend.py:10-255 with
frontend/base_frontend.py:12-119), so it drops into
``hdf5_dataset/kymatio_phase_scattering.py:92-95`` and
``hdf5_dataset/create_hdf5_dataset.py:360``.  The transform itself is one call
into libtebscat.so (include/tebscat.h); there is no torch/CPU implementation
behind it.

Differences from the fork, all on the lenient side:
 * ``T=None`` means ``2**J`` (the fork crashes);
 * ``Q`` may be an int or a ``(Q1, 1)`` tuple (the fork takes ints only);
 * ``average=False`` (un-averaged U1/U2 outputs, ``out_type='list'`` or ``vectorize=False``) is one launch too; ``oversampling``,
   ``out_type='list'`` and ``vectorize=False`` are supported (the latter two are views of the
   fused array result).
"""
import ctypes
import math
import numbers
import warnings

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from . import filterbank as fbk
from .meta import compute_meta, output_size
from .schedule import LOG2_NP_MAX, build_plan, build_plan_unaveraged


class _DevicePlan:
    """Owner of one tebscat_plan handle (one per device)."""

    def __init__(self, plan, device_index):
        lib = _lib.load()
        desc = _lib.PlanDesc()
        desc.abi_version = _lib.ABI_VERSION
        desc.N = plan.N
        desc.log2_Np = plan.geo.J_pad
        desc.pad_left = plan.geo.pad_left
        desc.n_paths = plan.n_paths
        desc.n_out = plan.n_out
        desc.n_threads = plan.n_threads
        desc.smem_complex = plan.smem_complex
        desc.n_tasks = plan.tasks.shape[0]
        desc.n_steps = plan.steps.shape[0]
        desc.border_mode = int(getattr(plan, 'border', 0))     # phase plans only; the transform always reflects
        desc.scratch_complex = int(getattr(plan, 'scratch_complex', 0))   # per-CTA global scratch of schedules that park U0
        arena = np.ascontiguousarray(plan.arena, np.float32)
        tasks = np.ascontiguousarray(plan.tasks, np.int32)
        steps = np.ascontiguousarray(plan.steps, np.int32)
        chan = np.ascontiguousarray(plan.chan, np.int32)
        handle = ctypes.c_void_p()
        i32p = ctypes.POINTER(ctypes.c_int32)
        rc = lib.tebscat_plan_create(
            ctypes.byref(desc), arena.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), arena.size,
            tasks.ctypes.data_as(i32p), steps.ctypes.data_as(i32p), chan.ctypes.data_as(i32p), chan.size,
            int(device_index), ctypes.byref(handle))
        _lib.check(rc)
        self.handle = handle
        self._lib = lib

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                self._lib.tebscat_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class _ScatteringFunction(torch.autograd.Function):
    """Autograd node of the array-output transform: forward = the product path, backward = the transposed
    cascade in CUDA (no torch ops on either)."""

    @staticmethod
    def forward(ctx, x2, module):
        ctx.module = module
        ctx.save_for_backward(x2)
        return module._forward_array(x2)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gS):
        (x2,) = ctx.saved_tensors
        return ctx.module._backward_array(x2, gS), None


class _UnaveragedFunction(torch.autograd.Function):
    """Autograd node of the average=False transform: the row of un-averaged moduli; backward = the transposed cascade
    with the cotangents entering at the moduli (LargeDevicePlan.backward_unaveraged)."""

    @staticmethod
    def forward(ctx, x2, module):
        ctx.module = module
        ctx.save_for_backward(x2)
        return module._unaveraged_row(x2)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grow):
        (x2,) = ctx.saved_tensors
        return ctx.module._backward_unaveraged(x2, grow), None


class Scattering1D(nn.Module):
    def __init__(self, J, shape, Q=1, max_order=2, average=True, oversampling=0, vectorize=True,
                 out_type='array', backend='torch', T=None):
        super().__init__()
        self.frontend_name = 'torch'
        self.J = J
        self.shape = shape
        self.Q = Q
        self.max_order = max_order
        self.average = average
        self.oversampling = oversampling
        self.vectorize = vectorize
        self.out_type = out_type
        self.T = T
        if isinstance(backend, str):
            if not backend.startswith('torch'):
                raise ImportError('The backend ' + backend + ' can not be called from the frontend torch.')
        self.backend = 'tebscat'
        self.build()
        self.create_filters()
        self.register_filters()
        self._plans = {}
        self._host_plan = None
        self._sched = None

    # ---- base_frontend.py:27-77 ------------------------------------------------
    def build(self):
        self.r_psi = fbk.R_PSI
        self.sigma0 = fbk.SIGMA0
        self.alpha = fbk.ALPHA
        self.P_max = fbk.P_MAX
        self.eps = fbk.EPS
        self.criterion_amplitude = fbk.CRITERION_AMPLITUDE
        self.normalize = 'l1'
        if isinstance(self.shape, numbers.Integral):
            self.N = int(self.shape)
        elif isinstance(self.shape, tuple):
            self.N = self.shape[0]
            if len(self.shape) > 1:
                raise ValueError("If shape is specified as a tuple, it must "
                                 "have exactly one element")
        else:
            raise ValueError("shape must be an integer or a 1-tuple")
        if self.T is None:
            self.T = 2 ** self.J
        elif self.T > 2 ** self.J:
            raise ValueError("The temporal support T of the low-pass filter "
                             "cannot exceed 2**J (got {} > {})".format(self.T, 2 ** self.J))
        if self.max_order not in (1, 2):
            raise ValueError('max_order must be 1 or 2, got {}'.format(self.max_order))
        self._Q1 = fbk._as_Q1(self.Q)
        geo = fbk.build_geometry(self.N, self.J, self._Q1, self.T)
        self.J_pad = geo.J_pad
        self.pad_left, self.pad_right = geo.pad_left, geo.pad_right
        self.ind_start, self.ind_end = geo.ind_start, geo.ind_end
        self._geo = geo

    # ---- base_frontend.py:79-85 + torch_frontend.py:75-97 ----------------------
    def create_filters(self):
        self._bank = fbk.build_filter_bank(self.J_pad, self.J, self._Q1, self.T)

    def register_filters(self):
        """Filters as fp32 ``(L, 1)`` buffers ``tensor0..`` in the reference's order
        (phi levels, psi1, psi2 levels) and the dict views ``phi_f/psi1_f/psi2_f``."""
        n = 0

        def reg(a):
            nonlocal n
            t = torch.from_numpy(a).float().view(-1, 1)
            self.register_buffer('tensor' + str(n), t)
            n += 1
            return t

        b = self._bank
        self.phi_f = {'levels': [reg(a) for a in b.phi.levels], 'xi': 0, 'sigma': b.phi.sigma, 'j': 0}
        self.psi1_f = [{'levels': [reg(a) for a in p.levels], 'xi': p.xi, 'sigma': p.sigma, 'j': p.j}
                       for p in b.psi1]
        self.psi2_f = [{'levels': [reg(a) for a in p.levels], 'xi': p.xi, 'sigma': p.sigma, 'j': p.j}
                       for p in b.psi2]

    def load_filters(self):
        buffers = dict(self.named_buffers())
        n = 0
        for f in [self.phi_f] + self.psi1_f + self.psi2_f:
            for level in range(len(f['levels'])):
                f['levels'][level] = buffers['tensor' + str(n)]
                n += 1

    # ---- metadata ----------------------------------------------------------------
    def meta(self):
        return compute_meta(self.J, self._Q1, self.T, max_order=self.max_order)

    def output_size(self, detail=False):
        return output_size(self.J, self._Q1, self.T, max_order=self.max_order, detail=detail)

    # ---- plan handling -------------------------------------------------------------
    def _schedule(self):
        """Fused schedule of the current (J, N, Q, T, max_order, oversampling); raises NotImplementedError when the
        configuration does not fit the single-kernel design.  Both outcomes are cached under that key, so mutating
        `T` / `oversampling` afterwards re-decides (the reference re-reads its attributes on every call too)."""
        key = (self.J, self.N, self._Q1, self.T, self.max_order, int(self.oversampling))
        if self._sched is None or self._sched[0] != key:
            try:
                plan = build_plan(self.J, self.N, self._Q1, self.T, self.max_order, oversampling=int(self.oversampling))
            except NotImplementedError as e:
                plan = e
            self._sched = (key, plan)
            self._plans = {}
        if isinstance(self._sched[1], NotImplementedError):
            raise NotImplementedError(str(self._sched[1]))
        return self._sched[1]

    @property
    def _op_by_op(self):
        """True when the fused single-kernel schedule cannot hold the current configuration in shared memory (output-rate
        lengths of 2048 samples and more, i.e. T <= 4 at a padded length of 8192): the op-by-op level of tebscat/large.py
        serves it -- the same CUDA ops, one launch each.  Part of the schedule key, not a sticky flag."""
        if self.J_pad > LOG2_NP_MAX:
            return False
        try:
            self._schedule()
            return False
        except NotImplementedError:
            return True

    def _plan_for(self, device_index):
        sched = self._schedule()
        if device_index not in self._plans:
            self._plans[device_index] = _DevicePlan(sched, device_index)
            self._apply_window(self._plans[device_index])
        return self._plans[device_index]

    def set_window(self, window):
        """Extension: analysis window w[t] applied to the samples as the kernels load them (x[t] * w[t] before
        padding), e.g. the Tukey taper of KymatioPhaseScattering1D (hdf5_dataset/kymatio_phase_scattering.py:405-407).
        `window`: N values or None.  Acts on every level (fused, op-by-op, large support) and on the backward pass."""
        if window is not None:
            window = np.ascontiguousarray(np.asarray(window, dtype=np.float32).reshape(-1))
            if window.shape[0] != self.N:
                raise ValueError('window of {} samples on a transform of shape={}'.format(window.shape[0], self.N))
        self._window = window
        for p in list(self._plans.values()) + list(getattr(self, '_uplans', {}).values()):
            self._apply_window(p)
        for p in getattr(self, '_lplans', {}).values():
            self._apply_window_large(p)

    def _apply_window(self, dev_plan):
        w = getattr(self, '_window', None)
        ptr = w.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if w is not None else None
        _lib.check(_lib.load().tebscat_plan_set_window(dev_plan.handle, ptr))

    def _apply_window_large(self, ldp):
        w = getattr(self, '_window', None)
        ptr = w.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if w is not None else None
        _lib.check(_lib.load().tebscat_large_set_window(ldp.handle, ptr, int(self.N)))
        ldp.invalidate_graphs()

    def _check_options(self):
        if self.out_type not in ('array', 'list'):
            raise RuntimeError("The out_type must be one of 'array' or 'list'.")
        if not self.average and self.out_type == 'array' and self.vectorize:
            raise ValueError("Options average=False, out_type='array' and "
                             "vectorize=True are mutually incompatible. "
                             "Please set out_type to 'list' or vectorize to "
                             "False.")
        if not self.vectorize:
            warnings.warn("The vectorize option is deprecated and will be "
                          "removed in version 0.3. Please set "
                          "out_type='list' for equivalent functionality.", DeprecationWarning)
        if int(self.oversampling) < 0:
            raise ValueError('oversampling must be >= 0')

    # ---- forward ----------------------------------------------------------------------
    def forward(self, x):
        if x is None:
            raise TypeError('The input should be not empty.')
        if not x.is_contiguous():
            raise RuntimeError('Tensors must be contiguous.')
        return self.scattering(x)

    def scattering(self, x):
        if len(x.shape) < 1:
            raise ValueError('Input tensor x should have at least one axis, got {}'.format(len(x.shape)))
        self._check_options()
        if x.shape[-1] != self.N:
            raise ValueError('Input length {} does not match shape={}'.format(x.shape[-1], self.N))
        if x.dtype is not torch.float32:
            raise TypeError('Input and filter must be of the same dtype.')
        if x.device.type != 'cuda':
            raise TypeError('Input must be on GPU.')
        batch_shape = x.shape[:-1]
        x2 = x.reshape(-1, self.N)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        B = x2.shape[0]
        needs_grad = torch.is_grad_enabled() and x.requires_grad
        if not self.average:
            return self._scattering_unaveraged(x2, batch_shape, needs_grad)
        # differentiable like the reference's torch backend (ModulusStable, kymatio/backend/torch_backend.py:5-96):
        # the forward is the same fused launch, the backward the transposed cascade of tebscat/large.py
        S = _ScatteringFunction.apply(x2, self) if needs_grad else self._forward_array(x2)
        C, n_out = S.shape[1], S.shape[2]
        P = S.view(B, 1, C, n_out)                       # core/scattering1d.py:395-397: P is S before the reshape
        if self.out_type == 'array' and self.vectorize:
            return [S.reshape(batch_shape + (C, n_out)), P]
        # the other output conventions are views of the same fused result (core :379-384,
        # torch_frontend.py:240-253); in the reference P aliases S in these cases too
        meta = self.meta()
        if self.out_type == 'array':                     # vectorize=False: dict keyed by the filter indices
            out = {meta['key'][c]: S[:, c:c + 1, :].reshape(batch_shape + (1, n_out)) for c in range(C)}
            return [out, out]
        out = []                                         # out_type == 'list'
        for c in range(C):
            key = meta['key'][c]
            j = tuple(int(v) for v in meta['j'][c][:len(key)])
            out.append({'coef': S[:, c, :].reshape(batch_shape + (n_out,)), 'j': j})
        return [out, out]

    def _forward_array(self, x2):
        """S (B, C, n_out) of x2 (B, N): the fused single-launch cascade up to padded lengths of 2^13, the
        large-support level of tebscat/large.py above."""
        dev = x2.device
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        B = x2.shape[0]
        if self.J_pad > LOG2_NP_MAX or self._op_by_op:
            lp, ldp = self._large_plan_for(index)
            S = torch.empty((B, lp.n_paths, lp.n_out), dtype=torch.float32, device=dev)
            chunk = self._large_chunk(ldp, backward=False)
            for b0 in range(0, B, chunk):
                ldp.forward(x2[b0:b0 + chunk], S[b0:b0 + chunk], direct=B <= chunk)
            return S
        plan = self._plan_for(index)
        sched = self._sched[1]
        S = torch.empty((B, sched.n_paths, sched.n_out), dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.load().tebscat_scat1d_forward(plan.handle, x2.data_ptr(), B, S.data_ptr(), stream))
        return S

    def _large_chunk(self, ldp, backward):
        """Signals per chunk of the op-by-op levels, bounded by workspace memory: B * Np <= 2^28 forward (U0 and W1 2 GB
        each, W2 <= 2 GB, leaves: <= 6.2 GB), 2^25 backward (eight buffers of 256 MB + leaves).  As large as memory
        allows on purpose: every op is one launch over the chunk and its tile jobs run at the interpreter's per-step
        latency, so the level is launch/latency-bound, not HBM-bound -- measured (tools/l2_chunk_sweep.sh,
        profiles/r02_l2_chunk_sweep.txt): chunks sized for their workspace to stay in L2 (64 MB) run at 2.4 instead of
        5.7 TFLOP/s at a padded length of 2^14, and the backward at 19 k instead of 42 k signals/s."""
        import os
        if os.environ.get('TEBSCAT_LARGE_L2_MB'):              # experiment switch of the sweep tool
            budget = int(float(os.environ['TEBSCAT_LARGE_L2_MB']) * (1 << 20))
            return int(max(16, min(1 << 16, budget // ldp.bytes_per_signal(backward))))
        return max(1, ((1 << 25) if backward else (1 << 28)) >> self.J_pad)

    def _large_plan_for(self, index):
        """Op-list plan of the large-support level: the forward above 2^13 and the backward pass at any length."""
        from .large import LargeDevicePlan, LargePlan
        key = (self.J, self.N, self._Q1, self.T, self.max_order, int(self.oversampling))
        if getattr(self, '_lsched', None) is None or self._lsched[0] != key:
            self._lsched = (key, LargePlan(self.J, self.N, self._Q1, self.T, self.max_order, int(self.oversampling)))
            self._lplans = {}
        if index not in self._lplans:
            self._lplans[index] = LargeDevicePlan(self._lsched[1], index)
            if getattr(self, '_window', None) is not None:
                self._apply_window_large(self._lplans[index])
        return self._lsched[1], self._lplans[index]

    def _backward_array(self, x2, gS):
        """(dS/dx)^T gS -- see LargeDevicePlan.backward."""
        dev = x2.device
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        lp, ldp = self._large_plan_for(index)
        gS = gS.contiguous()
        if gS.dtype is not torch.float32:
            raise TypeError('Input and filter must be of the same dtype.')
        gx = torch.empty_like(x2)
        chunk = self._large_chunk(ldp, backward=True)
        for b0 in range(0, x2.shape[0], chunk):
            ldp.backward(x2[b0:b0 + chunk], gS[b0:b0 + chunk], gx[b0:b0 + chunk], direct=x2.shape[0] <= chunk)
        return gx

    def _backward_unaveraged(self, x2, grow):
        """(d row / dx)^T grow of the average=False transform, see LargeDevicePlan.backward_unaveraged."""
        dev = x2.device
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        lp, ldp = self._large_plan_for(index)
        sched = self._usched[1]
        grow = grow.contiguous()
        gx = torch.empty_like(x2)
        chunk = self._large_chunk(ldp, backward=True)
        for b0 in range(0, x2.shape[0], chunk):
            ldp.backward_unaveraged(x2[b0:b0 + chunk], grow[b0:b0 + chunk], gx[b0:b0 + chunk], sched.segments,
                                    direct=x2.shape[0] <= chunk)
        return gx

    def _scattering_unaveraged(self, x2, batch_shape, needs_grad=False):
        """average=False (core/scattering1d.py:293-294, :329-330, :366-367): the input itself, then the unpadded
        moduli U1 / U2 at their own rates.  One launch writes all paths back to back into one row per signal;
        the coefficients returned are views of it.  Differentiable like the reference (the order-0 coefficient IS
        the input; the row's backward is the transposed cascade in CUDA)."""
        row = _UnaveragedFunction.apply(x2, self) if needs_grad else self._unaveraged_row(x2)
        sched = self._usched[1]
        meta = self.meta()
        j_of = {tuple(meta['key'][c]): tuple(int(v) for v in meta['j'][c][:len(meta['key'][c])]) for c in range(len(meta['key']))}
        coefs = [((), x2.reshape(batch_shape + (self.N,)))]                       # :294  S_0 = x
        coefs += [(k, row[:, off:off + ln].reshape(batch_shape + (ln,))) for k, off, ln in sched.segments]
        if self.out_type == 'array':                     # vectorize=False: dict keyed by the filter indices (:380-381)
            out = {k: v for k, v in coefs}
            return [out, out]
        out = [{'coef': v, 'j': j_of[k]} for k, v in coefs]                      # :382-385
        return [out, out]

    def _unaveraged_row(self, x2):
        key = (self.J, self.N, self._Q1, self.T, self.max_order, int(self.oversampling))
        if getattr(self, '_usched', None) is None or self._usched[0] != key:
            self._usched = (key, build_plan_unaveraged(self.J, self.N, self._Q1, self.T, self.max_order,
                                                       oversampling=int(self.oversampling)))
            self._uplans = {}
        sched = self._usched[1]
        dev = x2.device
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        if index not in self._uplans:
            self._uplans[index] = _DevicePlan(sched, index)
            self._apply_window(self._uplans[index])
        B = x2.shape[0]
        row = torch.empty((B, sched.n_out), dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.load().tebscat_scat1d_forward(self._uplans[index].handle, x2.data_ptr(), B, row.data_ptr(), stream))
        return row

    def forward_normalized(self, x, mean, variance, log_channels='all_except_0', asinh_channels=None,
                           log_epsilon=1e-6, trim=0, time_major=True):
        """The transform with the dataset's post-processing fused into its stores: returns
        exactly what ``CombinedHDF5Dataset.__getitem__`` hands the model for a stored ``fhr_st`` record --
        trimmed by `trim` decimated samples at both ends (hdf5_dataset/hdf5_dataset.py:733-741), normalised
        like ``normalize_tensor_data`` (:18-137: ``log(clamp(x, 0) + eps)`` / ``asinh`` on the configured
        channels, then ``(x - mean) / (sqrt(variance) + 1e-8)``) and laid out ``(time, channels)``
        (:758-759).  `mean` / `variance` are the per-channel statistics of DatasetStatsCalculator;
        `log_channels` / `asinh_channels` take the values of the reference's config ('all_except_0', 'all',
        or a list of channel indices)."""
        self._check_options()
        if self.out_type != 'array' or not self.vectorize:
            raise NotImplementedError('forward_normalized produces the array output only')
        if x.shape[-1] != self.N:
            raise ValueError('Input length {} does not match shape={}'.format(x.shape[-1], self.N))
        if x.dtype is not torch.float32:
            raise TypeError('Input and filter must be of the same dtype.')
        if x.device.type != 'cuda':
            raise TypeError('Input must be on GPU.')
        batch_shape = x.shape[:-1]
        x2 = x.reshape(-1, self.N)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        B = x2.shape[0]
        plan = self._plan_for(x.device.index if x.device.index is not None else torch.cuda.current_device())
        sched = self._sched[1]
        C, n_out = sched.n_paths, sched.n_out
        trim = int(trim)
        if trim < 0 or 2 * trim >= n_out:
            raise ValueError('trim={} leaves nothing of {} samples'.format(trim, n_out))
        mean_t = torch.as_tensor(np.asarray(mean, np.float32)).reshape(-1)
        std_t = torch.as_tensor(np.asarray(np.sqrt(np.asarray(variance)), np.float32)).reshape(-1)   # :62-66
        if mean_t.numel() != C or std_t.numel() != C:
            raise ValueError('statistics must have one entry per channel ({})'.format(C))
        mode = np.zeros(C, np.uint8)
        if log_channels == 'all_except_0':                                        # :86-91
            mode[1:] = 1
        elif isinstance(log_channels, (list, tuple)):
            mode[list(log_channels)] = 1
        if asinh_channels == 'all':                                               # :94-99 (applied after the log)
            if mode.any():
                raise NotImplementedError('a channel with both log and asinh normalisation is not supported')
            mode[:] = 2
        elif isinstance(asinh_channels, (list, tuple)) and len(asinh_channels):
            if mode[list(asinh_channels)].any():
                raise NotImplementedError('a channel with both log and asinh normalisation is not supported')
            mode[list(asinh_channels)] = 2
        dev = x.device
        mean_d, std_d, mode_d = mean_t.to(dev), std_t.to(dev), torch.from_numpy(mode).to(dev)
        keep = n_out - 2 * trim
        out = torch.empty((B, keep, C) if time_major else (B, C, keep), dtype=torch.float32, device=dev)
        ep = _lib.Epilogue(mean_d.data_ptr(), std_d.data_ptr(), mode_d.data_ptr(), float(log_epsilon), trim,
                           1 if time_major else 0)
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = _lib.load().tebscat_scat1d_forward_ex(plan.handle, x2.data_ptr(), B, out.data_ptr(), ctypes.byref(ep), stream)
        _lib.check(rc)
        return out.reshape(batch_shape + tuple(out.shape[1:]))

    def scattering_host(self, x, out=None, device=0):
        """End-to-end path on HOST tensors: pinned staging, chunked H2D / kernel / D2H
        overlap inside the library (tebscat_scat1d_forward_host).  Returns S on the host.
        Served by the fused single-kernel level only (padded lengths up to 2^13 that fit its schedule); other
        configurations raise NotImplementedError -- use forward() on device tensors for them."""
        self._check_options()
        if x.device.type != 'cpu' or x.dtype is not torch.float32:
            raise TypeError('scattering_host expects a float32 CPU tensor')
        if x.shape[-1] != self.N:
            raise ValueError('Input length {} does not match shape={}'.format(x.shape[-1], self.N))
        batch_shape = x.shape[:-1]
        x2 = x.reshape(-1, self.N).contiguous()
        B = x2.shape[0]
        plan = self._plan_for(device)
        sched = self._sched[1]
        C, n_out = sched.n_paths, sched.n_out
        if out is None:
            out = torch.empty((B, C, n_out), dtype=torch.float32, pin_memory=True)
        elif (not torch.is_tensor(out) or out.device.type != 'cpu' or out.dtype is not torch.float32 or
              not out.is_contiguous() or out.numel() != B * C * n_out):
            # the library writes B * C * n_out floats through this pointer
            raise ValueError('out must be a contiguous float32 CPU tensor of {} elements'.format(B * C * n_out))
        rc = _lib.load().tebscat_scat1d_forward_host(plan.handle, x2.data_ptr(), B, out.data_ptr())
        _lib.check(rc)
        return out.reshape(batch_shape + (C, n_out))


ScatteringTorch1D = Scattering1D
__all__ = ['Scattering1D', 'ScatteringTorch1D']
