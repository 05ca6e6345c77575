"""Host-side plan builder: filter arena + step schedule for the fused cascade kernel.

The CUDA kernel (csrc/scat_core.cuh) is an interpreter of STEPS; each step is a
set of TASKS on disjoint, warp-aligned thread ranges of one 512-thread CTA
followed by one barrier.  This module turns the cascade of
``kymatio/scattering1d/core/scattering1d.py:269-370`` into that form.

Design (what makes the steps full):

* PAIR PACKING: every forward transform acts on a real signal (a modulus), so two of them share one
  complex transform (FFT_PACK); phi leaves of a packed spectrum need no separation, psi2 children are
  separated on the fly by MULFOLD2 (mirrored bins).

* Same-length transforms are BATCHED: first-order filters that share the
  subsampling 2^k1 are processed ``8192 / L1`` at a time in one contiguous buffer,
  their children (second order) are grouped by length the same way, and every FFT
  pass of a batch is ONE task whose butterflies span all blocks of the buffer.
* The phi low-pass leaves (one per output channel) are tiny: their spectra are
  collected in a ping-pong POOL and inverse-transformed / unpadded / stored
  (core :287-292, :320-327, :357-364) by one batched task per pass.
* "Filter multiply + periodise" (MULFOLD) visits only the 4-bin chunks where the
  filter is not negligible (below 1e-9 of its peak for every output bin).
* Chains of tasks that do not depend on each other are packed side by side into
  steps by a list scheduler under two resources: the 512 threads of the CTA and
  the shared-memory slots of their buffers.
* The signal's spectrum U0 is only ever READ after its transform: it is parked in a per-CTA
  global scratch (L2-resident) and its consumers are global-source multiplies (OP_GMULFOLD,
  OP_GMULFOLD2 for the two partners of a packed pair), which frees 8192 slots for the rest
  of the signal (`u0_in_scratch`).  The same multiplies let the subtrees of at
  most 8192 samples under a longer spectrum run here (`build_hybrid_plans`).
* The scheduler is greedy, so `build_plan` builds several candidate schedules (chain priority
  depth first / by weight, subtree reservation 'sum' / 'peak') and keeps the one the cost
  model prefers; the model is calibrated with on-device step timings (tools/step_profile.py).

Filters are cast to fp32 exactly like ``register_filters``
(``frontend/torch_frontend.py:75-97``), permuted to bit-reversed bin order and
laid out in one flat arena.  Task encoding (12 x int32) must match
``csrc/scat_core.cuh``.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import filterbank as fbk

OP_NOP, OP_LOAD, OP_FFT, OP_MULFOLD, OP_STOREB, OP_STOREZ, OP_TINY, OP_MULFOLD2, OP_LOADPAIR, OP_STOREU = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9
OP_LOADC, OP_STOREC = 10, 11
OP_GMULFOLD = 12                  # MULFOLD with its source spectrum in global memory (fused subtrees of the large-support level)
OP_GMULFOLD2 = 13                 # two filters (the partners of a packed pair) on one read of the global source
FFT_INV, FFT_MOD, FFT_FUSE_FWD, FFT_PACK = 1, 2, 4, 8
TASK_INTS = 12

N_THREADS = 512
LOG2_NP_MAX = 13                  # single-CTA limit of the kernel (kLog2TwMax)
SMEM_BYTES_MAX = 227 * 1024
TW_SLOTS = 68 + 136 + 4 + 48 + 288   # padded twiddle tables + the kernel's static shared memory (context, record ring)
MASK_THRESHOLD = 1e-9             # relative filter magnitude below which a 4-bin chunk is skipped
BATCH_SLOTS = 8192                # target size of one batch buffer (complex slots)
POOL_SLOTS = 1024                 # size of one half of the leaf pool


def pack_enabled() -> bool:
    """Pair packing (two real moduli per forward transform); TEBSCAT_PACK=0 builds the unpacked plan (A/B)."""
    return os.environ.get('TEBSCAT_PACK', '1') != '0'


def bitrev_indices(n: int) -> np.ndarray:
    """perm[p] = bit-reversal of p on log2(n) bits."""
    bits = int(math.log2(n))
    p = np.arange(n, dtype=np.int64)
    r = np.zeros(n, dtype=np.int64)
    for b in range(bits):
        r |= ((p >> b) & 1) << (bits - 1 - b)
    return r


def radix_split(n: int) -> List[int]:
    """Passes of a length-2^n forward (DIF) transform as log2-radices: radix 16 as often as
    possible, the remainder LAST -- the last DIF pass (first DIT pass: inverse transforms use
    the reversed list) has unit stride and needs no twiddles, so it is the cheapest place
    for a small radix."""
    if n <= 0:
        return []
    return [4] * (n // 4) + ([n % 4] if n % 4 else [])


def _round16(n: int) -> int:
    return (n + 15) & ~15


# ------------------------------------------------------------------------------------
# symbolic tasks / chains
# ------------------------------------------------------------------------------------
@dataclass
class Buf:
    """A shared-memory buffer of `size` complex slots (allocated by the scheduler)."""
    size: int
    name: str = ''
    off: int = -1
    readers_left: int = 0          # chains that still have to read it before it can be freed
    producer_done: bool = False


LEAF = 'leaf'                      # destination marker: a slot of the leaf pool


@dataclass
class TaskSpec:
    op: int
    work: int                      # independent work items
    lat: float                     # cycles of one work item on an otherwise idle SM
    instr: float                   # warp instructions of one work item
    a: object = 0                  # int, or (Buf, offset)
    b: int = 0
    c: int = 0
    d: object = 0                  # int, (Buf, offset) or LEAF
    e: int = 0
    f: object = 0                  # int, or (Buf, offset): partner region of a packed pass
    g: object = 0                  # int, or (Buf, offset): second destination of a MULFOLD2
    h: int = 0
    trip: float = -1.0             # latency one more trip of a thread adds (default: instr)
    sexp: int = 0
    pad: int = 0                   # field 11: bit 0 = relaxed barrier (post-pass), bits 4..15 = third chained FFT pass
    channel: object = -1           # output channel(s) of a leaf MULFOLD: int or (real-part, imaginary-part or -1)
    tpi: int = 1                   # threads per work item (32 for warp-local FFT tasks)


@dataclass
class Chain:
    name: str
    stages: List[List[TaskSpec]]
    after: List['Chain'] = field(default_factory=list)
    waits: List['Chain'] = field(default_factory=list)  # ordering only: must be complete before this chain starts
    reads: List[Buf] = field(default_factory=list)     # buffers read by stage 0
    owns: List[Buf] = field(default_factory=list)      # allocated when the chain starts
    frees_own_at_end: bool = False
    depth: int = 0
    pool_half: int = -1            # >= 0 for the flush chains of the leaf pool
    shrink: List[Tuple[int, Buf, int]] = field(default_factory=list)   # (stage, buffer, new size): tail freed after the stage
    # scheduler state
    stage: int = 0
    issued: List[bool] = field(default_factory=list)
    done_step: int = -1
    priority: float = 0.0


# Cost model calibrated on B200 with tools/step_floor.py / tools/step_profile.py:
# a step lasts  STEP_OVERHEAD + max_t(lat_t) + sum_t warps_t/4 * trips_t * instr_t  cycles, where lat is
# the dependent-issue latency of one work item on an idle SM and instr the issue cycles one
# more warp per scheduler adds (one R16 butterfly alone: 1850 cycles; 4 warps/scheduler: 3330).
_FFT_LAT = {1: 350.0, 2: 450.0, 3: 600.0, 4: 800.0}
_FFT_INSTR = {1: 40.0, 2: 90.0, 3: 180.0, 4: 330.0}
STEP_OVERHEAD = 450.0


def _global_pass(ref, n: int, count: int, logB: int, r: int, flags: int, partner=0, n_paired: int = 0) -> TaskSpec:
    bfly = count << (n - r)
    if r <= 3 and logB == r and not (flags & FFT_MOD) and ((count << n) & 15) == 0:
        # unit-stride remainder pass: the kernel takes 16 slots per thread and trip
        return TaskSpec(OP_FFT, (count << n) >> 4, 500.0, 300.0, a=ref, b=bfly, c=logB, d=r, e=flags)
    lat, instr = _FFT_LAT[r], _FFT_INSTR[r]
    if flags & FFT_FUSE_FWD:
        k = 1.8 + (0.9 * n_paired / count if flags & FFT_PACK else 0.0)
        lat, instr = k * lat, k * instr
    return TaskSpec(OP_FFT, bfly, lat, instr, a=ref, b=bfly, c=logB, d=r, e=flags, f=partner, g=n_paired)


def _fft_stages(ref, n: int, count: int, kind: str, hi: int = 0) -> List[List[TaskSpec]]:
    """Stages of `count` in-place length-2^n transforms stored back to back at `ref`.

    kind: 'fwd' (DIF), 'inv' (DIT), 'inv_mod' (DIT ending in the modulus) or 'pair'
    (inverse, modulus, forward -- core/scattering1d.py:312-318 / :350-355).  In a 'pair' the
    last inverse and the first forward pass touch the same 16 elements per thread and are
    fused into one task (one shared-memory round trip and one barrier less).

    Packed 'pair' (hi > 0, n >= 4): `count + hi` transforms are inverted; block i (< hi) of the second
    group -- stored right behind the first `count` blocks -- is the PARTNER of block i: the fused pass
    takes both moduli as real and imaginary part, so the forward half runs on `count` blocks only.

    Consecutive passes over blocks of <= 512 slots are chained into one task afterwards
    (_merge_local_passes)."""
    if n < 1:
        raise NotImplementedError('length-1 transforms are not supported')
    if n < 4:
        # 2, 4 or 8 samples: one thread per transform does the whole thing (csrc: tiny_task)
        flags = {'fwd': 0, 'inv': FFT_INV, 'inv_mod': FFT_INV | FFT_MOD, 'pair': FFT_INV | FFT_FUSE_FWD}[kind]
        return [[TaskSpec(OP_TINY, count, 400.0 + 60.0 * (1 << n), 8.0 * (1 << (2 * n)), a=ref, b=count, c=n, e=flags)]]
    dif = []
    logB = n
    for r in radix_split(n):
        dif.append((logB, r))
        logB -= r
    dit = dif[::-1]
    if kind == 'fwd':
        return [[_global_pass(ref, n, count, b, r, 0)] for b, r in dif]
    if kind in ('inv', 'inv_mod'):
        mod = FFT_MOD if kind == 'inv_mod' else 0
        return [[_global_pass(ref, n, count, b, r, FFT_INV | (mod if i == len(dit) - 1 else 0))]
                for i, (b, r) in enumerate(dit)]
    if kind == 'pair':
        out = [[_global_pass(ref, n, count + hi, b, r, FFT_INV)] for b, r in dit[:-1]]
        b, r = dit[-1]
        if hi > 0:
            out.append([_global_pass(ref, n, count, b, r, FFT_INV | FFT_MOD | FFT_FUSE_FWD | FFT_PACK,
                                     partner=(ref[0], ref[1] + (count << n)), n_paired=hi)])
        else:
            out.append([_global_pass(ref, n, count, b, r, FFT_INV | FFT_MOD | FFT_FUSE_FWD)])
        out += [[_global_pass(ref, n, count, b2, r2, 0)] for b2, r2 in dif[1:]]
        return out
    raise ValueError(kind)


MAX_CHAIN = 4                     # passes per FFT task (the first + three packed into h / pad)
LOCAL_LOG2 = 9                    # blocks of <= 512 slots: 32 consecutive work items own them in every pass


def chain_enabled() -> bool:
    return os.environ.get('TEBSCAT_CHAIN', '1') != '0'


def _merge_local_passes(stages: List[List[TaskSpec]]) -> List[List[TaskSpec]]:
    """Chain consecutive single-task FFT stages whose blocks are at most 512 slots into ONE task: its
    passes run back to back separated by warp fences (csrc: fft_task), not by CTA barriers and trips through
    the dispatcher.  Valid because such passes -- radix-16 butterflies or 16-slot unit-stride groups, the same
    number of work items in each -- map 32 consecutive work items onto the same 512 slots."""
    if not chain_enabled():
        return stages

    def local(st) -> bool:
        if len(st) != 1:
            return False
        t = st[0]
        if t.op != OP_FFT or t.c > LOCAL_LOG2 or (t.e & (FFT_FUSE_FWD | FFT_PACK)) or t.h or t.pad:
            return False
        return ((t.b << t.d) & 15) == 0 and (t.d == 4 or t.c == t.d)

    out: List[List[TaskSpec]] = []
    for st in stages:
        if out and local(st) and len(out[-1]) == 1 and getattr(out[-1][0], '_chain', 0) and \
                out[-1][0]._chain < MAX_CHAIN and out[-1][0].a == st[0].a and \
                (out[-1][0].b << out[-1][0].d) == (st[0].b << st[0].d):
            head, t = out[-1][0], st[0]
            code = t.c | (t.d << 4) | (t.e << 7)
            if head._chain == 1:
                head.h |= code
            elif head._chain == 2:
                head.h |= code << 12
            else:
                head.pad |= code << 4
            head._chain += 1
            head.lat += t.lat
            head.instr += t.instr
            continue
        if local(st):
            st[0]._chain = 1
        out.append(st)
    return out


class _Arena:
    def __init__(self):
        self.chunks: List[np.ndarray] = []
        self.size = 0
        self.by_off: Dict[int, np.ndarray] = {}
        self._masks: Dict[Tuple[int, int], int] = {}

    def add(self, f64: np.ndarray) -> int:
        f32 = f64.astype(np.float32)                       # the reference's .float() cast
        perm = f32[bitrev_indices(f32.shape[0])]
        off = self.size
        pad = (-perm.shape[0]) % 4
        self.chunks.append(np.concatenate([perm, np.zeros(pad, np.float32)]))
        self.by_off[off] = perm
        self.size += perm.shape[0] + pad
        return off

    @staticmethod
    def chunk_log2(logk: int) -> int:
        """log2 of the chunk width: 4 bins, or k/32 when the fold has more than 32 four-bin chunks."""
        return max(2, logk - 5)

    def chunk_mask(self, off: int, logk: int) -> int:
        """Bit c set <=> some output bin has a non-negligible filter value in chunk c (<= 32 chunks)."""
        key = (off, logk)
        if key not in self._masks:
            f = np.abs(self.by_off[off].astype(np.float64))
            k, cw = 1 << logk, 1 << self.chunk_log2(logk)
            peak = f.max()
            sig = (f.reshape(-1, k // cw, cw).max(axis=2) > MASK_THRESHOLD * peak).any(axis=0)
            mask = 0
            for c, on in enumerate(sig):
                if on:
                    mask |= 1 << c
            self._masks[key] = mask if mask else 1
        return self._masks[key]

    def compact(self, off: int, logk: int, mask: int) -> int:
        """Copy of filter `off` restricted to the active chunks of a k=2^logk fold, laid out
        [output bin][active chunk][chunk width] so that consecutive outputs read consecutive memory."""
        key = ('c', off, logk, mask)
        if key not in self._masks:
            k, cw = 1 << logk, 1 << self.chunk_log2(logk)
            f = self.by_off[off].reshape(-1, k // cw, cw)
            sel = [c for c in range(k // cw) if (mask >> c) & 1]
            comp = np.ascontiguousarray(f[:, sel, :]).reshape(-1)
            new = self.size
            self.chunks.append(comp)
            self.size += comp.shape[0]
            self._masks[key] = new
        return self._masks[key]

    def finish(self) -> np.ndarray:
        return np.concatenate(self.chunks) if self.chunks else np.zeros(4, np.float32)


def _fuse_first_inverse_pass(mulfolds: List[TaskSpec], stages: List[List[TaskSpec]], n: int):
    """If the transform's first inverse pass is the unit-stride remainder pass (radix 2 or 4) and
    every MULFOLD feeding it is a plain k=1 product, let the MULFOLD do that pass on the four
    slots each of its threads owns and drop the pass (core :307-312 in one round trip)."""
    r0 = radix_split(n)[-1]
    if r0 <= 2 and len(radix_split(n)) > 1 and all(m.op in (OP_MULFOLD, OP_GMULFOLD, OP_GMULFOLD2) and m.c == 0 for m in mulfolds):
        first = stages[0][0]
        if first.op == OP_FFT and first.c == r0 and first.d == r0 and (first.e & FFT_INV) and not (first.e & (FFT_MOD | FFT_FUSE_FWD)):
            for m in mulfolds:
                if m.op == OP_GMULFOLD2:
                    m.h |= r0 << 8
                else:
                    m.g = r0
                m.instr += 10.0
            return stages[1:]
    return stages


def _mulfold(arena: _Arena, src, log_src: int, logk: int, dst, filt_off: int, channel: int = -1) -> TaskSpec:
    log_dst = log_src - logk
    trip = -1.0
    if logk >= 2:
        mask = arena.chunk_mask(filt_off, logk)
        logcw = _Arena.chunk_log2(logk)
        nch = bin(mask).count('1') << (logcw - 2)     # in units of four bins
        filt_off = arena.compact(filt_off, logk, mask)
        # the kernel takes four outputs per thread and trip; every trip is one L2 round trip when at most two
        # chunks are active, one per chunk otherwise (measured: tools/step_floor.py, tools/step_phases.py)
        work = -(-(1 << log_dst) // 4)
        if nch <= 2 and logcw == 2:
            lat, instr, trip = 1300.0 + 250.0 * nch, 100.0 + 120.0 * nch, 600.0 + 150.0 * nch
        else:
            lat, instr, trip = 800.0 + 1000.0 * nch, 100.0 + 120.0 * nch, 300.0 + 900.0 * nch
    else:
        mask = 0
        # four 4-slot items per thread are in flight together (one L2 round trip per group of four)
        work, lat, instr, trip = -(-(1 << (log_src - 2)) // 4), 2300.0, 320.0, 2000.0
    # mean over k blocks (2^-logk) and the 1/L of the following inverse transform (2^-log_dst)
    if mask >= 1 << 31:
        mask -= 1 << 32                                    # the task table holds int32 fields
    return TaskSpec(OP_MULFOLD, work, lat, instr, a=src, b=log_src, c=logk, d=dst, e=filt_off, f=mask,
                    h=_Arena.chunk_log2(logk) if logk >= 2 else 0, sexp=logk + log_dst, channel=channel, trip=trip)


def _mulfold2(arena: _Arena, src, log_src: int, logk: int, dst_a, dst_b, filt_off: int) -> TaskSpec:
    """MULFOLD on a packed source (spectrum of u_a + i u_b): one task, both children (csrc: mulfold2_task)."""
    log_dst = log_src - logk
    if logk >= 2:
        mask = arena.chunk_mask(filt_off, logk)
        logcw = _Arena.chunk_log2(logk)
        nch = bin(mask).count('1') << (logcw - 2)
        filt_off = arena.compact(filt_off, logk, mask)
        if nch <= 2 and logcw == 2:                        # four outputs per thread and trip, loads up front
            work, lat, instr, trip = -(-(1 << log_dst) // 4), 1700.0 + 500.0 * nch, 150.0 + 250.0 * nch, 1200.0 + 500.0 * nch
        else:                                              # two outputs per trip, one L2 round trip per chunk
            work, lat, instr, trip = -(-(1 << log_dst) // 2), 900.0 + 1100.0 * nch, 100.0 + 150.0 * nch, 300.0 + 1000.0 * nch
    else:
        mask = 0
        work, lat, instr, trip = max(1, -(-(1 << (log_src - 2)) // (1 << MF2_LOG_ITEMS))), 2600.0 * (1 << (MF2_LOG_ITEMS - 2)) ** 0.5, \
            500.0 * (1 << (MF2_LOG_ITEMS - 2)), 2400.0 * (1 << (MF2_LOG_ITEMS - 2)) ** 0.5     # 2^MF2_LOG_ITEMS 4-slot items per thread and trip
    lat, trip = lat * MF2_SCALE, trip * MF2_SCALE
    if mask >= 1 << 31:
        mask -= 1 << 32
    # mean over k blocks, 1/L of the inverse transform, and the 1/2 of the pair separation
    return TaskSpec(OP_MULFOLD2, work, lat, instr, a=src, b=log_src, c=logk, d=dst_a, e=filt_off, f=mask, g=dst_b,
                    h=_Arena.chunk_log2(logk) if logk >= 2 else 0, sexp=logk + log_dst + 1, trip=trip)


MF2_LOG_ITEMS = int(os.environ.get('TEBSCAT_MF2_LOG_ITEMS', '2'))    # csrc: kMf2Items = 4 items per thread and trip (8 measured: +0.4 % at H, -3 % at J = 4)
MF2_SCALE = float(os.environ.get('TEBSCAT_MF2_SCALE', '1.0'))    # cost-model knob: latency of the packed-source multiplies
AUTO_SCRATCH_GAIN = 0.97          # 'auto': the scratch layout must be modelled at least 3 % faster to be chosen
GSRC_PAIRS = os.environ.get('TEBSCAT_GSRC_PAIRS', '1') != '0'     # partners of a packed pair share one read of a global source
# Cost-model knob: bytes per cycle at which a global-source multiply is charged for its source.  Measured, the
# source costs about 64-100 B/cycle/SM of L2 port time -- but the schedules built AS IF it were free are the fastest
# on the device (tools/ab_u0_scratch.py: 507 k signals/s against 493 k with 64 B/cycle at the headline configuration):
# the list scheduler otherwise spreads those multiplies over thinner steps to "hide" a cost it cannot hide.
GSRC_BYTES_PER_CYCLE = float(os.environ.get('TEBSCAT_GSRC_BPC', 'inf'))


def _gmulfold(arena: _Arena, src_off: int, log_src: int, logk: int, dst, filt_off: int) -> TaskSpec:
    """MULFOLD whose source is a spectrum of 2^log_src bins in GLOBAL memory (csrc: gmulfold_task): the parent is
    too long for one SM, the result (<= 8192 bins) lands in shared memory.  Every trip has its filter and source
    loads in flight together: one trip to L2 / HBM (modelled like the L2-bound trips of _mulfold, a bit longer)."""
    log_dst = log_src - logk
    if logk >= 2:
        mask = arena.chunk_mask(filt_off, logk)
        logcw = _Arena.chunk_log2(logk)
        nch = bin(mask).count('1') << (logcw - 2)
        filt_off = arena.compact(filt_off, logk, mask)
        work, lat, instr, trip = -(-(1 << log_dst) // 4), 900.0 + 1100.0 * nch, 110.0 + 130.0 * nch, 350.0 + 1000.0 * nch
    else:
        mask = 0
        work, lat, instr, trip = -(-(1 << (log_src - 2)) // 4), 2400.0, 330.0, 2100.0
    # the source comes through the SM's L2 port at about 64 bytes per cycle (measured on the headline and the
    # production configuration, tools/ab_u0_scratch.py: +1000 cycles per 64 KB read): a per-trip cost when every
    # thread of the CTA works on it
    src_bytes = 8.0 * ((1 << log_src) if logk < 2 else (1 << log_dst) * 4 * nch)
    per_trip = (src_bytes / GSRC_BYTES_PER_CYCLE) / max(1.0, work / 512.0)
    lat, trip = lat + per_trip, trip + per_trip
    if mask >= 1 << 31:
        mask -= 1 << 32
    return TaskSpec(OP_GMULFOLD, work, lat, instr, a=int(src_off), b=log_src, c=logk, d=dst, e=filt_off, f=mask,
                    h=_Arena.chunk_log2(logk) if logk >= 2 else 0, sexp=logk + log_dst, trip=trip)


def _gmulfold2(arena: _Arena, log_src: int, logk: int, dst_a, dst_b, filt_a: int, filt_b: int) -> TaskSpec:
    """Two filters of one scale on ONE read of the global source (csrc: gmulfold2_task): the partners of a packed
    pair.  Both filters are compacted to the union of their active chunks."""
    log_dst = log_src - logk
    if logk >= 2:
        mask = arena.chunk_mask(filt_a, logk) | arena.chunk_mask(filt_b, logk)
        logcw = _Arena.chunk_log2(logk)
        nch = bin(mask).count('1') << (logcw - 2)
        filt_a, filt_b = arena.compact(filt_a, logk, mask), arena.compact(filt_b, logk, mask)
        if nch == 1 and logcw == 2:        # four outputs per thread and trip, one trip to L2
            work, lat, instr, trip = -(-(1 << log_dst) // 4), 2200.0, 420.0, 1900.0
        elif nch == 2 and logcw == 2:      # two outputs per trip, both chunks in flight
            work, lat, instr, trip = -(-(1 << log_dst) // 2), 2200.0, 420.0, 1900.0
        else:
            work, lat, instr, trip = -(-(1 << log_dst) // 2), 900.0 + 1100.0 * nch, 120.0 + 180.0 * nch, 350.0 + 1000.0 * nch
    else:
        mask, logcw = 0, 0
        work, lat, instr, trip = -(-(1 << (log_src - 2)) // 4), 2600.0, 560.0, 2300.0    # four items per thread and trip
    src_bytes = 8.0 * ((1 << log_src) if logk < 2 else (1 << log_dst) * 4 * nch)
    per_trip = (src_bytes / GSRC_BYTES_PER_CYCLE) / max(1.0, work / 512.0)
    lat, trip = lat + per_trip, trip + per_trip
    if mask >= 1 << 31:
        mask -= 1 << 32
    return TaskSpec(OP_GMULFOLD2, work, lat, instr, a=dst_b, b=log_src, c=logk, d=dst_a, e=filt_a, f=mask, g=filt_b,
                    h=logcw, sexp=logk + log_dst, trip=trip)


# ------------------------------------------------------------------------------------
# the plan
# ------------------------------------------------------------------------------------
@dataclass
class ScatPlan:
    J: int
    Q: int
    T: int
    N: int
    max_order: int
    geo: fbk.Geometry
    bank: fbk.FilterBank
    keys: List[Tuple[int, ...]]
    n_out: int
    arena: np.ndarray              # float32
    tasks: np.ndarray              # int32 [n_tasks, 12]
    steps: np.ndarray              # int32 [n_steps, 2]
    chan: np.ndarray               # int32 channel table of the batched stores
    smem_complex: int              # PHYSICAL complex slots (logical slots + 1 pad per 16)
    n_threads: int = N_THREADS
    stats: Dict[str, float] = field(default_factory=dict)
    scratch_complex: int = 0       # > 0: per-CTA global scratch (complex elements) the schedule parks U0 in

    @property
    def n_paths(self) -> int:
        return len(self.keys)


class _Allocator:
    """First-fit allocator over [0, capacity) in 16-slot granules."""

    def __init__(self, capacity: int):
        self.capacity = capacity
        self.free: List[Tuple[int, int]] = [(0, capacity)]
        self.high_water = 0

    def alloc(self, size: int) -> int:
        size = _round16(size)
        for i, (o, s) in enumerate(self.free):
            if s >= size:
                if s == size:
                    self.free.pop(i)
                else:
                    self.free[i] = (o + size, s - size)
                self.high_water = max(self.high_water, o + size)
                return o
        return -1

    def release(self, off: int, size: int) -> None:
        size = _round16(size)
        self.free.append((off, size))
        self.free.sort()
        merged: List[Tuple[int, int]] = []
        for o, s in self.free:
            if merged and merged[-1][0] + merged[-1][1] == o:
                merged[-1] = (merged[-1][0], merged[-1][1] + s)
            else:
                merged.append((o, s))
        self.free = merged


def build_chains(bank: fbk.FilterBank, geo: fbk.Geometry, T: int, max_order: int, arena: _Arena,
                 batch_slots: int = BATCH_SLOTS, oversampling: int = 0, child_slots: Optional[int] = None,
                 global_u0: bool = False):
    """The cascade as a forest of chains of batched tasks, in the reference's channel order.

    global_u0 (large-support level, padded lengths above 2^13): the signal's spectrum U0 lives in GLOBAL memory
    -- no root chain, no order-0 leaf -- and only the first-order filters whose subsampled length fits one SM
    (<= 8192 samples) are scheduled, their psi1 multiply + periodisation reading U0 through OP_GMULFOLD."""
    n = geo.J_pad
    log2_T = int(math.floor(math.log2(T)))
    os_ = int(oversampling)
    if child_slots is None:
        child_slots = batch_slots // 2
    kf = max(log2_T - os_, 0)                             # final subsampling (:285, :322, :358)
    lf = n - kf                                           # log2 of the final (output-rate) length
    if lf < 0:
        raise ValueError('T is larger than the padded support')
    if lf < 1:
        raise NotImplementedError('output-rate length below 2 samples is not supported')
    i0, i1 = geo.ind_start[kf], geo.ind_end[kf]
    n_out = i1 - i0

    phi_off = [arena.add(a) for a in bank.phi.levels]
    psi1_off = [arena.add(p.levels[0]) for p in bank.psi1]
    psi2_off = [[arena.add(a) for a in p.levels] for p in bank.psi2]

    keys: List[Tuple[int, ...]] = [()] + [(i,) for i in range(len(bank.psi1))]
    if max_order == 2:
        for n1, p1 in enumerate(bank.psi1):
            for n2, p2 in enumerate(bank.psi2):
                if p2.j > p1.j:
                    keys.append((n1, n2))
    channel = {k: c for c, k in enumerate(keys)}
    chains: List[Chain] = []

    def leaf(src, log_src: int, level: int, key, key_im=None) -> TaskSpec:
        """phi[level] multiply + periodise down to 2^lf into a pool slot (core :287-289).  phi is real and
        even, so the leaf of a PACKED spectrum U_a + i U_b is the packed pair of leaves: after the inverse
        transform the real part is channel `key`, the imaginary part channel `key_im`."""
        ch = (channel[key], channel[key_im] if key_im is not None else -1)
        return _mulfold(arena, src, log_src, log_src - lf, LEAF, phi_off[level], ch)

    def pack_stage(stages) -> int:
        for si, st in enumerate(stages):
            if any(t.op == OP_FFT and (t.e & FFT_PACK) for t in st):
                return si
        return -1

    def paired(items, pack: bool):
        """Adjacent filters of the same scale j share one forward transform: they carry similar energy
        and have the same children.  Pairs first, singles last (the packed pass wants partners in front)."""
        if not pack:
            return [(x, None) for x in items]
        pairs, singles, i = [], [], 0
        while i < len(items):
            if i + 1 < len(items) and bank.psi1[items[i]].j == bank.psi1[items[i + 1]].j:
                pairs.append((items[i], items[i + 1]))
                i += 2
            else:
                singles.append((items[i], None))
                i += 1
        return pairs + singles

    packing = pack_enabled()

    # root: pad + forward transform of the signal (core/scattering1d.py:278-280)
    u0 = Buf(1 << n, 'U0')
    scratch = global_u0 == 'scratch'
    split = global_u0 == 'split'
    if split:
        # U0 stays in shared memory for the order-0 leaf and the SUBSAMPLED first-order filters, is copied to the
        # per-CTA global scratch on the side, and is dropped from shared memory before the full-length filters
        # (k1 = 0: the pairs whose working set fills the SM) start: those read it through OP_GMULFOLD.
        root = Chain('root', [[TaskSpec(OP_LOAD, 1 << n, 300.0, 16.0, a=(u0, 0))]] +
                     _merge_local_passes(_fft_stages((u0, 0), n, 1, 'fwd')), owns=[u0], depth=0)
        chains.append(root)
        chains.append(Chain('S0', [[leaf((u0, 0), n, 0, ())]], after=[root], reads=[u0], depth=1))   # :285-292
        park = Chain('park', [[TaskSpec(OP_STOREC, 1 << n, 300.0, 12.0, a=(u0, 0), b=1 << n)]], after=[root], reads=[u0],
                     depth=1)
        chains.append(park)
    elif scratch:
        # U0 is computed here but PARKED in a per-CTA global scratch (L2-resident) right after its transform: its
        # 2^n slots are free for the rest of the signal, and every consumer reads it through OP_GMULFOLD
        root = Chain('root', [[TaskSpec(OP_LOAD, 1 << n, 300.0, 16.0, a=(u0, 0))]] +
                     _merge_local_passes(_fft_stages((u0, 0), n, 1, 'fwd')) +
                     [[TaskSpec(OP_STOREC, 1 << n, 300.0, 12.0, a=(u0, 0), b=1 << n)]],
                     owns=[u0], frees_own_at_end=True, depth=0)
        chains.append(root)
        s0 = _gmulfold(arena, 0, n, n - lf, LEAF, phi_off[0])
        s0.channel = (channel[()], -1)
        chains.append(Chain('S0', [[s0]], after=[root], depth=1))                                    # :285-292
    elif not global_u0:
        root = Chain('root', [[TaskSpec(OP_LOAD, 1 << n, 300.0, 16.0, a=(u0, 0))]] +
                     _merge_local_passes(_fft_stages((u0, 0), n, 1, 'fwd')), owns=[u0], depth=0)
        chains.append(root)
        chains.append(Chain('S0', [[leaf((u0, 0), n, 0, ())]], after=[root], reads=[u0], depth=1))   # :285-292
    from_u0 = [root] if (scratch or split or not global_u0) else []
    reads_u0 = [] if (global_u0 and not split) else [u0]
    subsampled_first: List[Chain] = []                 # split mode: the chains that still read U0 from shared memory

    def first_mulfold(k1: int, dst, filt_off: int) -> TaskSpec:
        if split:
            return _gmulfold(arena, 0, n, k1, dst, filt_off) if k1 == 0 else _mulfold(arena, (u0, 0), n, k1, dst, filt_off)
        return _gmulfold(arena, 0, n, k1, dst, filt_off) if global_u0 else _mulfold(arena, (u0, 0), n, k1, dst, filt_off)

    # first order, batched by subsampling k1 (:300-318)
    groups: Dict[int, List[int]] = {}
    for n1, p1 in enumerate(bank.psi1):
        k1 = max(min(p1.j - os_, log2_T - os_), 0)                     # :304
        if not p1.xi < 0.5 / (2 ** k1):
            raise AssertionError('psi1 aliasing assertion of the reference violated')
        if global_u0 and n - k1 > LOG2_NP_MAX:
            continue                                                   # stays on the level's own kernels
        groups.setdefault(k1, []).append(n1)
    for k1 in sorted(groups, reverse=split):           # (split mode: the subsampled filters are built -- and start -- first)
        l1 = n - k1
        pack1 = packing and l1 >= 4
        # entries (a, b): filters a and b share the forward transform; one batch = `per_batch` entries
        entries_all = paired(groups[k1], pack1)
        per_batch = max(1, (batch_slots // 2 if pack1 else batch_slots) >> l1)
        for s in range(0, len(entries_all), per_batch):
            batch = sorted(entries_all[s:s + per_batch], key=lambda e: e[1] is None)
            lo = len(batch)
            hi = sum(1 for _, b in batch if b is not None)             # pairs come first
            x1 = Buf((lo + hi) << l1, 'U1[k1=%d:%d]' % (k1, batch[0][0]))
            if global_u0 and not split and GSRC_PAIRS:
                # global source: the partners of a pair share ONE read of U0
                mf = [_gmulfold2(arena, n, k1, (x1, i << l1), (x1, (lo + i) << l1), psi1_off[a], psi1_off[b])
                      if b is not None else first_mulfold(k1, (x1, i << l1), psi1_off[a]) for i, (a, b) in enumerate(batch)]
            else:
                mf = [first_mulfold(k1, (x1, i << l1), psi1_off[a]) for i, (a, _) in enumerate(batch)]
                mf += [first_mulfold(k1, (x1, (lo + i) << l1), psi1_off[b])
                       for i, (_, b) in enumerate(batch) if b is not None]
            st = [mf] + _merge_local_passes(
                _fuse_first_inverse_pass(mf, _fft_stages((x1, 0), l1, lo, 'pair', hi=hi), l1))             # :307-318
            if split and k1 == 0:
                c1 = Chain(x1.name, st, after=[root], waits=[park] + list(subsampled_first), owns=[x1], depth=1)
            else:
                c1 = Chain(x1.name, st, after=list(from_u0), reads=list(reads_u0), owns=[x1], depth=1)
                if split:
                    subsampled_first.append(c1)
            if hi:
                c1.shrink.append((pack_stage(st), x1, lo << l1))
            chains.append(c1)
            # low-pass leaves of the batch (:320-327)
            chains.append(Chain('S1' + x1.name,
                                [[leaf((x1, i << l1), l1, k1, (a,), (b,) if b is not None else None)
                                  for i, (a, b) in enumerate(batch)]],
                                after=[c1], reads=[x1], depth=2))
            if max_order != 2:
                continue
            # second order, grouped by child length (:337-364); the children of a packed parent are a pair again
            kids: Dict[int, List[Tuple[int, int, Optional[int], int]]] = {}
            for i, (a, b) in enumerate(batch):
                p1 = bank.psi1[a]
                for n2, p2 in enumerate(bank.psi2):
                    if p2.j > p1.j:
                        if not p2.xi < p1.xi:
                            raise AssertionError('psi2 ordering assertion of the reference violated')
                        k2 = max(min(p2.j - k1 - os_, log2_T - k1 - os_), 0)   # :344-345
                        kids.setdefault(k2, []).append((i, a, b, n2))
            for k2 in sorted(kids):
                l2 = l1 - k2
                pack2 = packing and l2 >= 4
                fam = sorted(kids[k2], key=lambda e: e[2] is None)            # pairs first (stable)
                per = max(1, child_slots >> l2)
                for s2 in range(0, len(fam), per):
                    sub = fam[s2:s2 + per]
                    lo2 = len(sub)
                    hi2 = sum(1 for e in sub if e[2] is not None)
                    x2 = Buf((lo2 + hi2) << l2, 'U2[%d,k2=%d]' % (batch[0][0], k2))
                    mf = []
                    for c, (i, a, b, n2) in enumerate(sub):                                 # :347-348
                        if b is not None:
                            mf.append(_mulfold2(arena, (x1, i << l1), l1, k2, (x2, c << l2), (x2, (lo2 + c) << l2),
                                                psi2_off[n2][k1]))
                        else:
                            mf.append(_mulfold(arena, (x1, i << l1), l1, k2, (x2, c << l2), psi2_off[n2][k1]))
                    if pack2:
                        st = [mf] + _merge_local_passes(
                            _fuse_first_inverse_pass(mf, _fft_stages((x2, 0), l2, lo2, 'pair', hi=hi2), l2))   # :350-355
                        lv = [leaf((x2, c << l2), l2, k1 + k2, (a, n2), (b, n2) if b is not None else None)
                              for c, (i, a, b, n2) in enumerate(sub)]
                    else:
                        st = [mf] + _merge_local_passes(
                            _fuse_first_inverse_pass(mf, _fft_stages((x2, 0), l2, lo2 + hi2, 'pair'), l2))
                        lv = [leaf((x2, c << l2), l2, k1 + k2, (a, n2)) for c, (i, a, b, n2) in enumerate(sub)]
                        lv += [leaf((x2, (lo2 + c) << l2), l2, k1 + k2, (b, n2))
                               for c, (i, a, b, n2) in enumerate(sub) if b is not None]
                    c2 = Chain(x2.name, st, after=[c1], reads=[x1], owns=[x2], depth=2)
                    if pack2 and hi2:
                        c2.shrink.append((pack_stage(st), x2, lo2 << l2))
                    chains.append(c2)
                    chains.append(Chain('S2' + x2.name, [lv], after=[c2], reads=[x2], depth=3))   # :358-364
    return chains, keys, n_out, lf, i0


# ------------------------------------------------------------------------------------
# list scheduler
# ------------------------------------------------------------------------------------
def _want_threads(work: int, tpi: int = 1) -> int:
    return min(N_THREADS, max(32, (work * tpi + 31) & ~31))


def _step_time(items: List[Tuple[TaskSpec, int]]) -> float:
    """Cost model of one step (cycles): fixed overhead + the longest dependent latency + the
    issue cycles all warps of the step add on the four schedulers."""
    lat = 0.0
    issue = 0.0
    for t, nt in items:
        trips = math.ceil(t.work * t.tpi / nt)
        lat = max(lat, t.lat + (trips - 1) * (t.trip if t.trip >= 0 else t.instr))
        issue += (nt / 128.0) * trips * t.instr
    return STEP_OVERHEAD + lat + issue


def _split_threads(tasks: List[TaskSpec]) -> Optional[List[int]]:
    """Thread counts (multiples of 32, sum <= 512) for the tasks of one step: start from one
    warp each and keep giving warps to the task with the most trips left."""
    n = len(tasks)
    if 32 * n > N_THREADS:
        return None
    nts = [32] * n
    left = N_THREADS - 32 * n
    while left > 0:
        times = [(math.ceil(t.work * t.tpi / nt) - 1) * (t.trip if t.trip >= 0 else t.instr) + t.lat
                 for t, nt in zip(tasks, nts)]
        order = sorted(range(n), key=lambda i: -times[i])
        grew = False
        for i in order:
            want = _want_threads(tasks[i].work, tasks[i].tpi)
            if nts[i] >= want:
                continue
            it = math.ceil(tasks[i].work * tasks[i].tpi / nts[i])
            need = nts[i] + 32
            while need < want and math.ceil(tasks[i].work * tasks[i].tpi / need) >= it:
                need += 32
            if need - nts[i] > left:
                continue
            left -= need - nts[i]
            nts[i] = need
            grew = True
            break
        if not grew:
            break
    return nts


class _LeafPool:
    """Ping-pong pool of 2^lf-slot leaf spectra; a full half is flushed by one batched
    inverse transform + unpad + store (the flush chain)."""

    def __init__(self, lf: int, i0: int, n_out: int, pool_slots: int = POOL_SLOTS):
        self.lf, self.i0, self.n_out = lf, i0, n_out
        self.per_half = max(1, pool_slots >> lf)
        self.bufs = [Buf(self.per_half << lf, 'pool0'), Buf(self.per_half << lf, 'pool1')]
        self.fill = [0, 0]
        self.channels: List[List[int]] = [[], []]
        self.busy = [False, False]                      # half is being flushed
        self.cur = 0
        self.chan_table: List[int] = []

    def room(self) -> int:
        return 0 if self.busy[self.cur] else self.per_half - self.fill[self.cur]

    def take(self, channel: int) -> int:
        h = self.cur
        slot = self.fill[h]
        self.fill[h] += 1
        self.channels[h].append(channel if isinstance(channel, tuple) else (channel, -1))
        return self.bufs[h].off + (slot << self.lf)

    def flush_chain(self, h: int) -> Chain:
        cnt = self.fill[h]
        table_off = len(self.chan_table) // 2          # in slots: the table holds (real, imaginary) channel pairs
        for pair in self.channels[h]:
            self.chan_table += [pair[0], pair[1]]
        ref = (self.bufs[h], 0)
        st = _merge_local_passes(_fft_stages(ref, self.lf, cnt, 'inv'))
        st.append([TaskSpec(OP_STOREB, cnt * self.n_out, 150.0, 14.0, a=ref, b=cnt, c=self.i0, d=self.n_out,
                            e=table_off, f=self.lf)])
        ch = Chain('flush%d@%d' % (h, table_off), st, depth=9, pool_half=h)
        self.busy[h] = True
        self.fill[h] = 0
        self.channels[h] = []
        return ch

    def rotate(self):
        """Make `cur` point at a half that can take leaves, if there is one."""
        if self.room() == 0:
            o = 1 - self.cur
            if not self.busy[o] and self.fill[o] < self.per_half:
                self.cur = o


# Priority of a chain in the list scheduler = depth * PRIO_DEPTH + weight of its subtree.  Depth first (1e12) finishes
# what is started before it opens something new; by weight only (0) lets the heavy full-length pairs start while the
# tails of the subsampled groups still run.  build_plan schedules BOTH ways and keeps the one the cost model prefers
# (headline configuration: 100 steps / 570 k modelled cycles against 107 / 586 k; measured 529 000 against 513 900
# signals/s, profiles/r02n_ab_priority.txt); TEBSCAT_PRIO_DEPTH pins one.
PRIO_DEPTH = os.environ.get('TEBSCAT_PRIO_DEPTH')
RESERVE_MODES = os.environ.get('TEBSCAT_RESERVE', 'sum,peak').split(',')


def schedule_chains(chains: List[Chain], capacity: int, lf: int, i0: int, n_out: int,
                    max_parallel: int = 64, pack_gain: float = 0.97, pool_slots: int = POOL_SLOTS,
                    open_demand: float = 2.0, depth_weight: float = 1e12, reserve: str = 'sum',
                    defer_thin: bool = False):
    """Greedy list scheduling of chains into steps (see module docstring)."""
    children: Dict[int, List[Chain]] = {}
    for ch in chains:
        for a in ch.after:
            children.setdefault(id(a), []).append(ch)
    for ch in chains:
        for b in ch.reads:
            b.readers_left += 1

    own_size = {id(c): sum(_round16(b.size) for b in c.owns) for c in chains}
    weight: Dict[int, float] = {}
    need: Dict[int, int] = {}

    def visit(c: Chain):
        w = sum(t.work * t.instr for st in c.stages for t in st)
        nd = 0
        for k in children.get(id(c), []):
            if id(k) not in weight:
                visit(k)
            w += weight[id(k)]
            nd += own_size[id(k)] + need[id(k)]
        if reserve == 'peak':
            # what the subtree needs BEYOND the chain's own buffers if its children run one after another and the
            # chain's own buffer has shrunk (the partner half of a packed batch is dead after the packed pass):
            # a weaker guarantee than the sum over all children -- a subtree may start while another one's tail
            # still runs -- so a schedule built this way can dead-end; the caller then falls back to 'sum'
            own = own_size[id(c)]
            after = own
            for _, buf, new_size in c.shrink:
                after -= _round16(buf.size) - _round16(new_size)
            kid_peak = max((own_size[id(k)] + need[id(k)] for k in children.get(id(c), [])), default=0)
            nd = max(own, after + kid_peak) - own
        weight[id(c)] = w
        need[id(c)] = nd if c.depth > 0 else 0          # the root does not reserve for the whole tree

    for c in chains:
        if id(c) not in weight:
            visit(c)
    parent_of = {id(c): (c.after[0] if c.after else None) for c in chains}
    for c in chains:
        c.priority = c.depth * depth_weight + weight[id(c)]
        c.stage = 0
        c.issued = [False] * len(c.stages[0])
        c.done_step = -1

    alloc = _Allocator(capacity)
    pool = _LeafPool(lf, i0, n_out, pool_slots)
    has_leaves = any(t.d is LEAF for c in chains for st in c.stages for t in st)
    if has_leaves:
        for b in pool.bufs:
            b.off = alloc.alloc(b.size)
            assert b.off >= 0
    free_slots = capacity - (sum(_round16(b.size) for b in pool.bufs) if has_leaves else 0)
    reserve_left: Dict[int, int] = {}
    pending = list(chains)
    active: List[Chain] = []
    steps: List[List[List[int]]] = []
    step_idx = 0
    est_time = 0.0
    est_issue = 0.0
    by_id = {id(c): c for c in chains}

    def release(buf: Buf):
        nonlocal free_slots
        alloc.release(buf.off, buf.size)
        free_slots += _round16(buf.size)
        buf.off = -2

    def release_if_dead(buf: Buf):
        if buf.readers_left == 0 and buf.producer_done and buf.off >= 0:
            release(buf)

    def ancestors(c: Chain):
        a = parent_of.get(id(c))
        while a is not None:
            yield a
            a = parent_of.get(id(a))

    def subtree_done(ch: Chain) -> bool:
        return ch.done_step >= 0 and all(subtree_done(k) for k in children.get(id(ch), []))

    def try_start(c: Chain, force: bool) -> bool:
        nonlocal free_slots
        mine = own_size.get(id(c), 0)
        reserved_others = sum(reserve_left.values()) - sum(reserve_left.get(id(a), 0) for a in ancestors(c))
        if not force and free_slots - reserved_others < mine + need.get(id(c), 0):
            return False
        offs = []
        for b in c.owns:
            o = alloc.alloc(b.size)
            if o < 0:
                for bb, oo in zip(c.owns, offs):
                    alloc.release(oo, bb.size)
                return False
            offs.append(o)
        for b, o in zip(c.owns, offs):
            b.off = o
        free_slots -= mine
        for a in ancestors(c):
            if id(a) in reserve_left:
                reserve_left[id(a)] = max(0, reserve_left[id(a)] - mine)
        if need.get(id(c), 0) > 0:
            reserve_left[id(c)] = need[id(c)]
        return True

    def resolve(ref):
        if isinstance(ref, tuple):
            assert ref[0].off >= 0, 'task touches a released buffer (%s)' % ref[0].name
            return ref[0].off + ref[1]
        return int(ref)

    def start_flush(h: int):
        ch = pool.flush_chain(h)
        ch.priority = 9e12
        ch.stage = 0
        ch.issued = [False] * len(ch.stages[0])
        by_id[id(ch)] = ch
        active.append(ch)
        pool.rotate()

    guard = 0
    while pending or active or pool.fill[0] or pool.fill[1]:
        guard += 1
        if guard > 100000:
            raise RuntimeError('scheduler did not terminate')
        if not pending and not active:
            for h in (0, 1):
                if pool.fill[h] and not pool.busy[h]:
                    start_flush(h)
        # ---- start chains whose inputs are complete ------------------------------------
        startable = [c for c in pending if all(a.done_step >= 0 and a.done_step < step_idx for a in c.after + c.waits)]
        startable.sort(key=lambda c: -c.priority)
        demand = sum(sum(_want_threads(t.work, t.tpi) for t, done in zip(c.stages[c.stage], c.issued) if not done)
                     for c in active)
        for c in startable:
            if len(active) >= max_parallel:
                break
            if demand >= open_demand * N_THREADS and c.depth <= 1:
                continue                          # enough queued work; do not open new subtrees
            if try_start(c, force=False):
                pending.remove(c)
                active.append(c)
                demand += sum(_want_threads(t.work, t.tpi) for t in c.stages[0])
        if not active:
            for c in startable:                   # progress guarantee
                if try_start(c, force=True):
                    pending.remove(c)
                    active.append(c)
                    break
        if not active:
            raise RuntimeError('schedule deadlock: %d chains cannot be placed in %d slots'
                               % (len(pending), capacity))

        # ---- choose the tasks of this step ------------------------------------------------
        active.sort(key=lambda c: -c.priority)
        cands: List[Tuple[Chain, int, TaskSpec]] = []
        for c in active:
            for ti, t in enumerate(c.stages[c.stage]):
                if not c.issued[ti]:
                    cands.append((c, ti, t))
        def pick(order):
            chosen_: List[Tuple[Chain, int, TaskSpec]] = []
            nts_: List[int] = []
            cur = 0.0
            room = pool.room()
            skipped = []
            for cand in order:
                t = cand[2]
                if t.d is LEAF and room <= 0:
                    continue
                trial = chosen_ + [cand]
                split = _split_threads([k[2] for k in trial])
                if split is None:
                    break
                tt = _step_time(list(zip([k[2] for k in trial], split)))
                alone = _step_time([(t, _want_threads(t.work, t.tpi))])
                if chosen_ and tt > pack_gain * (cur + alone):
                    skipped.append(cand)
                    continue
                chosen_, nts_, cur = trial, split, tt
                if t.d is LEAF:
                    room -= 1
            return chosen_, nts_, cur, skipped

        chosen, nts, cur_time, skipped = pick(cands)
        if defer_thin and chosen and sum(nts) <= N_THREADS // 4:
            # A thin step (a few warps of leftovers) while a task that wants most of the CTA had to be skipped: run
            # the wide task now, the leftovers ride along with a later step that has room for them.
            wide = [k for k in skipped if _want_threads(k[2].work, k[2].tpi) >= 3 * N_THREADS // 4]
            if wide:
                again = pick(wide + [k for k in cands if not any(k is w for w in wide)])
                if again[0] and sum(again[1]) > sum(nts):
                    chosen, nts, cur_time, skipped = again
        if not chosen:
            raise RuntimeError('scheduler stalled on the leaf pool')
        this_step: List[List[int]] = []
        used = 0
        for (c, ti, t), nt in zip(chosen, nts):
            dst = pool.take(t.channel) if t.d is LEAF else resolve(t.d)
            this_step.append([t.op | (t.sexp << 8), used, nt, resolve(t.a), t.b, t.c, dst, t.e, resolve(t.f), resolve(t.g),
                              t.h, t.pad])
            used += nt
            est_issue += (nt // 32) * math.ceil(t.work * t.tpi / nt) * t.instr
            c.issued[ti] = True
        est_time += cur_time
        steps.append(this_step)

        # ---- advance chains --------------------------------------------------------------------
        for c in list({id(k[0]): k[0] for k in chosen}.values()):
            if not all(c.issued):
                continue
            if c.stage == 0:
                for b in c.reads:
                    b.readers_left -= 1
                    release_if_dead(b)
            c.stage += 1
            for si, buf, new_size in c.shrink:
                if si == c.stage - 1 and buf.off >= 0 and _round16(new_size) < _round16(buf.size):
                    # the partner blocks of a packed pass are dead once the pass has run: free the tail
                    alloc.release(buf.off + _round16(new_size), _round16(buf.size) - _round16(new_size))
                    free_slots += _round16(buf.size) - _round16(new_size)
                    buf.size = new_size
            if c.stage == len(c.stages):
                c.done_step = step_idx
                active.remove(c)
                for b in c.owns:
                    b.producer_done = True
                    if c.frees_own_at_end:
                        release(b)
                    else:
                        release_if_dead(b)
                if c.pool_half >= 0:
                    pool.busy[c.pool_half] = False
                    pool.rotate()
            else:
                c.issued = [False] * len(c.stages[c.stage])
        for cid in list(reserve_left.keys()):
            if subtree_done(by_id[cid]):
                reserve_left.pop(cid)
        if pool.fill[pool.cur] == pool.per_half and not pool.busy[pool.cur]:
            start_flush(pool.cur)
        step_idx += 1
    sched_stats = dict(est_cycles=est_time, est_issue=est_issue)
    return steps, alloc.high_water, pool.chan_table, sched_stats


# ------------------------------------------------------------------------------------
# barrier relaxation (post-pass over the emitted schedule)
# ------------------------------------------------------------------------------------
def task_accesses(t, log2_Np):
    """(slots, warps, is_write) arrays of one task row: a superset of what each warp touches."""
    op = t[0] & 0xff
    t0, nt = int(t[1]), int(t[2])
    a, b, c, d, e, f, g, h = (int(v) for v in t[3:11])
    out = []
    def add(slots, lt, write):
        slots = np.asarray(slots); lt = np.asarray(lt)
        if slots.ndim == 2:
            lt = np.broadcast_to(lt[:, None], slots.shape)
        w = (t0 + lt) >> 5
        out.append((slots.reshape(-1).astype(np.int64), w.reshape(-1).astype(np.int64), write))
    if op == OP_NOP:
        pass
    elif op == OP_LOAD:
        i = np.arange(1 << log2_Np)
        add(a + i, i % nt, True)
    elif op == OP_FFT:
        logB, logR, flags = c, d, e
        if logR <= 3 and logB == logR and not (flags & FFT_MOD) and ((b << logR) & 15) == 0:
            gidx = np.arange((b << logR) >> 4)
            s = a + 16 * gidx[:, None] + np.arange(16)[None, :]
            add(s, gidx % nt, False); add(s, gidx % nt, True)
        else:
            u = np.arange(b)
            logs = logB - logR
            i0 = u & ((1 << logs) - 1); blk = u >> logs
            s = (a + (blk << logB) + i0)[:, None] + (np.arange(1 << logR) << logs)[None, :]
            add(s, u % nt, False); add(s, u % nt, True)
            if flags & FFT_PACK:
                sel = blk < g
                add(s[sel] - a + f, (u % nt)[sel], False)
        more = (h & 0xffffff) | (((int(t[11]) >> 4) & 0xfff) << 24)
        while more & 0xfff:                             # chained passes touch the same slots, warp by warp
            sub = t.copy()
            sub[5], sub[6], sub[7] = more & 15, (more >> 4) & 7, (more >> 7) & 15
            sub[4] = (b << logR) >> int(sub[6])
            sub[10] = 0; sub[11] = 0
            out += task_accesses(sub, log2_Np)
            more >>= 12
    elif op == OP_MULFOLD:
        log_src, logk = b, c
        if logk >= 2:
            m = np.arange(1 << (log_src - logk))
            add(a + (m[:, None] << logk) + np.arange(1 << logk)[None, :], m % nt, False)
            add(d + m, m % nt, True)
        else:
            it = np.arange(1 << (log_src - 2))
            add(a + 4 * it[:, None] + np.arange(4)[None, :], it % nt, False)
            if logk == 0:
                add(d + 4 * it[:, None] + np.arange(4)[None, :], it % nt, True)
            else:
                add(d + 2 * it[:, None] + np.arange(2)[None, :], it % nt, True)
    elif op == OP_GMULFOLD:                             # the source is global: only the destination slots count
        log_src, logk = b, c
        if logk >= 2:
            m = np.arange(1 << (log_src - logk))
            add(d + m, m % nt, True)
        else:
            it = np.arange(1 << (log_src - 2))
            w_ = 4 >> logk
            add(d + w_ * it[:, None] + np.arange(w_)[None, :], it % nt, True)
    elif op == OP_GMULFOLD2:                            # two destinations: d (filter A) and a (filter B)
        log_src, logk = b, c
        if logk >= 2:
            m = np.arange(1 << (log_src - logk))
            add(d + m, m % nt, True); add(a + m, m % nt, True)
        else:
            it = np.arange(1 << (log_src - 2))
            w_ = 4 >> logk
            add(d + w_ * it[:, None] + np.arange(w_)[None, :], it % nt, True)
            add(a + w_ * it[:, None] + np.arange(w_)[None, :], it % nt, True)
    elif op == OP_LOADPAIR:
        s_ = a + np.arange(1 << log2_Np)
        for w in range(nt // 32):
            add(s_, np.full(s_.shape, 32 * w), True)
    elif op == OP_MULFOLD2:
        log_src, logk = b, c
        src = a + np.arange(1 << log_src)
        for w in range(nt // 32):                       # mirrored reads: any slot of the source
            add(src, np.full(src.shape, 32 * w), False)
        if logk >= 2:
            m = np.arange(1 << (log_src - logk))
            add(d + m, m % nt, True); add(g + m, m % nt, True)
        else:
            it = np.arange(1 << (log_src - 2))
            w_ = 4 >> logk
            add(d + w_ * it[:, None] + np.arange(w_)[None, :], it % nt, True)
            add(g + w_ * it[:, None] + np.arange(w_)[None, :], it % nt, True)
    elif op == OP_STOREB:
        s = a + np.arange(b << f)
        for w in range(nt // 32):
            add(s, np.full(s.shape, 32 * w), False)
    elif op == OP_STOREZ:
        i = np.arange(d)
        add(a + c + i, i % nt, False)
    elif op == OP_STOREU:
        i = np.arange(d)
        add(a + c + i, i % nt, False)
    elif op == OP_LOADC:
        i = np.arange(b)
        add(a + i, i % nt, True)
    elif op == OP_STOREC:
        i = np.arange(b)
        add(a + i, i % nt, False)
    elif op == OP_TINY:
        u = np.arange(b)
        s = a + (u[:, None] << c) + np.arange(1 << c)[None, :]
        add(s, u % nt, False); add(s, u % nt, True)
    else:
        raise ValueError(op)
    return out

def elide_barriers(tasks, steps, capacity, log2_Np):
    """keep[s] = the barrier after step s must be a CTA barrier.  Between two CTA barriers the warps run
    unsynchronised (warp-level fences only), so a step may join the current region only if none of its
    accesses touches a slot that ANOTHER warp has written -- or, for writes, read -- since the region began."""
    n_steps = steps.shape[0]
    keep = np.ones(n_steps, bool)            # keep[s]: CTA barrier after step s
    writer = np.full(capacity + 16, -1, np.int64)       # -1 none, else warp
    readers = np.zeros(capacity + 16, np.int64)         # bitmask of warps
    def record(acc):
        for s, w, wr in acc:
            if wr:
                writer[s] = w
            else:
                np.bitwise_or.at(readers, s, 1 << w)
    prev = None
    for st in range(n_steps):
        acc = []
        for ti in range(steps[st, 0], steps[st, 1]):
            acc += task_accesses(tasks[ti], log2_Np)
        conflict = False
        for s, w, wr in acc:
            lw = writer[s]
            if np.any((lw >= 0) & (lw != w)):
                conflict = True; break
            if wr and np.any(readers[s] & ~(1 << w)):
                conflict = True; break
        if st > 0:
            keep[st - 1] = conflict
        if conflict or st == 0:
            writer[:] = -1; readers[:] = 0
        record(acc)
    keep[n_steps - 1] = True
    return keep


def build_chains_unaveraged(bank: fbk.FilterBank, geo: fbk.Geometry, T: int, max_order: int, arena: _Arena,
                            batch_slots: int = BATCH_SLOTS, oversampling: int = 0):
    """average=False (core/scattering1d.py:293-294, :329-330, :366-367): no phi low-pass; every path is the
    unpadded modulus at its own rate.  Returns the chains and the segments [(key, offset, length)] of the
    output row (order 0 is the input itself and is not produced here)."""
    n = geo.J_pad
    log2_T = int(math.floor(math.log2(T)))
    os_ = int(oversampling)
    keys: List[Tuple[int, ...]] = [(i,) for i in range(len(bank.psi1))]
    if max_order == 2:
        for n1, p1 in enumerate(bank.psi1):
            for n2, p2 in enumerate(bank.psi2):
                if p2.j > p1.j:
                    keys.append((n1, n2))
    seg_len: Dict[Tuple[int, ...], Tuple[int, int]] = {}                 # key -> (first index, length)
    k_of: Dict[Tuple[int, ...], int] = {}
    for n1, p1 in enumerate(bank.psi1):
        k1 = max(min(p1.j - os_, log2_T - os_), 0)
        k_of[(n1,)] = k1
        if max_order == 2:
            for n2, p2 in enumerate(bank.psi2):
                if p2.j > p1.j:
                    k_of[(n1, n2)] = k1 + max(min(p2.j - k1 - os_, log2_T - k1 - os_), 0)
    offsets: Dict[Tuple[int, ...], int] = {}
    total = 0
    for key in keys:
        k = k_of[key]
        i0, i1 = geo.ind_start[k], geo.ind_end[k]
        seg_len[key] = (i0, i1 - i0)
        offsets[key] = total
        total += i1 - i0

    def store(ref, key) -> TaskSpec:
        i0, ln = seg_len[key]
        return TaskSpec(OP_STOREU, ln, 200.0, 12.0, a=ref, c=i0, d=ln, e=offsets[key])

    psi1_off = [arena.add(p.levels[0]) for p in bank.psi1]
    psi2_off = [[arena.add(a) for a in p.levels] for p in bank.psi2]
    chains: List[Chain] = []
    u0 = Buf(1 << n, 'U0')
    root = Chain('root', [[TaskSpec(OP_LOAD, 1 << n, 300.0, 16.0, a=(u0, 0))]] +
                 _merge_local_passes(_fft_stages((u0, 0), n, 1, 'fwd')), owns=[u0], depth=0)
    chains.append(root)
    groups: Dict[int, List[int]] = {}
    for n1, p1 in enumerate(bank.psi1):
        groups.setdefault(k_of[(n1,)], []).append(n1)
    for k1 in sorted(groups):
        l1 = n - k1
        per_batch = max(1, batch_slots >> l1)
        members = groups[k1]
        for s0 in range(0, len(members), per_batch):
            batch = members[s0:s0 + per_batch]
            nb = len(batch)
            x1 = Buf(nb << l1, 'U1[k1=%d:%d]' % (k1, batch[0]))
            mf = [_mulfold(arena, (u0, 0), n, k1, (x1, i << l1), psi1_off[n1]) for i, n1 in enumerate(batch)]
            st = [mf] + _merge_local_passes(_fuse_first_inverse_pass(mf, _fft_stages((x1, 0), l1, nb, 'inv_mod'), l1))   # :307-315
            st.append([store((x1, i << l1), (n1,)) for i, n1 in enumerate(batch)])                                   # :329-330
            kids: Dict[int, List[Tuple[int, int, int]]] = {}
            if max_order == 2:
                for i, n1 in enumerate(batch):
                    for n2, p2 in enumerate(bank.psi2):
                        if p2.j > bank.psi1[n1].j:
                            kids.setdefault(k_of[(n1, n2)] - k1, []).append((i, n1, n2))
            if kids:
                st += _merge_local_passes(_fft_stages((x1, 0), l1, nb, 'fwd'))                                       # :317-318
            c1 = Chain(x1.name, st, after=[root], reads=[u0], owns=[x1], depth=1)
            chains.append(c1)
            for k2 in sorted(kids):
                l2 = l1 - k2
                fam = kids[k2]
                per = max(1, (batch_slots // 2) >> l2)
                for s2 in range(0, len(fam), per):
                    sub = fam[s2:s2 + per]
                    x2 = Buf(len(sub) << l2, 'U2[%d,k2=%d]' % (batch[0], k2))
                    mf2 = [_mulfold(arena, (x1, i << l1), l1, k2, (x2, c << l2), psi2_off[n2][k1])
                           for c, (i, n1, n2) in enumerate(sub)]                                               # :347-348
                    st2 = [mf2] + _merge_local_passes(
                        _fuse_first_inverse_pass(mf2, _fft_stages((x2, 0), l2, len(sub), 'inv_mod'), l2))           # :350-353
                    st2.append([store((x2, c << l2), (n1, n2)) for c, (i, n1, n2) in enumerate(sub)])             # :366-367
                    chains.append(Chain(x2.name, st2, after=[c1], reads=[x1], owns=[x2], frees_own_at_end=True, depth=2))
    segments = [(key, offsets[key], seg_len[key][1]) for key in keys]
    return chains, segments, total


def build_plan_unaveraged(J: int, N: int, Q, T: int, max_order: int = 2, oversampling: int = 0):
    """Plan of the average=False transform: one output row of `n_out` floats per signal holding the paths
    back to back (`segments`); order 0 (the input itself, core :293-294) is added by the frontend."""
    Q1 = fbk._as_Q1(Q)
    geo = fbk.build_geometry(N, J, Q1, T)
    if geo.J_pad > LOG2_NP_MAX:
        raise NotImplementedError('padded length 2**%d exceeds the single-CTA shared-memory design' % geo.J_pad)
    bank = fbk.build_filter_bank(geo.J_pad, J, Q1, T)
    capacity = smem_capacity()
    last_err = None
    for batch_slots in (BATCH_SLOTS, 4096, 2048, 1024, 512):
        arena = _Arena()
        chains, segments, total = build_chains_unaveraged(bank, geo, T, max_order, arena, batch_slots, oversampling)
        try:
            steps, high, chan, sched = schedule_chains(chains, capacity, 1, 0, 1)
            break
        except (RuntimeError, AssertionError) as e:
            last_err = e
    else:
        raise NotImplementedError('no schedule fits shared memory for this configuration: %s' % last_err)
    tasks, ranges = emit(steps)
    if os.environ.get('TEBSCAT_RELAX', '1') != '0':
        keep = elide_barriers(tasks, ranges, capacity, geo.J_pad)
        for st in range(ranges.shape[0]):
            if not keep[st]:
                tasks[ranges[st, 0]:ranges[st, 1], 11] |= 1
    logical = _round16(high)

    class _U:
        pass
    u = _U()
    u.J, u.Q, u.T, u.N, u.max_order, u.geo, u.bank = J, Q1, T, N, max_order, geo, bank
    u.n_paths, u.n_out, u.segments = 1, total, segments
    u.arena, u.tasks, u.steps = arena.finish(), tasks, ranges
    u.chan = np.zeros(2, np.int32)
    u.smem_complex, u.n_threads = logical + logical // 16, N_THREADS
    u.stats = dict(n_steps=len(steps), n_tasks=tasks.shape[0], smem_logical=high, **sched)
    return u


TILE_SLOTS = (8192, 3 * 8192)     # complex elements one tile job moves through shared memory: small jobs keep every
                                  # SM busy on short buffers; with three 8192-blocks every task has three trips per
                                  # thread, which amortises the per-step latency of the interpreter


def build_tile_plan(n: int, inverse, kind: Optional[str] = None, tile_slots: int = TILE_SLOTS[0]):
    """Large-support level: in-place transforms of length 2^n (n <= 13) on a GLOBAL buffer, one
    job = one tile of 8192 elements = 8192 >> n transforms: LOADC -> chained passes -> STOREC.
    Forward: natural -> bit-reversed; inverse: bit-reversed -> natural, unnormalised (like the cascade's own);
    kind='pair': inverse -> modulus -> forward in one job (core/scattering1d.py:312-318)."""
    if not 1 <= n <= LOG2_NP_MAX:
        raise ValueError('tile transforms have 2 .. 8192 samples')
    count = max(1, tile_slots >> n)
    slots = count << n
    buf = Buf(slots, 'tile')
    st = [[TaskSpec(OP_LOADC, -(-slots // 4), 900.0, 30.0, a=(buf, 0), b=slots)]]
    st += _merge_local_passes(_fft_stages((buf, 0), n, count, kind or ('inv' if inverse else 'fwd')))
    st.append([TaskSpec(OP_STOREC, slots, 300.0, 12.0, a=(buf, 0), b=slots)])
    steps, high, chan, sched = schedule_chains([Chain('tile', st, owns=[buf], depth=0)], smem_capacity(), 1, 0, 1)
    tasks, ranges = emit(steps)
    logical = _round16(high)

    class _T:
        pass
    t = _T()
    t.N, t.n_paths, t.n_out, t.slots = 1 << LOG2_NP_MAX, 1, 1, slots
    t.geo = type('G', (), dict(J_pad=LOG2_NP_MAX, pad_left=0))()
    t.arena, t.tasks, t.steps, t.chan = np.zeros(4, np.float32), tasks, ranges, np.zeros(2, np.int32)
    t.smem_complex, t.n_threads = logical + logical // 16, N_THREADS
    return t


def smem_capacity() -> int:
    """Logical complex slots a schedule may address (the kernel pads one slot per 16)."""
    return ((SMEM_BYTES_MAX // 8 - TW_SLOTS) * 16 // 17) & ~15


def emit(steps) -> Tuple[np.ndarray, np.ndarray]:
    rows, ranges = [], []
    for st in steps:
        ranges.append([len(rows), len(rows) + len(st)])
        rows.extend(st)
    return (np.asarray(rows, dtype=np.int32).reshape(-1, TASK_INTS),
            np.asarray(ranges, dtype=np.int32).reshape(-1, 2))


def u0_in_scratch():
    """TEBSCAT_U0_GLOBAL: where the consumers of the signal's spectrum U0 find it -- '1' (default) a per-CTA global
    scratch for every consumer ('scratch': measured +2.7 .. +4.4 % on the headline, production and J=4 configurations,
    profiles/r02h_ab_u0_scratch2.txt), '0' shared memory (False, the round-1 layout), 'split' shared memory for the
    subsampled first-order filters and the scratch for the full-length ones, 'auto' by the cost model.  A/B switch;
    build_plan's `tune` overrides it."""
    v = os.environ.get('TEBSCAT_U0_GLOBAL', '1')
    return {'0': False, '1': 'scratch', 'split': 'split', 'auto': 'auto'}.get(v, False)


def build_plan(J: int, N: int, Q, T: int, max_order: int = 2, max_parallel: int = 64,
               oversampling: int = 0, tune: Optional[dict] = None) -> ScatPlan:
    """`tune` overrides scheduler knobs (tools/sweep_sched.py): batch_slots, child_slots, pool_slots,
    pack_gain, open_demand, u0_scratch."""
    tune = dict(tune or {})
    scratch = tune.get('u0_scratch', u0_in_scratch())      # False, True ('scratch': every consumer), 'split' or 'auto'
    if scratch == 'auto':
        # both layouts through the cost model: the scratch pays for the L2 traffic of its multiplies, shared memory
        # for the slots U0 occupies; second-order cascades with full-length pairs gain, first-order ones lose
        if max_order != 2:
            scratch = False
        else:
            t2 = dict(tune)
            a = build_plan(J, N, Q, T, max_order, max_parallel, oversampling, dict(t2, u0_scratch=False))
            try:
                b = build_plan(J, N, Q, T, max_order, max_parallel, oversampling, dict(t2, u0_scratch='scratch'))
            except NotImplementedError:
                return a
            return b if b.stats['est_cycles'] < AUTO_SCRATCH_GAIN * a.stats['est_cycles'] else a
    if scratch is True:
        scratch = 'scratch'
    if scratch == 'split' and max_order != 2:
        scratch = False
    Q1 = fbk._as_Q1(Q)
    geo = fbk.build_geometry(N, J, Q1, T)
    if geo.J_pad > LOG2_NP_MAX:
        raise NotImplementedError(
            'padded length 2**%d exceeds the single-CTA shared-memory design (max 2**%d); '
            'the large-support path is not built yet' % (geo.J_pad, LOG2_NP_MAX))
    bank = fbk.build_filter_bank(geo.J_pad, J, Q1, T)
    capacity = smem_capacity()
    # batches as large as shared memory allows: retry with smaller batches / pool when the
    # buffers of a configuration (large output-rate lengths, T << 2**J) do not fit
    last_err = None
    ladder = [(BATCH_SLOTS, POOL_SLOTS), (BATCH_SLOTS, 512), (4096, 1024), (2048, 1024), (1024, 512), (512, 256)]
    if 'batch_slots' in tune or 'pool_slots' in tune:
        ladder.insert(0, (tune.get('batch_slots', BATCH_SLOTS), tune.get('pool_slots', POOL_SLOTS)))
    if 'depth_weight' in tune:
        depth_modes = [float(tune['depth_weight'])]
    elif PRIO_DEPTH is not None:
        depth_modes = [float(PRIO_DEPTH)]
    else:
        depth_modes = [1e12, 0.0]
    # ... and with both reservation rules (schedule_chains: 'sum' always finishes; 'peak' lets a subtree start while
    # another one's tail still runs and may dead-end, in which case that candidate is simply dropped)
    reserve_modes = [tune['reserve']] if 'reserve' in tune else (RESERVE_MODES if max_order == 2 else ['sum'])
    best = None
    for batch_slots, pool_slots in ladder:
        for dw, rs in [(d, r) for r in reserve_modes for d in depth_modes]:
            arena = _Arena()
            chains, keys, n_out, lf, i0 = build_chains(bank, geo, T, max_order, arena, batch_slots, oversampling,
                                                       child_slots=tune.get('child_slots'),
                                                       global_u0=scratch or False)
            try:
                cand = schedule_chains(chains, capacity, lf, i0, n_out, max_parallel,
                                       pool_slots=pool_slots, pack_gain=tune.get('pack_gain', 0.97),
                                       open_demand=tune.get('open_demand', 2.0), depth_weight=dw, reserve=rs,
                                       defer_thin=bool(tune.get('defer_thin', rs == 'peak')))
            except (RuntimeError, AssertionError) as e:
                last_err = e
                continue
            # (the round-1 rules unless another combination is modelled at least 1 % faster)
            if best is None or cand[3]['est_cycles'] < 0.99 * best[0][3]['est_cycles']:
                best = (cand, arena, keys, n_out, lf, i0, (dw, rs))
        if best is not None:
            break
    if best is None:
        raise NotImplementedError('no schedule fits shared memory for this configuration: %s' % last_err)
    (steps, high, chan, sched), arena, keys, n_out, lf, i0, depth_used = best
    sched = dict(sched, depth_weight=depth_used[0], reserve=depth_used[1])
    tasks, ranges = emit(steps)
    n_tasks = tasks.shape[0]
    n_relaxed = 0
    if os.environ.get('TEBSCAT_RELAX', '1') != '0':
        keep = elide_barriers(tasks, ranges, capacity, geo.J_pad)
        for st in range(ranges.shape[0]):
            # (the barrier analysis tracks shared-memory slots only: the step that parks U0 in the global scratch
            # must end in a CTA barrier, the global-source multiplies of other warps read what it wrote)
            if scratch and np.any((tasks[ranges[st, 0]:ranges[st, 1], 0] & 0xff) == OP_STOREC):
                keep[st] = True
            if not keep[st]:
                tasks[ranges[st, 0]:ranges[st, 1], 11] |= 1
                n_relaxed += 1
    stats = dict(n_steps=len(steps), n_tasks=n_tasks, n_relaxed=n_relaxed, smem_logical=high,
                 mean_tasks_per_step=n_tasks / max(1, len(steps)), **sched)
    logical = _round16(high)
    plan = ScatPlan(J, Q1, T, N, max_order, geo, bank, keys, n_out, arena.finish(), tasks, ranges,
                    np.asarray(chan, dtype=np.int32), logical + logical // 16, N_THREADS, stats)
    plan.scratch_complex = (1 << geo.J_pad) if scratch else 0
    plan.u0_mode = scratch or 'shared'
    return plan


# ------------------------------------------------------------------------------------
# large-support level: the subtrees that fit one SM, fused
# ------------------------------------------------------------------------------------
def build_kid_chains(bank: fbk.FilterBank, geo: fbk.Geometry, T: int, arena: _Arena, n1: int,
                     oversampling: int = 0, child_slots: int = BATCH_SLOTS):
    """Second-order subtrees of ONE first-order filter whose own modulus is longer than 8192 samples: the spectrum
    of |u1| (2^l1 bins, bit-reversed) lives in global memory; every child of at most 8192 samples runs here --
    psi2 multiply + periodisation from global memory (OP_GMULFOLD), iFFT -> modulus -> FFT, phi leaf
    (core/scattering1d.py:337-364).  Children are not pair-packed (partners would be different psi2 octaves of
    one parent, whose energies can differ widely).  Returns (chains, channels written, children left out)."""
    n = geo.J_pad
    log2_T = int(math.floor(math.log2(T)))
    os_ = int(oversampling)
    kf = max(log2_T - os_, 0)
    lf = n - kf
    i0, i1 = geo.ind_start[kf], geo.ind_end[kf]
    phi_off = [arena.add(a) for a in bank.phi.levels]
    psi2_off = [[arena.add(a) for a in p.levels] for p in bank.psi2]
    keys: List[Tuple[int, ...]] = [()] + [(i,) for i in range(len(bank.psi1))]
    for m1, q1 in enumerate(bank.psi1):
        for n2, p2 in enumerate(bank.psi2):
            if p2.j > q1.j:
                keys.append((m1, n2))
    channel = {k: c for c, k in enumerate(keys)}
    p1 = bank.psi1[n1]
    k1 = max(min(p1.j - os_, log2_T - os_), 0)
    l1 = n - k1
    kids: Dict[int, List[int]] = {}
    left_out: List[int] = []
    for n2, p2 in enumerate(bank.psi2):
        if p2.j > p1.j:
            k2 = max(min(p2.j - k1 - os_, log2_T - k1 - os_), 0)
            if l1 - k2 <= LOG2_NP_MAX:
                kids.setdefault(k2, []).append(n2)
            else:
                left_out.append(n2)
    chains: List[Chain] = []
    written: List[int] = []
    for k2 in sorted(kids):
        l2 = l1 - k2
        per = max(1, child_slots >> l2)
        fam = kids[k2]
        for s2 in range(0, len(fam), per):
            sub = fam[s2:s2 + per]
            x2 = Buf(len(sub) << l2, 'U2[%d,k2=%d:%d]' % (n1, k2, sub[0]))
            mf = [_gmulfold(arena, 0, l1, k2, (x2, c << l2), psi2_off[n2][k1]) for c, n2 in enumerate(sub)]      # :347-348
            st = [mf] + _merge_local_passes(
                _fuse_first_inverse_pass(mf, _fft_stages((x2, 0), l2, len(sub), 'pair'), l2))                   # :350-355
            c2 = Chain(x2.name, st, owns=[x2], depth=1)
            chains.append(c2)
            lv = [_mulfold(arena, (x2, c << l2), l2, l2 - lf, LEAF, phi_off[k1 + k2], (channel[(n1, n2)], -1))
                  for c, n2 in enumerate(sub)]                                                                 # :358-364
            chains.append(Chain('S2' + x2.name, [lv], after=[c2], reads=[x2], depth=2))
            written += [channel[(n1, n2)] for n2 in sub]
    return chains, written, left_out, len(keys), i1 - i0, lf, i0


def _finish_fused_plan(J, Q1, T, N, max_order, geo, bank, n_paths, n_out, lf, i0, chains, arena, pool_slots):
    capacity = smem_capacity()
    # the same candidates as build_plan: chain priority depth first / by weight, reservation rule 'sum' / 'peak'
    # (schedule_chains resets the chains' and buffers' scheduling state, so one set of chains serves all of them)
    best, last_err = None, None
    depth_modes = [float(PRIO_DEPTH)] if PRIO_DEPTH is not None else [1e12, 0.0]
    for rs in RESERVE_MODES:
        for dw in depth_modes:
            for c in chains:
                for b in c.owns + c.reads:
                    b.off, b.readers_left, b.producer_done = -1, 0, False
            sizes = {id(b): b.size for c in chains for b in c.owns}
            try:
                cand = schedule_chains(chains, capacity, lf, i0, n_out, pool_slots=pool_slots, depth_weight=dw, reserve=rs,
                                       defer_thin=rs == 'peak')
            except (RuntimeError, AssertionError) as e:
                last_err = e
                cand = None
            for c in chains:                                   # (a packed pass shrinks its buffer while scheduling)
                for b in c.owns:
                    b.size = sizes[id(b)]
            if cand is not None and (best is None or cand[3]['est_cycles'] < 0.99 * best[0][3]['est_cycles']):
                best = (cand, dw, rs)
    if best is None:
        raise RuntimeError('no schedule: %s' % last_err)
    (steps, high, chan, sched), dw, rs = best
    sched = dict(sched, depth_weight=dw, reserve=rs)
    tasks, ranges = emit(steps)
    if os.environ.get('TEBSCAT_RELAX', '1') != '0':
        keep = elide_barriers(tasks, ranges, capacity, LOG2_NP_MAX)
        for st in range(ranges.shape[0]):
            if not keep[st]:
                tasks[ranges[st, 0]:ranges[st, 1], 11] |= 1
    logical = _round16(high)

    class _F:
        pass
    f = _F()
    # the device plan only needs a geometry the kernel accepts: these schedules have no LOAD (their source is global)
    f.J, f.Q, f.T, f.max_order, f.bank = J, Q1, T, max_order, bank
    f.N, f.geo = 1 << LOG2_NP_MAX, type('G', (), dict(J_pad=LOG2_NP_MAX, pad_left=0))()
    f.n_paths, f.n_out = n_paths, n_out
    f.arena, f.tasks, f.steps = arena.finish(), tasks, ranges
    f.chan = np.asarray(chan, dtype=np.int32)
    f.smem_complex, f.n_threads = logical + logical // 16, N_THREADS
    f.stats = dict(n_steps=len(steps), n_tasks=tasks.shape[0], smem_logical=high, **sched)
    return f


def build_hybrid_plans(J: int, N: int, Q, T: int, max_order: int = 2, oversampling: int = 0):
    """Padded lengths above 2^13 (the large-support level): the parts of the cascade that fit one SM as schedules
    of the fused kernel with a global source.

    Returns dict(first=plan or None, first_n1=[...], kids=[(plan, [n1, ...], [n2 done...])]):
      * `first`: every first-order filter whose subsampled length is <= 8192 samples, with its whole subtree
        (pair-packed like the cascade proper) -- ONE launch per batch reading the signal's spectrum U0;
      * `kids`: per group of longer first-order filters of one scale j1 (same children, channels a constant shift
        apart), the second-order children of <= 8192 samples -- one launch per filter reading the spectrum of |u1|.
    A part whose buffers do not fit shared memory is left out (None / missing): the level's own kernels serve it."""
    Q1 = fbk._as_Q1(Q)
    geo = fbk.build_geometry(N, J, Q1, T)
    n = geo.J_pad
    if n <= LOG2_NP_MAX:
        raise ValueError('the fused cascade serves this padded length as a whole')
    bank = fbk.build_filter_bank(n, J, Q1, T)
    log2_T = int(math.floor(math.log2(T)))
    os_ = int(oversampling)
    k1_of = [max(min(p.j - os_, log2_T - os_), 0) for p in bank.psi1]
    out = dict(first=None, first_n1=[], kids=[])
    small = [n1 for n1 in range(len(bank.psi1)) if n - k1_of[n1] <= LOG2_NP_MAX]
    if small:
        for batch_slots, pool_slots in [(BATCH_SLOTS, POOL_SLOTS), (BATCH_SLOTS, 512), (4096, 1024), (2048, 1024), (1024, 512)]:
            arena = _Arena()
            try:
                chains, keys, n_out, lf, i0 = build_chains(bank, geo, T, max_order, arena, batch_slots, oversampling,
                                                           global_u0=True)
                out['first'] = _finish_fused_plan(J, Q1, T, N, max_order, geo, bank, len(keys), n_out, lf, i0, chains,
                                                  arena, pool_slots)
                out['first_n1'] = small
                break
            except (RuntimeError, AssertionError, NotImplementedError):
                continue
    if max_order == 2:
        groups: Dict[Tuple[int, int], List[int]] = {}
        for n1, p1 in enumerate(bank.psi1):
            if n - k1_of[n1] > LOG2_NP_MAX:
                groups.setdefault((p1.j, k1_of[n1]), []).append(n1)
        for (_, _), members in sorted(groups.items()):
            head = members[0]
            for child_slots, pool_slots in [(BATCH_SLOTS, POOL_SLOTS), (BATCH_SLOTS, 512), (4096, 1024), (2048, 512)]:
                arena = _Arena()
                try:
                    chains, written, left_out, n_paths, n_out, lf, i0 = build_kid_chains(bank, geo, T, arena, head,
                                                                                         oversampling, child_slots)
                    if not chains:
                        break
                    plan = _finish_fused_plan(J, Q1, T, N, max_order, geo, bank, n_paths, n_out, lf, i0, chains, arena,
                                              pool_slots)
                    plan.head, plan.written = head, written
                    done = [n2 for n2, p2 in enumerate(bank.psi2) if p2.j > bank.psi1[head].j and n2 not in left_out]
                    out['kids'].append((plan, members, done))
                    break
                except (RuntimeError, AssertionError, NotImplementedError):
                    continue
    return out
