"""Host-side plan builder: filter arena + step schedule for the fused cascade kernel.

The CUDA kernel (csrc/scat_core.cuh) is an interpreter of STEPS; each step is a
set of TASKS on disjoint thread ranges of one 512-thread CTA followed by one
barrier.  This module turns the cascade of
``kymatio/scattering1d/core/scattering1d.py:269-370`` into that form:

* every transform of the cascade becomes a CHAIN of tasks
  (MULFOLD -> inverse FFT passes [+ modulus] -> forward FFT passes, or
  MULFOLD -> inverse FFT passes -> STORE for the phi low-pass leaves);
* chains that do not depend on each other are packed side by side into the same
  steps by a list scheduler under two resources: the 512 threads of the CTA and
  the shared-memory slots their buffers need;
* filters are cast to fp32 exactly like ``register_filters``
  (``frontend/torch_frontend.py:75-97``), permuted to bit-reversed bin order and
  laid out in one flat arena.

Task encoding (8 x int32) must match ``csrc/scat_core.cuh``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import filterbank as fbk

OP_NOP, OP_LOAD, OP_FFT, OP_MULFOLD, OP_STORE = 0, 1, 2, 3, 4
FFT_INV, FFT_MOD = 1, 2
ST_IMAG = 1

N_THREADS = 512
LOG2_NP_MAX = 13                  # single-CTA limit of the kernel (kLog2TwMax)
SMEM_BYTES_MAX = 227 * 1024
TW_SLOTS = 64 + 128               # twiddle tables live behind the schedule's slots


def bitrev_indices(n: int) -> np.ndarray:
    """perm[p] = bit-reversal of p on log2(n) bits."""
    bits = int(math.log2(n))
    p = np.arange(n, dtype=np.int64)
    r = np.zeros(n, dtype=np.int64)
    for b in range(bits):
        r |= ((p >> b) & 1) << (bits - 1 - b)
    return r


def radix_split(n: int) -> List[int]:
    """Split log2(L)=n into the fewest passes of log2-radix <= 4, as evenly as possible."""
    if n <= 0:
        return []
    m = -(-n // 4)
    base, extra = divmod(n, m)
    return [base + 1] * extra + [base] * (m - extra)


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


@dataclass
class TaskSpec:
    op: int
    work: int                      # independent work items (butterflies / outputs / samples)
    cost: float                    # relative cost of one work item
    a: object = 0                  # ints, or Buf (resolved to .off at emission)
    b: int = 0
    c: int = 0
    d: object = 0
    e: int = 0
    sexp: int = 0


@dataclass
class Chain:
    name: str
    tasks: List[TaskSpec]
    reads: Optional[Buf] = None            # buffer read by the first task (must be complete)
    after: Optional['Chain'] = None        # chain that produces `reads`
    owns: List[Buf] = field(default_factory=list)      # allocated when the chain starts
    frees_own_at_end: bool = True
    priority: float = 0.0
    # scheduler state
    pos: int = 0
    done_step: int = -1


def _fft_tasks(buf: Buf, n: int, inverse: bool, modulus: bool = False) -> List[TaskSpec]:
    """Passes of one in-place length-2^n transform (forward: DIF, inverse: DIT)."""
    cost = {1: 30.0, 2: 70.0, 3: 170.0, 4: 400.0}
    out = []
    if not inverse:
        logB = n
        for r in radix_split(n):
            out.append(TaskSpec(OP_FFT, 1 << (n - r), cost[r], a=buf, b=n, c=logB, d=r, e=0))
            logB -= r
    else:
        logB = 0
        split = radix_split(n)[::-1]
        for i, r in enumerate(split):
            logB += r
            last = i == len(split) - 1
            flags = FFT_INV | (FFT_MOD if (modulus and last) else 0)
            out.append(TaskSpec(OP_FFT, 1 << (n - r), cost[r], a=buf, b=n, c=logB, d=r, e=flags))
    return out


def _mulfold(src: Buf, log_src: int, logk: int, dst: Buf, filt_off: int) -> TaskSpec:
    log_dst = log_src - logk
    # mean over k blocks (2^-logk) and the 1/L of the following inverse transform (2^-log_dst)
    return TaskSpec(OP_MULFOLD, 1 << log_dst, 8.0 + 6.0 * (1 << logk), a=src, b=log_src, c=logk,
                    d=dst, e=filt_off, sexp=logk + log_dst)


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
    tasks: np.ndarray              # int32 [n_tasks, 8]
    steps: np.ndarray              # int32 [n_steps, 2]
    smem_complex: int              # PHYSICAL complex slots (logical slots + 1 pad per 16)
    n_threads: int = N_THREADS
    stats: Dict[str, float] = field(default_factory=dict)

    @property
    def n_paths(self) -> int:
        return len(self.keys)


class _Arena:
    def __init__(self):
        self.chunks: List[np.ndarray] = []
        self.size = 0

    def add(self, f64: np.ndarray) -> int:
        f32 = f64.astype(np.float32)                       # the reference's .float() cast
        perm = f32[bitrev_indices(f32.shape[0])]
        off = self.size
        pad = (-perm.shape[0]) % 4
        self.chunks.append(np.concatenate([perm, np.zeros(pad, np.float32)]))
        self.size += perm.shape[0] + pad
        return off

    def finish(self) -> np.ndarray:
        return np.concatenate(self.chunks) if self.chunks else np.zeros(4, np.float32)


def _round16(n: int) -> int:
    return (n + 15) & ~15


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


def build_chains(bank: fbk.FilterBank, geo: fbk.Geometry, T: int, max_order: int,
                 arena: _Arena) -> Tuple[List[Chain], List[Tuple[int, ...]], int]:
    """The cascade as a forest of chains, in the reference's channel order."""
    n = geo.J_pad
    log2_T = int(math.floor(math.log2(T)))
    lf = n - log2_T                                       # log2 of the final (output-rate) length
    if lf < 0:
        raise ValueError('T is larger than the padded support')
    i0, i1 = geo.ind_start[log2_T], geo.ind_end[log2_T]
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

    def lowpass(src: Buf, log_src: int, level: int, key, parent: Chain) -> Chain:
        """phi[level] multiply, periodise down to 2^lf, inverse transform, unpad, store."""
        y = Buf(1 << lf, 'Y%s' % (key,))
        tasks = [_mulfold(src, log_src, log_src - lf, y, phi_off[level])]
        tasks += _fft_tasks(y, lf, inverse=True)
        tasks.append(TaskSpec(OP_STORE, n_out, 6.0, a=y, b=channel[key], c=i0, d=n_out, e=0))
        ch = Chain('S%s' % (key,), tasks, reads=src, after=parent, owns=[y])
        chains.append(ch)
        return ch

    # root: pad + forward transform of the signal (core/scattering1d.py:278-280)
    u0 = Buf(1 << n, 'U0')
    root = Chain('root', [TaskSpec(OP_LOAD, 1 << n, 8.0, a=u0)] + _fft_tasks(u0, n, inverse=False),
                 owns=[u0], frees_own_at_end=False)
    chains.append(root)
    lowpass(u0, n, 0, (), root)                                        # S0 (:285-292)

    for n1, p1 in enumerate(bank.psi1):                                # :300
        k1 = max(min(p1.j, log2_T), 0)                                 # :304
        if not p1.xi < 0.5 / (2 ** k1):
            raise AssertionError('psi1 aliasing assertion of the reference violated')
        l1 = n - k1
        x1 = Buf(1 << l1, 'U1[%d]' % n1)
        t = [_mulfold(u0, n, k1, x1, psi1_off[n1])]                    # :307-311
        t += _fft_tasks(x1, l1, inverse=True, modulus=True)            # :312-315
        t += _fft_tasks(x1, l1, inverse=False)                         # :318
        c1 = Chain('U1[%d]' % n1, t, reads=u0, after=root, owns=[x1], frees_own_at_end=False)
        chains.append(c1)
        lowpass(x1, l1, k1, (n1,), c1)                                 # :320-327
        if max_order == 2:
            for n2, p2 in enumerate(bank.psi2):                        # :337
                if p2.j > p1.j:
                    if not p2.xi < p1.xi:
                        raise AssertionError('psi2 ordering assertion of the reference violated')
                    k2 = max(min(p2.j - k1, log2_T - k1), 0)           # :344-345
                    l2 = l1 - k2
                    x2 = Buf(1 << l2, 'U2[%d,%d]' % (n1, n2))
                    t = [_mulfold(x1, l1, k2, x2, psi2_off[n2][k1])]   # :347-348
                    t += _fft_tasks(x2, l2, inverse=True, modulus=True)
                    t += _fft_tasks(x2, l2, inverse=False)             # :355
                    c2 = Chain('U2[%d,%d]' % (n1, n2), t, reads=x1, after=c1, owns=[x2],
                               frees_own_at_end=False)
                    chains.append(c2)
                    lowpass(x2, l2, k1 + k2, (n1, n2), c2)             # :358-364
    return chains, keys, n_out


def _want_threads(work: int) -> int:
    return min(N_THREADS, max(32, (work + 31) & ~31))


def _task_time(t: TaskSpec, nt: int) -> float:
    return math.ceil(t.work / nt) * t.cost


def _split_threads(tasks: List[TaskSpec]) -> Optional[List[int]]:
    """Thread counts (multiples of 32, sum <= 512) that minimise the slowest task."""
    n = len(tasks)
    if 32 * n > N_THREADS:
        return None
    nts = [32] * n
    left = N_THREADS - 32 * n
    while left > 0:
        times = [_task_time(t, nt) for t, nt in zip(tasks, nts)]
        order = sorted(range(n), key=lambda i: -times[i])
        grew = False
        for i in order:
            if nts[i] >= _want_threads(tasks[i].work):
                if i == order[0]:
                    break                                # the slowest task cannot go faster
                continue
            # smallest increment that removes one iteration
            it = math.ceil(tasks[i].work / nts[i])
            need = nts[i] + 32
            while need < N_THREADS and math.ceil(tasks[i].work / need) >= it and \
                    need < _want_threads(tasks[i].work):
                need += 32
            if need - nts[i] > left:
                if i == order[0]:
                    break
                continue
            left -= need - nts[i]
            nts[i] = need
            grew = True
            break
        if not grew:
            break
    return nts


def schedule_chains(chains: List[Chain], capacity: int, max_parallel: int = 64,
                    stretch: float = 1.15, step_overhead: float = 150.0):
    """Greedy list scheduling of chains into steps.

    Resources: the 512 threads of the CTA (per step) and `capacity` shared-memory
    slots (over buffer lifetimes).  A chain may start once the chain producing the
    buffer it reads has finished and its own buffers fit; a buffer is released when
    every chain reading it has executed its first task.  Starting a chain that has
    dependants also RESERVES room for their buffers so that a subtree, once begun,
    can always be finished and its memory returned (no deadlock, depth first).
    """
    children: Dict[int, List[Chain]] = {}
    for ch in chains:
        if ch.after is not None:
            children.setdefault(id(ch.after), []).append(ch)
    readers: Dict[int, int] = {}
    for ch in chains:
        if ch.reads is not None:
            readers[id(ch.reads)] = readers.get(id(ch.reads), 0) + 1
    bufs = {id(b): b for ch in chains for b in ch.owns}
    for k, v in readers.items():
        bufs[k].readers_left = v

    own_size = {id(c): sum(_round16(b.size) for b in c.owns) for c in chains}
    depth: Dict[int, int] = {}
    weight: Dict[int, float] = {}
    need: Dict[int, int] = {}

    def visit(c: Chain, d: int):
        depth[id(c)] = d
        w = sum(t.work * t.cost for t in c.tasks)
        nd = 0
        for k in children.get(id(c), []):
            visit(k, d + 1)
            w += weight[id(k)]
            nd += own_size[id(k)] + need[id(k)]
        weight[id(c)] = w
        need[id(c)] = nd if d > 0 else 0          # the root does not reserve for the whole tree

    for c in chains:
        if c.after is None:
            visit(c, 0)
    parent_of = {id(c): c.after for c in chains}
    for c in chains:
        c.priority = depth[id(c)] * 1e12 + weight[id(c)]
        c.pos = 0
        c.done_step = -1

    alloc = _Allocator(capacity)
    free_slots = capacity
    reserve_left: Dict[int, int] = {}              # chain id -> slots still reserved for its subtree
    pending = list(chains)
    active: List[Chain] = []
    steps: List[List[List[int]]] = []
    producer_done: Dict[int, bool] = {}
    step_idx = 0
    est_time = 0.0
    est_work = 0.0

    def release(buf: Buf):
        nonlocal free_slots
        alloc.release(buf.off, buf.size)
        free_slots += _round16(buf.size)
        buf.off = -2

    def release_if_dead(buf: Buf):
        if buf.readers_left == 0 and producer_done.get(id(buf), False) and buf.off >= 0:
            release(buf)

    def ancestors(c: Chain):
        a = parent_of[id(c)]
        while a is not None:
            yield a
            a = parent_of[id(a)]

    def try_start(c: Chain, force: bool) -> bool:
        nonlocal free_slots
        mine = own_size[id(c)]
        reserved_others = sum(reserve_left.values()) - sum(reserve_left.get(id(a), 0) for a in ancestors(c))
        if not force and free_slots - reserved_others < mine + need[id(c)]:
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
        if need[id(c)] > 0:
            reserve_left[id(c)] = need[id(c)]
        return True

    while pending or active:
        startable = [c for c in pending if c.after is None or
                     (c.after.done_step >= 0 and c.after.done_step < step_idx)]
        startable.sort(key=lambda c: -c.priority)
        demand = sum(_want_threads(c.tasks[c.pos].work) for c in active)
        for c in startable:
            if len(active) >= max_parallel:
                break
            if demand >= 2 * N_THREADS and depth[id(c)] <= 1:
                continue                          # enough queued work; do not open new subtrees
            if try_start(c, force=False):
                pending.remove(c)
                active.append(c)
                demand += _want_threads(c.tasks[0].work)
        if not active:
            for c in startable:                   # progress guarantee
                if try_start(c, force=True):
                    pending.remove(c)
                    active.append(c)
                    break
        if not active:
            raise RuntimeError('schedule deadlock: %d chains cannot be placed in %d slots'
                               % (len(pending), capacity))

        # choose the tasks of this step
        active.sort(key=lambda c: -c.priority)
        chosen: List[Chain] = []
        nts: List[int] = []
        cur_max = 0.0
        for c in active:
            trial = chosen + [c]
            split = _split_threads([k.tasks[k.pos] for k in trial])
            if split is None:
                break
            tmax = max(_task_time(k.tasks[k.pos], nt) for k, nt in zip(trial, split))
            if chosen and tmax > stretch * cur_max + 1e-9:
                continue
            chosen, nts, cur_max = trial, split, max(tmax, cur_max) if chosen else tmax
        this_step: List[List[int]] = []
        used = 0
        for c, nt in zip(chosen, nts):
            t = c.tasks[c.pos]
            a = t.a.off if isinstance(t.a, Buf) else t.a
            d = t.d.off if isinstance(t.d, Buf) else t.d
            assert a >= 0 and d >= 0, 'task touches a released buffer'
            this_step.append([t.op | (t.sexp << 8), used, nt, int(a), t.b, t.c, int(d), t.e])
            used += nt
            est_work += t.work * t.cost
        est_time += cur_max + step_overhead
        steps.append(this_step)
        for c in chosen:
            if c.pos == 0 and c.reads is not None:
                c.reads.readers_left -= 1
                release_if_dead(c.reads)
            c.pos += 1
            if c.pos == len(c.tasks):
                c.done_step = step_idx
                active.remove(c)
                reserve_left.pop(id(c), None) if not children.get(id(c)) else None
                for b in c.owns:
                    producer_done[id(b)] = True
                    if c.frees_own_at_end:
                        release(b)
                    else:
                        release_if_dead(b)
        # a reservation ends when the whole subtree of its chain has finished
        for cid in list(reserve_left.keys()):
            def subtree_done(ch: Chain) -> bool:
                return ch.done_step >= 0 and all(subtree_done(k) for k in children.get(id(ch), []))
            owner = next(c for c in chains if id(c) == cid)
            if subtree_done(owner):
                reserve_left.pop(cid)
        step_idx += 1
    sched_stats = dict(est_time=est_time, est_work=est_work,
                       est_fill=est_work / (N_THREADS * max(est_time, 1.0)))
    return steps, alloc.high_water, sched_stats


def emit(steps) -> Tuple[np.ndarray, np.ndarray]:
    rows, ranges = [], []
    for st in steps:
        ranges.append([len(rows), len(rows) + len(st)])
        rows.extend(st)
    return np.asarray(rows, dtype=np.int32).reshape(-1, 8), np.asarray(ranges, dtype=np.int32).reshape(-1, 2)


def build_plan(J: int, N: int, Q, T: int, max_order: int = 2, max_parallel: int = 64) -> ScatPlan:
    Q1 = fbk._as_Q1(Q)
    geo = fbk.build_geometry(N, J, Q1, T)
    if geo.J_pad > LOG2_NP_MAX:
        raise NotImplementedError(
            'padded length 2**%d exceeds the single-CTA shared-memory design (max 2**%d); '
            'the large-support path is not built yet' % (geo.J_pad, LOG2_NP_MAX))
    bank = fbk.build_filter_bank(geo.J_pad, J, Q1, T)
    arena = _Arena()
    chains, keys, n_out = build_chains(bank, geo, T, max_order, arena)
    # logical slots; the kernel pads one slot per 16 (scat_core.cuh: swz)
    capacity = ((SMEM_BYTES_MAX // 8 - TW_SLOTS) * 16 // 17) & ~15
    steps, high, sched = schedule_chains(chains, capacity, max_parallel)
    tasks, ranges = emit(steps)
    n_tasks = tasks.shape[0]
    stats = dict(n_steps=len(steps), n_tasks=n_tasks, smem_complex=high,
                 mean_tasks_per_step=n_tasks / max(1, len(steps)), **sched)
    logical = _round16(high)
    return ScatPlan(J, Q1, T, N, max_order, geo, bank, keys, n_out, arena.finish(), tasks, ranges,
                    logical + logical // 16, N_THREADS, stats)
