#!/usr/bin/env python
"""bench.py -- Scattering1D signals/s (J=6, Q=8, T=64, N=4800) on 1..8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the fused cascade over one batch of BASELINE.json's
configs[1]: 8192 two-channel CTG-shaped samples = 16384 signals PER GPU (weak
scaling; the batch shards with no collective).  `value` is device-resident
throughput (CUDA events on the launch stream, max over ranks); `e2e` is the same
metric through the host-buffer C-ABI entry point with pinned host tensors, H2D
and D2H inside the timed region.  `--impl reference` times the CPU port of the
reference (oracle/scattering1d_torch_port.py, all host threads) on a bounded sample of the
same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'vae-teb_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

J, Q, T, N = 6, 8, 64, 4800
SAMPLES_PER_GPU = 8192                    # two-channel samples -> 2 signals each
METRIC = 'Scattering1D signals/s (J=6,Q=8,N=4800)'
UNIT = 'signals/s'
BYTES_PER_SIGNAL = N * 4 + 126 * 75 * 4   # 57 000 B algorithmic
FLOPS_PER_SIGNAL = 31.56e6                # reference-equivalent FFT flops


def measured_traffic(n_sig):
    """DRAM bytes of one launch of the dominant kernel from the committed ncu capture (profiles/)."""
    for name in ('r02m_traffic.json', 'r01c_traffic.json', 'r01b_traffic.json'):     # newest capture first
        path = os.path.join(ROOT, 'profiles', name)
        if os.path.exists(path):
            break
    else:
        return None
    with open(path) as f:
        t = json.load(f)
    if t.get('signals_per_launch') != n_sig:
        return None
    return t['dram_bytes_read'] + t['dram_bytes_write']


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def bf16_sustained_peak():
    """Measured dense bf16 TFLOP/s of this pool's B200 inside a long step (MEASURED_PEAKS.json), 0.0 if absent."""
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f).get('bf16_tflops_sustained', 0.0))
    return 0.0


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.sm, self.reasons, self.sm_max = [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8): 'hw_slowdown',
                getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
                getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
                getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4): 'sw_power_cap',
            }
            while not self.stop_flag.is_set():
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.1)
        except Exception as e:                     # clocks are evidence, not a dependency
            self.reasons.add('unavailable:%s' % type(e).__name__)

    def summary(self):
        sm = sorted(self.sm)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': self.sm_max,
                'reasons': sorted(self.reasons)}


def cpu_port_rate(n_signals, workers):
    """signals/s of the torch-CPU port of the reference (oracle/scattering1d_torch_port.py: the
    reference's own torch calls, op for op) on `workers` host threads, batches of 64 signals
    (BASELINE configs[0]: batch 32 x 2 channels)."""
    import torch
    from oracle.scattering1d_torch_port import TorchPort
    from tebscat.synth import ctg_batch
    torch.set_num_threads(workers)
    x = ctg_batch(n_signals // 2, N, seed=1234).reshape(n_signals, N)
    port = TorchPort(J, N, Q, T, 2)
    port(x[:8])                                    # warm-up
    t0 = time.perf_counter()
    for b0 in range(0, n_signals, 64):
        port(x[b0:b0 + 64])
    dt = time.perf_counter() - t0
    return n_signals / dt, dt


def run_reference(args):
    """The reference arm: CPU port of kymatio's torch-CPU path, all host threads, bounded sample."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = 1024
    rates = []
    for _ in range(max(1, args.warmup)):
        cpu_port_rate(64, cores)
    for _ in range(args.steps):
        r, _ = cpu_port_rate(sample, cores)
        rates.append(r)
    value = sum(rates) / len(rates)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * sample / value,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': 'Scattering1D J=6 Q=8 T=64 N=4800 orders 0-2 (CPU port of the reference, '
                               'bounded sample of %d signals per step)' % sample},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d CTG signals per step in batches of 64, torch-CPU port of the reference (oracle/scattering1d_torch_port.py)' % sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from tebscat import Scattering1D, _lib
    from tebscat.synth import ctg_batch
    from tebscat.sharding import bind_host_to_device, max_over_ranks
    import ctypes

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    # host side of the end-to-end path: this rank's pinned buffers live on the GPU's own NUMA node
    prev_affinity = bind_host_to_device(local) if world > 1 else None
    S = Scattering1D(J, N, Q, T=T).to(dev)
    n_sig = 2 * SAMPLES_PER_GPU
    # a few hundred distinct synthetic records tiled up to the batch (generation is host-side and slow)
    base = ctg_batch(256, N, seed=1234 + rank)                      # (256, 2, N)
    x_host = base.repeat(SAMPLES_PER_GPU // 256, 1, 1).contiguous().pin_memory()   # (8192, 2, N)
    x_dev = x_host.to(dev)
    out, _ = S(x_dev)                                               # builds the plan, warms up
    torch.cuda.synchronize()
    C, n_out = out.shape[-2], out.shape[-1]
    del out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ------------------------------------------------
    for _ in range(args.warmup):
        S(x_dev)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches = 0
    ev[0].record()
    for i in range(args.steps):
        S(x_dev)
        launches += _lib.load().tebscat_last_launch_count()
        ev[i + 1].record()
    barrier()
    sampler.stop_flag.set()
    sampler.join()
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps))
    total_ms_max = max_over_ranks(total_ms, dev)
    value = world * n_sig * args.steps / (total_ms_max * 1e-3)

    # ---- end to end through the host-buffer C-ABI entry point -------------------------
    out_host = torch.empty((n_sig, C, n_out), dtype=torch.float32, pin_memory=True)
    S.scattering_host(x_host, out=out_host, device=local)           # warm-up (allocates the pipeline)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        S.scattering_host(x_host, out=out_host, device=local)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_value = world * n_sig * args.steps / max_over_ranks(e2e_s, dev)
    # the same pipeline without its kernel: what this box's host <-> device path can move for exactly this traffic,
    # with all ranks copying at once (tools/host_copy_ceiling.py sweeps it on its own)
    plan_h = S._plan_for(local).handle
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _lib.check(_lib.load().tebscat_scat1d_host_copies_only(plan_h, x_host.data_ptr(), n_sig, out_host.data_ptr()))
    torch.cuda.synchronize()
    copy_s = max_over_ranks(time.perf_counter() - t0, dev)
    copy_value = world * n_sig * args.steps / copy_s
    local_cpus = len(os.sched_getaffinity(0))
    if prev_affinity is not None:
        os.sched_setaffinity(0, prev_affinity)                       # the CPU baseline below uses every host core

    # ---- secondary: cross-channel phase scattering (BASELINE configs[2]) on a bounded batch ----
    phase = None
    if rank == 0 and not args.no_phase:
        from tebscat import KymatioPhaseScattering1D
        pm = KymatioPhaseScattering1D(J=J, Q=Q, T=T, shape=N, device=dev)
        PB = args.phase_batch                                          # BASELINE configs[2]: 8192 FHR x UP pairs
        xb = x_dev[:PB]
        pm(xb[:512], compute_phase=False, compute_cross_phase=True)    # builds the plans
        pm(xb, compute_phase=False, compute_cross_phase=True)          # warm-up at the full batch (workspaces, allocator)
        torch.cuda.synchronize()
        ph = pm._dev_plan(local).handle
        lib = _lib.load()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        _lib.check(lib.tebscat_phase_plan_profile(ph, 1))
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        p0.record()
        marks[0].record()
        for r in range(reps):
            po = pm(xb, compute_phase=False, compute_cross_phase=True)['cross_phase_corr']
            marks[r + 1].record()
        p1.record()
        torch.cuda.synchronize()
        pms = p0.elapsed_time(p1) / reps
        rep_ms = [marks[r].elapsed_time(marks[r + 1]) for r in range(reps)]
        a_ms, b_ms, n_chunks = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_int(0)
        _lib.check(lib.tebscat_phase_plan_profile_read(ph, ctypes.byref(a_ms), ctypes.byref(b_ms), ctypes.byref(n_chunks)))
        _lib.check(lib.tebscat_phase_plan_profile(ph, 0))
        n_pairs, n_po = int(po.shape[1]), int(po.shape[2])
        del po
        stage_b_s = b_ms.value * 1e-3 / reps
        # dominant kernel: the dense contraction on tcgen05 -- rows x (2 N) x 80 columns, three TF32 products each
        tensor_flops = PB * n_pairs * (2 * N) * 80 * 2 * 3
        bf16_sustained = bf16_sustained_peak()
        tf32_peak = bf16_sustained / 2 if bf16_sustained else 1125.0
        ach = tensor_flops / stage_b_s / 1e12
        phase = {'metric': 'cross-channel phase scattering FHR x UP pairs/s (J=6,Q=8,N=4800, 741 pairs)',
                 'value': PB / (pms * 1e-3), 'unit': 'signal-pairs/s', 'batch': PB, 'ms': pms, 'ms_per_pass': rep_ms,
                 'config': 'BASELINE configs[2]: batch %d two-channel signals, N=4800, all 741 pairs; output %.2f GB per pass' % (
                     PB, PB * n_pairs * n_po * 4 / 1e9),
                 'stage_ms': {'stage_a': a_ms.value / reps, 'stage_b': b_ms.value / reps, 'chunks': n_chunks.value // reps,
                              'how': 'CUDA events on the launch stream around the stages of every workspace chunk '
                                     '(tebscat_phase_plan_profile)'},
                 'roofline': {'bound': 'tensor', 'kernel': 'phase_pair_tc_kernel (stage B, dense form)',
                              'achieved': ach, 'peak': tf32_peak, 'unit': 'TFLOP/s', 'frac': ach / tf32_peak,
                              'peak_source': 'dense TF32 = half of MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)'
                                             if bf16_sustained else 'fallback: half of the nominal 2250 TFLOP/s bf16',
                              'executed': 'rows x 2N x 80 x 2 x 3 (3xTF32: Ah Bh + Al Bh + Ah Bl) / stage-B time',
                              'fp32_equivalent_tflops': ach / 3,
                              'reference_equivalent_tflops': PB / (pms * 1e-3) * 440e6 / 1e12,
                              'reference_equivalent_note': 'SURVEY 8d: 440 MFLOP per pair of the reference\'s FFT formulation, whole call',
                              'hbm': {'algorithmic_bytes_per_pair': 260700,
                                      'achieved_gbs': PB / (pms * 1e-3) * 260700 / 1e9, 'peak_gbs': peaks()[0],
                                      'frac': PB / (pms * 1e-3) * 260700 / 1e9 / peaks()[0]},
                              'traffic': None},
                 'note': 'includes the scattering transform of the FHR channel and stage A of both channels'}
        del pm
        # production dataset step (create_hdf5_dataset.py:360-441): S + 44 within + 130 cross pairs in one pass
        from tebscat.synth import ctg_batch as _ctg
        pmp = KymatioPhaseScattering1D(J=11, Q=4, T=16, shape=5760, device=dev, max_order=1)
        sel = pmp.get_optimal_coefficients_for_fhr(11, 4, 16)['recommendations']
        xp = _ctg(512, 5760, seed=99).to(dev)
        torch.cuda.empty_cache()
        pmp.forward_dataset(xp, sel['use_phase_mask'], sel['use_cross_mask'])
        torch.cuda.synchronize()
        p0.record()
        for _ in range(3):
            pmp.forward_dataset(xp, sel['use_phase_mask'], sel['use_cross_mask'])
        p1.record()
        torch.cuda.synchronize()
        dms = p0.elapsed_time(p1) / 3
        phase['dataset_step'] = {'metric': 'production dataset step segments/s (J=11,Q=4,T=16,N=5760: S + 44 within + 130 cross pairs)',
                                 'value': 512 / (dms * 1e-3), 'unit': 'segments/s', 'batch': 512, 'ms': dms,
                                 'stage_b': 'transform form' if pmp._dev_plan(local).uses_fft_pairs else 'dense operator'}
        del pmp

    # ---- secondary: backward pass on a bounded batch ---------------------------------
    backward = None
    if rank == 0 and not args.no_phase:
        xg = x_dev[:1024].reshape(2048, N).clone().requires_grad_(True)
        og, _ = S(xg)
        wg = torch.ones_like(og)
        og.backward(wg)                                               # builds the op-list plan, warms up
        torch.cuda.synchronize()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(3):
            xg.grad = None
            og, _ = S(xg)
            og.backward(wg)
        b1.record()
        torch.cuda.synchronize()
        bms = b0.elapsed_time(b1) / 3
        backward = {'metric': 'Scattering1D forward + backward signals/s (J=6,Q=8,N=4800)', 'value': 2048 / (bms * 1e-3),
                    'unit': 'signals/s', 'batch': 2048, 'ms': bms,
                    'note': 'forward = the fused launch, backward = recompute + transposed cascade (tebscat/large.py)'}
        del xg, og, wg

    if rank == 0:
        hbm_peak, peak_src = peaks()
        kernel_ms = sum(per_launch_ms) / len(per_launch_ms)
        achieved_gbs = n_sig * BYTES_PER_SIGNAL / (kernel_ms * 1e-3) / 1e9
        fp32 = ctypes.c_double(0.0)
        _lib.load().tebscat_bench_fp32_peak(local, ctypes.byref(fp32))
        achieved_tf = n_sig * FLOPS_PER_SIGNAL / (kernel_ms * 1e-3) / 1e12
        cores = os.cpu_count() or 1
        cpu_rate, cpu_dt = cpu_port_rate(1024, cores)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': total_ms_max / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'Scattering1D J=6 Q=8 T=64 N=4800 orders 0-2, batch 8192 two-channel '
                                   'signals (16384 signals) per GPU -- BASELINE configs[1]',
                       'signals_per_gpu': n_sig, 'l2': 'inputs+outputs 934 MB per step, larger than L2',
                       'parallelism': 'batch-sharded x%d, no collective' % world,
                       'host_binding': ('NUMA-local, %d CPUs per rank' % local_cpus) if prev_affinity is not None else 'none'},
            'clocks': sampler.summary(),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': n_sig * N * 4,
                    'd2h_bytes_per_step': n_sig * C * n_out * 4,
                    'host_copy_ceiling': {'value': copy_value, 'unit': UNIT,
                                          'gbs': copy_value * BYTES_PER_SIGNAL / 1e9,
                                          'note': 'the same pinned H2D + D2H chunks on the same streams with the kernel '
                                                  'left out, all ranks at once (tebscat_scat1d_host_copies_only)'},
                    'frac_of_copy_ceiling': e2e_value / copy_value,
                    'pipeline': '3 slots of up to 1184 signals (chunk sizes ramp 148, 296, 592 up and down at the ends): H2D, kernel and D2H on separate streams'},
            'gpu_launches': launches,
            # 554 flop/B -- the cascade is bound by the FP32 pipe (with shared-memory bandwidth as the
            # co-limit), not by HBM; the HBM figures are reported beside it
            'roofline': {'bound': 'fp32', 'achieved': achieved_tf, 'peak': fp32.value, 'unit': 'TFLOP/s',
                         'frac': achieved_tf / fp32.value if fp32.value else None,
                         'traffic': measured_traffic(n_sig),
                         'flops_per_signal': FLOPS_PER_SIGNAL,
                         'peak_source': 'FMA microbenchmark in this run (tebscat_bench_fp32_peak)',
                         'kernel': 'scat1d_kernel', 'kernel_ms': kernel_ms,
                         'note': 'reference-equivalent FFT flops (5 L log2 L per transform the reference performs) / launch '
                                 'duration; the kernel itself executes ~25% fewer (pair-packed forward transforms)',
                         'hbm': {'bound': 'hbm', 'achieved': achieved_gbs, 'peak': hbm_peak, 'unit': 'GB/s',
                                 'frac': achieved_gbs / hbm_peak, 'algorithmic_bytes': n_sig * BYTES_PER_SIGNAL,
                                 'peak_source': peak_src}},
            'cpu_baseline': {'value': cpu_rate, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                             'sample': '1024 CTG signals in batches of 64, %.1f s, torch-CPU port of the reference (oracle/scattering1d_torch_port.py)' % cpu_dt},
            'phase': phase,
            'backward': backward,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='tebscat', choices=['tebscat', 'reference'])
    ap.add_argument('--no-phase', action='store_true', help='skip the secondary phase-scattering measurement')
    ap.add_argument('--phase-batch', type=int, default=8192, help='two-channel samples of the phase secondary (BASELINE configs[2]: 8192)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'tebscat' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
