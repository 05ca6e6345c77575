"""Time the phase path (H and production configs) on the GPU."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
import torch
from tebscat import KymatioPhaseScattering1D
from tebscat.synth import ctg_batch

def timeit(f, n=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
m = KymatioPhaseScattering1D(J=6, Q=8, T=64, shape=4800, device=torch.device('cuda'))
x = ctg_batch(B, 4800, seed=1).cuda()
ms = timeit(lambda: m(x, compute_phase=False, compute_cross_phase=True))
print('H cross  B=%d: %.1f ms -> %.0f pairs-of-signals/s' % (B, ms, B / ms * 1e3))
ms = timeit(lambda: m(x, compute_phase=True, phase_channels=[0]))
print('H within B=%d: %.1f ms -> %.0f signals/s' % (B, ms, B / ms * 1e3))
ms = timeit(lambda: m.scattering(x[:, 0].contiguous()))
print('H scattering only B=%d: %.2f ms' % (B, ms))
mp = KymatioPhaseScattering1D(J=11, Q=4, T=16, shape=5760, device=torch.device('cuda'), max_order=1)
sel = mp.get_optimal_coefficients_for_fhr(11, 4, 16)
pm, cm = sel['recommendations']['use_phase_mask'], sel['recommendations']['use_cross_mask']
xp = ctg_batch(B, 5760, seed=2).cuda()
def dataset_step():
    a = mp(xp, compute_phase=True, phase_channels=[0], phase_pairs=pm)
    b = mp(xp, compute_phase=False, compute_cross_phase=True, phase_channels=[0, 1], phase_pairs=cm)
    return a, b
ms = timeit(dataset_step)
print('P dataset step (S + 44 within + 130 cross) B=%d: %.1f ms -> %.0f segments/s' % (B, ms, B / ms * 1e3))
ms = timeit(lambda: mp(xp, compute_phase=False, compute_cross_phase=True, phase_channels=[0, 1]))
print('P cross all 903 pairs B=%d: %.1f ms -> %.0f /s' % (B, ms, B / ms * 1e3))
ms = timeit(lambda: mp.forward_dataset(xp, pm, cm))
print('P single-pass dataset entry B=%d: %.1f ms -> %.0f segments/s' % (B, ms, B / ms * 1e3))
