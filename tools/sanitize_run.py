"""Tiny driver for compute-sanitizer (racecheck / memcheck): one scattering forward per config on a
handful of signals, the normalisation epilogue, and the phase path in both forms of stage B."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
import numpy as np, torch
from tebscat import Scattering1D, KymatioPhaseScattering1D
from tebscat.synth import ctg_batch
what = sys.argv[1] if len(sys.argv) > 1 else 'scat'
if what == 'scat':
    for (J, N, Q, T, mo) in ((6, 4800, 8, 64, 2), (5, 700, 2, 8, 2), (11, 5760, 4, 16, 1)):
        S = Scattering1D(J, N, Q, max_order=mo, T=T).cuda()
        x = ctg_batch(2, N, seed=1).reshape(-1, N)[:3].cuda()
        out, _ = S(x)
        C = out.shape[1]
        f = S.forward_normalized(x, np.zeros(C), np.ones(C), trim=1)
        torch.cuda.synchronize()
        print('scat', J, N, Q, T, tuple(out.shape), tuple(f.shape), float(out.abs().sum()))
elif what == 'backward':
    for (J, N, Q, T, mo) in ((5, 700, 2, 8, 2), (4, 1000, 4, 16, 2)):
        S = Scattering1D(J, N, Q, max_order=mo, T=T).cuda()
        x = ctg_batch(2, N, seed=1).reshape(-1, N)[:3].cuda().requires_grad_(True)
        out, _ = S(x)
        out.sum().backward()
        torch.cuda.synchronize()
        print('backward', J, N, float(x.grad.abs().sum()))
elif what == 'phase_tc':
    os.environ['TEBSCAT_PHASE_FFT'] = '0'
    for (J, Q, T, N) in ((4, 4, 16, 1000), (4, 4, 16, 999)):
        m = KymatioPhaseScattering1D(J=J, Q=Q, T=T, shape=N, device=torch.device('cuda'))
        x = ctg_batch(2, N, seed=2).cuda()
        r = m(x, compute_phase=False, compute_cross_phase=True)
        torch.cuda.synchronize()
        print(what, N, float(r['cross_phase_corr'].abs().sum()))
else:
    os.environ['TEBSCAT_PHASE_FFT'] = '1' if what == 'phase_fft' else '0'
    m = KymatioPhaseScattering1D(J=11, Q=4, T=16, shape=5760, device=torch.device('cuda'), max_order=1)
    x = ctg_batch(2, 5760, seed=2).cuda()
    sel = m.get_optimal_coefficients_for_fhr(11, 4, 16)
    r = m.forward_dataset(x, sel['recommendations']['use_phase_mask'], sel['recommendations']['use_cross_mask'])
    torch.cuda.synchronize()
    print(what, {k: tuple(v.shape) for k, v in r.items() if hasattr(v, 'shape')})
