"""Small driver for ncu: one cross-channel phase forward of the headline config."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
import torch                                            # noqa: E402
from tebscat import KymatioPhaseScattering1D            # noqa: E402
from tebscat.synth import ctg_batch                     # noqa: E402
B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
m = KymatioPhaseScattering1D(J=6, Q=8, T=64, shape=4800, device=torch.device('cuda'))
x = ctg_batch(B, 4800, seed=3).cuda()
for _ in range(2):
    r = m(x, compute_phase=False, compute_cross_phase=True)
torch.cuda.synchronize()
print('ok', tuple(r['cross_phase_corr'].shape))
