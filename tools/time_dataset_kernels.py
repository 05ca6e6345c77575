"""Per-kernel times of the production dataset step (run under ncu --metrics gpu__time_duration.sum)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
import torch
from tebscat import KymatioPhaseScattering1D
from tebscat.synth import ctg_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
m = KymatioPhaseScattering1D(J=11, Q=4, T=16, shape=5760, device=torch.device('cuda'), max_order=1)
sel = m.get_optimal_coefficients_for_fhr(11, 4, 16)['recommendations']
x = ctg_batch(B, 5760, seed=12).cuda()
for _ in range(2):
    m.forward_dataset(x, sel['use_phase_mask'], sel['use_cross_mask'])
torch.cuda.synchronize()
