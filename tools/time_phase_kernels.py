"""Per-kernel times of the cross-channel phase path at the headline configuration (run under ncu --metrics gpu__time_duration.sum)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
import torch
from tebscat import KymatioPhaseScattering1D
from tebscat.synth import ctg_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = KymatioPhaseScattering1D(J=6, Q=8, T=64, shape=4800, device=torch.device('cuda'))
x = ctg_batch(B, 4800, seed=12).cuda()
for _ in range(2):
    m(x, compute_phase=False, compute_cross_phase=True)
torch.cuda.synchronize()
