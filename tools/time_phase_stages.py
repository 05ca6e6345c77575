"""Stage times of the cross-channel phase path at the headline configuration (events around the stages of every chunk,
tebscat_phase_plan_profile).  A/B a library variant with TEBSCAT_LIB=...   usage: time_phase_stages.py [batch]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
import torch
from tebscat import KymatioPhaseScattering1D, _lib
from tebscat.synth import ctg_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
m = KymatioPhaseScattering1D(J=6, Q=8, T=64, shape=4800, device=torch.device('cuda'))
x = ctg_batch(B, 4800, seed=1).cuda()
m(x[:512], compute_phase=False, compute_cross_phase=True); torch.cuda.synchronize()
ph, lib = m._dev_plan(0).handle, _lib.load()
for rep in range(2):
    _lib.check(lib.tebscat_phase_plan_profile(ph, 1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        m(x, compute_phase=False, compute_cross_phase=True)
    e1.record(); torch.cuda.synchronize()
    a, b, n = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_int(0)
    _lib.check(lib.tebscat_phase_plan_profile_read(ph, ctypes.byref(a), ctypes.byref(b), ctypes.byref(n)))
    _lib.check(lib.tebscat_phase_plan_profile(ph, 0))
print('B=%d: whole call %.2f ms (%.0f pairs/s) | stage A %.2f ms, stage B %.2f ms over %d chunks' % (
    B, e0.elapsed_time(e1) / 3, B / (e0.elapsed_time(e1) / 3) * 1e3, a.value / 3, b.value / 3, n.value // 3))
