"""Timeline of one CTA of the tcgen05 phase kernel (library built with -DTEBSCAT_TC_TRACE; TEBSCAT_LIB points at it):
who waits for whom -- producer groups (stage free / inputs landed / products stored / slab handed over), the MMA thread
(accumulator free / slab full / issued) and one epilogue warp (accumulator full / drained)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
import numpy as np, torch
from tebscat import KymatioPhaseScattering1D, _lib
from tebscat.synth import ctg_batch
m = KymatioPhaseScattering1D(J=6, Q=8, T=64, shape=4800, device=torch.device('cuda'))
x = ctg_batch(256, 4800, seed=1).cuda()
for _ in range(2):
    m(x, compute_phase=False, compute_cross_phase=True)
torch.cuda.synchronize()
lib = _lib.load()
NS, EV = 64, 6
t = np.zeros(6 * NS * EV, np.int64)
lib.tebscat_debug_tc_trace(t.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(t.size))
t = t.reshape(6, NS, EV)
t0 = t[t > 0].min()
r = lambda v: int(v - t0) if v > 0 else -1
print('producer groups: slab | wait-start, stage free, inputs landed, products stored, handed over  (cycles since the first event)')
for i in range(16, 40):
    g = i % 4
    print('  slab %2d group %d: %s   | wait %5d  inputs %5d  compute %5d  B+st-wait %5d  group barrier %5d  next copies issued %5d' % (
        i, g, ' '.join('%7d' % r(v) for v in t[g, i, :5]), t[g, i, 1] - t[g, i, 0], t[g, i, 2] - t[g, i, 1], t[g, i, 3] - t[g, i, 2], t[g, i, 4] - t[g, i, 3],
        t[g, i, 5] - t[g, i, 4], t[g, i + 4, 0] - t[g, i, 5]))
print('MMA issuer (kind 0, parity 0): drain group | start, accumulator free, slab 0 full, slab 0 issued, slab 1 full, slab 1 issued')
for g in range(8, 24, 2):
    print('  group %2d: %s   | acc wait %5d  full wait %5d  issue %5d  full wait %5d  issue %5d' % (g, ' '.join('%7d' % r(v) for v in t[4, g, :6]),
          t[4, g, 1] - t[4, g, 0], t[4, g, 2] - t[4, g, 1], t[4, g, 3] - t[4, g, 2], t[4, g, 4] - t[4, g, 3], t[4, g, 5] - t[4, g, 4]))
print('epilogue warp 0: group | wait-start, accumulator full, drained')
for g in range(8, 20):
    print('  group %2d: %s   | wait %5d  drain %5d' % (g, ' '.join('%7d' % r(v) for v in t[5, g, :3]), t[5, g, 1] - t[5, g, 0], t[5, g, 2] - t[5, g, 1]))
per_slab = (t[4, 24, 5] - t[4, 8, 5]) / 32
print('steady state: %.0f cycles per slab' % per_slab)
