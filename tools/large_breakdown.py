"""Where does the large-support level spend its time?  Per-op-class GPU time (events around every op, graph off)."""
import os, sys, ctypes, collections
os.environ['TEBSCAT_LARGE_GRAPH'] = '0'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
import torch
from tebscat import Scattering1D, _lib
J, N, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
S = Scattering1D(J, N, 8, T=2 ** J).cuda()
x = torch.randn(B, N, device='cuda')
S(x); torch.cuda.synchronize()
lib = _lib.load()
acc = collections.defaultdict(float); cnt = collections.Counter()
def wrap(name):
    f = getattr(lib, name)
    def g(*a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rc = f(*a); e1.record(); torch.cuda.synchronize()
        key = name.replace('tebscat_large_', '')
        if name == 'tebscat_scat1d_forward_gsrc': key = 'fused subtrees (source 2^%d)' % (int(a[2]).bit_length() - 1)
        if key in ('fft', 'pair'): key += '(2^%d)' % a[3]
        if key == 'mulfold': key += '(src 2^%d, k 2^%d)' % (a[5], a[6])
        acc[key] += e0.elapsed_time(e1); cnt[key] += 1
        return rc
    return g
class L: pass
proxy = L()
for n in dir(lib): pass
dp = list(S._lplans.values())[0]
class P:
    def __getattr__(self, n):
        return wrap(n) if n.startswith('tebscat_large_') or n == 'tebscat_scat1d_forward_gsrc' else getattr(lib, n)
dp._lib = P()
out = torch.empty(B, dp.plan.n_paths, dp.plan.n_out, device='cuda')
dp._run(x, out); torch.cuda.synchronize()
tot = sum(acc.values())
for k, v in sorted(acc.items(), key=lambda kv: -kv[1])[:25]:
    print('%-34s %4d calls %8.2f ms  %5.1f%%' % (k, cnt[k], v, 100 * v / tot))
print('total %.1f ms' % tot)
