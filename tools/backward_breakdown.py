"""Where does the backward pass spend its time?  Per-op-class GPU time (events around every op)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
import torch
from tebscat import Scattering1D, _lib
J, N, Q, T, B = (int(v) for v in sys.argv[1:6])
S = Scattering1D(J, N, Q, T=T).cuda()
x = torch.randn(B, N, device='cuda', requires_grad=True)
out, _ = S(x); out.sum().backward(); torch.cuda.synchronize()
lib = _lib.load()
acc = collections.defaultdict(float); cnt = collections.Counter()
def wrap(name):
    f = getattr(lib, name)
    def g(*a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rc = f(*a); e1.record(); torch.cuda.synchronize()
        key = name.replace('tebscat_large_', '')
        if key == 'fft': key += '(2^%d%s)' % (a[3], ' inv' if a[4] else '')
        if key in ('mulfold', 'unfold'): key += '(src 2^%d, k 2^%d)' % (a[5], a[6])
        if key in ('modulus_to', 'modulus_backward'): key += '(%d)' % (a[3] // B)
        acc[key] += e0.elapsed_time(e1); cnt[key] += 1
        return rc
    return g
dp = list(S._lplans.values())[0]
class P:
    def __getattr__(self, n):
        return wrap(n) if n.startswith('tebscat_large_') else getattr(lib, n)
dp._lib = P()
g = torch.empty_like(x)
w = torch.randn(B, dp.plan.n_paths, dp.plan.n_out, device='cuda')
dp._run_backward(x.detach().contiguous(), w, g); torch.cuda.synchronize()
tot = sum(acc.values())
for k, v in sorted(acc.items(), key=lambda kv: -kv[1])[:30]:
    print('%-40s %4d calls %8.2f ms  %5.1f%%' % (k, cnt[k], v, 100 * v / tot))
print('total %.1f ms, %d launches-ish calls' % (tot, sum(cnt.values())))
