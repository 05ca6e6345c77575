"""A/B scheduler knobs on the GPU: signals/s of the scattering kernel for each variant of build_plan(tune=...).

    python tools/sweep_sched.py [config]          (config: H or P)
"""
import itertools, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
import numpy as np, torch
from tebscat import _lib
from tebscat.schedule import build_plan
from tebscat.torch_frontend import _DevicePlan
from tebscat.synth import ctg_batch

CFG = {'H': (6, 4800, 8, 64, 2), 'P': (11, 5760, 4, 16, 1)}
name = sys.argv[1] if len(sys.argv) > 1 else 'H'
J, N, Q, T, mo = CFG[name]
B = 148 * 8
x = ctg_batch(B // 2, N, seed=3).reshape(-1, N)[:B].cuda()
lib = _lib.load()


def rate(tune):
    try:
        p = build_plan(J, N, Q, T, mo, tune=tune)
    except Exception as e:
        return None, str(e)[:60]
    dp = _DevicePlan(p, 0)
    out = torch.empty(B, p.n_paths, p.n_out, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    best = 1e9
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.tebscat_scat1d_forward(dp.handle, x.data_ptr(), B, out.data_ptr(), st))
        e1.record(); torch.cuda.synchronize()
        if rep:
            best = min(best, e0.elapsed_time(e1))
    return B / best * 1e3, p.stats['n_steps']


grid = {
    'batch_slots': [8192],
    'child_slots': [2048, 4096, 8192],
    'pool_slots': [512, 1024, 2048],
    'pack_gain': [0.9, 0.97, 1.05],
    'open_demand': [1.0, 2.0, 4.0],
}
if len(sys.argv) > 2:
    grid = json.loads(sys.argv[2])
keys = list(grid)
base, _ = rate({})
print('default: %.0f signals/s' % base)
res = []
for vals in itertools.product(*[grid[k] for k in keys]):
    tune = dict(zip(keys, vals))
    r, ns = rate(tune)
    res.append((r or 0, ns, tune))
    print('%9.0f  steps %s  %s' % (r or 0, ns, tune), flush=True)
res.sort(key=lambda t: -t[0])
print('best:', res[:5])
