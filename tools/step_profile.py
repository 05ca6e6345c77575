"""Per-step cycle counts of the headline schedule on the GPU (calibrates schedule.py's cost model).

    python tools/step_profile.py [config] > gpurun_out/steps.json
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
import numpy as np                              # noqa: E402
import torch                                    # noqa: E402
from tebscat import Scattering1D, _lib          # noqa: E402
from tebscat.synth import ctg_batch             # noqa: E402

CFG = {'H': (6, 4800, 8, 64, 2), 'P': (11, 5760, 4, 16, 1), 'K': (6, 512, 16, 64, 2), 'T': (5, 700, 2, 8, 2)}
name = sys.argv[1] if len(sys.argv) > 1 else 'H'
J, N, Q, T, mo = CFG[name]
S = Scattering1D(J, N, Q, max_order=mo, T=T).cuda()
B = 148 * 4
x = ctg_batch(B // 2, N, seed=3).reshape(-1, N)[:B].cuda()
out, _ = S(x)
torch.cuda.synchronize()
plan = S._plan_for(0)
sched = S._sched[1]
clk = np.zeros(sched.steps.shape[0] + 1, dtype=np.int64)
best = None
for rep in range(5):
    rc = _lib.load().tebscat_scat1d_profile_steps(plan.handle, x.data_ptr(), B, out.data_ptr(),
                                                  clk.ctypes.data, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)
    d = np.diff(clk)
    best = d if best is None else np.minimum(best, d)
steps = []
for (b, e), cyc in zip(sched.steps, best):
    steps.append({'cycles': int(cyc), 'tasks': [[int(v) for v in sched.tasks[i]] for i in range(b, e)]})
json.dump({'config': name, 'total_cycles': int(best.sum()), 'n_steps': len(steps), 'steps': steps}, sys.stdout)
