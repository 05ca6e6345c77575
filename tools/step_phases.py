"""Where does a step spend its time?  Needs a library built with -DTEBSCAT_PROF_PHASES
(TEBSCAT_LIB=build/variants/lib_phases.so): per step the clock of thread 0 at step start, after
the task record is decoded, after the task body and after the barrier."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
import numpy as np, torch
from tebscat import Scattering1D, _lib
from tebscat.synth import ctg_batch

J, N, Q, T, mo = 6, 4800, 8, 64, 2
S = Scattering1D(J, N, Q, max_order=mo, T=T).cuda()
B = 148 * 4
x = ctg_batch(B // 2, N, seed=3).reshape(-1, N)[:B].cuda()
out, _ = S(x); torch.cuda.synchronize()
plan = S._plan_for(0); sched = S._sched[1]
ns = sched.steps.shape[0]
clk = np.zeros(4 * ns + 1, dtype=np.int64)
rows = None
for rep in range(5):
    _lib.check(_lib.load().tebscat_scat1d_profile_steps(plan.handle, x.data_ptr(), B, out.data_ptr(), clk.ctypes.data,
                                                        torch.cuda.current_stream().cuda_stream))
    start = clk[:ns]; ph = clk[ns + 1:].reshape(ns, 3)
    r = np.stack([ph[:, 0] - start, ph[:, 1] - ph[:, 0], ph[:, 2] - ph[:, 1], clk[1:ns + 1] - start], 1)
    # keep the repetition with the smallest total (per-phase minima over repetitions do not add up to the step:
    # a warp that finishes its body early waits longer at the barrier)
    rows = r if rows is None or r[:, 3].sum() < rows[:, 3].sum() else rows
OPS = {0: 'NOP', 1: 'LOAD', 2: 'FFT', 3: 'MULFOLD', 4: 'STOREB', 5: 'STOREZ', 6: 'TINY', 7: 'MULFOLD2'}
tot = rows.sum(0)
print('totals: decode %d  body(warp0) %d  barrier wait(warp0) %d  step %d' % tuple(tot))
for s in range(min(ns, int(sys.argv[1]) if len(sys.argv) > 1 else 70)):
    b, e = sched.steps[s]
    t0 = [t for t in sched.tasks[b:e] if t[1] == 0]
    desc = [(OPS[t[0] & 255], int(t[2]), [int(v) for v in t[3:8]]) for t in sched.tasks[b:e]]
    print(s, rows[s].tolist(), desc[:3])
