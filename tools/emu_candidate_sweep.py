"""CPU-only robustness sweep of the schedule CANDIDATES: random configurations with padded lengths up
to 2^13; every combination of chain priority (depth first / by weight), reservation rule ('sum' / 'peak') and U0 layout
(global scratch / shared memory) that yields a schedule is run through the host emulator and compared with the float64
oracle (1e-5 per path) and, bit for bit, with the round-1 rules."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'vae-teb_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import numpy as np
from helpers import emu_forward
from oracle.scattering1d_oracle import ScatteringOracle
from tebscat.schedule import build_plan

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
count = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rng = np.random.default_rng(seed)
done = bad = 0
t_all = time.time()
while done < count:
    J = int(rng.integers(2, 10))
    Q = int(rng.choice([1, 2, 4, 8, 12, 16]))
    T = int(2 ** rng.integers(max(J - 5, 1), J + 1))
    N = int(rng.integers(100, 7000))
    mo = int(rng.choice([1, 2, 2]))
    os_ = int(rng.choice([0, 0, 0, 1, 2]))
    try:
        orc = ScatteringOracle(J, N, Q, T, mo, oversampling=os_)
        if orc.geo['J_pad'] > 13:
            continue
    except Exception:
        continue
    x = rng.standard_normal((1, N)).astype(np.float32)
    ref = orc(x)
    nr = np.linalg.norm(ref, axis=-1)
    base, line, ok = None, [], True
    for scratch in (False, 'scratch'):
        for rs in ('sum', 'peak'):
            for dw in (1e12, 0.0):
                try:
                    p = build_plan(J, N, Q, T, mo, oversampling=os_, tune=dict(u0_scratch=scratch, reserve=rs, depth_weight=dw))
                except NotImplementedError:
                    line.append('-')
                    continue
                out = emu_forward(p, x)
                err = np.linalg.norm(out.astype(np.float64) - ref, axis=-1)
                good = bool(np.all(err <= 1e-5 * nr + 1e-10 * nr.max())) and bool(np.isfinite(out).all())
                if base is None:
                    base = out
                same = bool(np.array_equal(out, base))
                ok = ok and good and same
                line.append('%d%s' % (p.stats['n_steps'], '' if good and same else '!'))
    done += 1
    bad += not ok
    print((J, Q, T, N, mo, os_), 'J_pad', orc.geo['J_pad'], 'C', ref.shape[1], 'steps', ' '.join(line), 'OK' if ok else 'FAIL', flush=True)
print('%d configurations, %d failures, %.0f s' % (done, bad, time.time() - t_all))
sys.exit(1 if bad else 0)
