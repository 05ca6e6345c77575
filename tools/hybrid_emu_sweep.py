"""Random configurations with padded lengths of 2^14 .. 2^16: the fused subtrees of the large-support level
(schedule.build_hybrid_plans) through the host emulator against the float64 oracle (CPU only).  Every schedule must
write exactly its own channels and meet 1e-5 per path; configurations whose subtrees do not fit are reported."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'vae-teb_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import numpy as np
from helpers import emu_forward_gsrc
from oracle.scattering1d_oracle import ScatteringOracle, reflect_pad, subsample_fourier
from tebscat import filterbank as fbk
from tebscat.schedule import bitrev_indices, build_hybrid_plans

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
count = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rng = np.random.default_rng(seed)
done = bad = 0
while done < count:
    J = int(rng.integers(3, 11))
    Q = int(rng.choice([1, 2, 4, 8, 12]))
    T = int(2 ** rng.integers(max(J - 6, 0), J + 1))
    N = int(rng.integers(6000, 40000))
    mo = int(rng.choice([1, 2, 2]))
    os_ = int(rng.choice([0, 0, 0, 1]))
    try:
        geo = fbk.build_geometry(N, J, fbk._as_Q1(Q), T)
    except Exception:
        continue
    if not 14 <= geo.J_pad <= 16:
        continue
    t0 = time.time()
    try:
        hyb = build_hybrid_plans(J, N, Q, T, mo, os_)
    except NotImplementedError as e:
        print((J, Q, T, N, mo, os_), 'NotImplemented:', e)
        continue
    orc = ScatteringOracle(J, N, Q, T, mo, oversampling=os_)
    x = rng.standard_normal((1, N)).astype(np.float32)
    ref = orc(x)
    g = orc.geo
    n = g['J_pad']
    U0 = np.fft.fft(reflect_pad(x.astype(np.float64), g['pad_left'], g['pad_right']), axis=-1)
    keys = list(orc.keys)
    chan = {k: c for c, k in enumerate(keys)}
    log2_T = int(np.floor(np.log2(T)))
    worst, parts = 0.0, []

    def check(out, want):
        global worst
        got = [c for c in range(len(keys)) if not np.isnan(out[:, c]).any()]
        assert got == sorted(want), (got, sorted(want))
        assert np.isnan(np.delete(out, got, axis=1)).all()
        nr = np.linalg.norm(ref[:, got], axis=-1)
        err = np.linalg.norm(out[:, got].astype(np.float64) - ref[:, got], axis=-1)
        worst = max(worst, float((err / np.maximum(nr, 1e-300)).max()))

    if hyb['first'] is not None:
        out = np.full((1, len(keys), ref.shape[-1]), np.nan, np.float32)
        emu_forward_gsrc(hyb['first'], U0[:, bitrev_indices(1 << n)], out)
        small = set(hyb['first_n1'])
        check(out, [c for c, k in enumerate(keys) if len(k) >= 1 and k[0] in small])
        parts.append('first %d n1 / %d steps' % (len(small), hyb['first'].stats['n_steps']))
    for plan, members, kids_done in hyb['kids']:
        for n1 in {members[0], members[-1]}:
            p1 = orc.psi1[n1]
            k1 = max(min(p1['j'] - os_, log2_T - os_), 0)
            u1 = np.abs(np.fft.ifft(subsample_fourier(U0 * p1['levels'][0], 2 ** k1), axis=-1))
            U1 = np.fft.fft(u1, axis=-1)[:, bitrev_indices(1 << (n - k1))]
            out = np.full((1, len(keys), ref.shape[-1]), np.nan, np.float32)
            emu_forward_gsrc(plan, U1, out, chan[(n1, kids_done[0])] - chan[(plan.head, kids_done[0])])
            check(out, [chan[(n1, n2)] for n2 in kids_done])
        parts.append('kids of %d filters' % len(members))
    ok = worst <= 1e-5
    bad += not ok
    done += 1
    print((J, Q, T, N, mo, os_), 'J_pad', n, 'C', len(keys), '; '.join(parts) or 'nothing fused', 'worst %.2e' % worst,
          'OK' if ok else 'FAIL', '%.1fs' % (time.time() - t0), flush=True)
print('%d configurations, %d failures' % (done, bad))
sys.exit(1 if bad else 0)
