"""One-off robustness sweep on the GPU: random (J, Q, T, N, max_order, oversampling) configurations -- forward against the
float64 oracle, gradients against the autograd oracle, the dense (tcgen05) phase path against the phase oracle.

The reference's own fp32 theta flips sign wherever an analytic sample lands on the negative real axis up to rounding
(SURVEY 8c).  PhaseOracle.align_branches offers the other branch at the two boundary samples (where reflect padding makes
it systematic) and at interior samples with |Im z| < 2e-6 |Re z|; for any phase case above 5e-5 the tool prints what is
needed to tell a defect from such a flip (repeatability, a fresh plan, the nearest negative-real analytic sample)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
import numpy as np, torch
from tebscat import Scattering1D, KymatioPhaseScattering1D
from oracle.scattering1d_oracle import ScatteringOracle
from oracle.scattering1d_grad_oracle import GradOracle
from oracle.phase_oracle import PhaseOracle

rng = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_fwd, n_bwd, n_ph = (int(v) for v in (sys.argv[2:5] if len(sys.argv) > 4 else (30, 8, 8)))
bad = 0

def config():
    while True:
        J = int(rng.randint(2, 9)); Q = int(rng.choice([1, 2, 4, 8, 12])); N = int(rng.randint(200, 7000 if rng.rand() < 0.8 else 20000))
        T = int(2 ** rng.randint(1, J + 1)); mo = int(rng.choice([1, 2])); os_ = int(rng.choice([0, 0, 0, 1, 2]))
        try:
            S = Scattering1D(J, N, Q, max_order=mo, T=T, oversampling=os_)
            return J, Q, T, N, mo, os_, S
        except (ValueError, NotImplementedError, AssertionError):
            continue

t0 = time.time()
for k in range(n_fwd):
    J, Q, T, N, mo, os_, S = config()
    S = S.cuda()
    x = torch.randn(int(rng.randint(1, 10)), N, generator=torch.Generator().manual_seed(k))
    out = S(x.cuda())[0].cpu().numpy().astype(np.float64)
    ref = ScatteringOracle(J, N, Q, T, mo, oversampling=os_)(x.numpy())
    nr = np.linalg.norm(ref, axis=-1); err = np.linalg.norm(out - ref, axis=-1)
    ok = out.shape == ref.shape and np.all(err <= 1e-5 * nr + 1e-10 * nr.max())
    bad += not ok
    print('fwd', (J, Q, T, N, mo, os_), 'op-by-op' if S._op_by_op else 'fused', 'J_pad', S.J_pad, 'C', out.shape[1], 'worst %.2e' % float((err / np.maximum(nr, 1e-30)).max()), 'OK' if ok else 'FAIL', flush=True)
for k in range(n_bwd):
    J, Q, T, N, mo, os_, S = config()
    S = S.cuda()
    x = torch.randn(int(rng.randint(1, 8)), N, generator=torch.Generator().manual_seed(100 + k)).cuda().requires_grad_(True)
    out, _ = S(x)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(200 + k))
    (out * w.cuda()).sum().backward()
    _, g64 = GradOracle(J, N, Q, T, mo, os_).vjp(x.detach().cpu().numpy(), w.numpy())
    e = np.linalg.norm(x.grad.cpu().numpy() - g64, axis=-1) / np.linalg.norm(g64, axis=-1)
    ok = e.max() < 1e-5
    bad += not ok
    print('bwd', (J, Q, T, N, mo, os_), 'J_pad', S.J_pad, 'worst %.2e' % e.max(), 'OK' if ok else 'FAIL', flush=True)
n_un = int(sys.argv[5]) if len(sys.argv) > 5 else 0
for k in range(n_un):                                   # average=False (un-averaged moduli) where the fused schedule exists
    while True:
        J, Q, T, N, mo, os_, S = config()
        try:
            if S.J_pad > 13:
                continue
            Su = Scattering1D(J, N, Q, max_order=mo, T=T, oversampling=os_, average=False, out_type='list').cuda()
            x = torch.randn(2, N, generator=torch.Generator().manual_seed(400 + k))
            out = Su(x.cuda())[0]
            break
        except NotImplementedError:
            continue
    ref = ScatteringOracle(J, N, Q, T, mo, oversampling=os_).unaveraged(x.numpy())
    worst, ok = 0.0, len(out) == len(ref)
    for o, (key, v) in zip(out, ref):
        got = o['coef'].cpu().numpy().astype(np.float64)
        if got.shape != v.shape:
            ok = False
            break
        e = (np.linalg.norm(got - v, axis=-1) / np.maximum(np.linalg.norm(v, axis=-1), 1e-30)).max()
        worst = max(worst, float(e))
    ok = ok and worst < 1e-5
    # and its backward (round 2): gradient of sum(coef_c * w_c) over the whole list output against the autograd oracle
    xg = x.cuda().requires_grad_(True)
    outg = Su(xg)[0]
    gen = torch.Generator().manual_seed(500 + k)
    ws = [torch.randn(o['coef'].shape, generator=gen) for o in outg]
    sum((o['coef'] * w.cuda()).sum() for o, w in zip(outg, ws)).backward()
    _, g64 = GradOracle(J, N, Q, T, mo, os_).vjp_unaveraged(x.numpy(), torch.cat(ws[1:], dim=-1).numpy())
    g64 = g64 + ws[0].numpy()
    eg = float((np.linalg.norm(xg.grad.cpu().numpy() - g64, axis=-1) / np.linalg.norm(g64, axis=-1)).max())
    ok = ok and eg < 1e-5
    bad += not ok
    print('unaveraged', (J, Q, T, N, mo, os_), 'paths', len(out), 'worst %.2e' % worst, 'grad %.2e' % eg, 'OK' if ok else 'FAIL', flush=True)
for k in range(n_ph):
    while True:
        J = int(rng.randint(3, 8)); Q = int(rng.choice([2, 4, 8])); T = int(2 ** rng.randint(2, J + 1))
        N = int(rng.randint(400, 5000)) if rng.rand() < 0.75 else int(rng.randint(6000, 14000))     # a quarter beyond 2^13 padded
        border = str(rng.choice(['reflect', 'reflect', 'constant', 'circular']))
        over = int(rng.choice([0, 0, 0, 1, 8]))                # 8 >= log2 T: no decimation (target length = N)
        try:
            m = KymatioPhaseScattering1D(J=J, Q=Q, T=T, shape=N, device=torch.device('cuda'), border_mode=border, oversampling=over)
            break
        except (ValueError, NotImplementedError, AssertionError):
            continue
    x = torch.randn(int(rng.randint(1, 6)), 2, N, generator=torch.Generator().manual_seed(300 + k))
    ours = m(x.cuda(), compute_phase=False, compute_cross_phase=True)['cross_phase_corr'].cpu().numpy().astype(np.float64)
    o = PhaseOracle(J, Q, T, N, m.scattering(x[:, 0].cuda().contiguous())[0].shape[-1], border_mode=border)
    ref = o.align_branches(x.numpy(), ours, mode='cross')
    rel = np.linalg.norm(ours - ref) / np.linalg.norm(ref)
    ok = ours.shape == ref.shape and rel < 5e-5
    bad += not ok
    wi = m(x.cuda(), compute_phase=True, phase_channels=[1])['phase_corr'].cpu().numpy().astype(np.float64)
    rw = o.align_branches(x.numpy()[:, 1], wi, mode='within')
    full = m(x.cuda(), compute_phase=False, compute_cross_phase=True, cross_phase_low_pass=False, cross_phase_same_pairs_only=True)['cross_phase_corr'].cpu().numpy()
    rf = o(x.numpy(), mode='cross', pair_subset=o.autoc_idx, low_pass=False)
    ok = ok and np.linalg.norm(wi - rw) / np.linalg.norm(rw) < 5e-5 and full.shape == rf.shape
    bad += (not ok) and rel < 5e-5
    print('phase', (J, Q, T, N, border, over), 'J_pad', m.J_pad, 'dec', m._plan.dec, 'pairs', ours.shape[1], 'n_out', ours.shape[2],
          'rel %.2e' % rel, 'OK' if ok else 'FAIL', flush=True)
    if rel >= 5e-5:                                            # same input again, same module and a fresh one: is it the input?
        again = m(x.cuda(), compute_phase=False, compute_cross_phase=True)['cross_phase_corr'].cpu().numpy().astype(np.float64)
        m2 = KymatioPhaseScattering1D(J=J, Q=Q, T=T, shape=N, device=torch.device('cuda'), border_mode=border, oversampling=over)
        fresh = m2(x.cuda(), compute_phase=False, compute_cross_phase=True)['cross_phase_corr'].cpu().numpy().astype(np.float64)
        pp = np.linalg.norm(ours - ref, axis=-1) / np.linalg.norm(ref, axis=-1)
        b, p = np.unravel_index(pp.argmax(), pp.shape)
        t = int(np.abs(ours - ref)[b, p].argmax())
        zi = o.analytic(x.numpy()[:, 0])[b, o.i_idx[p]]
        w = slice(max(t - 12, 0), t + 13)
        tt = int(np.argmin(np.where(zi[w].real < 0, np.abs(zi[w].imag) / np.abs(zi[w]), 1.0))) + w.start
        print('   input seed', 300 + k, 'rows', x.shape[0], '| same call again: identical' if np.array_equal(again, ours) else '| same call again: DIFFERENT',
              '| fresh module rel %.2e' % (np.linalg.norm(fresh - ref) / np.linalg.norm(ref)), '| bad paths', int((pp > 1e-4).sum()),
              'worst', (int(b), int(p)), 'i,j', int(o.i_idx[p]), int(o.j_idx[p]), 'peak err at t', t,
              '| nearest negative-real z_i: t', tt, 're %.3e im %.3e' % (zi[tt].real, zi[tt].imag), flush=True)
print('done in %.0f s, failures: %d' % (time.time() - t0, bad))
