"""tcgen05 form of phase stage B against the mma.sync form (two processes: the choice is read once per process)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
import numpy as np

CFGS = {'S': (4, 4, 16, 1000, 2, 3), 'odd': (4, 4, 16, 999, 2, 3), 'H': (6, 8, 64, 4800, 2, 4),
        'manyF': (6, 12, 64, 2000, 2, 3),      # F = 70 filters: two samples of a row tile no longer fit the direct 'j' slot map
        'Sbig': (4, 4, 16, 1000, 2, 901)}      # three stage-A chunks + a remainder, row tiles straddling samples

def child(tag, out):
    import torch
    from tebscat import KymatioPhaseScattering1D
    from tebscat.synth import ctg_batch
    res = {}
    for name, (J, Q, T, N, mo, B) in CFGS.items():
        m = KymatioPhaseScattering1D(J=J, Q=Q, T=T, shape=N, device=torch.device('cuda'), max_order=mo)
        assert not m._dev_plan(0).uses_fft_pairs
        x = ctg_batch(B, N, seed=11).cuda()
        y = m(x, compute_phase=False, compute_cross_phase=True)['cross_phase_corr']
        torch.cuda.synchronize()
        if name == 'H':
            xb = ctg_batch(256, N, seed=12).cuda()
            m(xb, compute_phase=False, compute_cross_phase=True); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                m(xb, compute_phase=False, compute_cross_phase=True)
            e1.record(); torch.cuda.synchronize()
            print(tag, 'H: %.2f ms per 256 samples -> %.0f pairs/s' % (e0.elapsed_time(e1) / 3, 256 / (e0.elapsed_time(e1) / 3e3)), flush=True)
        res[name] = y.cpu().numpy()[:64]
        if name == 'Sbig':
            res[name + '/tail'] = y.cpu().numpy()[-8:]
        # the line bookkeeping of the tcgen05 kernel under other row layouts: auto-correlation pairs only (F rows per
        # sample: a tile spans several samples), within-channel, random pair subsets of several sizes
        xs = x[:5]
        res[name + '/same'] = m(xs, compute_phase=False, compute_cross_phase=True, cross_phase_same_pairs_only=True)['cross_phase_corr'].cpu().numpy()
        res[name + '/within'] = m(xs, compute_phase=True, phase_channels=[1])['phase_corr'].cpu().numpy()
        rng = np.random.RandomState(3)
        P = int(y.shape[1])
        for k in (1, 7, min(44, P), min(200, P)):
            mask = np.zeros(P, bool); mask[rng.choice(P, k, replace=False)] = True
            res[name + '/subset%d' % k] = m(xs, compute_phase=False, compute_cross_phase=True, phase_pairs=torch.from_numpy(mask))['cross_phase_corr'].cpu().numpy()
        torch.cuda.synchronize()
    np.savez(out, **res)

if __name__ == '__main__':
    if len(sys.argv) > 1:
        child(sys.argv[1], sys.argv[2])
        sys.exit(0)
    outs = {}
    for tag in ('sync', 'tc'):
        env = dict(os.environ, TEBSCAT_PHASE_MMA=tag, TEBSCAT_PHASE_FFT='0')
        f = '/tmp/phase_%s.npz' % tag
        r = subprocess.run([sys.executable, __file__, tag, f], env=env, timeout=600)
        print(tag, 'rc', r.returncode, flush=True)
        if r.returncode == 0:
            outs[tag] = np.load(f)
    if len(outs) == 2:
        for name in sorted(outs['sync'].files):
            a, b = outs['sync'][name].astype(np.float64), outs['tc'][name].astype(np.float64)
            err = np.linalg.norm(a - b, axis=-1) / np.maximum(np.linalg.norm(a, axis=-1), 1e-30)
            strong = np.linalg.norm(a, axis=-1) > 1e-4 * np.linalg.norm(a, axis=-1).max()
            print(name, 'shape', a.shape, 'overall rel', np.linalg.norm(a - b) / np.linalg.norm(a), 'max per-row (strong)', err[strong].max(),
                  'nan', np.isnan(b).sum(), flush=True)
