"""Dense (tcgen05) vs transform form of phase stage B around the cross-over output length (one process per form)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
CFGS = [(8, 4, 32, 5760), (8, 4, 16, 4800), (8, 4, 16, 5760)]     # n_out = 180, 300, 360

def child(form):
    import torch
    from tebscat import KymatioPhaseScattering1D
    from tebscat.synth import ctg_batch
    for (J, Q, T, N) in CFGS:
        m = KymatioPhaseScattering1D(J=J, Q=Q, T=T, shape=N, device=torch.device('cuda'), max_order=1)
        x = ctg_batch(128, N, seed=3).cuda()
        m(x, compute_phase=False, compute_cross_phase=True); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            m(x, compute_phase=False, compute_cross_phase=True)
        e1.record(); torch.cuda.synchronize()
        print('form', form, (J, Q, T, N), 'n_out', m._plan.n_out, 'pairs', len(m.i_idx), 'fft' if m._dev_plan(0).uses_fft_pairs else 'dense',
              '%.2f ms per 128 samples' % (e0.elapsed_time(e1) / 3), flush=True)

if __name__ == '__main__':
    if len(sys.argv) > 1:
        child(sys.argv[1]); sys.exit(0)
    for form in ('0', '1'):
        subprocess.run([sys.executable, __file__, form], env=dict(os.environ, TEBSCAT_PHASE_FFT=form), timeout=280)
