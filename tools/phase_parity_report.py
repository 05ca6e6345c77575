"""Phase-path parity report (GPU): per fixture, stage-B form, mode and row, the worst per-path error of
the CUDA output against the branch-aligned float64 oracle, as rel-L2 and as a ratio to the tolerance of
tests/helpers.py:phase_path_tolerance.  Run on the GPU box:

    python tools/phase_parity_report.py [fixture ...]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'vae-teb_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)

from helpers import GOLDEN, phase_path_tolerance, rel_l2     # noqa: E402
from oracle.phase_oracle import PhaseOracle                   # noqa: E402

CFG = {'H': (6, 8, 64, 4800, 2), 'Hr': (6, 8, 64, 4800, 2), 'P': (11, 4, 16, 5760, 1), 'S': (4, 4, 16, 1000, 2)}
FORMS = (('dense/tcgen05', '0', 'tc'), ('dense/mma.sync', '0', 'sync'), ('transform', '1', 'tc'))


def main(names):
    from tebscat import KymatioPhaseScattering1D
    for name in names:
        d = np.load(os.path.join(GOLDEN, 'phase_%s.npz' % name))
        base = name.split('_')[0]
        J, Q, T, N, mo = CFG[base]
        border = str(d['border_mode']) if 'border_mode' in d.files else 'reflect'
        over = int(d['oversampling']) if 'oversampling' in d.files else 0
        subset = bool(d['subset']) if 'subset' in d.files else False
        n_ctg = int(d['n_ctg']) if 'n_ctg' in d.files else d['x'].shape[0] // 2
        o = PhaseOracle(J, Q, T, N, d['scattering'].shape[-1], border_mode=border)
        sub_w = np.nonzero(d['phase_mask'])[0] if subset else None
        sub_c = np.nonzero(d['cross_mask'])[0] if subset else None
        x = torch.from_numpy(d['x']).cuda()
        for form, fft, mma in FORMS:
            os.environ['TEBSCAT_PHASE_FFT'] = fft
            os.environ['TEBSCAT_PHASE_MMA'] = mma
            m = KymatioPhaseScattering1D(J=J, Q=Q, T=T, shape=N, device=torch.device('cuda'), max_order=mo,
                                         border_mode=border, oversampling=over)
            if fft == '1' and m._plan.pair_plan is None:
                continue
            for mode, ref, sub in (('within', d['within'], sub_w), ('cross', d['cross'], sub_c)):
                if mode == 'within':
                    ours = m(x, compute_phase=True, phase_channels=[0], phase_pairs=sub)['phase_corr']
                    xin = d['x'][:, 0]
                else:
                    ours = m(x, compute_phase=False, compute_cross_phase=True, phase_channels=[0, 1],
                             phase_pairs=sub)['cross_phase_corr']
                    xin = d['x']
                ours = ours.cpu().numpy().astype(np.float64)
                al_us = o.align_branches(xin, ours, mode=mode, pair_subset=sub)
                al_ref = o.align_branches(xin, ref, mode=mode, pair_subset=sub)
                tol = phase_path_tolerance(al_ref, ref)
                err = np.linalg.norm(ours - al_us, axis=-1)
                rel = err / np.linalg.norm(al_us, axis=-1)
                relref = np.linalg.norm(ref - al_ref, axis=-1) / np.linalg.norm(al_ref, axis=-1)
                for b in range(ours.shape[0]):
                    kind = 'ctg' if b < n_ctg else 'randn'
                    print('%-11s %-14s %-6s row %d %-5s overall %.2e  path rel max %.2e (reference: %.2e)  '
                          'n>1e-5 %3d/%d  worst err/tol %.3f' % (
                              name, form, mode, b, kind, rel_l2(ours[b], al_us[b]), rel[b].max(), relref[b].max(),
                              int((rel[b] > 1e-5).sum()), rel.shape[1], (err[b] / tol[b]).max()), flush=True)


if __name__ == '__main__':
    main(sys.argv[1:] or ['H', 'P', 'S', 'S_constant', 'S_circular', 'S_over1'])
