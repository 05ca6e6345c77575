"""Generate the golden fixtures under tests/golden/ from the LIVE reference.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

Imports the reference's own kymatio fork and phase module from /root/reference
(read-only), evaluates them on CPU on seeded inputs and stores inputs + outputs
as small .npz files.  Nothing from the reference's sources is copied; the
reference's known-answer *data* fixture test_data_1d.npz is re-saved as
kat_test_data_1d.npz so that the KAT travels to the GPU box.
"""
import hashlib
import os
import sys

import numpy as np
import scipy.special
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
sys.path.insert(0, os.path.join(REF, 'kymatio'))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
sys.dont_write_bytecode = True
if not hasattr(scipy.special, 'sph_harm'):
    scipy.special.sph_harm = None       # scattering3d import shim

from kymatio.scattering1d.frontend.torch_frontend import ScatteringTorch1D   # noqa: E402
from kymatio.scattering1d.filter_bank import scattering_filter_factory       # noqa: E402
import hdf5_dataset.kymatio_phase_scattering as kps                           # noqa: E402
from tebscat.synth import ctg_batch, randn_batch                              # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(8)

CONFIGS = {
    # name: (J, Q, T, N, max_order, B)
    'H': (6, 8, 64, 4800, 2, 3),        # BASELINE headline config
    'P': (11, 4, 16, 5760, 1, 2),       # production dataset config (create_hdf5_dataset.py:360)
    'S': (4, 4, 16, 1000, 2, 4),        # small ragged-length config
    'T': (5, 2, 8, 700, 2, 3),          # T < 2**J
    'O': (5, 4, 32, 1200, 2, 2, 1),     # oversampling = 1
    'L': (6, 4, 64, 9000, 2, 1),        # padded length 2**14: the large-support level
}


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def scat_fixture(name):
    J, Q, T, N, max_order, B = CONFIGS[name][:6]
    oversampling = CONFIGS[name][6] if len(CONFIGS[name]) > 6 else 0
    S = ScatteringTorch1D(J, N, Q, max_order=max_order, T=T, oversampling=oversampling)
    x = torch.cat([ctg_batch(B, N, seed=1234)[:, 0], randn_batch(B, N, 1, seed=4321)[:, 0]], 0)
    with torch.no_grad():
        out, _ = S(x)
    meta = S.meta()
    phi, psi1, psi2, t_max = scattering_filter_factory(S.J_pad, J, Q, T)
    filt = [np.asarray(a) for a in phi['levels']]
    for p in psi1 + psi2:
        filt += [np.asarray(a) for a in p['levels']]
    np.savez_compressed(
        os.path.join(OUT, 'scat_%s.npz' % name),
        J=J, Q=Q, T=T, N=N, max_order=max_order, oversampling=oversampling, x=x.numpy(), S=out.numpy(),
        J_pad=S.J_pad, pad_left=S.pad_left, pad_right=S.pad_right,
        ind_start=np.array([S.ind_start[j] for j in range(J + 1)]),
        ind_end=np.array([S.ind_end[j] for j in range(J + 1)]),
        keys=np.array([k + (-1,) * (2 - len(k)) for k in meta['key']], dtype=np.int64),
        order=meta['order'], meta_xi=meta['xi'], meta_sigma=meta['sigma'], meta_j=meta['j'],
        t_max_phi=t_max, n_filters=len(filt),
        filter_len=np.array([a.shape[0] for a in filt]),
        filter_sum=np.array([a.sum() for a in filt]),
        filter_l2=np.array([np.sqrt((a ** 2).sum()) for a in filt]),
        filter_sha256=np.array([digest(a) for a in filt]),
        psi1_0=psi1[0]['levels'][0], psi1_last=psi1[-1]['levels'][0],
        psi2_last_top=psi2[-1]['levels'][-1], phi_0=phi['levels'][0],
        output_size=np.array(S.output_size(detail=True)),
    )
    print(name, 'scat', tuple(out.shape), 'J_pad', S.J_pad)


def phase_fixture(name, B):
    J, Q, T, N, max_order, _ = CONFIGS[name][:6]
    m = kps.KymatioPhaseScattering1D(J=J, Q=Q, T=T, shape=N, device=torch.device('cpu'),
                                     max_order=max_order)
    x = torch.cat([ctg_batch(B, N, seed=77), randn_batch(B, N, 2, seed=78)], 0)
    with torch.no_grad():
        rw = m(x, compute_phase=True, phase_channels=[0])
        rc = m(x, compute_phase=False, compute_cross_phase=True, phase_channels=[0, 1])
    sel = m.get_optimal_coefficients_for_fhr(J, Q, T)
    pm = sel['recommendations']['use_phase_mask'].numpy()
    cm = sel['recommendations']['use_cross_mask'].numpy()
    within, cross = rw['phase_corr'].numpy(), rc['cross_phase_corr'].numpy()
    if within.shape[1] * within.shape[2] > 200000:      # keep big configs small: masked subset
        within, cross = within[:, pm], cross[:, cm]
        subset = True
    else:
        subset = False
    np.savez_compressed(
        os.path.join(OUT, 'phase_%s.npz' % name),
        J=J, Q=Q, T=T, N=N, max_order=max_order, x=x.numpy(),
        scattering=rw['scattering'].numpy(), within=within, cross=cross, subset=subset,
        phase_mask=pm, cross_mask=cm, center_freqs=m.center_freqs.numpy(),
        i_idx=m.i_idx.numpy(), j_idx=m.j_idx.numpy(), powers=m.powers.numpy(),
        autoc_idx=m.autoc_idx.numpy(), J_pad=m.J_pad, pad_left=m.pad_left, pad_right=m.pad_right,
    )
    print(name, 'phase', within.shape, cross.shape, 'masks', pm.sum(), cm.sum())


def phase_fixture_randn(tag, name, B, seed=578):
    """randn-only rows of the full pair set (no CTG rows): on these the live reference itself is within
    1e-5 per path of the float64 oracle, so north_star's bound is asserted on them without any relaxation."""
    J, Q, T, N, max_order, _ = CONFIGS[name][:6]
    m = kps.KymatioPhaseScattering1D(J=J, Q=Q, T=T, shape=N, device=torch.device('cpu'), max_order=max_order)
    x = randn_batch(B, N, 2, seed=seed)
    with torch.no_grad():
        rw = m(x, compute_phase=True, phase_channels=[0])
        rc = m(x, compute_phase=False, compute_cross_phase=True, phase_channels=[0, 1])
    np.savez_compressed(
        os.path.join(OUT, 'phase_%s.npz' % tag),
        J=J, Q=Q, T=T, N=N, max_order=max_order, x=x.numpy(), n_ctg=0, subset=False,
        scattering=rw['scattering'].numpy(), within=rw['phase_corr'].numpy(), cross=rc['cross_phase_corr'].numpy(),
    )
    print(tag, 'phase (randn rows)', rw['phase_corr'].shape, rc['cross_phase_corr'].shape)


def phase_option_fixture(tag, name, B, **opts):
    """Non-default options of the phase module: border_mode 'constant' / 'circular'
    (:162-173) and oversampling (target length of the scattering output, :445)."""
    J, Q, T, N, max_order, _ = CONFIGS[name][:6]
    m = kps.KymatioPhaseScattering1D(J=J, Q=Q, T=T, shape=N, device=torch.device('cpu'),
                                     max_order=max_order, **opts)
    x = torch.cat([ctg_batch(B, N, seed=177), randn_batch(B, N, 2, seed=178)], 0)
    with torch.no_grad():
        rw = m(x, compute_phase=True, phase_channels=[0])
        rc = m(x, compute_phase=False, compute_cross_phase=True, phase_channels=[0, 1])
    np.savez_compressed(
        os.path.join(OUT, 'phase_%s.npz' % tag),
        J=J, Q=Q, T=T, N=N, max_order=max_order, x=x.numpy(),
        border_mode=opts.get('border_mode', 'reflect'), oversampling=opts.get('oversampling', 0),
        tukey_alpha=float(opts.get('tukey_alpha') or 0.0),
        window=m._create_tukey_window(N, opts.get('tukey_alpha'), torch.device('cpu')).numpy(),
        window_odd=m._create_tukey_window(777, 0.5, torch.device('cpu')).numpy(),
        window_hann=m._create_tukey_window(64, 1.0, torch.device('cpu')).numpy(),
        scattering=rw['scattering'].numpy(), within=rw['phase_corr'].numpy(), cross=rc['cross_phase_corr'].numpy(),
    )
    print(tag, 'phase', rw['phase_corr'].shape, rc['cross_phase_corr'].shape)


def phase_option_fixtures():
    phase_option_fixture('S_constant', 'S', 1, border_mode='constant')
    phase_option_fixture('S_circular', 'S', 1, border_mode='circular')
    phase_option_fixture('S_over1', 'S', 1, oversampling=1)
    phase_option_fixture('S_tukey', 'S', 1, tukey_alpha=0.25)
    phase_option_fixture('S_nodec', 'S', 1, oversampling=4)        # target length = N: no decimation (:268-273)


if __name__ == '__main__':
    if sys.argv[1:] == ['phase-randn']:
        phase_fixture_randn('Hr', 'H', 4)
        sys.exit(0)
    if sys.argv[1:] == ['phase-large']:
        phase_fixture('L', 1)
        sys.exit(0)
    if sys.argv[1:] == ['phase-nodec']:
        phase_option_fixture('S_nodec', 'S', 1, oversampling=4)
        sys.exit(0)
    if sys.argv[1:] == ['phase-tukey']:
        phase_option_fixture('S_tukey', 'S', 1, tukey_alpha=0.25)
        sys.exit(0)
    if sys.argv[1:] == ['phase-options']:
        phase_option_fixtures()
        sys.exit(0)
    if len(sys.argv) == 3 and sys.argv[1] == 'scat':
        scat_fixture(sys.argv[2])
        sys.exit(0)
    for n in CONFIGS:
        scat_fixture(n)
    phase_fixture('H', 1)
    phase_fixture('P', 1)
    phase_fixture('S', 2)
    phase_fixture('L', 1)                       # padded length 2**14: stage A on the large-support level
    phase_fixture_randn('Hr', 'H', 4)
    phase_option_fixtures()
    kat = np.load(os.path.join(REF, 'kymatio/tests/scattering1d/test_data_1d.npz'))
    np.savez_compressed(os.path.join(OUT, 'kat_test_data_1d.npz'), **{k: kat[k] for k in kat})
    print('done')
