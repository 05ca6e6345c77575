"""Pin the oracle (oracle/*.py) against the reference's own KAT and against
fixtures generated from the live reference (oracle/make_golden.py)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import filters_oracle as fo
from oracle.scattering1d_oracle import ScatteringOracle
from oracle.phase_oracle import PhaseOracle
from helpers import phase_path_tolerance

SCAT = ['H', 'P', 'S', 'T', 'O', 'L']


def rel_l2(a, b, axis=None):
    return np.linalg.norm(a - b, axis=axis) / np.maximum(np.linalg.norm(b, axis=axis), 1e-30)


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_kat_test_data_1d(golden_dir):
    """kymatio/tests/scattering1d/test_torch_scattering1d.py:82-113 (test_sample_scattering)."""
    d = load(golden_dir, 'kat_test_data_1d.npz')
    J, Q = int(d['J']), int(d['Q'])
    o = ScatteringOracle(J, d['x'].shape[-1], Q, 2 ** J)
    Sx = o(d['x'])
    assert Sx.shape == d['Sx'].shape
    assert rel_l2(Sx, d['Sx']) < 5e-7
    assert np.allclose(Sx, d['Sx'], rtol=1e-5, atol=1e-8)      # torch.allclose defaults


@pytest.mark.parametrize('name', SCAT)
def test_geometry_filters_meta(golden_dir, name):
    d = load(golden_dir, 'scat_%s.npz' % name)
    J, Q, T, N = int(d['J']), int(d['Q']), int(d['T']), int(d['N'])
    g = fo.geometry(N, J, Q, T)
    assert (g['J_pad'], g['pad_left'], g['pad_right']) == (int(d['J_pad']), int(d['pad_left']), int(d['pad_right']))
    assert [g['ind_start'][j] for j in range(J + 1)] == list(d['ind_start'])
    assert [g['ind_end'][j] for j in range(J + 1)] == list(d['ind_end'])
    bank = fo.filter_factory(g['J_pad'], J, Q, T)
    filt = list(bank['phi'])
    for p in bank['psi1'] + bank['psi2']:
        filt += p['levels']
    assert len(filt) == int(d['n_filters'])
    assert [a.shape[0] for a in filt] == list(d['filter_len'])
    sha = [hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() for a in filt]
    assert sha == list(d['filter_sha256'])                       # bit-exact float64 filters
    assert bank['t_max_phi'] == int(d['t_max_phi'])
    keys = fo.path_keys(J, Q, T, int(d['max_order']))
    gk = [tuple(int(v) for v in row if v >= 0) for row in d['keys']]
    assert keys == gk


@pytest.mark.parametrize('name', SCAT)
def test_scattering_vs_reference(golden_dir, name):
    """Single-precision oracle == reference to 1e-5 per path; the float64 oracle is
    the 'truth' both are compared with.  On CTG-like inputs (baseline 140 bpm) some
    second-order paths carry ~1e-6 of the signal energy and the reference's own fp32
    rounding noise is ~2e-5 of those paths, hence the looser float64 bound."""
    d = load(golden_dir, 'scat_%s.npz' % name)
    args = (int(d['J']), int(d['N']), int(d['Q']), int(d['T']), int(d['max_order']))
    os_ = int(d['oversampling']) if 'oversampling' in d else 0
    S32 = ScatteringOracle(*args, cdtype=np.complex64, oversampling=os_)(d['x'])
    S64 = ScatteringOracle(*args, oversampling=os_)(d['x'])
    assert S32.shape == d['S'].shape and S64.shape == d['S'].shape
    assert rel_l2(S32, d['S'], axis=-1).max() < 1e-5
    assert rel_l2(S64, d['S'], axis=-1).max() < 5e-5
    assert rel_l2(S64, d['S']) < 1e-6
    randn = slice(d['x'].shape[0] // 2, None)               # second half of the batch is randn
    assert rel_l2(S64[randn], d['S'][randn], axis=-1).max() < 1e-5


@pytest.mark.parametrize('name', SCAT)
def test_torch_port_vs_reference(golden_dir, name):
    """The torch-CPU port timed as the CPU baseline issues the reference's own torch calls."""
    import torch
    from oracle.scattering1d_torch_port import TorchPort
    d = load(golden_dir, 'scat_%s.npz' % name)
    if 'oversampling' in d and int(d['oversampling']):
        pytest.skip('the timed port covers oversampling=0 (the benchmark configuration)')
    port = TorchPort(int(d['J']), int(d['N']), int(d['Q']), int(d['T']), int(d['max_order']))
    S = port(torch.from_numpy(d['x'])).numpy()
    assert S.shape == d['S'].shape
    assert rel_l2(S, d['S'], axis=-1).max() < 2e-6


def _pin_phase_oracle(o, x, ref, mode, sub, randn_rows, tag):
    """float64 oracle vs the live reference's fp32 output, per path: within 1e-5 wherever single-precision
    arithmetic allows it, and never further than 4x the distance between the single-precision and the float64
    evaluation of the oracle itself (the reference rounds p * theta in fp32 with p up to 108)."""
    xin = x[:, 0] if mode == 'within' else x
    assert o(xin, mode=mode, pair_subset=sub).shape == ref.shape
    aligned = o.align_branches(xin, ref, mode=mode, pair_subset=sub)
    o32 = PhaseOracle(o.J, o.Q, o.T, o.N, o.n_out, border_mode=o.border_mode, cdtype=np.complex64)
    y32 = o32(xin, mode=mode, pair_subset=sub).astype(np.float64)
    tol = phase_path_tolerance(o.align_branches(xin, y32, mode=mode, pair_subset=sub), y32)
    err = np.linalg.norm(aligned - ref, axis=-1)
    rel = err / np.linalg.norm(aligned, axis=-1)
    pw = o.powers if sub is None else o.powers[sub]
    for b in randn_rows:
        assert rel[b][pw < 32].max() <= 1e-5 and rel[b].max() <= 2.5e-5, (tag, mode, b, rel[b].max())
    assert (err / tol).max() <= 1.0, (tag, mode, (err / tol).max())
    assert rel_l2(aligned, ref) < 5e-5, (tag, mode, rel_l2(aligned, ref))


@pytest.mark.parametrize('name', ['H', 'Hr', 'P', 'S', 'L'])
def test_phase_vs_reference(golden_dir, name):
    d = load(golden_dir, 'phase_%s.npz' % name)
    J, Q, T, N = int(d['J']), int(d['Q']), int(d['T']), int(d['N'])
    n_out = d['scattering'].shape[-1]
    o = PhaseOracle(J, Q, T, N, n_out)
    if 'powers' in d.files:
        assert np.array_equal(o.center_freqs, d['center_freqs'])
        assert np.array_equal(o.i_idx, d['i_idx']) and np.array_equal(o.j_idx, d['j_idx'])
        assert np.array_equal(o.powers, d['powers'])
        assert np.array_equal(o.autoc_idx, d['autoc_idx'])
    B = d['x'].shape[0]
    n_ctg = int(d['n_ctg']) if 'n_ctg' in d.files else B // 2          # make_golden: CTG rows first, randn rows behind
    rows = [0, B - 1] if n_ctg else [0, 1]                              # one CTG and one randn row (bounded CPU time)
    x = d['x'][rows]
    sub_w = np.nonzero(d['phase_mask'])[0] if bool(d['subset']) else None
    sub_c = np.nonzero(d['cross_mask'])[0] if bool(d['subset']) else None
    randn_rows = [1] if n_ctg else [0, 1]
    _pin_phase_oracle(o, x, d['within'][rows], 'within', sub_w, randn_rows, name)
    _pin_phase_oracle(o, x, d['cross'][rows], 'cross', sub_c, randn_rows, name)


@pytest.mark.parametrize('tag', ['S_constant', 'S_circular', 'S_over1', 'S_tukey', 'S_nodec'])
def test_phase_options_vs_reference(golden_dir, tag):
    """border_mode 'constant' / 'circular' (kymatio_phase_scattering.py:162-173) and oversampling (the
    target length follows the scattering output, :445) against outputs of the live reference."""
    d = load(golden_dir, 'phase_%s.npz' % tag)
    J, Q, T, N = int(d['J']), int(d['Q']), int(d['T']), int(d['N'])
    o = PhaseOracle(J, Q, T, N, d['scattering'].shape[-1], border_mode=str(d['border_mode']))
    x = d['x']
    if 'tukey_alpha' in d.files and float(d['tukey_alpha']) > 0:       # the reference tapers its input first (:405-407)
        x = x * d['window']
    _pin_phase_oracle(o, x, d['within'], 'within', None, [1], tag)
    _pin_phase_oracle(o, x, d['cross'], 'cross', None, [1], tag)


@pytest.mark.parametrize('name', ['T', 'S', 'P1', 'O', 'H'])
def test_gradient_oracle_vs_reference(golden_dir, name):
    """The float64 autograd restatement against d/dx sum(S w) of the live reference's own autograd graph
    (ModulusStable, kymatio/backend/torch_backend.py:5-96)."""
    from oracle.scattering1d_grad_oracle import GradOracle
    d = load(golden_dir, 'backward_%s.npz' % name)
    o = GradOracle(int(d['J']), int(d['N']), int(d['Q']), int(d['T']), int(d['max_order']), int(d['oversampling']))
    S, gx = o.vjp(d['x'], d['w'])
    assert rel_l2(S, d['S'], axis=-1).max() < 2e-5
    # signal 0 is CTG-shaped (baseline 140 bpm): the reference's own fp32 rounding is ~2e-5 of its gradient there
    # (as for the forward); signal 1 is randn
    err = rel_l2(gx, d['gx'], axis=-1)
    assert err[1] < 1e-5 and err[0] < 5e-5, err


@pytest.mark.parametrize('name', ['Tu', 'P1u'])
def test_unaveraged_gradient_oracle_vs_reference(golden_dir, name):
    """average=False: gradient of the un-averaged outputs (order 0 = the input itself included) against the live
    reference's own autograd graph (oracle/make_golden_backward.py)."""
    from oracle.scattering1d_grad_oracle import GradOracle
    d = load(golden_dir, 'backward_%s.npz' % name)
    o = GradOracle(int(d['J']), int(d['N']), int(d['Q']), int(d['T']), int(d['max_order']), int(d['oversampling']))
    row, g = o.vjp_unaveraged(d['x'], d['w'])
    gx = g + d['w0']                                      # order 0: d sum(x w0) / dx = w0
    e_row = rel_l2(row, d['row'], axis=-1)                # row 0 is CTG-shaped (mean 140 bpm): the reference's fp32 moduli carry ~1e-5
    assert row.shape == d['row'].shape and e_row[1] < 2e-6 and e_row[0] < 2e-5, e_row
    err = rel_l2(gx, d['gx'], axis=-1)
    assert err[1] < 1e-5 and err[0] < 5e-5, err


def test_branch_alignment_covers_an_interior_sample_on_the_negative_real_axis():
    """The input the random sweep met (seed 32, tools/random_parity_sweep.py): analytic sample (row 0, filter 3,
    t = 2169) = -0.119 + 2.3e-8j, i.e. on the negative real axis to 2e-7 -- three ulps of |z| in float32.  Whichever
    sign an fp32 transform gives Im z there is legitimate; without decimation the other branch costs ~1e-3 on the
    paths of that filter.  align_branches picks the branch of the output under test, and only there."""
    import torch
    J, Q, T, N = 5, 8, 8, 3080
    x = torch.randn(3, 2, N, generator=torch.Generator().manual_seed(319)).numpy()[:1]
    o = PhaseOracle(J, Q, T, N, N)
    zi, zj = o.analytic(x[:, 0]), o.analytic(x[:, 1])
    z = zi[0, 3, 2169]
    assert z.real < 0 and abs(z.imag) < 2e-6 * abs(z.real)
    plain = o(x)
    zi[0, 3, 2169] = np.conj(z)                                            # the other branch at that one sample
    flipped = np.real(o.pair_stage(zi, zj))
    cost = np.linalg.norm(flipped - plain) / np.linalg.norm(plain)
    assert cost > 1e-4
    assert np.linalg.norm(flipped - o.align_branches(x, flipped, interior_rel_im=0)) / np.linalg.norm(plain) > 1e-4
    assert np.linalg.norm(flipped - o.align_branches(x, flipped)) / np.linalg.norm(plain) < 1e-12
    assert np.array_equal(o.align_branches(x, plain), plain)               # nothing to align: the plain oracle
    # an error anywhere else is NOT absorbed: perturb one unrelated sample of the output
    wrong = plain.copy()
    wrong[0, 100, 500] += 1e-2 * np.abs(plain[0, 100]).max()
    assert np.array_equal(o.align_branches(x, wrong), plain)
