"""CPU checks of the phase path's host logic: pair tables and masks against the golden
reference fixtures, the smoothing operator against the oracle, and stage A (analytic
signals) through the host emulator of the step interpreter."""
import os

import numpy as np
import pytest

from helpers import GOLDEN, emu_available, emu_forward, emu_pair_stage
from oracle.phase_oracle import PhaseOracle
from tebscat.phase import PhasePlan

CFG = {'H': (6, 8, 64, 4800), 'P': (11, 4, 16, 5760), 'S': (4, 4, 16, 1000)}
_plans = {}


def plan_of(name, n_out_scat):
    if name not in _plans:
        J, Q, T, N = CFG[name]
        _plans[name] = PhasePlan(J, Q, T, N, n_out_scat)
    return _plans[name]


@pytest.mark.parametrize('name', ['H', 'P', 'S'])
def test_pair_tables_match_reference(name):
    d = np.load(os.path.join(GOLDEN, 'phase_%s.npz' % name))
    p = plan_of(name, d['scattering'].shape[-1])
    assert np.array_equal(p.center_freqs, d['center_freqs'])          # same fp32 values -> same masks
    assert np.array_equal(p.i_idx, d['i_idx']) and np.array_equal(p.j_idx, d['j_idx'])
    assert np.array_equal(p.powers, d['powers'])
    assert np.array_equal(p.autoc_idx, d['autoc_idx'])
    assert (p.geo.J_pad, p.geo.pad_left, p.geo.pad_right) == (int(d['J_pad']), int(d['pad_left']), int(d['pad_right']))
    ref_len = d['within'].shape[-1]
    assert p.n_out == ref_len                                          # e.g. 66 != 63 for the ragged config


@pytest.mark.parametrize('name', ['H', 'S'])
def test_smoothing_operator_matches_oracle(name):
    d = np.load(os.path.join(GOLDEN, 'phase_%s.npz' % name))
    J, Q, T, N = CFG[name]
    p = plan_of(name, d['scattering'].shape[-1])
    o = PhaseOracle(J, Q, T, N, d['scattering'].shape[-1])
    rng = np.random.RandomState(3)
    c = rng.randn(4, N) + 1j * rng.randn(4, N)
    G = p.G[:, :p.n_out, 0].astype(np.float64) + 1j * p.G[:, :p.n_out, 1]
    ref = o._smooth(c)
    assert np.abs(c @ G - ref).max() < 1e-6 * np.abs(ref).max()
    assert np.all(p.G[:, p.n_out:] == 0)


@pytest.mark.skipif(not emu_available(), reason='host emulator not built')
@pytest.mark.parametrize('name', ['H', 'S'])
def test_stage_a_emulated_matches_oracle(name):
    d = np.load(os.path.join(GOLDEN, 'phase_%s.npz' % name))
    J, Q, T, N = CFG[name]
    p = plan_of(name, d['scattering'].shape[-1])
    o = PhaseOracle(J, Q, T, N, d['scattering'].shape[-1])
    x = np.random.RandomState(5).randn(2, N).astype(np.float32)
    z, zp = emu_forward(p.stage_a, x, stage_a=True)
    ref = o.analytic(x)
    assert not np.isnan(z).any()
    err = np.linalg.norm(z - ref, axis=-1) / np.linalg.norm(ref, axis=-1)
    assert err.max() < 1e-5, err.max()
    assert np.allclose(zp[..., 0], np.abs(z), rtol=1e-6, atol=1e-9)
    assert np.allclose(zp[..., 1], np.angle(z), atol=1e-6)


@pytest.mark.skipif(not emu_available(), reason='host emulator not built')
@pytest.mark.parametrize('name', ['H', 'P'])
def test_pair_stage_transform_form_matches_oracle(name):
    """Stage B as transforms on the interpreter (power-of-two decimation): LOADPAIR -> FFT -> phi on the kept
    bins -> reduced iFFT -> unpad must equal the oracle's _smooth of the same products."""
    d = np.load(os.path.join(GOLDEN, 'phase_%s.npz' % name))
    J, Q, T, N = CFG[name]
    p = plan_of(name, d['scattering'].shape[-1])
    assert p.pair_plan is not None and p.pair_plan.n_out == p.n_out
    o = PhaseOracle(J, Q, T, N, d['scattering'].shape[-1])
    rng = np.random.RandomState(9)
    rows = 5                                                  # odd: the last job repeats its row
    zi = (rng.randn(rows, N) + 1j * rng.randn(rows, N)).astype(np.complex64)
    zj = (rng.randn(rows, N) + 1j * rng.randn(rows, N)).astype(np.complex64)
    pw = np.array([1.0, 1.5, 2.0, 3.25, 7.0], np.float32)
    zp = np.stack([np.abs(zi), np.angle(zi)], -1).astype(np.float32)
    zc = np.stack([zj.real, zj.imag], -1).astype(np.float32)
    out = emu_pair_stage(p.pair_plan, zp, zc, pw)
    assert out.shape == (rows, p.n_out) and not np.isnan(out).any()
    theta = zp[..., 1].astype(np.float32) * pw[:, None]       # fp32 product like the reference (:215)
    c = zp[..., 0].astype(np.float64) * np.exp(1j * theta.astype(np.float64)) * np.conj(zj.astype(np.complex128))
    ref = o._smooth(c).real
    err = np.linalg.norm(out - ref, axis=-1) / np.linalg.norm(ref, axis=-1)
    assert err.max() < 1e-5, err


# ---- non-default options: border_mode (:162-173) and a power-of-two decimation on the ragged config ------------
@pytest.mark.parametrize('border', ['constant', 'circular'])
def test_smoothing_operator_border_modes(border):
    J, Q, T, N = CFG['S']
    p = PhasePlan(J, Q, T, N, 66, border_mode=border)
    o = PhaseOracle(J, Q, T, N, 66, border_mode=border)
    rng = np.random.RandomState(3)
    c = rng.randn(4, N) + 1j * rng.randn(4, N)
    G = p.G[:, :p.n_out, 0].astype(np.float64) + 1j * p.G[:, :p.n_out, 1]
    ref = o._smooth(c)
    assert np.abs(c @ G - ref).max() < 1e-6 * np.abs(ref).max()


@pytest.mark.skipif(not emu_available(), reason='host emulator not built')
@pytest.mark.parametrize('border', ['constant', 'circular'])
def test_stage_a_and_pair_stage_border_modes_emulated(border):
    """LOAD and LOADPAIR with zero / circular padding through the host emulator, against the oracle.
    Target length 125 (oversampling = 1 of the S config) gives the power-of-two decimation 8 the
    transform form of stage B needs."""
    J, Q, T, N = CFG['S']
    p = PhasePlan(J, Q, T, N, 125, border_mode=border)
    o = PhaseOracle(J, Q, T, N, 125, border_mode=border)
    assert p.dec == 8 and p.pair_plan is not None and p.stage_a.border == p.pair_plan.border != 0
    x = np.random.RandomState(5).randn(2, N).astype(np.float32) + 3.0      # an offset makes the padding rule visible
    z, zp = emu_forward(p.stage_a, x, stage_a=True)
    ref = o.analytic(x)
    err = np.linalg.norm(z - ref, axis=-1) / np.linalg.norm(ref, axis=-1)
    assert not np.isnan(z).any() and err.max() < 1e-5, err.max()
    # and it is not the reflect result
    assert np.abs(z - PhaseOracle(J, Q, T, N, 125).analytic(x)).max() > 1e-3

    rng = np.random.RandomState(9)
    rows = 3
    zi = (rng.randn(rows, N) + 1j * rng.randn(rows, N)).astype(np.complex64)
    zj = (rng.randn(rows, N) + 1j * rng.randn(rows, N)).astype(np.complex64)
    pw = np.array([1.0, 1.5, 3.25], np.float32)
    zpol = np.stack([np.abs(zi), np.angle(zi)], -1).astype(np.float32)
    zc = np.stack([zj.real, zj.imag], -1).astype(np.float32)
    out = emu_pair_stage(p.pair_plan, zpol, zc, pw)
    theta = zpol[..., 1].astype(np.float32) * pw[:, None]
    c = zpol[..., 0].astype(np.float64) * np.exp(1j * theta.astype(np.float64)) * np.conj(zj.astype(np.complex128))
    ref = o._smooth(c).real
    assert out.shape == ref.shape and not np.isnan(out).any()
    err = np.linalg.norm(out - ref, axis=-1) / np.linalg.norm(ref, axis=-1)
    assert err.max() < 1e-5, err


def test_tukey_window_matches_reference():
    """_create_tukey_window (kymatio_phase_scattering.py:362-392) of the live reference, stored by
    oracle/make_golden.py: interior taper, odd length, and the alpha >= 1 (Hann) branch; degenerate alphas."""
    from tebscat.phase import tukey_window
    d = np.load(os.path.join(GOLDEN, 'phase_S_tukey.npz'))
    for key, n, alpha in (('window', 1000, 0.25), ('window_odd', 777, 0.5), ('window_hann', 64, 1.0)):
        w = tukey_window(n, alpha)
        assert w.shape == d[key].shape and np.abs(w - d[key]).max() < 5e-7, key
    assert np.array_equal(tukey_window(50, None), np.ones(50)) and np.array_equal(tukey_window(50, 0.0), np.ones(50))
    assert np.array_equal(tukey_window(50, 1.5), np.ones(50))        # outside (0, 1]: rectangular, like the reference
    assert np.array_equal(tukey_window(5, 0.2), np.ones(5))          # taper shorter than one sample


@pytest.mark.skipif(not emu_available(), reason='host emulator not built')
def test_pair_stage_without_decimation_emulated():
    """Target length >= N (oversampling >= log2 T, or T = 1): the reference low-passes at full length,
    ifft(fft(pad(c)) phi)[pad_left : pad_left + N] (kymatio_phase_scattering.py:268-273).  Here: the transform form
    with one row per job (the dense operator would be N x N), through the host emulator against the oracle."""
    J, Q, T, N = CFG['S']
    p = PhasePlan(J, Q, T, N, N)                                # scattering output as long as the input
    assert p.dec == 1 and p.n_out == N and p.G is None and p.pair_plan is not None and p.pair_plan.n_paths == 1
    o = PhaseOracle(J, Q, T, N, N)
    rng = np.random.RandomState(19)
    rows = 3
    zi = (rng.randn(rows, N) + 1j * rng.randn(rows, N)).astype(np.complex64)
    zj = (rng.randn(rows, N) + 1j * rng.randn(rows, N)).astype(np.complex64)
    pw = np.array([1.0, 2.0, 4.5], np.float32)
    zp = np.stack([np.abs(zi), np.angle(zi)], -1).astype(np.float32)
    zc = np.stack([zj.real, zj.imag], -1).astype(np.float32)
    out = emu_pair_stage(p.pair_plan, zp, zc, pw)
    assert out.shape == (rows, N) and not np.isnan(out).any()
    theta = zp[..., 1].astype(np.float32) * pw[:, None]
    c = zp[..., 0].astype(np.float64) * np.exp(1j * theta.astype(np.float64)) * np.conj(zj.astype(np.complex128))
    ref = o._smooth(c).real
    assert ref.shape == out.shape
    err = np.linalg.norm(out - ref, axis=-1) / np.linalg.norm(ref, axis=-1)
    assert err.max() < 1e-5, err
