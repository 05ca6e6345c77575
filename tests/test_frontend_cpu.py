"""Host-side behaviour of the Scattering1D frontend (no GPU): geometry, filters, meta
ordering and the error behaviour of the reference's frontend, re-expressed from
kymatio/tests/scattering1d/test_torch_scattering1d.py and test_utils_scattering1d.py."""
import os
import re

import numpy as np
import pytest
import torch

from helpers import CONFIGS, GOLDEN, ROOT
from tebscat import Scattering1D
from tebscat import filterbank as fbk
from tebscat.meta import compute_meta, output_size


@pytest.mark.parametrize('name', ['H', 'P', 'S', 'T'])
def test_geometry_filters_meta_match_reference(name):
    d = np.load(os.path.join(GOLDEN, 'scat_%s.npz' % name))
    J, N, Q, T, mo = CONFIGS[name]
    S = Scattering1D(J, N, Q, max_order=mo, T=T)
    assert (S.J_pad, S.pad_left, S.pad_right) == (int(d['J_pad']), int(d['pad_left']), int(d['pad_right']))
    assert [S.ind_start[j] for j in range(J + 1)] == list(d['ind_start'])
    assert [S.ind_end[j] for j in range(J + 1)] == list(d['ind_end'])
    # buffers tensor0.. in the reference's registration order, fp32 (L,1)
    bufs = dict(S.named_buffers())
    assert len(bufs) == int(d['n_filters'])
    assert [bufs['tensor%d' % i].shape[0] for i in range(len(bufs))] == list(d['filter_len'])
    assert all(b.dtype == torch.float32 and b.shape[1] == 1 for b in bufs.values())
    np.testing.assert_array_equal(S.psi1_f[0]['levels'][0].numpy()[:, 0], d['psi1_0'].astype(np.float32))
    np.testing.assert_array_equal(S.psi1_f[-1]['levels'][0].numpy()[:, 0], d['psi1_last'].astype(np.float32))
    np.testing.assert_array_equal(S.psi2_f[-1]['levels'][-1].numpy()[:, 0], d['psi2_last_top'].astype(np.float32))
    np.testing.assert_array_equal(S.phi_f['levels'][0].numpy()[:, 0], d['phi_0'].astype(np.float32))
    m = S.meta()
    assert [tuple(int(v) for v in row if v >= 0) for row in d['keys']] == m['key']      # bit-exact ordering
    np.testing.assert_array_equal(m['order'], d['order'])
    np.testing.assert_array_equal(m['xi'], d['meta_xi'])
    np.testing.assert_array_equal(m['sigma'], d['meta_sigma'])
    np.testing.assert_array_equal(m['j'], d['meta_j'])
    assert tuple(d['output_size']) == S.output_size(detail=True)
    assert S.output_size() == len(m['key'])


def test_compute_padding_and_border_indices():
    """test_utils_scattering1d.py:8-53."""
    assert fbk.padding(5, 16) == (8, 8)
    with pytest.raises(ValueError) as ve:
        fbk.padding(3, 16)
    assert 'should be larger' in ve.value.args[0]
    with pytest.raises(ValueError) as ve:
        fbk.padding(6, 16)
    assert 'Too large padding value' in ve.value.args[0]
    rng = np.random.RandomState(42)
    J_signal, J = 10, 6
    n = 2 ** J_signal
    i0 = rng.randint(0, n // 2 + 1, 1)[0]
    i1 = rng.randint(i0 + 1, n, 1)[0]
    x = np.ones(n)
    x[i0:i1] = 0.
    start, end = fbk.border_indices(J, i0, i1)
    for j in range(J + 1):
        xs = x[::2 ** j]
        assert np.max(xs[start[j]:end[j]]) == 0.
        if start[j] > 0:
            assert np.min(xs[:start[j]]) > 0.
        if end[j] < xs.shape[-1]:
            assert np.min(xs[end[j]:]) > 0.


def test_filter_properties():
    """kymatio/tests/scattering1d/test_filters_scattering1d.py: periodisation == decimation,
    l1 norm, Morlet zero mean, Gaussian symmetry."""
    n = 1024
    psi = fbk.morlet_spectrum(n, 0.2, 0.02)
    assert abs(psi[0]) < 1e-12                                    # zero mean
    assert abs(np.abs(np.fft.ifft(psi)).sum() - 1.0) < 1e-9       # l1 normalised
    g = fbk.gauss_spectrum(n, 0.01)
    np.testing.assert_allclose(g[1:], g[1:][::-1], atol=1e-14)    # symmetric
    for k in (2, 4, 8):
        np.testing.assert_allclose(np.fft.ifft(fbk.fold(psi, k)), np.fft.ifft(psi)[::k], atol=1e-12)
    with pytest.raises(ValueError):
        fbk.morlet_spectrum(n, 0.2, 0.02, P_max=5.1)
    with pytest.raises(ValueError):
        fbk.gauss_spectrum(n, 0.02, P_max=-1)


def test_constructor_errors_and_options():
    with pytest.raises(ValueError) as ve:
        Scattering1D(4, (512, 2), 4)
    assert 'exactly one element' in ve.value.args[0]
    with pytest.raises(ValueError) as ve:
        Scattering1D(4, 512.0, 4)
    assert 'integer or a 1-tuple' in ve.value.args[0]
    with pytest.raises(ValueError) as ve:
        Scattering1D(4, 512, 4, T=64)
    assert 'cannot exceed 2**J' in ve.value.args[0]
    with pytest.raises(ImportError):
        Scattering1D(4, 512, 4, backend='numpy')
    S = Scattering1D(4, (512,), (4, 1))                           # tuple shape, (Q1, 1) and T=None accepted
    assert S.T == 16 and S.N == 512
    x = torch.zeros(2, 512)
    S.average = False
    with pytest.raises(ValueError) as ve:
        S(x)
    assert 'mutually incompatible' in ve.value.args[0]
    S.average = True
    S.out_type = 'doesnotexist'
    with pytest.raises(RuntimeError) as ve:
        S(x)
    assert "must be one of 'array' or 'list'" in ve.value.args[0]
    S.out_type = 'array'
    with pytest.raises(TypeError) as ve:
        S(None)
    assert 'should be not empty' in ve.value.args[0]
    with pytest.raises(RuntimeError) as ve:
        S(torch.zeros(512, 2).t())
    assert 'must be contiguous' in ve.value.args[0]
    with pytest.raises(ValueError):
        S(torch.zeros(2, 100))
    with pytest.raises(TypeError):
        S(torch.zeros(2, 512, dtype=torch.float64))
    with pytest.raises(TypeError) as ve:                          # no CPU path: the input must live on the GPU
        S(x)
    assert 'GPU' in ve.value.args[0]


def test_large_support_goes_to_the_large_level():
    """Padded lengths above 2^13 do not fit the single-CTA schedule (which must say so) and are served by the
    large-support level (tebscat/large.py): same channel order as meta(), same output geometry."""
    from tebscat.large import LargePlan
    S = Scattering1D(10, 2 ** 16, 8)
    assert S.J_pad == 17
    with pytest.raises(NotImplementedError):
        S._schedule()
    lp = LargePlan(8, 2 ** 13, 8, 256)
    S8 = Scattering1D(8, 2 ** 13, 8, T=256)
    assert lp.geo.J_pad == 14 and lp.n_paths == S8.output_size() == len(S8.meta()['key'])
    assert [tuple(k) for k in S8.meta()['key']] == lp.keys
    assert lp.n_out == 2 ** 13 // 256 and lp.lf == 14 - 8
    assert [e['ch'] for e in lp.first] == list(range(1, len(lp.first) + 1))
    chans = [0] + [e['ch'] for e in lp.first] + [k['ch'] for e in lp.first for k in e['kids']]
    assert sorted(chans) == list(range(lp.n_paths))                       # every channel produced exactly once
    for e in lp.first:
        off, log_src, logk, mask, logcw, sexp = e['mul']
        assert log_src == 14 and log_src - logk == e['l1'] and sexp == log_src
        for k in e['kids']:
            assert k['mul'][1] == e['l1'] and k['mul'][1] - k['mul'][2] == k['l2']
    with pytest.raises(NotImplementedError):
        LargePlan(12, 2 ** 18, 8, 4096)


def test_c_abi_exports_every_declared_symbol():
    """The C-ABI library loads and exports what include/tebscat.h declares (no compute)."""
    from tebscat import _lib
    header = open(os.path.join(ROOT, 'include', 'tebscat.h')).read()
    names = set(re.findall(r'\b(tebscat_[a-z0-9_]+)\s*\(', header))
    assert {'tebscat_plan_create', 'tebscat_scat1d_forward', 'tebscat_last_error'} <= names
    lib = _lib.load()
    for n in sorted(names):
        assert hasattr(lib, n), n
    assert lib.tebscat_abi_version() == _lib.ABI_VERSION


def test_missing_library_fails_loudly(monkeypatch):
    from tebscat import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libtebscat.so')
    with pytest.raises(_lib.TebscatError) as ve:
        _lib.load()
    assert 'no CPU fallback' in str(ve.value)


def test_op_by_op_plan_workspace_geometry():
    """Host logic of the op-by-op level: the second-order workspace follows the longest child -- half the padded length
    normally, the full length when oversampling leaves a child un-subsampled (core/scattering1d.py:344-345) -- and the
    configurations the fused schedule cannot hold are the ones with output-rate lengths of 2048 and more while U0
    lives in shared memory, of 8192 with U0 in the global scratch."""
    from tebscat.large import LargePlan
    from tebscat.schedule import build_plan
    assert LargePlan(6, 9000, 4, 64, 2).max_l2 == 13                      # Np = 2^14: children at most Np / 2
    lp = LargePlan(4, 5980, 1, 2, 2, oversampling=1)                      # T = 2, oversampling = 1: nothing is subsampled
    assert lp.geo.J_pad == 13 and lp.lf == 13 and lp.max_l2 == 13
    assert all(k['mul'][2] == 0 for e in lp.first for k in e['kids'])     # logk = 0 for every child
    assert LargePlan(6, 4800, 8, 64, 1).max_l2 == 1                       # first order only: no children
    with pytest.raises(NotImplementedError):
        build_plan(4, 4827, 8, 4, 2, tune=dict(u0_scratch=False))         # lf = 11 with U0 in shared memory (round 1)
    build_plan(4, 4827, 8, 8, 2, tune=dict(u0_scratch=False))             # lf = 10 fits
    # with U0 parked in the global scratch (the default) output-rate lengths of 2048 and 4096 fit too; 8192 does not
    assert build_plan(4, 4827, 8, 4, 2).scratch_complex == 8192           # lf = 11
    build_plan(4, 4827, 8, 2, 2)                                          # lf = 12
    with pytest.raises(NotImplementedError):
        build_plan(4, 4827, 8, 1, 2)                                      # lf = 13


def test_plan_validation_runs_before_any_device_call():
    """tebscat_plan_create validates the schedule on the host before it touches CUDA, so malformed plans are
    refused with TEBSCAT_EINVAL on a box without a GPU too: a thread range outside the CTA, and two tasks of
    one step on the same warps (the check that used to sit behind a double free)."""
    from tebscat.schedule import build_plan
    from tebscat.torch_frontend import _DevicePlan
    p = build_plan(5, 700, 2, 8, 2)
    p.tasks = p.tasks.copy()
    p.tasks[3, 1] = 5000
    with pytest.raises(ValueError) as e:
        _DevicePlan(p, 0)
    assert 'outside the CTA' in str(e.value)
    p = build_plan(5, 700, 2, 8, 2)
    p.tasks = p.tasks.copy()
    st = next(s for s in range(p.steps.shape[0]) if p.steps[s, 1] - p.steps[s, 0] >= 2)
    p.tasks[p.steps[st, 0] + 1, 1] = p.tasks[p.steps[st, 0], 1]
    with pytest.raises(ValueError) as e:
        _DevicePlan(p, 0)
    assert 'overlapping thread ranges' in str(e.value)


def test_global_source_tasks_are_validated_on_the_host():
    """The global-source multiplies (OP_GMULFOLD / OP_GMULFOLD2: U0 parked in the L2 scratch, fused subtrees of the
    large-support level) are checked like every other task before any device call: source extent against the
    scratch, alignment of the source offset, destination inside shared memory, filters inside the arena."""
    from tebscat.schedule import OP_GMULFOLD, OP_GMULFOLD2, build_hybrid_plans, build_plan
    from tebscat.torch_frontend import _DevicePlan

    def refused(plan, text):
        with pytest.raises(ValueError) as e:
            _DevicePlan(plan, 0)
        assert text in str(e.value), str(e.value)

    def fresh():
        p = build_plan(5, 700, 2, 8, 2)
        p.tasks = p.tasks.copy()
        return p

    p = fresh()
    assert p.scratch_complex == 1 << p.geo.J_pad
    ops = p.tasks[:, 0] & 0xff
    g2 = int(np.flatnonzero(ops == OP_GMULFOLD2)[0])
    g1 = int(np.flatnonzero(ops == OP_GMULFOLD)[0])
    p.scratch_complex = 512                         # smaller than the spectrum the schedule parks and reads
    refused(p, 'bad scratch size')
    p = fresh()
    p.tasks[g1, 3] = 2                              # source offset must be a multiple of four bins
    refused(p, 'bad GMULFOLD')
    p = fresh()
    p.tasks[g2, 3] = 1 << 20                        # destination of filter B outside shared memory
    refused(p, 'bad GMULFOLD2')
    p = fresh()
    p.tasks[g2, 9] = p.arena.size                   # filter B outside the arena
    refused(p, 'GMULFOLD2 filter outside the arena')
    p = fresh()
    p.tasks[g1, 4] = 20                             # a source of 2^20 bins
    refused(p, 'bad GMULFOLD')
    # a fused subtree of the large-support level: destinations are at most 8192 bins
    h = build_hybrid_plans(5, 9000, 4, 32)['first']
    h.tasks = h.tasks.copy()
    g = int(np.flatnonzero((h.tasks[:, 0] & 0xff) == OP_GMULFOLD2)[0])
    h.tasks[g, 5] = 0                               # no periodisation: 2^14 output bins
    refused(h, 'bad GMULFOLD2')


def test_level_selection_follows_the_current_options():
    """Which level serves a configuration (fused single kernel / op by op) is decided per schedule key, not latched:
    `oversampling` is read at call time like in the reference (core/scattering1d.py:260-261), so changing it on a
    live module must re-decide in both directions."""
    from tebscat import Scattering1D
    S = Scattering1D(6, 4800, 8, T=8)
    assert not S._op_by_op                      # output rate 1024 samples: fits the fused schedule
    S.oversampling = 1
    assert S._op_by_op                          # 2048 samples at the output rate: op-by-op level
    S.oversampling = 0
    assert not S._op_by_op


def test_plan_files_are_validated(tmp_path):
    """tebscat_plan_save / tebscat_plan_load (the ABI without the Python scheduler at run time): a saved plan reaches
    plan creation (on a box without a GPU that is the first CUDA call: RuntimeError), anything else is refused on the
    host with TEBSCAT_EINVAL -- wrong magic, truncation, trailing bytes, a flipped payload bit (checksum), and a
    schedule that does not validate is not even written."""
    import ctypes
    from tebscat import _lib
    from tebscat.export_plan import main as export_main, save_plan
    from tebscat.schedule import build_plan
    lib = _lib.load()
    path = str(tmp_path / 'T.tebplan')
    export_main(['--J', '5', '--shape', '700', '--Q', '2', '--T', '8', path])
    blob = open(path, 'rb').read()
    assert blob[:8] == b'TEBSCATP'

    def load(p):
        h = ctypes.c_void_p()
        rc = lib.tebscat_plan_load(p.encode(), 0, ctypes.byref(h))
        if rc == 0:
            lib.tebscat_plan_destroy(h)
        return rc, lib.tebscat_last_error().decode()

    rc, msg = load(path)
    assert rc in (_lib.TEBSCAT_OK, _lib.TEBSCAT_ECUDA), (rc, msg)       # valid file: only the device can object
    for name, data, word in (('magic', b'X' + blob[1:], 'not a tebscat plan file'),
                             ('short', blob[:-100], 'truncated'),
                             ('long', blob + b'\0', 'trailing'),
                             ('flip', blob[:5000] + bytes([blob[5000] ^ 1]) + blob[5001:], 'checksum')):
        bad = str(tmp_path / (name + '.tebplan'))
        open(bad, 'wb').write(data)
        rc, msg = load(bad)
        assert rc == _lib.TEBSCAT_EINVAL and word in msg, (name, rc, msg)
    rc, msg = load(str(tmp_path / 'missing.tebplan'))
    assert rc == _lib.TEBSCAT_EINVAL and 'cannot open' in msg
    p = build_plan(5, 700, 2, 8, 2)
    p.tasks = p.tasks.copy()
    p.tasks[3, 1] = 5000                                               # thread range outside the CTA
    with pytest.raises(ValueError):
        save_plan(p, str(tmp_path / 'bad.tebplan'))
    assert not (tmp_path / 'bad.tebplan').exists()


def test_large_plan_splits_the_cascade_between_the_levels(monkeypatch):
    """Padded lengths above 2^13: the host plan of the large-support level hands every subtree of at most
    8192 samples to the fused kernel -- the first-order filters whose subsampled length fits, and per longer filter the
    children that fit -- and keeps exactly the rest; filters of one scale share a schedule with channels a constant
    shift apart; TEBSCAT_HYBRID=0 keeps everything op by op; at 2^13 and below there is nothing to split."""
    from tebscat import large
    from tebscat.large import LargePlan
    lp = LargePlan(8, 2 ** 13, 8, 256)                                     # Np = 2^14, 54 first-order filters
    hyb = lp.hybrid_plans()
    assert hyb is lp.hybrid_plans()                                        # built once
    n = lp.geo.J_pad
    small = {e['n1'] for e in lp.first if e['l1'] <= 13}
    assert set(hyb['first_n1']) == small and 0 < len(small) < len(lp.first)
    assert hyb['first'].n_paths == lp.n_paths and hyb['first'].n_out == lp.n_out
    long_n1 = {e['n1'] for e in lp.first if e['l1'] > 13}
    covered = set()
    ch = lp._channel
    for plan, members, done in hyb['kids']:
        covered |= set(members)
        assert set(members) <= long_n1 and plan.head == members[0]
        shifts = {tuple(ch[(n1, n2)] - ch[(plan.head, n2)] for n2 in done) for n1 in members}
        assert all(len(set(s)) == 1 for s in shifts)                       # one shift per member
        for e in lp.first:
            if e['n1'] in members:                                         # the children left to the level are the long ones
                assert {k['n2'] for k in e['kids'] if k['l2'] <= 13} == set(done)
    assert covered == {e['n1'] for e in lp.first if e['l1'] > 13 and any(k['l2'] <= 13 for k in e['kids'])}
    assert LargePlan(6, 4800, 8, 64).hybrid_plans() is None                # Np = 2^13: the fused cascade serves it whole
    monkeypatch.setattr(large, 'HYBRID', False)
    assert LargePlan(8, 2 ** 13, 8, 256).hybrid_plans() is None
