"""Large-support level: padded lengths above 2^13 against the float64 oracle, and the pieces of the
level on their own (transforms of every length on global buffers against numpy)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle.scattering1d_oracle import ScatteringOracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('log_len', [3, 7, 10, 13, 14, 15, 16, 17])
def test_global_transforms_match_numpy(log_len):
    from tebscat import _lib
    from tebscat.large import LargeDevicePlan, LargePlan
    from tebscat.schedule import bitrev_indices
    dp = _get_ctx()
    L, n_tr = 1 << log_len, 5 if log_len > 10 else 37            # 37 transforms: the last tile is partial
    rng = np.random.RandomState(log_len)
    z = (rng.randn(n_tr, L) + 1j * rng.randn(n_tr, L)).astype(np.complex64)
    br = bitrev_indices(L)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    buf = torch.view_as_real(torch.from_numpy(z).cuda()).contiguous()
    _lib.check(_lib.load().tebscat_large_fft(dp.handle, ctypes.c_void_p(buf.data_ptr()), n_tr, log_len, 0, st))
    got = torch.view_as_complex(buf).cpu().numpy()
    ref = np.fft.fft(z.astype(np.complex128), axis=-1)[:, br]
    assert np.abs(got - ref).max() < 2e-6 * np.abs(ref).max()
    _lib.check(_lib.load().tebscat_large_fft(dp.handle, ctypes.c_void_p(buf.data_ptr()), n_tr, log_len, 1, st))
    back = torch.view_as_complex(buf).cpu().numpy() / L
    assert np.abs(back - z).max() < 2e-6 * np.abs(z).max()


_ctx = {}


def _get_ctx():
    from tebscat.large import LargeDevicePlan, LargePlan
    if 'dp' not in _ctx:
        class _P:                                            # every tile length, no cascade
            tile_lengths = list(range(1, 14))
            arena = np.zeros(4, np.float32)
        _ctx['dp'] = LargeDevicePlan(_P(), 0)
    return _ctx['dp']


@pytest.mark.parametrize('cfg', [(8, 8, 2 ** 13, 256, 2, 14), (6, 4, 12000, 64, 2, 14), (9, 2, 2 ** 14, 512, 1, 15),
                                 # the top of BASELINE configs[3] (J=10, Q=8, N = 2^15 and 2^16) and ragged lengths
                                 (10, 8, 2 ** 15, 1024, 2, 16), (10, 8, 2 ** 16, 1024, 2, 17),
                                 (8, 4, 50000, 256, 2, 16), (10, 2, 100000, 512, 2, 17)])
def test_large_support_matches_oracle(cfg):
    from tebscat import Scattering1D
    J, Q, N, T, mo, j_pad = cfg
    S = Scattering1D(J, N, Q, max_order=mo, T=T).cuda()
    assert S.J_pad == j_pad > 13
    x = torch.randn(3, N, generator=torch.Generator().manual_seed(J))
    out, P = S(x.cuda())
    torch.cuda.synchronize()
    out = out.cpu().numpy().astype(np.float64)
    ref = ScatteringOracle(J, N, Q, T, mo)(x.numpy())
    assert out.shape == ref.shape and P.shape == (3, 1) + ref.shape[1:]
    nr = np.linalg.norm(ref, axis=-1)
    err = np.linalg.norm(out - ref, axis=-1)
    assert np.all(err <= 1e-5 * nr + 1e-10 * nr.max()), float((err / nr).max())


def test_large_support_matches_the_live_reference(golden_dir):
    """tests/golden/scat_L.npz: outputs of the live reference at a padded length of 2^14 (oracle/make_golden.py)."""
    import os
    from tebscat import Scattering1D
    d = np.load(os.path.join(golden_dir, 'scat_L.npz'))
    S = Scattering1D(int(d['J']), int(d['N']), int(d['Q']), max_order=int(d['max_order']), T=int(d['T'])).cuda()
    assert S.J_pad == int(d['J_pad']) == 14
    out, _ = S(torch.from_numpy(d['x']).cuda())
    out = out.cpu().numpy().astype(np.float64)
    ref = d['S'].astype(np.float64)
    assert out.shape == ref.shape
    # signal 0 is CTG-shaped (the reference's own fp32 rounding dominates its weak paths), signal 1 randn
    err = np.linalg.norm(out - ref, axis=-1) / np.linalg.norm(ref, axis=-1)
    assert err[1].max() < 1e-5, err[1].max()
    assert np.linalg.norm(out[0] - ref[0]) / np.linalg.norm(ref[0]) < 2e-6


def test_large_support_output_conventions():
    """out_type='list' and vectorize=False at a padded length of 2^14: views of the same array (core :379-385)."""
    import warnings
    from tebscat import Scattering1D
    N = 9000
    x = torch.randn(2, N, generator=torch.Generator().manual_seed(3)).cuda()
    ref, _ = Scattering1D(6, N, 4, T=64).cuda()(x)
    lst, _ = Scattering1D(6, N, 4, T=64, out_type='list').cuda()(x)
    assert len(lst) == ref.shape[1]
    assert all(torch.equal(c['coef'], ref[:, i]) for i, c in enumerate(lst))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', DeprecationWarning)
        dct, _ = Scattering1D(6, N, 4, T=64, vectorize=False).cuda()(x)
    assert dct[()].shape == (2, 1, ref.shape[-1]) and torch.equal(dct[()][:, 0], ref[:, 0])


@pytest.mark.parametrize('graph', ['1', '0'])
def test_children_as_long_as_their_parent(graph, monkeypatch):
    """With oversampling a second-order path may not be subsampled at all (core/scattering1d.py:344-345): its
    transform is as long as its parent's, and the second-order workspace must hold it (regression: it was sized for
    half the padded length, and signals beyond the first half of the batch were overwritten)."""
    from tebscat import Scattering1D
    monkeypatch.setenv('TEBSCAT_LARGE_GRAPH', graph)
    J, Q, T, N, mo, os_ = 4, 1, 2, 5980, 2, 1
    S = Scattering1D(J, N, Q, max_order=mo, T=T, oversampling=os_).cuda()
    x = torch.randn(5, N, generator=torch.Generator().manual_seed(4))
    out = S(x.cuda())[0].cpu().numpy().astype(np.float64)
    assert S._op_by_op
    ref = ScatteringOracle(J, N, Q, T, mo, oversampling=os_)(x.numpy())
    assert out.shape == ref.shape
    assert (np.linalg.norm(out - ref, axis=-1) / np.linalg.norm(ref, axis=-1)).max() < 1e-5


@pytest.mark.parametrize('cfg', [(8, 8, 2 ** 13, 256, 2), (8, 8, 2 ** 14, 256, 2), (9, 2, 2 ** 14, 512, 1)])
def test_fused_subtrees_agree_with_the_op_by_op_level(cfg, monkeypatch):
    """above 2^13 the subtrees that fit one SM run on the fused kernel with a global source
    (tebscat_scat1d_forward_gsrc).  Same result as the level's own kernels, every channel written exactly once."""
    from tebscat import Scattering1D, large
    J, Q, N, T, mo = cfg
    x = torch.randn(5, N, generator=torch.Generator().manual_seed(11)).cuda()
    S = Scattering1D(J, N, Q, max_order=mo, T=T).cuda()
    a = S(x)[0]
    ldp = S._large_plan_for(0)[1]
    assert ldp._first is not None and (mo == 1 or ldp._kids)
    monkeypatch.setattr(large, 'HYBRID', False)
    S2 = Scattering1D(J, N, Q, max_order=mo, T=T).cuda()
    b = S2(x)[0]
    assert S2._large_plan_for(0)[1]._first is None and not S2._large_plan_for(0)[1]._kids
    torch.cuda.synchronize()
    a, b = a.cpu().numpy().astype(np.float64), b.cpu().numpy().astype(np.float64)
    assert np.isfinite(a).all()
    err = np.linalg.norm(a - b, axis=-1) / np.linalg.norm(b, axis=-1)
    assert err.max() < 3e-6, float(err.max())
    # a second call (the CUDA graph of the batch size) and another batch size give the same numbers
    assert torch.equal(S(x)[0].cpu(), torch.from_numpy(a.astype(np.float32)))
    assert torch.equal(S(x[:2])[0].cpu(), torch.from_numpy(a[:2].astype(np.float32)))


def test_plans_with_a_global_source_only_run_through_their_entry_point():
    from tebscat import Scattering1D, _lib
    lib = _lib.load()
    S = Scattering1D(8, 2 ** 13, 8, T=256).cuda()
    x = torch.randn(2, 2 ** 13).cuda()
    out = S(x)[0]
    ldp = S._large_plan_for(0)[1]
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    vp = ctypes.c_void_p
    with pytest.raises(ValueError, match='global source'):
        _lib.check(lib.tebscat_scat1d_forward(ldp._first.handle, vp(x.data_ptr()), 2, vp(out.data_ptr()), st))
    src = torch.zeros(2 << 14, 2, device='cuda')
    with pytest.raises(ValueError, match='stride'):
        _lib.check(lib.tebscat_scat1d_forward_gsrc(ldp._first.handle, vp(src.data_ptr()), 1 << 13, 2, vp(out.data_ptr()), st))
    with pytest.raises(ValueError, match='aligned'):
        _lib.check(lib.tebscat_scat1d_forward_gsrc(ldp._first.handle, vp(src.data_ptr() + 8), 1 << 14, 1, vp(out.data_ptr()), st))
    H = Scattering1D(6, 4800, 8, T=64).cuda()
    H(torch.randn(1, 4800).cuda())
    with pytest.raises(ValueError, match='no global-source task'):
        _lib.check(lib.tebscat_scat1d_forward_gsrc(H._plan_for(0).handle, vp(src.data_ptr()), 1 << 14, 1, vp(out.data_ptr()), st))
    # the library is still usable
    assert torch.isfinite(S(x)[0]).all()
