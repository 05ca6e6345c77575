"""average=False (un-averaged U1 / U2 outputs, core/scattering1d.py:293-294, :329-330, :366-367): the oracle against
the live reference's output (tests/golden/unaveraged_T.npz, oracle/make_golden_unaveraged.py), the kernel's
schedule through the host emulator, and the frontend on a GPU."""
import os

import numpy as np
import pytest

from helpers import GOLDEN, emu_available, emu_forward, rel_l2
from oracle.scattering1d_oracle import ScatteringOracle
from tebscat.schedule import build_plan_unaveraged


def _fixture():
    d = np.load(os.path.join(GOLDEN, 'unaveraged_T.npz'))
    J, Q, T, N, mo = (int(v) for v in d['config'])
    offs = np.concatenate([[0], np.cumsum(d['lengths'])])
    return d, (J, Q, T, N, mo), offs


def test_oracle_unaveraged_matches_reference():
    d, (J, Q, T, N, mo), offs = _fixture()
    got = ScatteringOracle(J, N, Q, T, mo).unaveraged(d['x'])
    assert [v.shape[-1] for _, v in got] == list(d['lengths'])
    assert np.array_equal(got[0][1], d['x'])                                   # order 0 is the input itself
    for c, (key, v) in enumerate(got):
        ref = d['flat'][:, offs[c]:offs[c + 1]]
        assert rel_l2(v, ref) < 2e-6, (key, rel_l2(v, ref))


@pytest.mark.skipif(not emu_available(), reason='host emulator not built')
@pytest.mark.parametrize('cfg', [(5, 2, 8, 700, 2), (6, 8, 64, 4800, 2), (4, 4, 16, 1000, 1)])
def test_emulated_unaveraged_matches_oracle(cfg):
    J, Q, T, N, mo = cfg
    u = build_plan_unaveraged(J, N, Q, T, mo)
    x = np.random.RandomState(4).randn(2, N).astype(np.float32)
    row = emu_forward(u, x)[:, 0, :]
    assert not np.isnan(row).any()
    ref = ScatteringOracle(J, N, Q, T, mo).unaveraged(x)[1:]
    assert [k for k, _, _ in u.segments] == [k for k, _ in ref]                # the reference's path order
    for (key, off, ln), (_, v) in zip(u.segments, ref):
        assert v.shape[-1] == ln
        e = np.linalg.norm(row[:, off:off + ln] - v, axis=-1) / np.linalg.norm(v, axis=-1)
        assert e.max() < 1e-5, (key, e.max())


@pytest.mark.gpu
def test_gpu_unaveraged_matches_reference_and_oracle():
    import torch
    from tebscat import Scattering1D
    d, (J, Q, T, N, mo), offs = _fixture()
    S = Scattering1D(J, N, Q, max_order=mo, average=False, out_type='list', T=T).cuda()
    x = torch.from_numpy(d['x']).cuda()
    out, P = S(x)
    torch.cuda.synchronize()
    assert len(out) == len(d['lengths']) and out is P
    for c, o in enumerate(out):
        assert set(o) == {'coef', 'j'} and o['j'] == tuple(int(v) for v in d['j'][c] if v >= 0)
        ref = d['flat'][:, offs[c]:offs[c + 1]]
        assert o['coef'].shape == ref.shape
        assert rel_l2(o['coef'].cpu().numpy(), ref) < 1e-5, (c, o['j'])
    # the headline configuration against the oracle, leading batch dimensions kept
    J, N, Q, T = 6, 4800, 8, 64
    S = Scattering1D(J, N, Q, average=False, out_type='list', T=T).cuda()
    xh = torch.randn(2, 2, N, device='cuda', generator=torch.Generator('cuda').manual_seed(3))
    out = S(xh)[0]
    ref = ScatteringOracle(J, N, Q, T, 2).unaveraged(xh.cpu().numpy())
    assert len(out) == len(ref) == 126
    for o, (key, v) in zip(out, ref):
        assert o['coef'].shape == v.shape == (2, 2, v.shape[-1])
        e = np.linalg.norm(o['coef'].cpu().numpy() - v, axis=-1) / np.linalg.norm(v, axis=-1)
        assert e.max() < 1e-5, (key, e.max())
    with pytest.raises(ValueError):                                           # torch_frontend.py:174
        Scattering1D(J, N, Q, average=False, T=T).cuda()(xh)
