"""Fused normalisation + layout epilogue: the oracle restatement against outputs of the live
reference's normalize_tensor_data (tests/golden/normalize_fhr_st.npz, made by oracle/make_golden_normalize.py),
the kernel's store path through the host emulator, and -- on a GPU -- the C-ABI entry point."""
import os

import numpy as np
import pytest

from helpers import GOLDEN, emu_available, emu_forward
from oracle.normalize_oracle import asinh_channels_of, log_channels_of, normalize_record
from tebscat.schedule import build_plan

FIX = os.path.join(GOLDEN, 'normalize_fhr_st.npz')


def _fixture():
    d = np.load(FIX)
    J, Q, T, N, mo = (int(v) for v in d['config'])
    return d, (J, Q, T, N, mo)


def test_oracle_matches_reference_normalisation():
    d, _ = _fixture()
    C = d['S'].shape[1]
    got = normalize_record(d['S'], d['mean'], d['variance'], log_channels_of('all_except_0', C), (), 1e-6,
                           trim=int(d['trim']), time_major=True)
    assert got.shape == d['out'].shape
    np.testing.assert_allclose(got, d['out'], rtol=2e-6, atol=2e-6)
    got2 = normalize_record(d['S'], d['mean'], d['variance'], list(d['log2']), asinh_channels_of(list(d['asinh2']), C),
                            1e-6, trim=0, time_major=False)
    np.testing.assert_allclose(got2, d['out2'], rtol=2e-6, atol=2e-6)
    assert log_channels_of('all_except_0', 4) == [1, 2, 3] and asinh_channels_of('all', 3) == [0, 1, 2]


def _tolerance(d, C, trim, log_ch, rel=5e-5):
    """Against the REFERENCE's output the comparison inherits the transform's own fp32 error (a few 1e-5
    relative on the weakest paths of baseline-dominated records, for the reference as for us -- DESIGN 2),
    amplified by the epilogue: d/dS log(S + eps) = 1 / (S + eps).  The epilogue's own arithmetic is checked
    to 2e-6 against the oracle applied to OUR transform output (`_check_fused`)."""
    S = d['S'][..., trim:d['S'].shape[-1] - trim] if trim else d['S']
    std = np.sqrt(d['variance']).astype(np.float32)[:, None] + 1e-8
    tol = rel * np.abs(S) / std
    tol[:, log_ch, :] = rel * (np.abs(S[:, log_ch, :]) / (np.maximum(S[:, log_ch, :], 0) + 1e-6)) / std[log_ch]
    return tol + 1e-5


def _check_fused(fused, plain, d, log_ch, asinh_ch, trim, time_major):
    want = normalize_record(plain, d['mean'], d['variance'], log_ch, asinh_ch, 1e-6, trim=trim, time_major=time_major)
    assert fused.shape == want.shape
    np.testing.assert_allclose(fused, want, rtol=2e-6, atol=2e-6)


@pytest.mark.skipif(not emu_available(), reason='host emulator not built')
def test_emulated_epilogue_matches_reference():
    d, (J, Q, T, N, mo) = _fixture()
    p = build_plan(J, N, Q, T, mo)
    C, trim = p.n_paths, int(d['trim'])
    mode = np.zeros(C, np.uint8); mode[1:] = 1
    ep = dict(mean=d['mean'], std=np.sqrt(d['variance']), mode=mode, log_eps=1e-6, trim=trim, time_major=True)
    out = emu_forward(p, d['x'], epilogue=ep)
    plain = emu_forward(p, d['x'])
    assert out.shape == d['out'].shape and not np.isnan(out).any()
    _check_fused(out, plain, d, list(range(1, C)), (), trim, True)
    tol = np.swapaxes(_tolerance(d, C, trim, list(range(1, C))), -1, -2)
    assert np.all(np.abs(out - d['out']) <= tol), float((np.abs(out - d['out']) / tol).max())
    # channel-major, no trim, explicit log list + asinh channels
    mode2 = np.zeros(C, np.uint8); mode2[list(d['log2'])] = 1; mode2[list(d['asinh2'])] = 2
    ep2 = dict(mean=d['mean'], std=np.sqrt(d['variance']), mode=mode2, log_eps=1e-6, trim=0, time_major=False)
    out2 = emu_forward(p, d['x'], epilogue=ep2)
    _check_fused(out2, plain, d, list(d['log2']), list(d['asinh2']), 0, False)
    tol2 = _tolerance(d, C, 0, list(d['log2']))
    assert out2.shape == d['out2'].shape
    assert np.all(np.abs(out2 - d['out2']) <= tol2), float((np.abs(out2 - d['out2']) / tol2).max())


@pytest.mark.gpu
def test_gpu_forward_normalized_matches_reference():
    import torch
    from tebscat import Scattering1D
    d, (J, Q, T, N, mo) = _fixture()
    S = Scattering1D(J, N, Q, max_order=mo, T=T).cuda()
    x = torch.from_numpy(d['x']).cuda()
    trim = int(d['trim'])
    out = S.forward_normalized(x, d['mean'], d['variance'], log_channels='all_except_0', log_epsilon=1e-6, trim=trim)
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    C = d['S'].shape[1]
    assert out.shape == d['out'].shape
    tol = np.swapaxes(_tolerance(d, C, trim, list(range(1, C))), -1, -2)
    assert np.all(np.abs(out - d['out']) <= tol), float((np.abs(out - d['out']) / tol).max())
    out2 = S.forward_normalized(x, d['mean'], d['variance'], log_channels=[int(v) for v in d['log2']],
                                asinh_channels=[int(v) for v in d['asinh2']], trim=0, time_major=False).cpu().numpy()
    tol2 = _tolerance(d, C, 0, list(d['log2']))
    assert np.all(np.abs(out2 - d['out2']) <= tol2), float((np.abs(out2 - d['out2']) / tol2).max())
    # and it is the plain transform followed by the oracle's restatement of the post-processing
    plain = S(x)[0].cpu().numpy()
    _check_fused(out, plain, d, list(range(1, C)), (), trim, True)
    _check_fused(out2, plain, d, list(d['log2']), list(d['asinh2']), 0, False)
    with pytest.raises(ValueError):
        S.forward_normalized(x, d['mean'], d['variance'], trim=10 ** 6)
    with pytest.raises(ValueError):
        S.forward_normalized(x, d['mean'][:-1], d['variance'][:-1])
