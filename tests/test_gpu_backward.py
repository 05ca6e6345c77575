"""Backward pass on the GPU: d/dx sum(S(x) w) through tebscat.Scattering1D's autograd node
(the transposed cascade of tebscat/large.py behind the C ABI) against

  * the float64 autograd oracle (oracle/scattering1d_grad_oracle.py), rel-L2 <= 1e-5 per signal, and
  * gradients of the live reference's own autograd graph (tests/golden/backward_*.npz, float32), <= 2e-5;

plus the reference's differentiability test (kymatio/tests/scattering1d/test_torch_scattering1d.py:292-315:
the gradient exists and is non-zero) and the adjoint identity <S'(x) v, w> = <v, S'(x)^T w> by finite differences."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, rel_l2
from oracle.scattering1d_grad_oracle import GradOracle

pytestmark = pytest.mark.gpu


def scat(d):
    from tebscat import Scattering1D
    return Scattering1D(int(d['J']), int(d['N']), int(d['Q']), max_order=int(d['max_order']), T=int(d['T']),
                        oversampling=int(d['oversampling'])).cuda()


@pytest.mark.parametrize('name', ['T', 'S', 'P1', 'O', 'H'])
def test_gradient_matches_oracle_and_reference(name):
    d = np.load(os.path.join(GOLDEN, 'backward_%s.npz' % name))
    S = scat(d)
    x = torch.from_numpy(d['x']).cuda().requires_grad_(True)
    out, _ = S(x)
    assert out.requires_grad
    assert rel_l2(out.detach().cpu().numpy(), d['S'], axis=-1).max() < 2e-5
    (out * torch.from_numpy(d['w']).cuda()).sum().backward()
    gx = x.grad.cpu().numpy()
    assert gx.shape == d['x'].shape and np.isfinite(gx).all()
    o = GradOracle(int(d['J']), int(d['N']), int(d['Q']), int(d['T']), int(d['max_order']), int(d['oversampling']))
    _, g64 = o.vjp(d['x'], d['w'])
    # signal 1 is randn: the north-star bound.  Signal 0 is CTG-shaped (baseline 140 bpm), where the reference's own
    # fp32 gradient is ~2e-5 away from the float64 one (tests/test_oracle_golden.py::test_gradient_oracle_vs_reference)
    e64, eref = rel_l2(gx, g64, axis=-1), rel_l2(gx, d['gx'], axis=-1)
    assert e64[1] < 1e-5 and e64[0] < 5e-5, e64
    assert eref[1] < 1e-5 and eref[0] < 5e-5, eref


def test_differentiability_like_the_reference():
    """test_differentiability_scattering (:292-315): J=6, Q=8, N=2**12, grad of the sum is not all zero."""
    from tebscat import Scattering1D
    S = Scattering1D(6, 2 ** 12, 8, T=64).cuda()
    x = torch.randn(2, 2 ** 12, device='cuda', requires_grad=True)
    s, _ = S.forward(x)
    loss = s.sum()
    loss.backward()
    assert torch.max(torch.abs(x.grad)) > 0.


def test_adjoint_identity_by_finite_differences():
    """<(S(x + h v) - S(x - h v)) / 2h, w> = <v, grad> to the accuracy of an fp32 central difference."""
    from tebscat import Scattering1D
    from tebscat.synth import randn_batch
    S = Scattering1D(5, 1500, 4, T=32).cuda()
    x = randn_batch(3, 1500, 1, seed=5)[:, 0].cuda()
    v = randn_batch(3, 1500, 1, seed=6)[:, 0].cuda()
    xg = x.clone().requires_grad_(True)
    out, _ = S(xg)
    w = torch.randn(out.shape, device='cuda', generator=torch.Generator(device='cuda').manual_seed(7))
    (out * w).sum().backward()
    lhs_grad = (xg.grad * v).sum(-1).double()
    h = 1e-2
    with torch.no_grad():
        sp, _ = S(x + h * v)
        sm, _ = S(x - h * v)
    fd = (((sp.double() - sm.double()) / (2 * h)) * w.double()).sum((-1, -2))
    assert torch.allclose(lhs_grad, fd, rtol=2e-3, atol=1e-4), (lhs_grad, fd)


def test_gradient_through_views_batch_shapes_and_large_level():
    """out_type='list' coefficients are views of the same node; leading batch dims; a padded length of 2^14 takes the
    large-support forward and the same backward."""
    from tebscat import Scattering1D
    S = Scattering1D(4, 1000, 4, T=16, out_type='list').cuda()
    x = torch.randn(2, 3, 1000, device='cuda', requires_grad=True)
    out, _ = S(x)
    sum(c['coef'].sum() for c in out).backward()
    ga = x.grad.clone()
    S2 = Scattering1D(4, 1000, 4, T=16).cuda()
    x2 = x.detach().clone().requires_grad_(True)
    S2(x2)[0].sum().backward()
    assert torch.allclose(ga, x2.grad, rtol=1e-5, atol=1e-6)

    N = 9000                                             # J_pad = 14
    L = Scattering1D(6, N, 4, T=64).cuda()
    assert L.J_pad == 14
    xl = torch.randn(2, N, device='cuda', requires_grad=True)
    sl, _ = L(xl)
    w = torch.randn(sl.shape, device='cuda')
    (sl * w).sum().backward()
    o = GradOracle(6, N, 4, 64)
    _, g64 = o.vjp(xl.detach().cpu().numpy(), w.cpu().numpy())
    assert rel_l2(xl.grad.cpu().numpy(), g64, axis=-1).max() < 1e-5


def test_unaveraged_backward_adjoint_identity():
    """average=False through the dict output (vectorize=False): <J v, w> = <v, J^T w> with J^T from the CUDA backward
    and J v from central differences of the CUDA forward (the map is piecewise smooth: |u| away from 0)."""
    from tebscat import Scattering1D
    S = Scattering1D(4, 1000, 4, T=16, average=False, vectorize=False).cuda()
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, 1000, generator=g).cuda()
    v = torch.randn(2, 1000, generator=g).cuda()
    with pytest.warns(DeprecationWarning):
        out, _ = S(x.clone().requires_grad_(True))
    keys = list(out.keys())
    ws = {k: torch.randn(out[k].shape, generator=g).cuda() for k in keys}
    xg = x.clone().requires_grad_(True)
    with pytest.warns(DeprecationWarning):
        o = S(xg)[0]
    sum((o[k] * ws[k]).sum() for k in keys).backward()
    lhs = float((xg.grad.double() * v.double()).sum())
    eps = 1e-2
    with torch.no_grad(), pytest.warns(DeprecationWarning):
        op, om = S(x + eps * v)[0], S(x - eps * v)[0]
    rhs = float(sum((((op[k] - om[k]) / (2 * eps)).double() * ws[k].double()).sum() for k in keys))
    assert abs(lhs - rhs) < 2e-3 * max(abs(lhs), abs(rhs)), (lhs, rhs)


def test_adjoint_entry_points_reject_bad_requests():
    """Status codes, never exceptions or crashes, across the C ABI (include/tebscat.h conventions)."""
    import ctypes
    from tebscat import _lib
    lib = _lib.load()
    ctx = ctypes.c_void_p()
    assert lib.tebscat_large_create(0, ctypes.byref(ctx)) == 0
    buf = torch.zeros(1 << 12, device='cuda')
    vp, st = ctypes.c_void_p, ctypes.c_void_p(0)
    ptr = vp(buf.data_ptr())
    assert lib.tebscat_large_unfold(ctx, ptr, ptr, ptr, 1, 8, 9, 0, 0, 0, 0, st) == _lib.TEBSCAT_EINVAL         # logk > log_src
    assert lib.tebscat_large_unfold(ctx, ptr, ptr, ptr, 1, 8, 4, 0, 2, 0, 0, st) == _lib.TEBSCAT_EINVAL         # empty chunk mask
    assert lib.tebscat_large_unstore(ctx, ptr, 1, 6, 60, 10, 3, 0, ptr, st) == _lib.TEBSCAT_EINVAL              # crop beyond the signal
    assert lib.tebscat_large_unstore(ctx, ptr, 1, 6, 0, 10, 3, 3, ptr, st) == _lib.TEBSCAT_EINVAL               # channel out of range
    assert lib.tebscat_large_pad_adjoint(ctx, ptr, 1, 100, 100, 8, ptr, st) == _lib.TEBSCAT_EINVAL              # pad >= N
    assert lib.tebscat_large_modulus_backward(ctx, None, ptr, 16, st) == _lib.TEBSCAT_EINVAL
    assert lib.tebscat_large_modulus_to(ctx, ptr, ptr, 0, st) == _lib.TEBSCAT_EINVAL
    assert b'padding' in lib.tebscat_last_error() or len(lib.tebscat_last_error()) > 0
    lib.tebscat_large_destroy(ctx)


@pytest.mark.parametrize('name', ['Tu', 'P1u'])
def test_unaveraged_backward_matches_reference_and_oracle(name):
    """average=False is differentiable like in the reference (its outputs are the moduli themselves): gradient of
    sum(coef_c * w_c) over the whole list output against the live reference's autograd (fixture) and the float64
    autograd oracle."""
    import os
    from helpers import GOLDEN
    from oracle.scattering1d_grad_oracle import GradOracle
    from tebscat import Scattering1D
    d = np.load(os.path.join(GOLDEN, 'backward_%s.npz' % name))
    J, N, Q, T, mo = int(d['J']), int(d['N']), int(d['Q']), int(d['T']), int(d['max_order'])
    S = Scattering1D(J, N, Q, max_order=mo, T=T, average=False, out_type='list').cuda()
    x = torch.from_numpy(d['x']).cuda().requires_grad_(True)
    out, _ = S(x)
    lengths = [int(v) for v in d['lengths']]
    assert [o['coef'].shape[-1] for o in out] == lengths
    ws = [torch.from_numpy(d['w0']).cuda()] + list(torch.split(torch.from_numpy(d['w']).cuda(), lengths[1:], dim=-1))
    sum((o['coef'] * w).sum() for o, w in zip(out, ws)).backward()
    gx = x.grad.cpu().numpy().astype(np.float64)
    _, g64 = GradOracle(J, N, Q, T, mo).vjp_unaveraged(d['x'], d['w'])
    g64 = g64 + d['w0']
    err = np.linalg.norm(gx - g64, axis=-1) / np.linalg.norm(g64, axis=-1)
    assert err[1] < 1e-5 and err[0] < 5e-5, err              # row 0 CTG-shaped, row 1 randn (as for average=True)
    ref = np.linalg.norm(gx - d['gx'], axis=-1) / np.linalg.norm(d['gx'], axis=-1)
    assert ref.max() < 5e-5, ref
    # second call: graph replay, fresh cotangents
    x2 = torch.from_numpy(d['x']).cuda().requires_grad_(True)
    out2, _ = S(x2)
    sum((o['coef'] * (2 * w)).sum() for o, w in zip(out2, ws)).backward()
    assert torch.allclose(x2.grad, 2 * x.grad, rtol=1e-5, atol=1e-6 * float(x.grad.abs().max()))
