"""Gradient fixtures from the LIVE reference (build container only; needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_backward.py

The reference's torch frontend is differentiable (ModulusStable; test_differentiability_scattering,
kymatio/tests/scattering1d/test_torch_scattering1d.py:292-315).  For seeded inputs x and cotangents w the
script stores d/dx sum(S(x) * w) as computed by the reference's own autograd graph on CPU (float32).
"""
import os
import sys

import numpy as np
import scipy.special
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
sys.path.insert(0, os.path.join(REF, 'kymatio'))
sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
sys.dont_write_bytecode = True
if not hasattr(scipy.special, 'sph_harm'):
    scipy.special.sph_harm = None

from kymatio.scattering1d.frontend.torch_frontend import ScatteringTorch1D   # noqa: E402
from tebscat.synth import ctg_batch, randn_batch                              # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
CONFIGS = {
    # name: (J, Q, T, N, max_order, oversampling)
    'T': (5, 2, 8, 700, 2, 0),
    'S': (4, 4, 16, 1000, 2, 0),
    'H': (6, 8, 64, 4800, 2, 0),
    'P1': (8, 4, 16, 2000, 1, 0),      # max_order = 1, T < 2**J
    'O': (5, 4, 32, 1200, 2, 1),       # oversampling = 1
}

if __name__ == '__main__':
    torch.set_num_threads(8)
    for name, (J, Q, T, N, mo, os_) in ([] if 'unaveraged' in sys.argv[1:] else CONFIGS.items()):
        S = ScatteringTorch1D(J, N, Q, max_order=mo, T=T, oversampling=os_)
        x = torch.cat([ctg_batch(1, N, seed=311)[:, 0], randn_batch(1, N, 1, seed=312)[:, 0]], 0).requires_grad_(True)
        out, _ = S(x)
        w = torch.randn(out.shape, generator=torch.Generator().manual_seed(313))
        (out * w).sum().backward()
        np.savez_compressed(os.path.join(OUT, 'backward_%s.npz' % name), J=J, Q=Q, T=T, N=N, max_order=mo, oversampling=os_,
                            x=x.detach().numpy(), w=w.numpy(), S=out.detach().numpy(), gx=x.grad.numpy())
        print(name, tuple(out.shape), float(x.grad.norm()))
    # average=False: the un-averaged moduli are the outputs (core/scattering1d.py:329-330, :366-367); the cotangent
    # covers every coefficient of the list output, order 0 (the input itself) included
    for name, (J, Q, T, N, mo, os_) in {'Tu': (5, 2, 8, 700, 2, 0), 'P1u': (6, 4, 16, 1000, 1, 0)}.items():
        S = ScatteringTorch1D(J, N, Q, max_order=mo, T=T, oversampling=os_, average=False, out_type='list')
        x = torch.cat([ctg_batch(1, N, seed=411)[:, 0], randn_batch(1, N, 1, seed=412)[:, 0]], 0).requires_grad_(True)
        out, _ = S(x)
        gen = torch.Generator().manual_seed(413)
        ws = [torch.randn(o['coef'].shape, generator=gen) for o in out]
        sum((o['coef'] * w).sum() for o, w in zip(out, ws)).backward()
        np.savez_compressed(os.path.join(OUT, 'backward_%s.npz' % name), J=J, Q=Q, T=T, N=N, max_order=mo, oversampling=os_,
                            x=x.detach().numpy(), w0=ws[0].numpy(), w=torch.cat(ws[1:], dim=-1).numpy(),
                            row=torch.cat([o['coef'] for o in out[1:]], dim=-1).detach().numpy(),
                            lengths=np.array([o['coef'].shape[-1] for o in out]), gx=x.grad.numpy())
        print(name, len(out), float(x.grad.norm()))
