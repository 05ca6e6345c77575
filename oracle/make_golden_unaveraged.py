"""Golden fixture for average=False from the LIVE reference (build container only).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_unaveraged.py

Runs the reference's ScatteringTorch1D(average=False, out_type='list') on CPU on a seeded input and stores the
input, every path's coefficients back to back, and the (j, length) of every path: tests/golden/unaveraged_T.npz.
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
from tebscat.synth import randn_batch                                         # noqa: E402

if __name__ == '__main__':
    J, Q, T, N, mo = 5, 2, 8, 700, 2
    x = randn_batch(3, N, seed=21)[:, 0, :].contiguous()
    out = ScatteringTorch1D(J, N, Q, max_order=mo, average=False, out_type='list', T=T)(x)[0]
    coefs = [o['coef'].numpy() for o in out]
    js = [tuple(int(v) for v in o['j']) for o in out]
    flat = np.concatenate([c.reshape(3, -1) for c in coefs], axis=1)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'unaveraged_T.npz'), x=x.numpy(), flat=flat,
                        lengths=np.asarray([c.shape[-1] for c in coefs]),
                        j=np.asarray([list(j) + [-1] * (2 - len(j)) for j in js]), config=np.asarray([J, Q, T, N, mo]))
    print('unaveraged_T.npz', flat.shape, len(coefs))
