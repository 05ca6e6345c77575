"""Golden fixture for the fused normalisation epilogue from the LIVE reference.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_normalize.py      (build container only)

`hdf5_dataset/hdf5_dataset.py` cannot be imported here (h5py is absent), so the reference's own
`normalize_tensor_data` is taken out of its source file at run time -- the FunctionDef node is compiled
and executed unmodified from /root/reference; nothing is copied into this repository -- and evaluated on the
live reference's scattering output of the production config.  The trim and the transposition of
`CombinedHDF5Dataset.__getitem__` (:733-741, :758-759) are applied around it the way __getitem__ does.
Writes tests/golden/normalize_fhr_st.npz.
"""
import ast
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
from tebscat.synth import ctg_batch                                           # noqa: E402


def reference_function(path, name):
    src = open(path).read()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    mod = ast.Module(body=[node], type_ignores=[])
    ns = {'torch': torch, 'np': np}
    exec('from typing import Union, Sequence, List, Tuple, Dict, Any, Optional', ns)
    exec(compile(mod, path, 'exec'), ns)
    return ns[name]


if __name__ == '__main__':
    normalize_tensor_data = reference_function(os.path.join(REF, 'hdf5_dataset', 'hdf5_dataset.py'), 'normalize_tensor_data')
    J, Q, T, N, mo = 11, 4, 16, 5760, 1                    # production config (create_hdf5_dataset.py:360)
    x = ctg_batch(3, N, seed=77)[:, 0, :].contiguous()     # three FHR records
    S = ScatteringTorch1D(J, N, Q, max_order=mo, T=T)(x)[0]            # (3, 43, 360)
    C = S.shape[1]
    rng = np.random.RandomState(5)
    # statistics of the transformed channels, like DatasetStatsCalculator produces them
    logS = torch.log(torch.clamp(S, min=0.0) + 1e-6)
    mean = S.mean(dim=(0, 2)).numpy().astype(np.float64)
    mean[1:] = logS.mean(dim=(0, 2)).numpy()[1:]
    var = S.var(dim=(0, 2)).numpy().astype(np.float64)
    var[1:] = logS.var(dim=(0, 2)).numpy()[1:]
    var *= rng.uniform(0.8, 1.25, size=C)
    stats = {'fhr_st': {'mean': mean, 'variance': var}}
    trim = 8                                               # trim_minutes=2 at 4 Hz / T=16 -> 30; any value works
    out = []
    for b in range(S.shape[0]):                            # __getitem__ works per sample: (C, L)
        rec = S[b][:, trim:-trim]                                                     # :738-741
        rec = normalize_tensor_data(rec, 'fhr_st', stats, {'fhr_st': 'all_except_0'}, {}, 1e-6)   # :563-573
        out.append(rec.transpose(0, 1).contiguous().numpy())                           # :758-759
    # a second variant: asinh on three channels, log on an explicit list, no trim, channel-major
    cfg_log, cfg_asinh = [1, 2, 5, 7], [3, 4, 6]
    out2 = np.stack([normalize_tensor_data(S[b], 'fhr_st', stats, {'fhr_st': cfg_log}, {'fhr_st': cfg_asinh}, 1e-6).numpy()
                     for b in range(S.shape[0])])
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'normalize_fhr_st.npz'),
                        x=x.numpy(), S=S.numpy(), mean=mean, variance=var, trim=trim, out=np.stack(out),
                        log2=np.asarray(cfg_log), asinh2=np.asarray(cfg_asinh), out2=out2,
                        config=np.asarray([J, Q, T, N, mo]))
    print('normalize_fhr_st.npz', np.stack(out).shape, out2.shape)
