"""Channel bookkeeping of the transform: ``meta()`` and ``output_size()``.

Mirrors ``compute_meta_scattering`` (kymatio/scattering1d/utils.py:190-289) and
``precompute_size_scattering`` (utils.py:136-187): channels are ordered
[order 0] + [order 1 by n1] + [order 2, n1-major, j2 > j1]; the numeric fields
are NaN-padded float arrays of shape (C, max_order).
"""
import math

import numpy as np

from .filterbank import calibrate


def path_table(J, Q, T, max_order=2):
    """[(order, xi-tuple, sigma-tuple, j-tuple, n-tuple)] in output-channel order."""
    cal = calibrate(J, Q, T)
    rows = [(0, (), (), (), ())]
    first = list(zip(cal.xi1, cal.sigma1, cal.j1))
    second = list(zip(cal.xi2, cal.sigma2, cal.j2))
    rows += [(1, (xi,), (sg,), (j,), (n1,)) for n1, (xi, sg, j) in enumerate(first)]
    if max_order >= 2:
        for n1, (xi1, sg1, j1) in enumerate(first):
            for n2, (xi2, sg2, j2) in enumerate(second):
                if j2 > j1:
                    rows.append((2, (xi1, xi2), (sg1, sg2), (j1, j2), (n1, n2)))
    return rows


def compute_meta(J, Q, T, max_order=2):
    rows = path_table(J, Q, T, max_order)
    pad = lambda t: t + (math.nan,) * (max_order - len(t))
    return {
        'order': np.array([r[0] for r in rows]),
        'xi': np.array([pad(r[1]) for r in rows]),
        'sigma': np.array([pad(r[2]) for r in rows]),
        'j': np.array([pad(r[3]) for r in rows]),
        'n': np.array([pad(r[4]) for r in rows]),
        'key': [r[4] for r in rows],
    }


def output_size(J, Q, T, max_order=2, detail=False):
    rows = path_table(J, Q, T, max_order)
    counts = tuple(sum(1 for r in rows if r[0] == o) for o in range(max_order + 1))
    return counts if detail else sum(counts)
