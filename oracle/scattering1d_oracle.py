"""ORACLE (test infrastructure only -- never imported by the product path).

numpy restatement of the reference's 1-D scattering cascade for
``average=True, vectorize=True, out_type='array'`` (any ``oversampling``):

  kymatio/kymatio/scattering1d/core/scattering1d.py:269-399   (cascade, ordering)
  kymatio/kymatio/scattering1d/backend/torch_backend.py:18-128 (pad, rfft, ifft,
      irfft, subsample_fourier, unpad)
  kymatio/kymatio/backend/torch_backend.py:137-219             (modulus, cdgmm)
  kymatio/kymatio/scattering1d/frontend/torch_frontend.py:163-255 (batch reshape,
      fp32 filter cast :82-96)

Default arithmetic is float64 on the *fp32-cast* filters (the reference casts its
float64 filters with ``.float()`` before use); ``dtype=np.complex64`` runs the
transforms in single precision like the reference does.

Pinned by tests/test_oracle_golden.py against (1) the reference's own KAT
``kymatio/tests/scattering1d/test_data_1d.npz`` and (2) fixtures generated from
the live reference (oracle/make_golden.py).
"""
import math

import numpy as np
import scipy.fft

from . import filters_oracle as fo


def reflect_pad(x, pad_left, pad_right):
    """torch_backend.py:50-78 -- F.pad(..., mode='reflect') (edge not repeated)."""
    if pad_left >= x.shape[-1] or pad_right >= x.shape[-1]:
        raise ValueError('Indefinite padding size (larger than tensor).')
    return np.pad(x, [(0, 0)] * (x.ndim - 1) + [(pad_left, pad_right)], mode='reflect')


def subsample_fourier(x_f, k):
    """torch_backend.py:18-48 -- periodisation by the mean of k blocks."""
    n = x_f.shape[-1]
    return x_f.reshape(x_f.shape[:-1] + (k, n // k)).mean(axis=-2)


class ScatteringOracle:
    def __init__(self, J, N, Q, T, max_order=2, cdtype=np.complex128, oversampling=0):
        self.J, self.N, self.Q, self.T, self.max_order = J, N, Q, T, max_order
        self.oversampling = oversampling
        self.cdtype = cdtype
        self.rdtype = np.float64 if cdtype == np.complex128 else np.float32
        self.geo = fo.geometry(N, J, Q, T)
        bank = fo.filter_factory(self.geo['J_pad'], J, Q, T)
        cast = lambda a: a.astype(np.float32).astype(self.rdtype)     # .float() then compute dtype
        self.phi = [cast(a) for a in bank['phi']]
        self.psi1 = [dict(p, levels=[cast(a) for a in p['levels']]) for p in bank['psi1']]
        self.psi2 = [dict(p, levels=[cast(a) for a in p['levels']]) for p in bank['psi2']]
        self.keys = fo.path_keys(J, Q, T, max_order)

    def _fft(self, u):
        return scipy.fft.fft(u.astype(self.cdtype), axis=-1)

    def _ifft(self, u_f):
        return scipy.fft.ifft(u_f.astype(self.cdtype), axis=-1)

    def __call__(self, x):
        """x: (..., N) real -> (..., C, N_out) in the compute real dtype."""
        x = np.asarray(x)
        batch_shape = x.shape[:-1]
        x = x.reshape(-1, x.shape[-1]).astype(self.rdtype)
        g = self.geo
        log2_T = math.floor(math.log2(self.T))
        os_ = self.oversampling
        i0, i1 = g['ind_start'], g['ind_end']
        out = []

        U0_f = self._fft(reflect_pad(x, g['pad_left'], g['pad_right']))          # :278-280
        k0 = max(log2_T - os_, 0)                                              # :285
        S0 = self._ifft(subsample_fourier(U0_f * self.phi[0], 2 ** k0)).real   # :288-290
        out.append(S0[:, i0[k0]:i1[k0]])                                       # :292
        order2 = []
        for n1, p1 in enumerate(self.psi1):                                    # :300
            j1 = p1['j']
            k1 = max(min(j1 - os_, log2_T - os_), 0)                           # :304
            assert p1['xi'] < 0.5 / (2 ** k1)                                  # :306
            U1 = np.abs(self._ifft(subsample_fourier(U0_f * p1['levels'][0], 2 ** k1)))   # :307-315
            U1_f = self._fft(U1)                                               # :318
            k1_J = max(log2_T - k1 - os_, 0)                                   # :322
            S1 = self._ifft(subsample_fourier(U1_f * self.phi[k1], 2 ** k1_J)).real       # :323-325
            out.append(S1[:, i0[k1_J + k1]:i1[k1_J + k1]])                     # :327
            if self.max_order == 2:
                for n2, p2 in enumerate(self.psi2):                            # :337
                    j2 = p2['j']
                    if j2 > j1:
                        assert p2['xi'] < p1['xi']                             # :341
                        k2 = max(min(j2 - k1 - os_, log2_T - k1 - os_), 0)     # :344-345
                        U2 = np.abs(self._ifft(subsample_fourier(U1_f * p2['levels'][k1], 2 ** k2)))
                        U2_f = self._fft(U2)                                   # :355
                        k2_J = max(log2_T - k2 - k1 - os_, 0)                  # :358
                        S2 = self._ifft(subsample_fourier(U2_f * self.phi[k1 + k2], 2 ** k2_J)).real
                        order2.append(S2[:, i0[k1 + k2 + k2_J]:i1[k1 + k2 + k2_J]])   # :364
        out.extend(order2)                                                     # :372-375
        S = np.stack(out, axis=1)                                              # :378 (dim 2 of (B,1,C,T))
        return S.reshape(batch_shape + S.shape[-2:]).astype(self.rdtype)       # torch_frontend.py:231-235

    def unaveraged(self, x):
        """average=False, out_type='list' (core :293-294, :329-330, :366-367): [(key, coef (..., len))] in the
        reference's order -- the input itself, then the unpadded moduli U1 and U2 at their own rates."""
        x = np.asarray(x)
        batch_shape = x.shape[:-1]
        x = x.reshape(-1, x.shape[-1]).astype(self.rdtype)
        g = self.geo
        log2_T = math.floor(math.log2(self.T))
        os_ = self.oversampling
        i0, i1 = g['ind_start'], g['ind_end']
        out = [((), x)]                                                         # :294
        order2 = []
        U0_f = self._fft(reflect_pad(x, g['pad_left'], g['pad_right']))
        for n1, p1 in enumerate(self.psi1):
            j1 = p1['j']
            k1 = max(min(j1 - os_, log2_T - os_), 0)
            U1 = np.abs(self._ifft(subsample_fourier(U0_f * p1['levels'][0], 2 ** k1)))
            out.append(((n1,), U1[:, i0[k1]:i1[k1]]))                          # :330
            if self.max_order == 2:
                U1_f = self._fft(U1)
                for n2, p2 in enumerate(self.psi2):
                    if p2['j'] > j1:
                        k2 = max(min(p2['j'] - k1 - os_, log2_T - k1 - os_), 0)
                        U2 = np.abs(self._ifft(subsample_fourier(U1_f * p2['levels'][k1], 2 ** k2)))
                        order2.append(((n1, n2), U2[:, i0[k1 + k2]:i1[k1 + k2]]))   # :367
        out.extend(order2)
        return [(k, v.reshape(batch_shape + v.shape[-1:]).astype(self.rdtype)) for k, v in out]
