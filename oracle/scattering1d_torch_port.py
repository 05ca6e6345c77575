"""ORACLE (test infrastructure only -- never imported by the product path).

torch-CPU port of the reference's Scattering1D forward, op for op, used as the CPU baseline
of bench.py (`cpu_baseline.kind = "port"`, `--impl reference`).  It issues exactly the torch
calls the reference's torch backend issues on CPU, so its speed is the reference's speed on
the same host cores:

  kymatio/kymatio/scattering1d/backend/torch_backend.py:50-78   F.pad(..., 'reflect')
  :106-114  rfft  = zero imaginary plane + full complex torch.fft.fft
  :18-48    subsample_fourier = view(k, L/k).mean
  :116-128  irfft = torch.fft.ifft(...).real ; ifft
  kymatio/kymatio/backend/torch_backend.py:137-141,206  modulus, cdgmm (A * B, B real (L,1))
  kymatio/kymatio/scattering1d/core/scattering1d.py:269-378    the cascade and its ordering

Pinned by tests/test_oracle_golden.py against the live reference's outputs.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import filters_oracle as fo


class TorchPort:
    def __init__(self, J, N, Q, T, max_order=2):
        self.J, self.N, self.Q, self.T, self.max_order = J, N, Q, T, max_order
        self.geo = fo.geometry(N, J, Q, T)
        bank = fo.filter_factory(self.geo['J_pad'], J, Q, T)
        t = lambda a: torch.from_numpy(a).float().view(-1, 1)              # register_filters :82-96
        self.phi = [t(a) for a in bank['phi']]
        self.psi1 = [dict(j=p['j'], levels=[t(a) for a in p['levels']]) for p in bank['psi1']]
        self.psi2 = [dict(j=p['j'], levels=[t(a) for a in p['levels']]) for p in bank['psi2']]

    @staticmethod
    def _fft(x):
        return torch.view_as_real(torch.fft.fft(torch.view_as_complex(x)))

    @staticmethod
    def _ifft(x):
        return torch.view_as_real(torch.fft.ifft(torch.view_as_complex(x)))

    @classmethod
    def _rfft(cls, x):
        x_r = torch.zeros(x.shape[:-1] + (2,), dtype=x.dtype)
        x_r[..., 0] = x[..., 0]
        return cls._fft(x_r)

    @staticmethod
    def _irfft(x):
        return torch.fft.ifft(torch.view_as_complex(x)).real[..., None]

    @staticmethod
    def _sub(x, k):
        n = x.shape[-2]
        return x.view(x.shape[:-2] + (k, n // k, 2)).mean(dim=-3)

    @staticmethod
    def _mod(x):
        return torch.linalg.vector_norm(x, dim=-1, keepdim=True)

    @torch.no_grad()
    def __call__(self, x):
        g = self.geo
        i0, i1 = g['ind_start'], g['ind_end']
        batch_shape = x.shape[:-1]
        x = x.reshape((-1, 1) + x.shape[-1:])
        log2_T = math.floor(math.log2(self.T))
        U0 = F.pad(x, (g['pad_left'], g['pad_right']), mode='reflect')[..., None]
        U0_hat = self._rfft(U0)
        out0 = [self._irfft(self._sub(U0_hat * self.phi[0], 2 ** log2_T)).reshape(x.shape[0], 1, -1)[..., i0[log2_T]:i1[log2_T]]]
        out1, out2 = [], []
        for p1 in self.psi1:
            j1 = p1['j']
            k1 = max(min(j1, log2_T), 0)
            U1_hat = self._rfft(self._mod(self._ifft(self._sub(U0_hat * p1['levels'][0], 2 ** k1))))
            k1_J = max(log2_T - k1, 0)
            S1 = self._irfft(self._sub(U1_hat * self.phi[k1], 2 ** k1_J))
            out1.append(S1.reshape(x.shape[0], 1, -1)[..., i0[k1_J + k1]:i1[k1_J + k1]])
            if self.max_order == 2:
                for p2 in self.psi2:
                    if p2['j'] > j1:
                        k2 = max(min(p2['j'] - k1, log2_T - k1), 0)
                        U2_hat = self._rfft(self._mod(self._ifft(self._sub(U1_hat * p2['levels'][k1], 2 ** k2))))
                        k2_J = max(log2_T - k2 - k1, 0)
                        S2 = self._irfft(self._sub(U2_hat * self.phi[k1 + k2], 2 ** k2_J))
                        out2.append(S2.reshape(x.shape[0], 1, -1)[..., i0[k1 + k2 + k2_J]:i1[k1 + k2 + k2_J]])
        S = torch.stack(out0 + out1 + out2, dim=2)
        return S.reshape(batch_shape + S.shape[-2:])
