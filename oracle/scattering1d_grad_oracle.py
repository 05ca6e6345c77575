"""TEST INFRASTRUCTURE -- gradient oracle of the scattering transform.

The reference is differentiable through torch autograd: every op of
  kymatio/kymatio/scattering1d/core/scattering1d.py:269-370
is a torch op and the modulus is ModulusStable (kymatio/kymatio/backend/torch_backend.py:5-96:
grad = x * grad_out / |x|, 0 where |x| = 0); the reference tests it in
  kymatio/tests/scattering1d/test_torch_scattering1d.py:292-315 (test_differentiability_scattering).
This file restates the cascade in float64 torch ops (torch.abs of a complex tensor has the same
sub-gradient, 0 at 0) on the numpy oracle's filters and lets autograd produce
    vjp(x, w) = d/dx sum(S(x) * w).
Pinned against gradients of the LIVE reference (oracle/make_golden_backward.py ->
tests/golden/backward_*.npz) in tests/test_oracle_golden.py.

Only tests/ import this module; nothing on the product path does.
"""
import math

import numpy as np
import torch

from .scattering1d_oracle import ScatteringOracle


class GradOracle:
    def __init__(self, J, N, Q, T, max_order=2, oversampling=0):
        self.o = ScatteringOracle(J, N, Q, T, max_order=max_order, oversampling=oversampling)

    @staticmethod
    def _periodise(u_f, k):                                  # torch_backend.py:18-48
        return u_f.reshape(u_f.shape[0], k, u_f.shape[-1] // k).mean(dim=1)

    def forward(self, x):
        """x: (B, N) float64 torch tensor -> S (B, C, n_out), differentiable."""
        o, g = self.o, self.o.geo
        t = lambda a: torch.from_numpy(np.asarray(a, np.float64))
        log2_T = math.floor(math.log2(o.T))
        os_ = o.oversampling
        i0, i1 = g['ind_start'], g['ind_end']
        xp = torch.nn.functional.pad(x[:, None, :], (g['pad_left'], g['pad_right']), mode='reflect')[:, 0]   # torch_backend.py:50-78
        U0_f = torch.fft.fft(xp.to(torch.complex128))                                                        # :278-280
        k0 = max(log2_T - os_, 0)
        out = [torch.fft.ifft(self._periodise(U0_f * t(o.phi[0]), 2 ** k0)).real[:, i0[k0]:i1[k0]]]          # :285-292
        order2 = []
        for p1 in o.psi1:
            j1 = p1['j']
            k1 = max(min(j1 - os_, log2_T - os_), 0)                                                         # :304
            U1 = torch.abs(torch.fft.ifft(self._periodise(U0_f * t(p1['levels'][0]), 2 ** k1)))              # :307-315
            U1_f = torch.fft.fft(U1.to(torch.complex128))                                                    # :318
            k1_J = max(log2_T - k1 - os_, 0)
            out.append(torch.fft.ifft(self._periodise(U1_f * t(o.phi[k1]), 2 ** k1_J)).real[:, i0[k1_J + k1]:i1[k1_J + k1]])
            if o.max_order == 2:
                for p2 in o.psi2:
                    j2 = p2['j']
                    if j2 > j1:
                        k2 = max(min(j2 - k1 - os_, log2_T - k1 - os_), 0)                                   # :344-345
                        U2 = torch.abs(torch.fft.ifft(self._periodise(U1_f * t(p2['levels'][k1]), 2 ** k2)))
                        U2_f = torch.fft.fft(U2.to(torch.complex128))
                        k2_J = max(log2_T - k2 - k1 - os_, 0)
                        S2 = torch.fft.ifft(self._periodise(U2_f * t(o.phi[k1 + k2]), 2 ** k2_J)).real
                        order2.append(S2[:, i0[k1 + k2 + k2_J]:i1[k1 + k2 + k2_J]])
        return torch.stack(out + order2, dim=1)

    def forward_unaveraged(self, x):
        """average=False (core/scattering1d.py:329-330, :366-367): the unpadded moduli U1 / U2 at their own rates,
        back to back in the reference's path order (all first-order paths, then all second-order ones) -> (B, total).
        Order 0 is the input itself and is left out."""
        o, g = self.o, self.o.geo
        t = lambda a: torch.from_numpy(np.asarray(a, np.float64))
        log2_T = math.floor(math.log2(o.T))
        os_ = o.oversampling
        i0, i1 = g['ind_start'], g['ind_end']
        xp = torch.nn.functional.pad(x[:, None, :], (g['pad_left'], g['pad_right']), mode='reflect')[:, 0]
        U0_f = torch.fft.fft(xp.to(torch.complex128))
        first, second = [], []
        for p1 in o.psi1:
            j1 = p1['j']
            k1 = max(min(j1 - os_, log2_T - os_), 0)
            U1 = torch.abs(torch.fft.ifft(self._periodise(U0_f * t(p1['levels'][0]), 2 ** k1)))
            first.append(U1[:, i0[k1]:i1[k1]])
            if o.max_order == 2:
                U1_f = torch.fft.fft(U1.to(torch.complex128))
                for p2 in o.psi2:
                    if p2['j'] > j1:
                        k2 = max(min(p2['j'] - k1 - os_, log2_T - k1 - os_), 0)
                        U2 = torch.abs(torch.fft.ifft(self._periodise(U1_f * t(p2['levels'][k1]), 2 ** k2)))
                        second.append(U2[:, i0[k1 + k2]:i1[k1 + k2]])
        return torch.cat(first + second, dim=1)

    def vjp_unaveraged(self, x, w):
        """x (B, N), w (B, total) -> (row float64, d sum(row w) / dx float64)."""
        xt = torch.from_numpy(np.asarray(x, np.float64)).clone().requires_grad_(True)
        row = self.forward_unaveraged(xt)
        (row * torch.from_numpy(np.asarray(w, np.float64))).sum().backward()
        return row.detach().numpy(), xt.grad.numpy()

    def vjp(self, x, w):
        """x (B, N), w (B, C, n_out) arrays -> (S float64, d sum(S w) / dx float64) as numpy arrays."""
        xt = torch.from_numpy(np.asarray(x, np.float64)).clone().requires_grad_(True)
        S = self.forward(xt)
        (S * torch.from_numpy(np.asarray(w, np.float64))).sum().backward()
        return S.detach().numpy(), xt.grad.numpy()
