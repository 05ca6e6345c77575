"""Morlet / Gaussian filter bank for the 1-D scattering cascade (host side, float64).

Replaces ``kymatio/scattering1d/filter_bank.py`` of the reference (calibration
:412-558, synthesis :74-216, factory :561-762) and the geometry helpers of
``kymatio/scattering1d/utils.py`` (:5-65, :67-133).  Everything here is a
one-off precompute in numpy float64; the result is cast to fp32 and uploaded
once into the device plan (``plan.py``).  The arithmetic order of the
calibration recurrences is kept the same as the reference so that ``xi``,
``sigma`` and therefore the fp32 ``center_freqs`` / ``powers`` of the phase
module come out as the very same floating-point values.
"""
from __future__ import annotations

import math
import warnings
from dataclasses import dataclass, field
from typing import List

import numpy as np
from scipy.fft import ifft as _ifft

# constants fixed by ScatteringBase1D.build (base_frontend.py:36-42)
R_PSI = math.sqrt(0.5)
SIGMA0 = 0.1
ALPHA = 5.0
P_MAX = 5
EPS = 1e-7
CRITERION_AMPLITUDE = 1e-3


# ----------------------------------------------------------------------------
# calibration of (xi, sigma, j)
# ----------------------------------------------------------------------------
def _dyadic_j(xi: float, sigma: float, alpha: float = ALPHA) -> int:
    """Largest j with xi + alpha*sigma < 2^-(j+1)   (filter_bank.py:344-346)."""
    return int(math.floor(-math.log2(min(xi + alpha * sigma, 0.5))) - 1)


def _sigma_of(xi: float, Q: int, r: float = R_PSI) -> float:
    """Bandwidth of a Morlet of centre xi in a Q-per-octave family (:250-253)."""
    step = 1.0 / math.pow(2, 1.0 / Q)
    return xi * ((1 - step) / (1 + step)) * (1.0 / math.sqrt(2 * math.log(1.0 / r)))


def wavelet_family(sigma_min: float, Q: int, r_psi: float = R_PSI, alpha: float = ALPHA):
    """(xi, sigma, j) lists of one band-pass family (filter_bank.py:412-487).

    Geometric part: xi_n = xi_max 2^{-n/Q} while sigma_n > sigma_min (the first
    filter always has j = 0, :469); then Q-1 filters of width sigma_min with
    linearly spaced centres below the last geometric one.
    """
    xi_max = max(1.0 / (1.0 + math.pow(2.0, 3.0 / Q)), 0.35)
    sigma_max = _sigma_of(xi_max, Q, r_psi)
    xis: List[float] = []
    sigmas: List[float] = []
    js: List[int] = []
    if sigma_max <= sigma_min:
        last_xi = sigma_max
    else:
        step = 1.0 / math.pow(2.0, 1.0 / Q)
        xi, sigma, j = xi_max, sigma_max, 0
        while sigma > sigma_min:
            xis.append(xi)
            sigmas.append(sigma)
            js.append(j)
            xi, sigma = xi * step, sigma * step
            j = _dyadic_j(xi, sigma, alpha)
        last_xi = xis[-1]
    for q in range(1, Q):
        xi = ((Q - 1 + 1.0 - q) / (Q - 1 + 1.0)) * last_xi
        xis.append(xi)
        sigmas.append(sigma_min)
        js.append(_dyadic_j(xi, sigma_min, alpha))
    return xis, sigmas, js


@dataclass
class Calibration:
    sigma_low: float
    xi1: List[float]
    sigma1: List[float]
    j1: List[int]
    xi2: List[float]
    sigma2: List[float]
    j2: List[int]


def _as_Q1(Q):
    """The fork takes an int Q (second order is hard-wired to Q=1,
    filter_bank.py:553); upstream-style ``(Q1, 1)`` tuples are accepted too."""
    if isinstance(Q, (tuple, list)):
        if len(Q) == 0 or len(Q) > 2 or (len(Q) == 2 and int(Q[1]) != 1):
            raise ValueError('Q must be an int or a tuple (Q1, 1); got {}'.format(Q))
        Q = Q[0]
    return int(Q)


def calibrate(J: int, Q, T: int, r_psi: float = R_PSI, sigma0: float = SIGMA0,
              alpha: float = ALPHA) -> Calibration:
    """filter_bank.py:490-558."""
    Q = _as_Q1(Q)
    if Q < 1:
        raise ValueError('Q should always be >= 1, got {}'.format(Q))
    sigma_min = sigma0 / math.pow(2, J)
    xi1, s1, j1 = wavelet_family(sigma_min, Q, r_psi, alpha)
    xi2, s2, j2 = wavelet_family(sigma_min, 1, r_psi, alpha)
    return Calibration(sigma0 / T, xi1, s1, j1, xi2, s2, j2)


# ----------------------------------------------------------------------------
# filter synthesis in the Fourier domain
# ----------------------------------------------------------------------------
def fold(h_f: np.ndarray, k: int) -> np.ndarray:
    """Fourier-domain periodisation by the *mean* of k blocks (:69-71)."""
    return h_f.reshape(k, h_f.shape[0] // k).mean(axis=0)


def _n_periods(sigma: float, P_max: int, eps: float) -> int:
    if type(P_max) != int:
        raise ValueError('P_max should be an int, got {}'.format(type(P_max)))
    if P_max < 1:
        raise ValueError('P_max should be non-negative, got {}'.format(P_max))
    return min(int(math.ceil(math.sqrt(-2 * (sigma ** 2) * math.log(eps)) + 1)), P_max)


def _unit_norm(h_f: np.ndarray, normalize: str) -> float:
    h = _ifft(h_f)
    l1 = np.abs(h).sum()
    if l1 < 1e-7:
        raise ValueError('Zero division error is very likely to occur, '
                         'aborting computations now.')
    if normalize == 'l1':
        return 1.0 / l1
    if normalize == 'l2':
        return 1.0 / np.sqrt((np.abs(h) ** 2).sum())
    raise ValueError("Supported normalizations only include 'l1' and 'l2'")


def gauss_spectrum(n: int, sigma: float, normalize: str = 'l1', P_max: int = P_MAX,
                   eps: float = EPS) -> np.ndarray:
    """Low-pass exp(-w^2/2sigma^2) sampled on fftfreq(n) (:168-216)."""
    P = _n_periods(sigma, P_max, eps)
    if P == 1:
        w = np.fft.fftfreq(n)
    else:
        w = np.arange((1 - P) * n, P * n, dtype=float) / float(n)
    g = fold(np.exp(-w ** 2 / (2 * sigma ** 2)), 2 * P - 1)
    g *= _unit_norm(g, normalize)
    return g


def morlet_spectrum(n: int, xi: float, sigma: float, normalize: str = 'l1',
                    P_max: int = P_MAX, eps: float = EPS) -> np.ndarray:
    """Gabor minus the multiple of the low-pass that zeroes bin 0 (:74-136)."""
    P = _n_periods(sigma, P_max, eps)
    w = np.arange((1 - P) * n, P * n, dtype=float) / float(n)
    w_low = np.fft.fftfreq(n) if P == 1 else w
    gabor = fold(np.exp(-(w - xi) ** 2 / (2 * sigma ** 2)), 2 * P - 1)
    low = fold(np.exp(-(w_low ** 2) / (2 * sigma ** 2)), 2 * P - 1)
    psi = gabor - (gabor[0] / low[0]) * low
    psi *= _unit_norm(psi, normalize)
    return psi


def half_support(h_f: np.ndarray, criterion_amplitude: float = CRITERION_AMPLITUDE) -> int:
    """Smallest half-support whose l1 tail is below the criterion (:256-310)."""
    h = np.abs(_ifft(h_f))
    half = h.shape[0] // 2
    tail = np.cumsum(h[:half][::-1])[::-1]
    ok = np.where(tail <= criterion_amplitude)[0]
    if ok.size:
        return int(ok.min()) + 1
    warnings.warn('Signal support is too small to avoid border effects')
    return half


@dataclass
class BandFilter:
    xi: float
    sigma: float
    j: int
    levels: List[np.ndarray] = field(default_factory=list)   # levels[l] has length Np / 2^l


@dataclass
class FilterBank:
    log2_Np: int
    phi: BandFilter
    psi1: List[BandFilter]
    psi2: List[BandFilter]
    t_max_phi: int
    calib: Calibration


def build_filter_bank(J_support: int, J: int, Q, T: int, normalize: str = 'l1',
                      criterion_amplitude: float = CRITERION_AMPLITUDE,
                      max_subsampling=None, r_psi: float = R_PSI, sigma0: float = SIGMA0,
                      alpha: float = ALPHA, P_max: int = P_MAX, eps: float = EPS) -> FilterBank:
    """All filters of the cascade (filter_bank.py:561-762).

    psi2[n2] exists at levels 0..max{j1 : j1 < j2}; psi1 only at level 0; phi at
    levels 0..max(j1, j2).  phi always uses the default P_max/eps (:749).
    """
    cal = calibrate(J, Q, T, r_psi, sigma0, alpha)
    n = 2 ** J_support

    psi2 = []
    for xi, sg, j2 in zip(cal.xi2, cal.sigma2, cal.j2):
        if max_subsampling is None:
            below = [j1 for j1 in cal.j1 if j2 > j1]
            top = max(below) if below else 0
        else:
            top = max_subsampling
        base = morlet_spectrum(n, xi, sg, normalize, P_max, eps)
        psi2.append(BandFilter(xi, sg, j2, [base] + [fold(base, 2 ** l) for l in range(1, top + 1)]))

    psi1 = [BandFilter(xi, sg, j1, [morlet_spectrum(n, xi, sg, normalize, P_max, eps)])
            for xi, sg, j1 in zip(cal.xi1, cal.sigma1, cal.j1)]

    top = max(max(cal.j1), max(cal.j2)) if max_subsampling is None else max_subsampling
    base = gauss_spectrum(n, cal.sigma_low)
    phi = BandFilter(0, cal.sigma_low, 0, [base] + [fold(base, 2 ** l) for l in range(1, top + 1)])

    return FilterBank(J_support, phi, psi1, psi2, half_support(base, criterion_amplitude), cal)


# ----------------------------------------------------------------------------
# geometry (utils.py:5-65,127-133; base_frontend.py:62-77)
# ----------------------------------------------------------------------------
def border_indices(J: int, i0: int, i1: int):
    start, end = {0: i0}, {0: i1}
    for j in range(1, J + 1):
        start[j] = (start[j - 1] // 2) + (start[j - 1] % 2)
        end[j] = (end[j - 1] // 2) + (end[j - 1] % 2)
    return start, end


def padding(J_pad: int, N: int):
    n_pad = 2 ** J_pad
    if n_pad < N:
        raise ValueError('Padding support should be larger than the original' +
                         'signal size!')
    extra = n_pad - N
    left = extra // 2
    right = extra - left
    if max(left, right) >= N:
        raise ValueError('Too large padding value, will lead to NaN errors')
    return left, right


def minimum_support_to_pad(N: int, J: int, Q, T: int, **kw) -> int:
    """3 x the phi half-support measured on 2^ceil(log2 N) points (utils.py:127-133)."""
    bank = build_filter_bank(int(np.ceil(np.log2(N))), J, Q, T, max_subsampling=0, **kw)
    return 3 * bank.t_max_phi


@dataclass
class Geometry:
    N: int
    J_pad: int
    pad_left: int
    pad_right: int
    ind_start: dict
    ind_end: dict


def build_geometry(N: int, J: int, Q, T: int, clamp_to_signal: bool = False) -> Geometry:
    """base_frontend.py:62-77.  ``clamp_to_signal`` reproduces the extra
    ``min(min_to_pad, N - 1)`` of the phase module (kymatio_phase_scattering.py:104)."""
    min_to_pad = minimum_support_to_pad(N, J, Q, T)
    if clamp_to_signal:
        min_to_pad = min(min_to_pad, N - 1)
    J_pad = min(int(np.ceil(np.log2(N + 2 * min_to_pad))), int(np.floor(np.log2(3 * N - 2))))
    left, right = padding(J_pad, N)
    start, end = border_indices(J, left, left + N)
    return Geometry(N, J_pad, left, right, start, end)
