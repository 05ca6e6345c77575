"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement, in plain numpy float64, of the reference's filter-bank and
geometry code for the scattering hot path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this package.

Reference (paths relative to /root/reference):
  kymatio/kymatio/scattering1d/filter_bank.py   (cited per function below)
  kymatio/kymatio/scattering1d/utils.py         (cited per function below)
  kymatio/kymatio/scattering1d/frontend/base_frontend.py:27-77 (geometry)

Pinned: ``tests/test_oracle_golden.py`` checks every array produced here
against fixtures generated from the live reference by ``oracle/make_golden.py``
(committed under ``tests/golden/``) and against the reference's own
known-answer fixture ``test_data_1d.npz``.
"""
import math

import numpy as np
import scipy.fft


# --- filter_bank.py:6-48 ------------------------------------------------------
def periods_needed(sigma, eps=1e-7):
    return int(math.ceil(math.sqrt(-2 * (sigma ** 2) * math.log(eps)) + 1))


# --- filter_bank.py:51-71 -----------------------------------------------------
def periodize(h_f, nperiods=1):
    n = h_f.shape[0] // nperiods
    return h_f.reshape(nperiods, n).mean(axis=0)


# --- filter_bank.py:139-165 ---------------------------------------------------
def l1_factor(h_f):
    l1 = np.abs(scipy.fft.ifft(h_f)).sum()
    if l1 < 1e-7:                                                       # :156-158
        raise ValueError('Zero division error is very likely to occur, aborting computations now.')
    return 1.0 / l1


# --- filter_bank.py:74-136 ----------------------------------------------------
def morlet(n, xi, sigma, P_max=5, eps=1e-7):
    P = min(periods_needed(sigma, eps), P_max)
    freqs = np.arange((1 - P) * n, P * n, dtype=float) / float(n)
    freqs_low = np.fft.fftfreq(n) if P == 1 else freqs
    gabor = periodize(np.exp(-(freqs - xi) ** 2 / (2 * sigma ** 2)), 2 * P - 1)
    low = periodize(np.exp(-(freqs_low ** 2) / (2 * sigma ** 2)), 2 * P - 1)
    kappa = gabor[0] / low[0]
    out = gabor - kappa * low
    out *= l1_factor(out)
    return out


# --- filter_bank.py:168-216 ---------------------------------------------------
def gauss(n, sigma, P_max=5, eps=1e-7):
    P = min(periods_needed(sigma, eps), P_max)
    if P == 1:
        freqs = np.fft.fftfreq(n)
    else:
        freqs = np.arange((1 - P) * n, P * n, dtype=float) / float(n)
    g = periodize(np.exp(-freqs ** 2 / (2 * sigma ** 2)), 2 * P - 1)
    g *= l1_factor(g)
    return g


# --- filter_bank.py:219-253, 313-347, 394-409 ---------------------------------
def sigma_psi(xi, Q, r=math.sqrt(0.5)):
    factor = 1. / math.pow(2, 1. / Q)
    return xi * ((1 - factor) / (1 + factor)) * (1. / math.sqrt(2 * math.log(1. / r)))


def max_dyadic_subsampling(xi, sigma, alpha=5.):
    return int(math.floor(-math.log2(min(xi + alpha * sigma, 0.5))) - 1)


def xi_max(Q):
    return max(1. / (1. + math.pow(2., 3. / Q)), 0.35)


# --- filter_bank.py:412-487 (incl. move_one_dyadic_step :350-391) --------------
def params_filterbank(sigma_min, Q, r_psi=math.sqrt(0.5), alpha=5.):
    xi0 = xi_max(Q)
    sg0 = sigma_psi(xi0, Q, r=r_psi)
    xi, sigma, j = [], [], []
    if sg0 <= sigma_min:
        last_xi = sg0
    else:
        cur_xi, cur_sg, cur_j = xi0, sg0, 0          # first j is hard-coded 0 (:469)
        while cur_sg > sigma_min:
            xi.append(cur_xi); sigma.append(cur_sg); j.append(cur_j)
            factor = 1. / math.pow(2., 1. / Q)
            cur_xi, cur_sg = cur_xi * factor, cur_sg * factor
            cur_j = max_dyadic_subsampling(cur_xi, cur_sg, alpha=alpha)
        last_xi = xi[-1]
    num_intermediate = Q - 1
    for q in range(1, num_intermediate + 1):
        factor = (num_intermediate + 1. - q) / (num_intermediate + 1.)
        xi.append(factor * last_xi)
        sigma.append(sigma_min)
        j.append(max_dyadic_subsampling(factor * last_xi, sigma_min, alpha=alpha))
    return xi, sigma, j


# --- filter_bank.py:490-558 ---------------------------------------------------
def calibrate(J, Q, T, r_psi=math.sqrt(0.5), sigma0=0.1, alpha=5.):
    sigma_min = sigma0 / math.pow(2, J)
    xi1, s1, j1 = params_filterbank(sigma_min, Q, r_psi=r_psi, alpha=alpha)
    xi2, s2, j2 = params_filterbank(sigma_min, 1, r_psi=r_psi, alpha=alpha)
    return sigma0 / T, xi1, s1, j1, xi2, s2, j2


# --- filter_bank.py:256-310 ---------------------------------------------------
def temporal_support(h_f, criterion_amplitude=1e-3):
    h = np.abs(scipy.fft.ifft(h_f))
    half = h.shape[0] // 2
    resid = np.cumsum(h[:half][::-1])[::-1]
    hits = np.where(resid <= criterion_amplitude)[0]
    return int(hits.min()) + 1 if hits.size else half


# --- filter_bank.py:561-762 ---------------------------------------------------
def filter_factory(J_support, J, Q, T, max_subsampling=None):
    """Returns dict(phi=[levels], psi1=[{xi,sigma,j,levels}], psi2=[...], t_max_phi)."""
    sigma_low, xi1s, s1s, j1s, xi2s, s2s, j2s = calibrate(J, Q, T)
    n = 2 ** J_support
    psi2 = []
    for xi2, s2, j2 in zip(xi2s, s2s, j2s):
        if max_subsampling is None:
            cands = [j1 for j1 in j1s if j2 > j1]
            top = max(cands) if cands else 0
        else:
            top = max_subsampling
        lv = [morlet(n, xi2, s2)]
        for level in range(1, top + 1):
            lv.append(periodize(lv[0], 2 ** level))
        psi2.append(dict(xi=xi2, sigma=s2, j=j2, levels=lv))
    psi1 = [dict(xi=a, sigma=b, j=c, levels=[morlet(n, a, b)]) for a, b, c in zip(xi1s, s1s, j1s)]
    top = max(max(j1s), max(j2s)) if max_subsampling is None else max_subsampling
    phi = [gauss(n, sigma_low)]
    for level in range(1, top + 1):
        phi.append(periodize(phi[0], 2 ** level))
    return dict(phi=phi, psi1=psi1, psi2=psi2, t_max_phi=temporal_support(phi[0]),
                sigma_low=sigma_low)


# --- base_frontend.py:62-77, utils.py:5-65,127-133 -----------------------------
def geometry(N, J, Q, T, clamp=False):
    """J_pad, pad_left, pad_right, ind_start, ind_end.  ``clamp`` adds the
    min(min_to_pad, N-1) of hdf5_dataset/kymatio_phase_scattering.py:104."""
    t_max = filter_factory(int(np.ceil(np.log2(N))), J, Q, T, max_subsampling=0)['t_max_phi']
    min_to_pad = 3 * t_max
    if clamp:
        min_to_pad = min(min_to_pad, N - 1)
    J_pad = min(int(np.ceil(np.log2(N + 2 * min_to_pad))), int(np.floor(np.log2(3 * N - 2))))
    to_add = 2 ** J_pad - N
    pad_left = to_add // 2
    pad_right = to_add - pad_left
    assert max(pad_left, pad_right) < N
    i0, i1 = {0: pad_left}, {0: pad_left + N}
    for j in range(1, J + 1):
        i0[j] = (i0[j - 1] // 2) + (i0[j - 1] % 2)
        i1[j] = (i1[j - 1] // 2) + (i1[j - 1] % 2)
    return dict(J_pad=J_pad, pad_left=pad_left, pad_right=pad_right, ind_start=i0, ind_end=i1)


# --- utils.py:190-289 (key ordering only) and :136-187 --------------------------
def path_keys(J, Q, T, max_order=2):
    _, xi1s, _, j1s, xi2s, _, j2s = calibrate(J, Q, T)
    keys = [()]
    keys += [(n1,) for n1 in range(len(xi1s))]
    if max_order >= 2:
        for n1, j1 in enumerate(j1s):
            for n2, j2 in enumerate(j2s):
                if j2 > j1:
                    keys.append((n1, n2))
    return keys
