"""ORACLE (test infrastructure only -- never imported by the product path).

numpy float64 restatement of the phase-harmonic correlation of
``hdf5_dataset/kymatio_phase_scattering.py`` (reference paths relative to
/root/reference):

  :100-113  geometry (with the min(min_to_pad, N-1) clamp)
  :115-160  psi1 / phi complex64 casts of the float64 level-0 filters, pair list,
            fp32 powers
  :211-218  _accelerate_phase       |z| exp(i p atan2(Im z, Re z))
  :220-231  _apply_filters          ifft(fft(reflect_pad(x)) psi1)[pad_left : pad_left+N]
  :233-273  _apply_phi_filter       fft(reflect_pad(c)) phi -> bins [0, Np/dec) -> ifft -> slice
  :275-360  within / cross channel pair products

PARITY UNPINNED by the reference's own tests (it has none for this module);
the oracle is pinned only against outputs of the live reference
(oracle/make_golden.py -> tests/golden/phase_*.npz).

Branch alignment: at t = 0 and t = N-1 the reflect padding
makes every analytic signal locally symmetric, so Im z there is rounding noise;
where Re z < 0 the reference's theta is +-pi at random and exp(i p theta) jumps.
``PhaseOracle.__call__`` therefore also returns, per pair, the two complex
impulse responses with which a test can flip the branch at those two samples
(``align_branches``).
"""
import numpy as np
import scipy.fft

from . import filters_oracle as fo
from .scattering1d_oracle import reflect_pad


def pad_signal(x, pad_left, pad_right, border_mode='reflect'):
    """:162-173 -- _pad_signal: 'reflect' (single fold, pad < N), 'constant' (zeros) or 'circular'."""
    if border_mode == 'reflect':
        return reflect_pad(x, pad_left, pad_right)
    widths = [(0, 0)] * (x.ndim - 1) + [(pad_left, pad_right)]
    if border_mode == 'constant':
        return np.pad(x, widths, mode='constant')
    if border_mode == 'circular':
        return np.pad(x, widths, mode='wrap')
    raise ValueError(f"Unsupported border_mode: {border_mode}")


class PhaseOracle:
    def __init__(self, J, Q, T, N, n_out, border_mode='reflect', cdtype=np.complex128):
        """n_out: temporal length of the scattering output (target_length, :445).
        cdtype=np.complex64 evaluates every step in single precision like the reference does (fp32 transforms,
        fp32 atan2, p * theta rounded in fp32): the distance of that output to the float64 one is the
        reference-class fp32 noise of a path, which the tests use as the noise floor where no output of the
        live reference is committed."""
        self.J, self.Q, self.T, self.N, self.n_out = J, Q, T, N, n_out
        self.border_mode = border_mode
        self.cdtype = np.dtype(cdtype)
        self.rdtype = np.float32 if self.cdtype == np.complex64 else np.float64
        self.geo = fo.geometry(N, J, Q, T, clamp=True)                      # :100-113
        bank = fo.filter_factory(self.geo['J_pad'], J, Q, T)                 # :117-120
        # complex64 cast keeps the fp32 value of the real float64 filters  (:123-125)
        self.psi1 = np.stack([p['levels'][0] for p in bank['psi1']]).astype(np.float32).astype(np.float64)
        self.phi = bank['phi'][0].astype(np.float32).astype(np.float64)
        self.center_freqs = np.array([p['xi'] for p in bank['psi1']], dtype=np.float32)   # :128
        F = len(self.center_freqs)
        pairs = [(i, j) for i in range(F) for j in range(F)
                 if self.center_freqs[j] >= self.center_freqs[i]]            # :141-146
        self.i_idx = np.array([p[0] for p in pairs], dtype=np.int64)
        self.j_idx = np.array([p[1] for p in pairs], dtype=np.int64)
        xi_i = self.center_freqs[self.i_idx]
        xi_j = self.center_freqs[self.j_idx]
        self.powers = np.where(xi_i > np.float32(1e-8), xi_j / xi_i, np.float32(1.0)).astype(np.float32)  # :148-152
        self.autoc_idx = np.array([k for k, (i, j) in enumerate(pairs) if i == j], dtype=np.int64)

    # :220-231
    def analytic(self, x):
        """x: (B, N) -> z (B, F, N) complex128."""
        g = self.geo
        xf = scipy.fft.fft(pad_signal(np.asarray(x, self.rdtype), g['pad_left'], g['pad_right'], self.border_mode)
                           .astype(self.cdtype), axis=-1)
        z = scipy.fft.ifft(xf[:, None, :] * self.psi1[None].astype(self.rdtype), axis=-1)
        return z[..., g['ind_start'][0]:g['ind_end'][0]]

    # :233-273 (decimation branch)
    def _smooth(self, c):
        g = self.geo
        Np = 2 ** g['J_pad']
        dec = self.N // self.n_out if (self.n_out > 0 and self.N > self.n_out) else 1   # :287-291
        cf = scipy.fft.fft(pad_signal(c, g['pad_left'], g['pad_right'], self.border_mode), axis=-1) * self.phi.astype(self.rdtype)
        if dec > 1:
            y = scipy.fft.ifft(cf[..., :max(Np // dec, 1)], axis=-1)          # :242-252
            s = g['pad_left'] // dec                                          # :258
            e = min(s + self.N // dec, y.shape[-1])                           # :259-266
            return y[..., s:e]
        return scipy.fft.ifft(cf, axis=-1)[..., g['ind_start'][0]:g['ind_end'][0]]

    def pair_stage(self, z_i, z_j, pair_subset=None, low_pass=True):
        """z_i, z_j: (B, F, N) analytic signals of the 'i' and 'j' channels."""
        ii, jj, pw = self.i_idx, self.j_idx, self.powers
        if pair_subset is not None:
            ii, jj, pw = ii[pair_subset], jj[pair_subset], pw[pair_subset]
        zi = z_i[:, ii, :]
        theta = np.arctan2(zi.imag, zi.real) * pw[None, :, None].astype(self.rdtype)  # :214-215
        acc = (np.abs(zi) * (np.cos(theta) + 1j * np.sin(theta))).astype(self.cdtype)   # :218
        c = acc * np.conj(z_j[:, jj, :])                                     # :283 / :339
        if not low_pass:
            return c.real                                                    # :357-360
        return self._smooth(c)                                               # complex; caller takes .real

    def __call__(self, x, mode='cross', pair_subset=None, low_pass=True):
        """mode 'within': x (B, N); mode 'cross': x (B, 2, N) -> (B, P, n_out) float64."""
        x = np.asarray(x, self.rdtype)
        if mode == 'within':
            z = self.analytic(x)
            y = self.pair_stage(z, z, pair_subset, low_pass)
        else:
            y = self.pair_stage(self.analytic(x[:, 0]), self.analytic(x[:, 1]), pair_subset, low_pass)
        return np.real(y)

    # ---- branch alignment helper (SURVEY 8c protocol) ----------------------------
    def align_branches(self, x, test_out, mode='cross', pair_subset=None, rel_im=1e-5, interior_rel_im=2e-6):
        """Return the oracle output where, at t=0 and t=N-1 of every 'i' filter whose
        analytic sample is on the negative real axis up to noise (|Im| < rel_im |Re|,
        Re < 0), the sign of theta is chosen (per pair, 4 combinations) to best match
        ``test_out``.  Everything else is the plain float64 oracle.

        interior_rel_im: the same choice at the rare INTERIOR samples that sit on the negative real axis by chance,
        within the rounding of a float32 transform (|Im| < 2e-6 |Re| is a few ulps of |z|: about one sample in a
        million).  The random sweep met one (tools/random_parity_sweep.py, seed 32: Im z / |z| = 1.9e-7, which the
        device FFT rounds to the other side of zero); without decimation the flip is not diluted and costs 5e-4 on
        the paths of that filter.  0 disables it."""
        x = np.asarray(x, np.float64)
        if mode == 'within':
            z_i = z_j = self.analytic(x)
        else:
            z_i, z_j = self.analytic(x[:, 0]), self.analytic(x[:, 1])
        ii, jj, pw = self.i_idx, self.j_idx, self.powers
        if pair_subset is not None:
            ii, jj, pw = ii[pair_subset], jj[pair_subset], pw[pair_subset]
        base = self.pair_stage(z_i, z_j, pair_subset)                        # complex (B,P,n)
        best = base.copy()
        N = self.N
        # impulse responses of the smoothing operator at the two boundary samples
        imp = np.zeros((2, N), np.complex128)
        imp[0, 0] = 1.0
        imp[1, N - 1] = 1.0
        g = self._smooth(imp)                                                # (2, n)
        p = pw.astype(np.float64)[None, :]

        def flip_delta(t):
            """Change of the pair product at sample t if theta takes the other sign there: (B, P) complex."""
            zi = z_i[:, ii, t]
            th = np.arctan2(zi.imag, zi.real)
            return np.abs(zi) * (np.exp(-1j * p * th) - np.exp(1j * p * th)) * np.conj(z_j[:, jj, t])

        deltas = []
        for t in (0, N - 1):
            zi = z_i[:, ii, t]
            amb = (zi.real < 0) & (np.abs(zi.imag) < rel_im * np.abs(zi.real))
            deltas.append(np.where(amb, flip_delta(t), 0.0))                 # (B,P)
        err = np.linalg.norm(best.real - test_out, axis=-1)
        for f0 in (0, 1):
            for f1 in (0, 1):
                if f0 == 0 and f1 == 0:
                    continue
                cand = base + f0 * deltas[0][..., None] * g[0] + f1 * deltas[1][..., None] * g[1]
                e = np.linalg.norm(cand.real - test_out, axis=-1)
                take = e < err
                best[take] = cand[take]
                err = np.where(take, e, err)
        if interior_rel_im > 0 and N > 2:
            used = np.unique(ii)
            zu = z_i[:, used, 1:N - 1]
            for b, f, t in np.argwhere((zu.real < 0) & (np.abs(zu.imag) < interior_rel_im * np.abs(zu.real))):
                f, t = used[f], t + 1
                imp_t = np.zeros((1, N), np.complex128)
                imp_t[0, t] = 1.0
                rows = np.nonzero(ii == f)[0]
                cand = best[b, rows] + flip_delta(t)[b, rows, None] * self._smooth(imp_t)[0]
                e = np.linalg.norm(cand.real - test_out[b, rows], axis=-1)
                take = e < err[b, rows]
                best[b, rows[take]] = cand[take]
                err[b, rows[take]] = e[take]
        return best.real
