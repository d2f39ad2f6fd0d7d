"""Chebyshev-in-k2a form of the SRTM2 convolution operator -- CHECKER for the CUDA kernel's hot path
(test infrastructure only; the product never imports oracle/).

kinetic_model.py:153-158 evaluates conv = estimate_continuous_convolution(t, c_r, exp(-k2a t)) = M e(k2a)
(oracle/forward.py:build_M).  For k2a in [lo, hi], s = (k2a - kmid)/h in [-1, 1]:

    exp(-k2a t_f) = e^{-kmid t_f} [ I_0(h t_f) + 2 sum_{d>=1} (-1)^d I_d(h t_f) T_d(s) ]      (Jacobi-Anger)

so conv = A T(s) with A = M C, C[f, d] = c_d (-1)^d e^{-kmid t_f} I_d(h t_f) depending on the frame grid only.
The kernel keeps ncols[b] columns for row block b (frames 18 b .. 18 b + 17)."""
import numpy as np
from scipy.special import iv

from . import forward

NCOLS = (6, 8, 12)            # petmh_device.cuh NCH0..2
KT_LO, KT_HI = 0.0, 6.0     # petmh_device.cuh PETMH_CHEB_KT_LO / _HI: range of k2a * t_last


def k2a_range(t):
    return KT_LO / np.max(t), KT_HI / np.max(t)


def cheb_table(t, lo, hi, ncol):
    """C (T, ncol)."""
    kmid, h = 0.5 * (lo + hi), 0.5 * (hi - lo)
    d = np.arange(ncol)
    c = np.where(d == 0, 1.0, 2.0) * (-1.0) ** d
    t = np.asarray(t, np.float64)
    return np.exp(-kmid * t)[:, None] * c[None, :] * iv(d[None, :], h * t[:, None])


def cheb_operator(t, c_r, lo=None, hi=None, ncols=NCOLS):
    """[A_b (18, ncols[b]) for b in 0..2] with conv[18b:18b+18] ~= A_b @ T_{0..}(s)."""
    if lo is None:
        lo, hi = k2a_range(t)
    M = forward.build_M(t, c_r)
    return [(M @ cheb_table(t, lo, hi, n))[18 * b:18 * b + 18] for b, n in enumerate(ncols)]


def cheb_T(s, n, dtype=np.float64):
    s = np.asarray(s, dtype)
    T = np.empty((n,) + s.shape, dtype)
    T[0] = 1
    if n > 1:
        T[1] = s
    for d in range(2, n):
        T[d] = dtype(2) * s * T[d - 1] - T[d - 2]
    return T


def conv_cheb(t, c_r, k2a, lo=None, hi=None, ncols=NCOLS):
    """(T, len(k2a)) convolution through the Chebyshev operator, fp64."""
    if lo is None:
        lo, hi = k2a_range(t)
    k2a = np.atleast_1d(np.asarray(k2a, np.float64))
    s = (2 * k2a - lo - hi) / (hi - lo)
    A = cheb_operator(t, c_r, lo, hi, ncols)
    return np.concatenate([A[b] @ cheb_T(s, ncols[b]) for b in range(3)], axis=0)
