"""SRTM2 forward model, restated from kinetic_model.py (reference file:line cited
per function).  fp64 numpy.  Two forms:

  * ``srtm2_tac``  -- reference-faithful: resample / convolve / interpolate on every
    call, the same work kinetic_model.SRTM2.create_activity_curve does
    (kinetic_model.py:142-158 -> :12-32 -> :35-57).
  * ``build_M`` + ``srtm2_tac_M`` -- the exact linear-operator form the CUDA kernels
    use: conv = M @ exp(-k2a t), M = dx * W_back * Toeplitz_lower(W_fwd c_r) * W_fwd.
"""
import numpy as np


def interp_weights(x, xp):
    """Dense linear-interpolation weight matrix W (len(x), len(xp)) with
    W @ fp == interpolation of (xp, fp) at x.  Follows interp1d_linear_vec
    (kinetic_model.py:41-49): searchsorted(left) index i, weight |xp[i-1]-x| on node i
    and |xp[i]-x| on node i-1, rows normalised.  For x == xp[0] the index -1 wraps
    (kinetic_model.py:47-48); after normalisation all weight sits on xp[0]."""
    x = np.asarray(x, np.float64)
    xp = np.asarray(xp, np.float64)
    W = np.zeros((x.size, xp.size))
    hi = np.searchsorted(xp, x)
    lo = hi - 1                       # may be -1 -> wraps to the last node, as in the reference
    rows = np.arange(x.size)
    W[rows, hi] = np.abs(xp[lo] - x)
    W[rows, lo] = np.abs(xp[hi] - x)  # written second, exactly like the reference
    W /= W.sum(axis=1, keepdims=True)
    return W


def resample_grid(t):
    """kinetic_model.py:13-18: 2*unique(t).size points on [min t, max t]."""
    n = 2 * np.unique(t).size
    x_rs = np.linspace(np.min(t), np.max(t), n)
    return x_rs, x_rs[1] - x_rs[0]


def continuous_convolution(t, y0, y1):
    """estimate_continuous_convolution (kinetic_model.py:12-32) for y1 of shape
    (T,) or (T, R).  The truncated causal discrete convolution is written as an
    explicit sum (equal to np.convolve(...)[:N] and to scipy convolve1d with
    origin=-N//2, mode='constant')."""
    t = np.asarray(t, np.float64)
    x_rs, dx = resample_grid(t)
    y0_rs = np.interp(x_rs, t, y0)                       # :21
    Wf = interp_weights(x_rs, t)
    y1_2d = y1.reshape(t.size, -1)
    y1_rs = Wf @ y1_2d                                   # :22
    n = x_rs.size
    conv = np.empty_like(y1_rs)
    for r in range(y1_rs.shape[1]):                      # :25-29
        conv[:, r] = np.convolve(y0_rs, y1_rs[:, r])[:n] * dx
    out = interp_weights(t, x_rs) @ conv                 # :32
    return out.reshape(y1.shape)


def srtm2_tac(t, c_r, DVR, R1, k2p):
    """SRTM2.create_activity_curve (kinetic_model.py:142-158). Returns (T, R)."""
    DVR = np.atleast_1d(np.asarray(DVR, np.float64))
    R1 = np.atleast_1d(np.asarray(R1, np.float64))
    k2 = k2p * R1
    k2a = k2 / DVR
    c_exp = np.exp(-k2a[None, :] * np.asarray(t)[:, None])       # :156, :191-196
    return R1[None, :] * np.asarray(c_r)[:, None] + (k2 - R1 * k2a)[None, :] * \
        continuous_convolution(t, c_r, c_exp)


def operator_factors(t):
    """Frame-grid-only factors of M: W_fwd (N,T), W_back (T,N), dx."""
    x_rs, dx = resample_grid(np.asarray(t, np.float64))
    return interp_weights(x_rs, t), interp_weights(t, x_rs), dx


def build_M(t, c_r):
    """M (T,T) with continuous_convolution(t, c_r, E) == M @ E for any E
    (SURVEY.md Appendix A.2).  np.interp (kinetic_model.py:21) and
    interp1d_linear_vec give identical weights, so c_rs = W_fwd @ c_r."""
    Wf, Wb, dx = operator_factors(t)
    c_rs = np.interp(resample_grid(np.asarray(t, np.float64))[0], t, c_r)
    n = c_rs.size
    idx = np.arange(n)[:, None] - np.arange(n)[None, :]
    L = np.where(idx >= 0, c_rs[np.clip(idx, 0, n - 1)], 0.0)
    return dx * (Wb @ (L @ Wf))


def active_columns(t):
    """Columns of M that can be non-zero and per-row prefix length n_j such that
    row j uses exactly active[:n_j] (pattern depends on the frame grid only)."""
    M = build_M(t, np.ones(len(t)))
    nz = M != 0
    active = np.where(nz.any(axis=0))[0]
    nrow = nz[:, active].sum(axis=1)
    # prefix property
    for j in range(len(t)):
        assert nz[j, active[:nrow[j]]].all() and not nz[j, active[nrow[j]:]].any()
    return active, nrow


def srtm2_tac_M(t, c_r, M, DVR, R1, k2p):
    """Same TAC via the operator form: R1 c_r + (k2 - R1 k2a) M exp(-k2a t)."""
    DVR = np.atleast_1d(np.asarray(DVR, np.float64))
    R1 = np.atleast_1d(np.asarray(R1, np.float64))
    k2 = k2p * R1
    k2a = k2 / DVR
    e = np.exp(-k2a[None, :] * np.asarray(t)[:, None])
    return R1[None, :] * np.asarray(c_r)[:, None] + (k2 - R1 * k2a)[None, :] * (M @ e)


def srtm_tac(t, c_r, DVR, k2, R1):
    """SRTM.forward_model (kinetic_model.py:69-84): k2 is a free parameter per ROI. Returns (T, R)."""
    DVR = np.atleast_1d(np.asarray(DVR, np.float64))
    R1 = np.atleast_1d(np.asarray(R1, np.float64))
    k2 = np.atleast_1d(np.asarray(k2, np.float64))
    k2a = k2 / DVR
    c_exp = np.exp(-k2a[None, :] * np.asarray(t)[:, None])
    return R1[None, :] * np.asarray(c_r)[:, None] + (k2 - R1 * k2a)[None, :] * continuous_convolution(t, c_r, c_exp)
