"""GPU-backed mirror of the reference's kinetic_model.py: the classes SRTM2 and SRTM with the same constructors and
``create_activity_curve`` / ``forward_model`` signatures (kinetic_model.py:62-84, 134-161) -- the resample-convolve-
interpolate "continuous convolution" runs on the B200 as the exact operator conv = M exp(-k2a t) -- and the module-level
helpers ``estimate_continuous_convolution`` (:12-32) and ``interp1d_linear_vec`` (:35-57) plus the classes' static
``make_time_exponential`` / ``make_time_func`` / ``convolve`` (:89-128, :163-201) on general grids in fp64
(csrc/petmh_conv.cuh).  No CPU fallback: every number comes from libpetmh.
"""
import atexit
import ctypes as C

import numpy as np

from . import _lib
from .sampler import MHSampler

_helpers = {}


@atexit.register
def _close_helpers():
    while _helpers:
        _helpers.popitem()[1].close()


def _helper(device=0):
    """One small handle per device for the stateless helper calls (they only need its device and stream)."""
    if device not in _helpers:
        _helpers[device] = MHSampler(n_chains=1, max_tacs=1, device=device)
    return _helpers[device]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def interp1d_linear_vec(x, xp, fp, dim=0, device=0):
    """kinetic_model.interp1d_linear_vec (kinetic_model.py:35-57): linear interpolation of fp (given at xp, along axis
    `dim`) at x, with the reference's weights (searchsorted node and its left neighbour; index -1 wraps for x <= xp[0]).
    x beyond xp[-1] raises IndexError like the reference."""
    x = np.ascontiguousarray(x, np.float64).reshape(-1)
    xp = np.ascontiguousarray(xp, np.float64).reshape(-1)
    fp = np.asarray(fp, np.float64)
    f2 = fp.reshape(-1, 1) if fp.ndim == 1 else np.moveaxis(fp, dim, 0)
    if f2.shape[0] != xp.size:
        raise ValueError("fp has %d points along dim %d, xp has %d" % (f2.shape[0], dim, xp.size))
    if x.size and not (x <= xp[-1]).all():
        raise IndexError("index %d is out of bounds for axis 1 with size %d" % (xp.size, xp.size))
    rest = f2.shape[1:]
    f2 = np.ascontiguousarray(f2.reshape(xp.size, -1))
    out = np.empty((x.size, f2.shape[1]), np.float64)
    s = _helper(device)
    s._ck(_lib.lib.petmh_interp1d_linear(s._h, x.size, _dp(x), xp.size, _dp(xp), _dp(f2), f2.shape[1], _dp(out)))
    if fp.ndim == 1:
        return out.reshape(x.size)
    return np.moveaxis(out.reshape((x.size,) + rest), 0, dim)


def estimate_continuous_convolution(x, y0, y1, num_points_resample=None, device=0):
    """kinetic_model.estimate_continuous_convolution (kinetic_model.py:12-32): y0 (n,) convolved with every column of y1
    ((n,) or (n, ...)) on the grid x: resampling onto num_points_resample (default 2 n) uniform points, truncated causal
    discrete convolution times the spacing, interpolation back."""
    x = np.ascontiguousarray(x, np.float64).reshape(-1)
    y0 = np.ascontiguousarray(y0, np.float64).reshape(-1)
    y1 = np.asarray(y1, np.float64)
    if y0.size != x.size or y1.shape[0] != x.size:
        raise ValueError("y0 and y1 must have len(x) points along axis 0")
    N = 0 if num_points_resample is None else int(num_points_resample)
    if y1.ndim > 1 and N % 2:       # scipy.ndimage.convolve1d (kinetic_model.py:28) has no origin -N//2 for an odd length
        raise ValueError("Invalid origin; origin must satisfy -(len(weights) // 2) <= origin <= (len(weights)-1) // 2")
    y2 = np.ascontiguousarray(y1.reshape(x.size, -1))
    out = np.empty_like(y2)
    s = _helper(device)
    s._ck(_lib.lib.petmh_continuous_convolution(s._h, x.size, _dp(x), _dp(y0), _dp(y2), y2.shape[1], N, _dp(out)))
    return out.reshape(y1.shape)


def _make_time_func(param, time_vector, func, time_scale=None, space_scale=None):
    """SRTM.make_time_func (kinetic_model.py:89-116): broadcast `func(param, t)` over a leading time axis and apply the
    optional scales.  `func` is the caller's Python callable (host); the model classes use make_time_exponential."""
    time_vector = np.asarray(time_vector, np.float64)
    if np.isscalar(param):
        output = func(param, time_vector)
        if time_scale is not None:
            output = output * time_scale
        if space_scale is not None:
            output = output * space_scale
        return output
    param = np.asarray(param, np.float64)
    tshape = [-1] + [1] * param.ndim
    output = func(param.reshape([1] + list(param.shape)), time_vector.reshape(tshape))
    if time_scale is not None:
        time_scale = np.asarray(time_scale)
        output = output * (time_scale.reshape(tshape) if time_scale.ndim == 1 else time_scale)
    if space_scale is not None:
        space_scale = np.asarray(space_scale)
        output = output * (space_scale if space_scale.ndim == output.ndim else space_scale[np.newaxis])
    return output


def _make_time_exponential(param, time_vector, time_scale=None, device=0, **kwargs):
    """SRTM.make_time_exponential (kinetic_model.py:118-122): exp(param * t) with a leading time axis, on the GPU."""
    t = np.ascontiguousarray(time_vector, np.float64).reshape(-1)
    p = np.ascontiguousarray(np.atleast_1d(np.asarray(param, np.float64)))
    flat = np.ascontiguousarray(p.reshape(-1))
    out = np.empty((t.size, flat.size), np.float64)
    s = _helper(device)
    s._ck(_lib.lib.petmh_time_exponential(s._h, flat.size, _dp(flat), t.size, _dp(t), _dp(out)))
    out = out[:, 0] if np.isscalar(param) else out.reshape((t.size,) + p.shape)
    if time_scale is None and not kwargs:
        return out
    return _make_time_func(param, t, lambda *_: out, time_scale=time_scale, **kwargs)


class SRTM2:
    def __init__(self, frame_time_list, frame_duration_list, tac_reference, device=0):
        self._frame_time_list = np.asarray(frame_time_list, np.float64)
        self.frame_duration_list = np.asarray(frame_duration_list, np.float64)
        self._tac_reference = np.asarray(tac_reference, np.float64)
        self._k2p = None
        self._s = MHSampler(n_chains=1, max_tacs=1, device=device)
        self._s.set_frames(self._frame_time_list, self.frame_duration_list)
        ident = np.eye(48)
        self._s.set_prior(np.zeros(48), ident, np.zeros(48), ident)   # priors are not used by the forward model

    # mcmc.py's Op holds its SRTM2 and is pickled to PyMC's worker processes (SURVEY.md 8 b): the object pickles as its
    # three arrays and its device; the library handle is rebuilt on the other side.
    def __getstate__(self):
        return {"frame_time_list": self._frame_time_list, "frame_duration_list": self.frame_duration_list,
                "tac_reference": self._tac_reference, "device": self._s.device}

    def __setstate__(self, st):
        self.__init__(st["frame_time_list"], st["frame_duration_list"], st["tac_reference"], device=st["device"])

    def _bind(self, k2p):
        if self._k2p != float(k2p):
            self._s.set_data(np.ones((1, 48, 54)), self._tac_reference[None], np.array([float(k2p)]), np.ones((48, 54)))
            self._k2p = float(k2p)

    def create_activity_curve(self, DVR=None, R1=None, k2p=None):
        """(54, n_roi) model TAC like the reference (n_roi <= 48; scalars give (54,))."""
        scalar = np.isscalar(DVR)
        d = np.atleast_1d(np.asarray(DVR, np.float64))
        r = np.atleast_1d(np.asarray(R1, np.float64))
        n = d.size
        if n > 48 or r.size != n:
            raise ValueError("DVR and R1 must have the same length <= 48")
        self._bind(np.asarray(k2p, np.float64).reshape(-1)[0])
        dd = np.ones(48); rr = np.ones(48)
        dd[:n] = d; rr[:n] = r
        out = self._s.forward(0, dd, rr)[:n].T          # (54, n)
        return out[:, 0] if scalar else out

    __call__ = create_activity_curve
    make_time_func = staticmethod(_make_time_func)                   # kinetic_model.py:163-196
    make_time_exponential = staticmethod(_make_time_exponential)
    convolve = staticmethod(lambda time_vector, c_0, c_1: estimate_continuous_convolution(time_vector, c_0, c_1))   # :198-201


class SRTM:
    """GPU-backed mirror of kinetic_model.SRTM (kinetic_model.py:62-84): k2 free, reference TAC per call."""

    def __init__(self, frame_time_list, frame_duration_list, device=0):
        self._frame_time_list = np.asarray(frame_time_list, np.float64)
        self._frame_duration_list = np.asarray(frame_duration_list, np.float64)
        self._s = MHSampler(n_chains=1, max_tacs=1, device=device)
        self._s.set_frames(self._frame_time_list, self._frame_duration_list)
        self._s.set_prior(np.zeros(48), np.eye(48), np.zeros(48), np.eye(48))

    def __getstate__(self):
        return {"frame_time_list": self._frame_time_list, "frame_duration_list": self._frame_duration_list, "device": self._s.device}

    def __setstate__(self, st):
        self.__init__(st["frame_time_list"], st["frame_duration_list"], device=st["device"])

    def forward_model(self, DVR=None, k2=None, R1=None, tac_ref=None):
        scalar = np.isscalar(DVR)
        d, k, r = (np.atleast_1d(np.asarray(x, np.float64)) for x in (DVR, k2, R1))
        n = d.size
        if n > 48 or k.size != n or r.size != n:
            raise ValueError("DVR, k2 and R1 must have the same length <= 48")
        self._tac_reference = np.asarray(tac_ref, np.float64)
        self._s.set_data(np.ones((1, 48, 54)), self._tac_reference[None], np.array([0.0]), np.ones((48, 54)))
        dd, kk, rr = np.ones(48), np.ones(48), np.ones(48)
        dd[:n], kk[:n], rr[:n] = d, k, r
        out = self._s.forward_srtm(0, dd, kk, rr)[:n].T
        return out[:, 0] if scalar else out

    __call__ = forward_model
    make_time_func = staticmethod(_make_time_func)                   # kinetic_model.py:89-122
    make_time_exponential = staticmethod(_make_time_exponential)
    convolve = staticmethod(lambda time_vector, c_0, c_1: estimate_continuous_convolution(time_vector, c_0, c_1))   # :124-128
