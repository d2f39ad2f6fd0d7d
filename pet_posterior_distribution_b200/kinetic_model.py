"""GPU-backed mirror of the part of the reference's kinetic_model.py that the MCMC path
uses: class SRTM2 with the same constructor and ``create_activity_curve`` signature
(kinetic_model.py:134-161).  The resample-convolve-interpolate "continuous convolution"
(kinetic_model.py:12-32) runs on the B200 as the exact operator conv = M exp(-k2a t).
No CPU fallback.
"""
import numpy as np

from .sampler import MHSampler


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


class SRTM:
    """GPU-backed mirror of kinetic_model.SRTM (kinetic_model.py:62-84): k2 free, reference TAC per call."""

    def __init__(self, frame_time_list, frame_duration_list, device=0):
        self._frame_time_list = np.asarray(frame_time_list, np.float64)
        self._frame_duration_list = np.asarray(frame_duration_list, np.float64)
        self._s = MHSampler(n_chains=1, max_tacs=1, device=device)
        self._s.set_frames(self._frame_time_list, self._frame_duration_list)
        self._s.set_prior(np.zeros(48), np.eye(48), np.zeros(48), np.eye(48))

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
