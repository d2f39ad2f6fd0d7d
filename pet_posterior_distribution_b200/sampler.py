"""MHSampler: thin Python owner of a libpetmh handle (one CUDA device, one stream).

Mirrors, for batches of test TACs, the body of the reference's per-sample loop
(mcmc.py:104-194): model set-up, pm.sample, chain extraction and summaries.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import N_COORD, N_FRAMES, N_ROI, N_STATS, PetmhError


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class MHSampler:
    """Batched element-wise Metropolis sampler for the SRTM2 posterior on one B200.

    n_chains: chains per TAC (the reference's `chains`, mcmc.py:58, honoured here);
    max_tacs: TACs resident at once; max_draws: stored thinned draws per chain (0 = moments only);
    tac_gid0: global index of the first local TAC, so Philox streams do not depend on sharding.
    """

    def __init__(self, n_chains=4, max_tacs=1, max_draws=0, seed=0, device=0, tac_gid0=0):
        self._h = C.c_void_p()
        cfg = _lib.Cfg(device, n_chains, max_tacs, max_draws, seed, tac_gid0)
        rc = _lib.lib.petmh_create(C.byref(cfg), C.byref(self._h))
        if rc:
            raise PetmhError(rc, (_lib.lib.petmh_last_error(None) or b"").decode())
        self.n_chains, self.max_tacs, self.max_draws = n_chains, max_tacs, max_draws
        self.device = device
        self.n_tac = 0

    # -- plumbing -----------------------------------------------------------------------
    def _ck(self, rc):
        if rc:
            raise PetmhError(rc, (_lib.lib.petmh_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib.petmh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- model --------------------------------------------------------------------------
    def set_frames(self, time_vector, dt):
        t = np.ascontiguousarray(time_vector, np.float64)
        d = np.ascontiguousarray(dt, np.float64)
        if t.shape != (N_FRAMES,) or d.shape != (N_FRAMES,):
            raise ValueError("time_vector and dt must have shape (54,)")
        self._ck(_lib.lib.petmh_set_frames(self._h, _d(t), _d(d)))

    def set_prior(self, mu_DVR, Cov_DVR, mu_R1, Cov_R1):
        a = [np.ascontiguousarray(x, np.float64) for x in (mu_DVR, Cov_DVR, mu_R1, Cov_R1)]
        if a[0].shape != (N_ROI,) or a[1].shape != (N_ROI, N_ROI) or a[2].shape != (N_ROI,) or a[3].shape != (N_ROI, N_ROI):
            raise ValueError("prior must be 48-dimensional")
        self._ck(_lib.lib.petmh_set_prior(self._h, _d(a[0]), _d(a[1]), _d(a[2]), _d(a[3])))

    def set_data(self, y_obs, tac_ref, k2p, sigma_noise=None):
        """y_obs (S,48,54) = tac_noisy/dt; tac_ref (S,54); k2p (S,); sigma_noise (48,54).
        float32 inputs take the bulk path (no conversion copy)."""
        y = np.asarray(y_obs)
        if y.ndim == 2:
            y = y[None]
        S = y.shape[0]
        if y.shape[1:] != (N_ROI, N_FRAMES):
            raise ValueError("y_obs must have shape (S,48,54)")
        if y.dtype == np.float32:
            y = np.ascontiguousarray(y)
            cr = np.ascontiguousarray(np.asarray(tac_ref, np.float32).reshape(S, N_FRAMES))
            k = np.ascontiguousarray(np.broadcast_to(np.asarray(k2p, np.float32).reshape(-1), (S,)))
            sn = None if sigma_noise is None else np.ascontiguousarray(sigma_noise, np.float32)
            self._ck(_lib.lib.petmh_set_data_f32(self._h, S, _f(y), _f(cr), _f(k), None if sn is None else _f(sn)))
        else:
            y = np.ascontiguousarray(y, np.float64)
            cr = np.ascontiguousarray(np.asarray(tac_ref, np.float64).reshape(S, N_FRAMES))
            k = np.ascontiguousarray(np.broadcast_to(np.asarray(k2p, np.float64).reshape(-1), (S,)))
            sn = None if sigma_noise is None else np.ascontiguousarray(sigma_noise, np.float64)
            self._ck(_lib.lib.petmh_set_data(self._h, S, _d(y), _d(cr), _d(k), None if sn is None else _d(sn)))
        self.n_tac = S

    def set_data_ptr(self, n_tac, y_ptr, tac_ref_ptr, k2p_ptr, sigma_ptr=None):
        """Bulk float32 path from raw HOST pointers (e.g. pinned torch tensors' data_ptr())."""
        fp = C.POINTER(C.c_float)
        cast = lambda p: C.cast(C.c_void_p(int(p)), fp) if p else None
        self._ck(_lib.lib.petmh_set_data_f32(self._h, int(n_tac), cast(y_ptr), cast(tac_ref_ptr), cast(k2p_ptr), cast(sigma_ptr)))
        self.n_tac = int(n_tac)

    def summary_ptr(self, out_ptr):
        """Summary (S,96,8) f32 into a raw HOST pointer (e.g. pinned memory)."""
        self._ck(_lib.lib.petmh_get_summary(self._h, C.cast(C.c_void_p(int(out_ptr)), C.POINTER(C.c_float))))

    def synth(self, n_tac, seed, mu_tac_ref, Cov_tac_ref, k2p, sigma_noise):
        """K4: generate n_tac training-style synthetic TACs on the GPU and bind them as this sampler's data."""
        mu = np.ascontiguousarray(mu_tac_ref, np.float64)
        cov = np.ascontiguousarray(Cov_tac_ref, np.float64)
        sn = np.ascontiguousarray(sigma_noise, np.float64)
        if mu.shape != (N_FRAMES,) or cov.shape != (N_FRAMES, N_FRAMES) or sn.shape != (N_ROI, N_FRAMES):
            raise ValueError("mu_tac_ref (54,), Cov_tac_ref (54,54), sigma_noise (48,54) expected")
        rc = _lib.lib.petmh_synth(self._h, int(n_tac), int(seed), _d(mu), _d(cov), float(k2p), _d(sn))
        if rc in (0, -6):                 # PETMH_ESYNTH: the data IS bound, some TACs hit a rejection cap (attempts < 0)
            self.n_tac = int(n_tac)
        self._ck(rc)

    def synth_test_rule(self, alpha, Cov_DVR=None, Cov_R1=None, Cov_tac_ref=None, dof=N_ROI):
        """Test-style rejection rule for the following synth() calls (sample_sim_data.py:128-133): keep a drawn vector only
        if chi2.cdf(Mahalanobis^2, dof) < alpha, with np.linalg.inv(Cov) as the reference computes it (:106,110,117) and
        dof = 48 for all three variables (:132).  alpha = None switches the rule off (the training-style set)."""
        if alpha is None:
            self._ck(_lib.lib.petmh_synth_set_test_rule(self._h, 0.0, None, None, None))
            return
        from scipy import stats
        inv = [np.ascontiguousarray(np.linalg.inv(np.asarray(c, np.float64))) for c in (Cov_DVR, Cov_R1, Cov_tac_ref)]
        if inv[0].shape != (N_ROI, N_ROI) or inv[1].shape != (N_ROI, N_ROI) or inv[2].shape != (N_FRAMES, N_FRAMES):
            raise ValueError("Cov_DVR (48,48), Cov_R1 (48,48), Cov_tac_ref (54,54) expected")
        self._ck(_lib.lib.petmh_synth_set_test_rule(self._h, float(stats.chi2.ppf(alpha, dof)), _d(inv[0]), _d(inv[1]), _d(inv[2])))

    def synth_get(self, fields=("DVR", "R1", "tac_ref", "tac_clean", "y", "attempts")):
        """dict(DVR (n,48), R1 (n,48), tac_ref (n,54), tac_clean (n,48,54), y (n,48,54), attempts (n,)); `fields` limits
        what is copied back (a million TACs are 10 GB per (n,48,54) array)."""
        n = self.n_tac
        want = set(fields)
        dr = np.empty((n, N_COORD), np.float32) if want & {"DVR", "R1"} else None
        cr = np.empty((n, N_FRAMES), np.float64) if "tac_ref" in want else None
        cl = np.empty((n, N_ROI, N_FRAMES), np.float32) if "tac_clean" in want else None
        y = np.empty((n, N_ROI, N_FRAMES), np.float32) if "y" in want else None
        at = np.empty(n, np.int32) if "attempts" in want else None
        self._ck(_lib.lib.petmh_synth_get(self._h, None if dr is None else _f(dr), None if cr is None else _d(cr),
                                          None if cl is None else _f(cl), None if y is None else _f(y),
                                          None if at is None else at.ctypes.data_as(C.POINTER(C.c_int))))
        out = dict(tac_ref=cr, tac_clean=cl, y=y, attempts=at)
        if dr is not None:
            out.update(DVR=dr[:, :N_ROI].copy(), R1=dr[:, N_ROI:].copy())
        return {k: v for k, v in out.items() if v is not None}

    # -- parity hooks ---------------------------------------------------------------------
    def forward(self, tac, DVR, R1):
        """(48,54) model TAC, == SRTM2.create_activity_curve(DVR,R1,k2p).T (mcmc.py:38-39)."""
        a = np.ascontiguousarray(DVR, np.float64)
        b = np.ascontiguousarray(R1, np.float64)
        out = np.empty((N_ROI, N_FRAMES), np.float64)
        self._ck(_lib.lib.petmh_forward(self._h, int(tac), _d(a), _d(b), _d(out)))
        return out

    def forward_srtm(self, tac, DVR, k2, R1):
        """(48,54) SRTM TAC with k2 free, == SRTM.forward_model(DVR, k2, R1, tac_ref).T (kinetic_model.py:69-84)."""
        a, k, b = (np.ascontiguousarray(x, np.float64) for x in (DVR, k2, R1))
        out = np.empty((N_ROI, N_FRAMES), np.float64)
        self._ck(_lib.lib.petmh_forward_srtm(self._h, int(tac), _d(a), _d(k), _d(b), _d(out)))
        return out

    def loglik(self, tac, DVR, R1):
        """(ll per ROI (48,), (logprior_DVR, logprior_R1)) of the pymc model (mcmc.py:148-155)."""
        a = np.ascontiguousarray(DVR, np.float64)
        b = np.ascontiguousarray(R1, np.float64)
        ll = np.empty(N_ROI, np.float64)
        lp = np.empty(2, np.float64)
        self._ck(_lib.lib.petmh_loglik(self._h, int(tac), _d(a), _d(b), _d(ll), _d(lp)))
        return ll, lp

    def operator(self, tac):
        m = np.empty((N_FRAMES, N_FRAMES), np.float64)
        self._ck(_lib.lib.petmh_get_operator(self._h, int(tac), _d(m)))
        return m

    def cheb_operator(self, tac):
        """The Chebyshev form of the operator as the sweep kernel uses it: list of three (18, ncols) float32 blocks
        A_b with conv[18 b : 18 b + 18] = A_b @ T_0..(s), and the k2a range (lo, hi) it is valid on."""
        a = np.empty(26 * 20, np.float32)
        nc = (C.c_int * 3)()
        lo, hi = C.c_double(), C.c_double()
        self._ck(_lib.lib.petmh_get_cheb_operator(self._h, int(tac), _f(a), nc, C.byref(lo), C.byref(hi)))
        blocks, off = [], 0
        for b in range(3):
            blocks.append(a[off:off + nc[b] * 20].reshape(nc[b], 20)[:, :18].T.copy())
            off += nc[b] * 20
        return blocks, (lo.value, hi.value)

    def philox_raw(self, chain_gid, sweep, block):
        out = np.empty((N_ROI, 4), np.uint32)
        self._ck(_lib.lib.petmh_philox_raw(self._h, int(chain_gid), int(sweep), int(block),
                                           out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out

    # -- sampling ---------------------------------------------------------------------------
    def run(self, draws, tune, thin=1):
        self._ck(_lib.lib.petmh_run(self._h, int(draws), int(tune), int(thin)))

    def reset(self):
        self._ck(_lib.lib.petmh_reset(self._h))

    def plan(self, draws, tune, thin=1):
        self._ck(_lib.lib.petmh_plan(self._h, int(draws), int(tune), int(thin)))

    def advance(self, n_sweeps):
        self._ck(_lib.lib.petmh_advance(self._h, int(n_sweeps)))

    def run_taped(self, tac, normals, logu, rank, tune):
        """normals/logu (c,s,2,48) f32, rank (c,s,2,48) u8 -> dict(draws, delta, accept, scale)."""
        n = np.ascontiguousarray(normals, np.float32)
        lu = np.ascontiguousarray(logu, np.float32)
        rk = np.ascontiguousarray(rank, np.uint8)
        c, s = n.shape[:2]
        draws = np.empty((c, s, 2, N_ROI), np.float32)
        delta = np.empty((c, s, 2, N_ROI), np.float32)
        acc = np.empty((c, s, 2, N_ROI), np.uint8)
        scale = np.empty((c, 2, N_ROI), np.float32)
        u8 = C.POINTER(C.c_uint8)
        self._ck(_lib.lib.petmh_run_taped(self._h, int(tac), c, s, int(tune), _f(n), _f(lu), rk.ctypes.data_as(u8),
                                          _f(draws), _f(delta), acc.ctypes.data_as(u8), _f(scale)))
        return dict(draws=draws, delta=delta, accept=acc.astype(bool), scale=scale)

    def sample_srtm(self, mu_k2, Cov_k2, draws, tune, thin=1, tape=None, tac=0):
        """The k2-free SRTM (kinetic_model.py:62-84) as a sampled model: element-wise Metropolis over the blocks DVR, R1, k2
        with an MvNormal(mu_k2, Cov_k2) prior on k2.  Free run: dict(DVR, R1, k2: (S, chains, n, 48),
        accept_rate (S, chains, 3, 48)).  tape = (normals, logu, rank), each (c, draws + tune, 3, 48): teacher-forcing
        mode for TAC `tac`: dict(draws (c, s, 3, 48), delta, accept)."""
        mu = np.ascontiguousarray(mu_k2, np.float64)
        cov = np.ascontiguousarray(Cov_k2, np.float64)
        if mu.shape != (N_ROI,) or cov.shape != (N_ROI, N_ROI):
            raise ValueError("k2 prior must be 48-dimensional")
        u8 = C.POINTER(C.c_uint8)
        if tape is None:
            n_out = (int(draws) + int(thin) - 1) // int(thin)
            out = np.empty((self.n_tac, self.n_chains, n_out, 3, N_ROI), np.float32)
            acc = np.empty((self.n_tac, self.n_chains, 3, N_ROI), np.float32)
            self._ck(_lib.lib.petmh_srtm_sample(self._h, _d(mu), _d(cov), int(draws), int(tune), int(thin), 0, 0, None, None, None,
                                                _f(out), None, None, _f(acc)))
            return dict(DVR=out[..., 0, :], R1=out[..., 1, :], k2=out[..., 2, :], accept_rate=acc)
        n, lu, rk = (np.ascontiguousarray(tape[0], np.float32), np.ascontiguousarray(tape[1], np.float32),
                     np.ascontiguousarray(tape[2], np.uint8))
        c, s = n.shape[:2]
        if s != draws + tune or n.shape != (c, s, 3, N_ROI):
            raise ValueError("tape arrays must have shape (chains, draws + tune, 3, 48)")
        out = np.empty((c, s, 3, N_ROI), np.float32)
        delta = np.empty((c, s, 3, N_ROI), np.float32)
        acc = np.empty((c, s, 3, N_ROI), np.uint8)
        self._ck(_lib.lib.petmh_srtm_sample(self._h, _d(mu), _d(cov), int(draws), int(tune), 1, c, int(tac), _f(n), _f(lu),
                                            rk.ctypes.data_as(u8), _f(out), _f(delta), acc.ctypes.data_as(u8), None))
        return dict(draws=out, delta=delta, accept=acc.astype(bool))

    # -- outputs ------------------------------------------------------------------------------
    @property
    def n_stored(self):
        return _lib.lib.petmh_n_stored(self._h)

    def chains(self):
        """(DVR_mcmc, R1_mcmc), each (S, chains, n_stored, 48) float32 (mcmc.py:162-163 layout per TAC)."""
        ns = self.n_stored
        shape = (self.n_tac, self.n_chains, ns, N_ROI)
        dvr = np.empty(shape, np.float32)
        r1 = np.empty(shape, np.float32)
        self._ck(_lib.lib.petmh_get_chains(self._h, _f(dvr), _f(r1)))
        return dvr, r1

    def summary(self):
        """(S, 96, 8) float32; columns _lib.STAT_NAMES; rows DVR[0..47] then R1[0..47]."""
        out = np.empty((self.n_tac, N_COORD, N_STATS), np.float32)
        self._ck(_lib.lib.petmh_get_summary(self._h, _f(out)))
        return out

    def summary_ext(self):
        """(S, 96, 4) float32: hdi_3%, hdi_97%, mcse_sd, ess_sd -- the pm.summary columns beyond summary()'s, from the
        stored draws on the GPU."""
        out = np.empty((self.n_tac, N_COORD, 4), np.float32)
        self._ck(_lib.lib.petmh_get_summary_ext(self._h, _f(out)))
        return out

    def ess_cross_chain(self):
        """(S, 96) float32: tfp.mcmc.effective_sample_size(..., cross_chain_dims=-1) of the stored draws
        (what main_script.py:807-810 computes from DVR_mcmc / R1_mcmc), on the GPU."""
        out = np.empty((self.n_tac, N_COORD), np.float32)
        self._ck(_lib.lib.petmh_get_ess_cross_chain(self._h, _f(out)))
        return out

    def posterior_cov(self):
        """(cov, corr), each (S, 2, 48, 48) float64 (block 0 = DVR, 1 = R1): np.cov / np.corrcoef of the pooled stored draws
        across ROIs, the matrices main_script.py:717-738 builds from DVR_mcmc / R1_mcmc, on the GPU."""
        cov = np.empty((self.n_tac, 2, N_ROI, N_ROI), np.float64)
        corr = np.empty_like(cov)
        self._ck(_lib.lib.petmh_get_posterior_cov(self._h, _d(cov), _d(corr)))
        return cov, corr

    def summary_into(self, device_ptr, stream=None):
        """Write the (S,96,8) f32 summary to DEVICE memory (e.g. an all-gather slot)."""
        self._ck(_lib.lib.petmh_summary_device(self._h, C.c_void_p(int(device_ptr)),
                                               C.c_void_p(int(stream)) if stream else None))

    def state(self):
        q = np.empty((self.n_tac, self.n_chains, N_COORD), np.float32)
        sc = np.empty_like(q)
        self._ck(_lib.lib.petmh_get_state(self._h, _f(q), _f(sc)))
        return q, sc

    def set_state(self, q=None, scale=None, sweep=0):
        """Warm start / resume: (S, chains, 96) float32 positions and scalings."""
        qq = None if q is None else np.ascontiguousarray(q, np.float32)
        ss = None if scale is None else np.ascontiguousarray(scale, np.float32)
        for a in (qq, ss):
            if a is not None and a.shape != (self.n_tac, self.n_chains, N_COORD):
                raise ValueError("state arrays must have shape (n_tac, n_chains, 96)")
        self._ck(_lib.lib.petmh_set_state(self._h, None if qq is None else _f(qq), None if ss is None else _f(ss), int(sweep)))

    def checkpoint(self):
        """Everything needed to continue this run exactly (positions, scalings, tune counters, accepted-move counters,
        running moments, stored draws, schedule position) as a uint8 array."""
        n = int(_lib.lib.petmh_checkpoint_bytes(self._h))
        buf = np.empty(n, np.uint8)
        self._ck(_lib.lib.petmh_get_checkpoint(self._h, buf.ctypes.data_as(C.c_void_p), n))
        return buf

    def restore(self, blob):
        """Continue from a checkpoint() blob (bind the same data with set_data first)."""
        buf = np.ascontiguousarray(blob, np.uint8)
        self._ck(_lib.lib.petmh_set_checkpoint(self._h, buf.ctypes.data_as(C.c_void_p), buf.size))

    def set_global_ids(self, tac_gids=None, chain_gid0=0, chains_per_tac_global=0):
        """Global identity of the local TACs / chains (Philox streams independent of batching and sharding):
        tac_gids (n_tac,) uint64 or None (= tac_gid0 + local index); chains of a TAC split over ranks pass the
        first local chain's global index and the global chain count."""
        g = None if tac_gids is None else np.ascontiguousarray(tac_gids, np.uint64)
        self._ck(_lib.lib.petmh_set_global_ids(self._h, 0 if g is None else g.size,
                                               None if g is None else g.ctypes.data_as(C.POINTER(C.c_uint64)),
                                               int(chain_gid0), int(chains_per_tac_global)))

    def set_stream(self, stream):
        self._ck(_lib.lib.petmh_set_stream(self._h, C.c_void_p(int(stream))))

    def synchronize(self):
        self._ck(_lib.lib.petmh_synchronize(self._h))

    def last_kernel_ms(self):
        ms = C.c_float()
        n = C.c_int()
        self._ck(_lib.lib.petmh_last_kernel_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value
