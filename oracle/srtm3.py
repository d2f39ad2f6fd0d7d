"""The k2-free SRTM (kinetic_model.py:62-84) as a sampled three-block model -- CHECKER for petmh_srtm_sample
(test infrastructure only).  Forward model: SRTM.forward_model(DVR, k2, R1, tac_ref) restated through the
pinned operator M (oracle/forward.py): TAC = R1 c_r + (k2 - R1 k2a) M exp(-k2a t), k2a = k2 / DVR (kinetic_model.py:76-84).
Likelihood block and DVR / R1 priors as mcmc.py:148-155; the MvNormal prior on k2 is the caller's (the reference ships none).
Sampler: the element-wise Metropolis of oracle/mh.py (pymc semantics, parity unpinned) over the blocks DVR, R1, k2, off an
explicit tape, with the same fp32 state arithmetic."""
import numpy as np

from .logp import loglik_roi_reduced
from .mh import TUNE_INTERVAL, tune_factor


class Model3:
    def __init__(self, m, mu_k2, Cov_k2):
        """m: an oracle.logp.Model (frames, reference TAC, data, DVR / R1 priors)."""
        self.m = m
        self.mu = [m.mu[0], m.mu[1], np.asarray(mu_k2, np.float64)]
        P = np.linalg.inv(np.asarray(Cov_k2, np.float64))
        self.P = [m.P[0], m.P[1], 0.5 * (P + P.T)]

    def tac_roi(self, dvr, r1, k2):
        with np.errstate(all="ignore"):
            k2a = k2 / dvr
            return r1 * self.m.c_r + (k2 - r1 * k2a) * (self.m.M @ np.exp(-k2a * self.m.t))

    def ll_roi(self, roi, dvr, r1, k2):
        return float(loglik_roi_reduced(self.m.y[roi], self.tac_roi(dvr, r1, k2), self.m.sigma_noise[roi]))


def random_tape(n_sweeps, rng, n=48):
    normals = rng.standard_normal((n_sweeps, 3, n)).astype(np.float32)
    logu = np.log(1.0 - rng.random((n_sweeps, 3, n))).astype(np.float32)
    rank = np.stack([[rng.permutation(n) for _ in range(3)] for _ in range(n_sweeps)]).astype(np.uint8)
    return normals, logu, rank


def run_chain(model, tape, n_tune, n_draws, forced_draws=None):
    """As oracle.mh.run_chain (mode 'lean'), three blocks.  forced_draws (n_sweeps, 3, 48): teacher forcing."""
    normals, logu, rank = tape
    n_sweeps, n = n_tune + n_draws, 48
    q = [mu.astype(np.float32).copy() for mu in model.mu]
    scale = np.ones((3, n), np.float32)
    counts = np.zeros((3, n), np.int64)
    draws = np.empty((n_sweeps, 3, n), np.float32)
    accept = np.zeros((n_sweeps, 3, n), bool)
    delta = np.full((n_sweeps, 3, n), np.nan)
    forced_accept = np.zeros((n_sweeps, 3, n), bool)
    undecidable = np.zeros((n_sweeps, 3, n), bool)
    ll = np.array([model.ll_roi(i, float(q[0][i]), float(q[1][i]), float(q[2][i])) for i in range(n)])
    for s in range(n_sweeps):
        for b in range(3):
            if s < n_tune and s > 0 and s % TUNE_INTERVAL == 0:
                for i in range(n):
                    scale[b, i] = np.float32(scale[b, i] * np.float32(tune_factor(counts[b, i])))
                counts[b, :] = 0
            prop = (q[b] + (normals[s, b] * scale[b]).astype(np.float32)).astype(np.float32)
            r = model.P[b] @ (q[b].astype(np.float64) - model.mu[b])
            for i in np.argsort(rank[s, b], kind="stable"):
                qo, qn = q[b][i], prop[i]
                d = float(qn) - float(qo)
                args = [float(q[0][i]), float(q[1][i]), float(q[2][i])]
                args[b] = float(qn)
                ll_new = model.ll_roi(i, *args)
                with np.errstate(all="ignore"):
                    dl = (ll_new - ll[i]) - (d * r[i] + 0.5 * d * d * model.P[b][i, i])
                delta[s, b, i] = dl
                acc = bool(np.isfinite(dl) and float(logu[s, b, i]) < dl)
                accept[s, b, i] = acc
                took = acc
                if forced_draws is not None:
                    if qn == qo:
                        undecidable[s, b, i] = True
                    else:
                        took = bool(forced_draws[s, b, i] == qn)
                        assert took or forced_draws[s, b, i] == qo, "forced trajectory is not a valid MH path"
                    forced_accept[s, b, i] = took
                if took:
                    q[b][i] = qn
                    counts[b, i] += 1
                    if forced_draws is not None and not acc:
                        ll_new = model.ll_roi(i, float(q[0][i]), float(q[1][i]), float(q[2][i]))
                    ll[i] = ll_new
                    r = r + model.P[b][:, i] * d
            draws[s, b] = q[b]
    return dict(draws=draws, accept=accept, delta=delta, scale=scale, forced_accept=forced_accept, undecidable=undecidable)
