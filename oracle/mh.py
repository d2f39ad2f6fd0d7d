"""Element-wise Metropolis sampler of mcmc.py:156-157, restated (numpy).

PARITY UNPINNED: follows pymc==5.12.0 `Metropolis.astep` / `tune` / `metrop_select` /
`CompoundStep` semantics as published; pymc is
third-party and not available offline, and the reference ships no chain fixtures.

One `draw` (sweep) = for block in (DVR, R1): [tune every 100 sweeps while tuning];
delta = N(0,1)^48 * scale; fresh random visit order; for each visited coordinate i:
Delta = logp(q with q_i + delta_i) - logp(q); accept iff isfinite(Delta) and log U < Delta.

All randomness comes from an explicit *tape* so that the CUDA kernel (taped mode) and
this oracle consume identical normals / log-uniforms / visit ranks.  The chain state
and the proposal arithmetic are float32 (exactly the kernel's: q' = fl32(q + fl32(n*scale)));
every log-probability is evaluated in float64 at those float32 states.
"""
import numpy as np

TUNE_INTERVAL = 100  # pymc Metropolis default tune_interval


def tune_factor(count, interval=TUNE_INTERVAL):
    """pymc.step_methods.metropolis.tune as a function of the accept count."""
    acc = count / float(interval)
    if acc < 0.001:
        return 0.1
    if acc < 0.05:
        return 0.5
    if acc < 0.2:
        return 0.9
    if acc > 0.95:
        return 10.0
    if acc > 0.75:
        return 2.0
    if acc > 0.5:
        return 1.1
    return 1.0


class Tape:
    """normals, logu: float32 (n_sweeps, 2, 48); rank: uint8 (n_sweeps, 2, 48) where
    rank[s, b, i] is the visit position of coordinate i in block b of sweep s."""

    def __init__(self, normals, logu, rank):
        self.normals = np.ascontiguousarray(normals, np.float32)
        self.logu = np.ascontiguousarray(logu, np.float32)
        self.rank = np.ascontiguousarray(rank, np.uint8)

    @staticmethod
    def random(n_sweeps, rng, n=48):
        normals = rng.standard_normal((n_sweeps, 2, n)).astype(np.float32)
        logu = np.log(1.0 - rng.random((n_sweeps, 2, n))).astype(np.float32)
        rank = np.empty((n_sweeps, 2, n), np.uint8)
        for s in range(n_sweeps):
            for b in range(2):
                rank[s, b] = rng.permutation(n)
        return Tape(normals, logu, rank)


def run_chain(model, tape, n_tune, n_draws, mode="lean", forced_draws=None, init=None):
    """Run one chain for n_tune + n_draws sweeps off `tape`.

    mode 'lean'     : Delta = [ll_i(new) - ll_i(old)] - (d r_i + d^2 P_ii / 2)   (A.3)
    mode 'faithful' : Delta = logp_full(q') - logp_full(q), two full 48-ROI model
                      evaluations per chain-step -- the work pymc's delta_logp does.
    forced_draws    : (n_sweeps, 2, 48) float32 trajectory of another implementation
                      (teacher forcing): decisions are evaluated here but the state
                      follows `forced_draws`; returns agreement statistics.

    Returns dict(draws (n_sweeps,2,48) f32, accept (n_sweeps,2,48) bool,
                 delta (n_sweeps,2,48) f64, scale (2,48) f32, forced_accept, undecidable)
    """
    n_sweeps = n_tune + n_draws
    n = model.mu[0].size
    q = [model.mu[0].astype(np.float32).copy(), model.mu[1].astype(np.float32).copy()]
    if init is not None:
        q = [np.asarray(init[0], np.float32).copy(), np.asarray(init[1], np.float32).copy()]
    scale = np.ones((2, n), np.float32)
    counts = np.zeros((2, n), np.int64)
    draws = np.empty((n_sweeps, 2, n), np.float32)
    accept = np.zeros((n_sweeps, 2, n), bool)
    delta = np.full((n_sweeps, 2, n), np.nan)
    forced_accept = np.zeros((n_sweeps, 2, n), bool) if forced_draws is not None else None
    undecidable = np.zeros((n_sweeps, 2, n), bool) if forced_draws is not None else None

    if mode == "lean":
        ll = model.ll_all(q[0].astype(np.float64), q[1].astype(np.float64))

    for s in range(n_sweeps):
        for b in range(2):
            if s < n_tune and s > 0 and s % TUNE_INTERVAL == 0:
                for i in range(n):
                    scale[b, i] = np.float32(scale[b, i] * np.float32(tune_factor(counts[b, i])))
                counts[b, :] = 0
            step = (tape.normals[s, b] * scale[b]).astype(np.float32)   # fl32(n * scale)
            prop = (q[b] + step).astype(np.float32)                      # fl32(q + step)
            order = np.argsort(tape.rank[s, b], kind="stable")
            if mode == "lean":
                r = model.P[b] @ (q[b].astype(np.float64) - model.mu[b])
            for i in order:
                qi_old = q[b][i]
                qi_new = prop[i]
                d = float(qi_new) - float(qi_old)
                if mode == "lean":
                    if b == 0:
                        ll_new = model.ll_roi(i, float(qi_new), float(q[1][i]))
                    else:
                        ll_new = model.ll_roi(i, float(q[0][i]), float(qi_new))
                    with np.errstate(all="ignore"):
                        dl = (ll_new - ll[i]) - (d * r[i] + 0.5 * d * d * model.P[b][i, i])
                else:
                    q_old64 = [q[0].astype(np.float64), q[1].astype(np.float64)]
                    q_new64 = [q_old64[0].copy(), q_old64[1].copy()]
                    q_new64[b][i] = float(qi_new)
                    with np.errstate(all="ignore"):
                        dl = model.logp_full(*q_new64) - model.logp_full(*q_old64)
                delta[s, b, i] = dl
                acc = bool(np.isfinite(dl) and float(tape.logu[s, b, i]) < dl)
                accept[s, b, i] = acc
                if forced_draws is not None:
                    f_new = forced_draws[s, b, i]
                    if qi_new == qi_old:
                        undecidable[s, b, i] = True      # proposal == state: no information
                        took = acc
                    else:
                        took = bool(f_new == qi_new)
                        assert took or f_new == qi_old, "forced trajectory is not a valid MH path"
                    forced_accept[s, b, i] = took
                    acc_apply = took
                else:
                    acc_apply = acc
                if acc_apply:
                    q[b][i] = qi_new
                    counts[b, i] += 1
                    if mode == "lean":
                        if forced_draws is not None and not acc:
                            # the oracle would have rejected: recompute ll at the forced state
                            ll_new = model.ll_roi(i, float(q[0][i]), float(q[1][i]))
                        ll[i] = ll_new
                        r = r + model.P[b][:, i] * d
            draws[s, b] = q[b]
    return dict(draws=draws, accept=accept, delta=delta, scale=scale,
                forced_accept=forced_accept, undecidable=undecidable)


def run_chain_rng(model, n_tune, n_draws, seed, mode="lean"):
    """Free-running chain with its own numpy random tape (statistical comparisons)."""
    rng = np.random.default_rng(seed)
    return run_chain(model, Tape.random(n_tune + n_draws, rng), n_tune, n_draws, mode=mode)
