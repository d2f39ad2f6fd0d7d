"""ArviZ posterior summaries as mcmc.py uses them (pm.summary, pm.rhat: mcmc.py:181,186-187),
restated in numpy.

PARITY UNPINNED: arviz is third-party, unpinned by the reference's requirements.txt and not
available offline; this follows the published algorithms (Vehtari, Gelman, Simpson, Carpenter,
Buerkner 2021: rank-normalised split R-hat, bulk/tail ESS with Geyer's initial sequences) as
ArviZ implements them (arviz.stats.diagnostics: _rhat_rank, _ess_bulk, _ess_tail, _ess_mean,
_mcse_mean, _z_scale, _split_chains).  Inputs are (n_chain, n_draw) arrays.
"""
import numpy as np
from scipy import stats


def split_chains(a):
    """arviz _split_chains: first and last n_draw // 2 draws of every chain -> (2C, half)."""
    a = np.asarray(a, np.float64)
    half = a.shape[1] // 2
    return np.vstack([a[:, :half], a[:, -half:]])


def z_scale(a):
    """arviz _z_scale: average ranks over all values, (r - 3/8) / (n + 1/4), inverse normal cdf."""
    a = np.asarray(a, np.float64)
    r = stats.rankdata(a, method="average").reshape(a.shape)
    return stats.norm.ppf((r - 0.375) / (a.size + 0.25))


def _rhat(a):
    n = a.shape[1]
    cm = a.mean(axis=1)
    cv = a.var(axis=1, ddof=1)
    between = n * cm.var(ddof=1)
    within = cv.mean()
    return np.sqrt((between / within + n - 1) / n)


def rhat_rank(a):
    a = np.asarray(a, np.float64)
    bulk = _rhat(z_scale(split_chains(a)))
    folded = np.abs(a - np.median(a))
    tail = _rhat(z_scale(split_chains(folded)))
    return max(bulk, tail)


def autocov(x):
    """Biased autocovariance (divide by n) of one series, all lags (what arviz's FFT computes)."""
    x = np.asarray(x, np.float64)
    n = x.size
    xc = x - x.mean()
    m = 1 << int(np.ceil(np.log2(2 * n)))
    f = np.fft.rfft(xc, m)
    return np.fft.irfft(f * np.conj(f), m)[:n] / n


def ess(a):
    """arviz _ess on (n_chain, n_draw) -- no splitting, no rank transform here."""
    a = np.asarray(a, np.float64)
    n_chain, n_draw = a.shape
    if n_draw < 4:
        return np.nan
    acov = np.stack([autocov(c) for c in a])
    cm = a.mean(axis=1)
    mean_var = acov[:, 0].mean() * n_draw / (n_draw - 1.0)
    var_plus = mean_var * (n_draw - 1.0) / n_draw
    if n_chain > 1:
        var_plus += cm.var(ddof=1)
    if not var_plus > 0:
        return np.nan
    rho = np.zeros(n_draw)
    rho_even = 1.0
    rho[0] = rho_even
    rho_odd = 1.0 - (mean_var - acov[:, 1].mean()) / var_plus
    rho[1] = rho_odd
    t = 1
    while t < (n_draw - 3) and (rho_even + rho_odd) > 0.0:
        rho_even = 1.0 - (mean_var - acov[:, t + 1].mean()) / var_plus
        rho_odd = 1.0 - (mean_var - acov[:, t + 2].mean()) / var_plus
        if (rho_even + rho_odd) >= 0:
            rho[t + 1] = rho_even
            rho[t + 2] = rho_odd
        t += 2
    max_t = t - 2
    if rho_even > 0:
        rho[max_t + 1] = rho_even
    t = 1
    while t <= max_t - 2:
        if (rho[t + 1] + rho[t + 2]) > (rho[t - 1] + rho[t]):
            rho[t + 1] = (rho[t - 1] + rho[t]) / 2.0
            rho[t + 2] = rho[t + 1]
        t += 2
    n_tot = n_chain * n_draw
    tau = -1.0 + 2.0 * rho[: max_t + 1].sum() + rho[max_t + 1: max_t + 2].sum()
    tau = max(tau, 1.0 / np.log10(n_tot))
    return n_tot / tau


def ess_bulk(a):
    return ess(z_scale(split_chains(a)))


def ess_tail(a):
    a = np.asarray(a, np.float64)
    q05, q95 = np.quantile(a, [0.05, 0.95])
    return min(ess(split_chains((a <= q05).astype(np.float64))), ess(split_chains((a <= q95).astype(np.float64))))


def ess_mean(a):
    return ess(split_chains(a))


def mcse_mean(a):
    a = np.asarray(a, np.float64)
    return a.std(ddof=1) / np.sqrt(ess_mean(a))


def ess_sd(a):
    """arviz _ess_sd (arviz/stats/diagnostics.py, 0.12 .. 0.17): ``ary = _split_chains(ary); return min(_ess(ary),
    _ess(ary ** 2))`` -- the mean-ESS of the draws and of their RAW squares (Stan's original definition; newer
    `posterior` releases centre first).  Unpinned like the rest of this file: tests/test_pymc_pin.py compares it with
    az.ess(method="sd") wherever ArviZ is importable."""
    a = np.asarray(a, np.float64)
    return min(ess_mean(a), ess_mean(a ** 2))


def mcse_sd(a):
    """arviz _mcse_sd: sd * sqrt(e (1 - 1/ess)^(ess-1) - 1), ess = ess_sd."""
    a = np.asarray(a, np.float64)
    e = ess_sd(a)
    return a.std(ddof=1) * np.sqrt(np.exp(1) * (1 - 1 / e) ** (e - 1) - 1)


def hdi(a, prob=0.94):
    """arviz.hdi / _hdi (unimodal): narrowest interval holding floor(prob n) + 1 of the pooled sorted draws."""
    x = np.sort(np.asarray(a, np.float64).ravel())
    n = x.size
    inc = int(np.floor(prob * n))
    w = x[inc:] - x[: n - inc]
    i = int(np.argmin(w))
    return x[i], x[i + inc]


def batch_means_ess(a, n_batch=8):
    """The moments-mode estimator of the CUDA library (petmh_diag.cuh), restated: split halves, n_batch batches of
    B = half // n_batch draws each, sigma2_inf = B * var(batch means about their grand mean),
    ESS = N var_plus / sigma2_inf.  Not an ArviZ quantity."""
    a = split_chains(a)
    m, n = a.shape
    B = max(1, n // min(n_batch, n))
    nb = min(n_batch, n // B)
    bm = a[:, : nb * B].reshape(m, nb, B).mean(axis=2)
    s2inf = B * ((bm - bm.mean()) ** 2).sum() / (bm.size - 1)
    var_plus = (n - 1) / n * a.var(axis=1, ddof=1).mean() + a.mean(axis=1).var(ddof=1)
    return min(m * n * var_plus / s2inf, m * n * np.log10(m * n))


def summary_row(a):
    """mean, sd, mcse_mean, ess_bulk, ess_tail, r_hat -- the GPU summary's first six columns."""
    a = np.asarray(a, np.float64)
    return np.array([a.mean(), a.std(ddof=1), mcse_mean(a), ess_bulk(a), ess_tail(a), rhat_rank(a)])


def tfp_ess_cross_chain(a):
    """tfp.mcmc.effective_sample_size(x, cross_chain_dims=...) with its defaults (filter_threshold=0.,
    filter_beyond_lag=None, filter_beyond_positive_pairs=False) for one scalar quantity -- what the
    consumer computes from DVR_mcmc / R1_mcmc (main_script.py:807-810).  a: (n_chain, n_draw).

    PARITY UNPINNED (tensorflow_probability==0.24.0 is third-party and absent).  Restated from TFP's
    published algorithm: per-chain auto-covariance with the unbiased 1/(N-k) normalisation
    (tfp.stats.auto_correlation, normalize=False, center=True), W = mean biased within-chain variance,
    B/N = unbiased variance of the chain means, rho_k = 1 - (W - mean_c acov_k) / (W + B/N), every lag from
    the first rho_k < 0 on dropped, ESS = C N / (-1 + 2 sum_k (N-k)/N rho_k).  One chain: rho_k = acov_k/acov_0.
    """
    a = np.asarray(a, np.float64)
    n_chain, n = a.shape
    k = np.arange(n)
    acov = np.stack([autocov(row) * n / (n - k) for row in a])          # (C, N), 1/(N-k) normalisation
    if n_chain > 1:
        w_biased = acov[:, 0].mean()
        b_div_n = a.mean(axis=1).var(ddof=1)
        rho = 1.0 - (w_biased - acov.mean(axis=0)) / (w_biased + b_div_n)
    else:
        rho = acov[0] / acov[0, 0]
    mask = np.maximum(1.0 - np.cumsum(rho < 0.0), 0.0)
    return n_chain * n / (-1.0 + 2.0 * np.sum((n - k) / n * rho * mask))
