"""Numpy restatement of ArviZ's diagnostics (oracle/diagnostics.py): arviz itself is absent
(parity unpinned), so these are property checks against known answers."""
import numpy as np

from oracle import diagnostics as dg


def _ar1(rng, n_chain, n, rho):
    x = np.zeros((n_chain, n))
    e = rng.standard_normal((n_chain, n))
    x[:, 0] = e[:, 0]
    for t in range(1, n):
        x[:, t] = rho * x[:, t - 1] + np.sqrt(1 - rho ** 2) * e[:, t]
    return x


def test_iid_draws():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((4, 4000))
    assert abs(dg.rhat_rank(x) - 1) < 0.01
    for f in (dg.ess_bulk, dg.ess_mean, dg.ess_tail):
        assert 0.8 * x.size < f(x) < 1.25 * x.size
    assert abs(dg.mcse_mean(x) - 1 / np.sqrt(x.size)) < 0.2 / np.sqrt(x.size)


def test_ar1_effective_sample_size():
    rng = np.random.default_rng(1)
    rho = 0.9
    x = _ar1(rng, 8, 8000, rho)
    expect = x.size * (1 - rho) / (1 + rho)
    assert 0.8 * expect < dg.ess_mean(x) < 1.25 * expect
    assert 0.7 * expect < dg.ess_bulk(x) < 1.4 * expect
    assert dg.rhat_rank(x) < 1.02


def test_unconverged_chains_flagged():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((4, 1000))
    x[0] += 3.0                                     # one chain elsewhere
    assert dg.rhat_rank(x) > 1.3
    y = rng.standard_normal((4, 1000)) * np.array([1, 1, 1, 5.0])[:, None]   # same mean, different scale: folded R-hat
    assert dg.rhat_rank(y) > 1.05
    assert dg.ess_bulk(x) < 50


def test_ties_and_split():
    """Metropolis chains repeat values: ranks are averaged; splitting keeps 2*(n//2) draws per chain."""
    x = np.repeat(np.arange(50.0), 4)[None, :].repeat(2, axis=0)      # (2, 200) with ties of 4
    z = dg.z_scale(x)
    assert np.allclose(z[0, :4], z[0, 0]) and z[0, 4] > z[0, 3]
    assert dg.split_chains(np.zeros((3, 101))).shape == (6, 50)
    assert np.isfinite(dg.summary_row(np.random.default_rng(3).standard_normal((2, 64)))).all()


def test_autocov_matches_direct_sum():
    rng = np.random.default_rng(4)
    x = rng.standard_normal(257)
    a = dg.autocov(x)
    xc = x - x.mean()
    for lag in (0, 1, 7, 100):
        assert abs(a[lag] - (xc[: x.size - lag] * xc[lag:]).sum() / x.size) < 1e-12


def test_tfp_cross_chain_ess_ar1():
    """TFP-style cross-chain ESS (main_script.py:807-810) on AR(1) chains: N C (1-phi)/(1+phi) within MC error;
    single-chain fallback; lag truncation at the first negative autocorrelation (direct-sum check)."""
    rng = np.random.default_rng(3)
    for phi in (0.5, 0.9):
        x = _ar1(rng, 4, 20000, phi)
        got = dg.tfp_ess_cross_chain(x)
        want = 4 * 20000 * (1 - phi) / (1 + phi)
        assert abs(got / want - 1) < 0.08, (phi, got, want)
        one = dg.tfp_ess_cross_chain(x[:1])
        assert abs(one / (want / 4) - 1) < 0.15
    # direct restatement with explicit loops on a short series
    x = _ar1(rng, 3, 300, 0.7)
    C, N = x.shape
    xc = x - x.mean(axis=1, keepdims=True)
    w = (xc ** 2).mean(axis=1).mean()
    b = x.mean(axis=1).var(ddof=1)
    tot = 0.0
    for k in range(N):
        acov = np.mean([np.dot(xc[c, :N - k], xc[c, k:]) / (N - k) for c in range(C)])
        rho = 1 - (w - acov) / (w + b)
        if rho < 0:
            break
        tot += (N - k) / N * rho
    assert abs(dg.tfp_ess_cross_chain(x) / (C * N / (-1 + 2 * tot)) - 1) < 1e-10
    # chains that disagree: between-chain variance inflates the autocorrelation -> much smaller ESS
    y = x + np.arange(3)[:, None] * 5.0
    assert dg.tfp_ess_cross_chain(y) < 0.1 * dg.tfp_ess_cross_chain(x)
