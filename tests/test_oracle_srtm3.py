"""oracle/srtm3.py (checker of the sampled k2-free SRTM): its forward model is the live reference's SRTM.forward_model
(golden vectors), and with k2 = k2p R1 it reduces to the SRTM2 oracle."""
import numpy as np

from oracle import srtm3


def test_forward_matches_reference_srtm_golden(forward_golden, prior):
    from oracle.logp import Model
    g = forward_golden
    for c in range(g["c_r"].shape[0]):
        m = Model(g["t"], g["c_r"][c], g["k2p"][c], np.ones((48, 54)), np.ones((48, 54)), prior["mu_DVR"], prior["Cov_DVR"],
                  prior["mu_R1"], prior["Cov_R1"])
        m3 = srtm3.Model3(m, np.full(48, 0.01), np.eye(48))
        ref = g["tac_srtm"][c]                                   # (54, 48) = SRTM.forward_model(DVR, k2, R1, c_r)
        got = np.stack([m3.tac_roi(g["DVR"][c][i], g["R1"][c][i], g["k2"][c][i]) for i in range(48)], axis=1)
        assert np.abs(got - ref).max() <= 1e-10 * np.abs(ref).max()


def test_reduces_to_srtm2_when_k2_is_k2p_r1(models):
    m = models[0]
    m3 = srtm3.Model3(m, np.full(48, 0.01), np.eye(48))
    rng = np.random.default_rng(0)
    for _ in range(5):
        i = int(rng.integers(48))
        dvr, r1 = m.mu[0][i] * (1 + 0.05 * rng.standard_normal()), m.mu[1][i] * (1 + 0.05 * rng.standard_normal())
        assert abs(m3.ll_roi(i, dvr, r1, m.k2p * r1) - m.ll_roi(i, dvr, r1)) < 1e-9
