"""The C oracle (oracle/c/mh_oracle.c) is the numpy oracle, faster: identical trajectories on the same tape."""
import numpy as np

from oracle import cmh, mh


def test_c_oracle_equals_numpy_oracle(models):
    cm = cmh.CModel(models[1])
    tape = mh.Tape.random(230, np.random.default_rng(21))
    a = mh.run_chain(models[1], tape, 200, 30)
    b = cm.run_taped(tape, 200, 30)
    assert np.array_equal(a["draws"], b["draws"]) and np.array_equal(a["accept"], b["accept"])
    assert np.array_equal(a["scale"], b["scale"])
    fin = np.isfinite(a["delta"]) & (np.abs(a["delta"]) < 1e3)
    assert np.array_equal(np.isfinite(a["delta"]), np.isfinite(b["delta"]))
    assert np.abs(a["delta"][fin] - b["delta"][fin]).max() < 1e-9
    for roi, dvr, r1 in ((0, 0.9, 0.8), (17, 1.3, 1.1), (47, 0.7, 0.6)):
        assert abs(cm.ll_roi(roi, dvr, r1) - models[1].ll_roi(roi, dvr, r1)) < 1e-10 * abs(models[1].ll_roi(roi, dvr, r1))


def test_c_oracle_forced_mode_matches_numpy(models):
    cm = cmh.CModel(models[0])
    tape = mh.Tape.random(60, np.random.default_rng(22))
    free = mh.run_chain(models[0], tape, 60, 0)
    forced_np = mh.run_chain(models[0], tape, 60, 0, forced_draws=free["draws"])
    forced_c = cm.run_forced(tape, 60, 0, free["draws"])
    for k in ("accept", "forced_accept", "undecidable"):
        assert np.array_equal(forced_np[k], forced_c[k]), k
    dec = ~forced_c["undecidable"]
    assert np.array_equal(forced_c["accept"][dec], forced_c["forced_accept"][dec])


def test_c_oracle_free_chains_are_sane(models):
    cm = cmh.CModel(models[2])
    draws, nacc = cm.run_free(4, 300, 100, seed=5, threads=4)
    assert draws.shape == (4, 400, 2, 48) and np.isfinite(draws).all() and nacc > 0
    assert not np.array_equal(draws[0], draws[1])
    assert np.abs(draws[:, 300:].mean(axis=(0, 1)) - np.stack(models[2].mu)).max() < 1.5   # stays near the prior scale
