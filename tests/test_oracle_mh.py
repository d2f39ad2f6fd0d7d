"""Restated pymc element-wise Metropolis (oracle/mh.py) -- internal consistency."""
import numpy as np

from oracle import mh, philox


def test_tune_table():
    """pymc.step_methods.metropolis.tune thresholds on the accept count of 100 steps."""
    exp = {0: 0.1, 1: 0.5, 4: 0.5, 5: 0.9, 19: 0.9, 20: 1.0, 50: 1.0, 51: 1.1, 75: 1.1, 76: 2.0, 95: 2.0, 96: 10.0, 100: 10.0}
    for c, f in exp.items():
        assert mh.tune_factor(c) == f, c


def test_lean_equals_reference_faithful(models):
    """A.3 incremental form == two full-model log-probs per step (what pymc evaluates)."""
    tape = mh.Tape.random(4, np.random.default_rng(3))
    a = mh.run_chain(models[0], tape, 4, 0, mode="lean")
    b = mh.run_chain(models[0], tape, 4, 0, mode="faithful")
    assert np.array_equal(a["accept"], b["accept"]) and np.array_equal(a["draws"], b["draws"])
    fin = np.isfinite(b["delta"])
    assert np.array_equal(fin, np.isfinite(a["delta"]))
    big = np.abs(b["delta"][fin]) < 1e6
    assert np.abs(a["delta"][fin][big] - b["delta"][fin][big]).max() < 1e-7 * (1 + np.abs(b["delta"][fin][big]).max())


def test_deterministic_given_tape_and_tuning(models):
    tape = mh.Tape.random(230, np.random.default_rng(4))
    a = mh.run_chain(models[2], tape, 200, 30)
    b = mh.run_chain(models[2], tape, 200, 30)
    assert np.array_equal(a["draws"], b["draws"])
    assert (a["scale"] != 1).all()                       # two tune events happened (sweeps 100 and 200 > n_tune? only 100)
    assert a["draws"].dtype == np.float32
    # start at the prior mean (pymc initial point), first sweep moves at most by the proposal
    assert np.isfinite(a["draws"]).all()
    # visit order is a permutation each (sweep, block)
    assert (np.sort(tape.rank, axis=-1) == np.arange(48)).all()


def test_teacher_forcing_self_consistency(models):
    tape = mh.Tape.random(60, np.random.default_rng(5))
    free = mh.run_chain(models[1], tape, 60, 0)
    forced = mh.run_chain(models[1], tape, 60, 0, forced_draws=free["draws"])
    dec = ~forced["undecidable"]
    assert np.array_equal(forced["accept"][dec], forced["forced_accept"][dec])
    assert np.array_equal(forced["draws"], free["draws"])


def test_nonfinite_proposals_are_rejected(models):
    """scale 1.0 at the start proposes DVR <= 0 etc.; isfinite guard rejects (metrop_select)."""
    tape = mh.Tape.random(3, np.random.default_rng(6))
    out = mh.run_chain(models[0], tape, 3, 0)
    bad = ~np.isfinite(out["delta"])
    assert not out["accept"][bad].any()


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for c, k, out in kat:
        got = philox.philox4x32_10(np.array(c, np.uint32), np.array(k, np.uint32))
        assert got.tolist() == out


def test_philox_draw_layout():
    n, lu, rk = philox.draws(seed=42, chain_gid=2 ** 33 + 5, sweep=7, block=1)
    assert n.shape == (48,) and sorted(rk.tolist()) == list(range(48)) and (lu <= 0).all()
    a = philox.raw_draws(42, 1, 0, 0); b = philox.raw_draws(42, 1, 0, 1); c = philox.raw_draws(42, 2, 0, 0)
    assert not np.array_equal(a, b) and not np.array_equal(a, c)
    big = np.concatenate([philox.draws(1, g, s, 0)[0] for g in range(20) for s in range(20)])
    assert abs(big.mean()) < 0.05 and abs(big.std() - 1) < 0.05
