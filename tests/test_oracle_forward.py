"""Oracle forward model vs the golden vectors of the LIVE reference kinetic_model.py
(tools/make_golden.py) -- this is what pins the oracle for the forward-model rows."""
import os
import re

import numpy as np

from oracle import forward, frames

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_grid_matches_golden(forward_golden):
    t, dt = frames.frame_grid()
    assert np.array_equal(t, forward_golden["t"]) and np.array_equal(dt, forward_golden["dt"])
    assert t.size == 54 and abs(t[-1] - 120.0) < 1e-12 and abs(dt.sum() - 120.0) < 1e-9


def test_tac_matches_reference(forward_golden):
    g = forward_golden
    for c in range(g["c_r"].shape[0]):
        out = forward.srtm2_tac(g["t"], g["c_r"][c], g["DVR"][c], g["R1"][c], g["k2p"][c])
        assert np.abs(out - g["tac"][c]).max() <= 1e-12 * np.abs(g["tac"][c]).max()


def test_srtm_k2_free_matches_reference(forward_golden):
    g = forward_golden
    for c in range(g["c_r"].shape[0]):
        out = forward.srtm_tac(g["t"], g["c_r"][c], g["DVR"][c], g["k2"][c], g["R1"][c])
        assert np.abs(out - g["tac_srtm"][c]).max() <= 1e-12 * np.abs(g["tac_srtm"][c]).max()


def test_operator_matches_reference(forward_golden):
    g = forward_golden
    for c in range(g["c_r"].shape[0]):
        M = forward.build_M(g["t"], g["c_r"][c])
        assert np.abs(M - g["M"][c]).max() <= 1e-13 * np.abs(g["M"][c]).max()
        out = forward.srtm2_tac_M(g["t"], g["c_r"][c], M, g["DVR"][c], g["R1"][c], g["k2p"][c])
        assert np.abs(out - g["tac"][c]).max() <= 1e-12 * np.abs(g["tac"][c]).max()


def test_operator_is_linear_in_reference_tac(forward_golden):
    g = forward_golden
    a, b = g["c_r"][0], g["c_r"][3]
    Ma, Mb, Mab = forward.build_M(g["t"], a), forward.build_M(g["t"], b), forward.build_M(g["t"], 2 * a + 0.5 * b)
    assert np.abs(Mab - (2 * Ma + 0.5 * Mb)).max() < 1e-12 * np.abs(Mab).max()


def test_scalar_and_1d_paths(forward_golden):
    g = forward_golden
    full = forward.srtm2_tac(g["t"], g["c_r"][1], g["DVR"][1], g["R1"][1], g["k2p"][1])
    one = forward.srtm2_tac(g["t"], g["c_r"][1], g["DVR"][1][5], g["R1"][1][5], g["k2p"][1])
    assert np.abs(one[:, 0] - full[:, 5]).max() < 1e-13


def test_interp_weights_edge_cases():
    xp = np.array([0.0, 1.0, 3.0])
    W = forward.interp_weights(np.array([0.0, 0.5, 1.0, 2.5, 3.0]), xp)
    assert np.allclose(W, [[1, 0, 0], [.5, .5, 0], [0, 1, 0], [0, .25, .75], [0, 0, 1]])


def test_compiled_schedule_matches_oracle_pattern():
    """tools/gen_schedule.py derives the sparsity independently; it must equal the oracle's."""
    inc = open(os.path.join(ROOT, "pet_posterior_distribution_b200", "csrc", "m_schedule.inc")).read()
    get = lambda name: [int(v) for v in re.search(r"#define %s \{([^}]*)\}" % name, inc).group(1).split(",")]
    act, nrow = forward.active_columns(frames.frame_grid()[0])
    assert get("PETMH_ACTIVE_COLS") == act.tolist()
    assert get("PETMH_NROW_PREFIX") == nrow.tolist()
    assert nrow.sum() == 1122 and act.size == 45
    # packed layout: every non-zero of M appears exactly once
    src = np.array(get("PETMH_MPACK_SRC"))
    used = src[src >= 0]
    assert used.size == 1122 and np.unique(used).size == 1122
    rows, cols = used >> 6, used & 63
    assert all(cols[i] < nrow[rows[i]] for i in range(used.size))


import os as _os
import pytest as _pytest


@_pytest.mark.skipif(not _os.path.isfile("/root/reference/kinetic_model.py"), reason="the live reference exists in the build container only")
def test_golden_regenerates_from_the_live_reference(forward_golden):
    """Provenance: the LIVE /root/reference kinetic_model.py, called on the fixture's inputs, returns the fixture's outputs bit
    for bit (SRTM2.create_activity_curve, estimate_continuous_convolution, SRTM.forward_model).  Child process: keeps the
    reference's module names out of this one."""
    import subprocess
    import sys
    path = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden", "forward_golden.npz")
    code = r"""
import sys, numpy as np
sys.path.insert(0, "/root/reference")
import kinetic_model as km
g = np.load(%r)
t, dt = g["t"], g["dt"]
for c in range(g["c_r"].shape[0]):
    m = km.SRTM2(frame_time_list=t, frame_duration_list=dt, tac_reference=g["c_r"][c])
    assert np.array_equal(m.create_activity_curve(DVR=g["DVR"][c], R1=g["R1"][c], k2p=g["k2p"][c]), g["tac"][c]), c
    assert np.array_equal(km.estimate_continuous_convolution(t, g["c_r"][c], np.eye(54)), g["M"][c]), c
    s = km.SRTM(frame_time_list=t, frame_duration_list=dt)
    assert np.array_equal(s.forward_model(DVR=g["DVR"][c], k2=g["k2"][c], R1=g["R1"][c], tac_ref=g["c_r"][c]), g["tac_srtm"][c]), c
print("ok")
""" % path
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (r.stdout[-300:], r.stderr[-800:])
