"""Host-side logic of the drop-in (no GPU): file naming / discovery / skip rule of mcmc.py,
summary CSV layout, sharding, and the 2-rank gloo all-gather of summaries."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mcmc_module_keeps_reference_names():
    from pet_posterior_distribution_b200 import mcmc
    for name in ("n_ROI_test", "n_samples_test", "mean_sigma_noise_load", "iter_mcmc", "burn_mcmc", "chains",
                 "CreateTAC_SRTM2", "NP_DTYPE", "FLAG_PLOT", "CUR_DIR"):
        assert hasattr(mcmc, name), name
    assert (mcmc.n_ROI_test, mcmc.n_samples_test, mcmc.iter_mcmc, mcmc.burn_mcmc, mcmc.chains) == (48, 100, 200, 400, 4)


def test_generator_module_keeps_reference_names():
    """sample_sim_data.py:88-95: same configuration names and shipped defaults (the training set)."""
    from pet_posterior_distribution_b200 import sample_sim_data as gen
    assert (gen.n_samples, gen.n_ROI, gen.flag_testing_data, gen.mean_sigma_noise_save, gen.alpha) == (100000, 48, False, 0.1, 0.8)
    from pet_posterior_distribution_b200.frames import MK_HALF_T
    assert MK_HALF_T == 109.8


def test_kinetic_model_module_keeps_reference_names():
    """kinetic_model.py's public surface: two module-level functions, SRTM and SRTM2 with their static helpers."""
    from pet_posterior_distribution_b200 import kinetic_model as km
    for name in ("estimate_continuous_convolution", "interp1d_linear_vec", "SRTM", "SRTM2"):
        assert hasattr(km, name), name
    for cls in (km.SRTM, km.SRTM2):
        for name in ("make_time_func", "make_time_exponential", "convolve", "__call__"):
            assert callable(getattr(cls, name)), (cls.__name__, name)
    assert hasattr(km.SRTM, "forward_model") and hasattr(km.SRTM2, "create_activity_curve")
    # make_time_func is host-side broadcasting around the caller's function (kinetic_model.py:89-116)
    t, p = np.linspace(1, 5, 4), np.array([[0.1, 0.2], [0.3, 0.4]])
    out = km.SRTM.make_time_func(p, t, lambda x, tt: x * tt, time_scale=np.arange(4.0) + 1, space_scale=np.full((2, 2), 2.0))
    assert out.shape == (4, 2, 2) and np.allclose(out, p[None] * t[:, None, None] * (np.arange(4.0) + 1)[:, None, None] * 2.0)
    assert np.allclose(km.SRTM2.make_time_func(0.5, t, lambda x, tt: x * tt, time_scale=3.0), 1.5 * t)


def test_save_name_matches_reference_pattern():
    """mcmc.py:119-123 and the shipped file name MH_MCMC_nROI48_it2.0e+04_brn4.0e+04_km_obs-0.842-0.833-0.013.pik."""
    from pet_posterior_distribution_b200 import mcmc
    old = (mcmc.iter_mcmc, mcmc.burn_mcmc)
    try:
        mcmc.iter_mcmc, mcmc.burn_mcmc = 20000, 40000
        name = mcmc.save_name({"DVR": np.array([0.8424]), "R1": np.array([0.8331]), "k2p": np.array([0.0126])})
        assert name == "MH_MCMC_nROI48_it2.0e+04_brn4.0e+04_km_obs-0.842-0.833-0.013.pik"
    finally:
        mcmc.iter_mcmc, mcmc.burn_mcmc = old


def test_find_test_file_takes_latest(tmp_path):
    from pet_posterior_distribution_b200 import mcmc
    for d in ("25-01-01_00-00-00_test", "25-07-10_15-38-57_test", "25-09-09_00-00-00_train"):
        p = tmp_path / "sim_data" / "nROI48" / d
        p.mkdir(parents=True)
        (p / "data_nROI48_n100_s1.0e-01.pik").write_bytes(b"x")
    d, f = mcmc.find_test_file(str(tmp_path / "sim_data"))
    assert d.endswith("25-07-10_15-38-57_test") and f == "data_nROI48_n100_s1.0e-01.pik"
    with pytest.raises(IndexError):
        mcmc.find_test_file(str(tmp_path / "nowhere"))


def test_summary_csv_layout():
    from pet_posterior_distribution_b200 import diagnostics
    rng = np.random.default_rng(0)
    dvr = 1 + 0.01 * rng.standard_normal((4, 50, 48))
    r1 = 0.8 + 0.01 * rng.standard_normal((4, 50, 48))
    g = np.zeros((96, 8), np.float32)
    g[:, 0] = np.concatenate([dvr.mean((0, 1)), r1.mean((0, 1))]); g[:, 1] = 0.01; g[:, 2] = 1e-3
    g[:, 3] = 180.4; g[:, 4] = 150.6; g[:, 5] = 1.004
    txt = diagnostics.summary_csv(dvr, r1, [0.0126], g)
    lines = txt.strip().split("\n")
    assert lines[0] == ",mean,sd,hdi_3%,hdi_97%,mcse_mean,mcse_sd,ess_bulk,ess_tail,r_hat"
    assert len(lines) == 1 + 97 and lines[1].startswith("var_DVR[0],") and lines[49].startswith("var_R1[0],")
    assert lines[97] == "var_k2p,0.013,0.0,0.013,0.013,0.0,0.0,200.0,200.0,"      # a constant Deterministic in az.summary
    f = lines[1].split(",")
    assert f[2] == "0.01" and f[7] == "180.0" and f[8] == "151.0" and f[9] == "1.0"   # pandas' float formatting, ArviZ's rounding
    import io
    import pandas as pd
    df = pd.read_csv(io.StringIO(txt), index_col=0)                               # what a consumer of _summary.csv does
    assert list(df.columns) == list(diagnostics.COLUMNS) and df.shape == (97, 9) and df.loc["var_R1[3]", "r_hat"] == 1.0
    # the pandas-free fallback writes the same text
    import builtins, sys
    real_import = builtins.__import__
    def no_pandas(name, *a, **k):
        if name == "pandas":
            raise ImportError(name)
        return real_import(name, *a, **k)
    builtins.__import__ = no_pandas
    try:
        assert diagnostics.summary_csv(dvr, r1, [0.0126], g) == txt
    finally:
        builtins.__import__ = real_import
    lo, hi = diagnostics.hdi(rng.standard_normal(100000))
    assert abs(lo + 1.88) < 0.05 and abs(hi - 1.88) < 0.05


def test_rhat_log(tmp_path):
    from pet_posterior_distribution_b200 import diagnostics
    g = np.ones((96, 8), np.float32)
    g[:, 5] = 1.01
    assert diagnostics.append_rhat_log(str(tmp_path), "f.pik", 3, g) < 1.02
    assert not (tmp_path / "rhat_less_than_102.txt").exists()
    g[5, 5] = 1.0345
    diagnostics.append_rhat_log(str(tmp_path), "f.pik", 3, g)
    assert (tmp_path / "rhat_less_than_102.txt").read_text() == "f.pik - sample 3 - rhat_max = 1.0345\n"


def test_idata_stand_in():
    from pet_posterior_distribution_b200 import diagnostics
    idata = diagnostics.make_idata(np.zeros((4, 10, 48)), np.ones((4, 10, 48)), [0.0126])
    assert np.asarray(idata.posterior["var_DVR"]).shape == (4, 10, 48)       # mcmc.py:162
    assert np.asarray(idata.posterior["var_k2p"]).shape == (4, 10)
    pickle.loads(pickle.dumps(idata))


def test_plots_write_the_reference_file_names(tmp_path):
    """mcmc.py:198-258 (FLAG_PLOT): four PNGs per sample, named like the reference's."""
    pytest.importorskip("matplotlib")
    from pet_posterior_distribution_b200 import plots
    rng = np.random.default_rng(0)
    ch = {"DVR": 1 + 0.01 * rng.standard_normal((2, 50, 48)), "R1": 0.8 + 0.01 * rng.standard_normal((2, 50, 48))}
    km = {"DVR": np.ones(48), "R1": 0.8 * np.ones(48), "k2p": np.array([0.0126])}
    out = plots.plot_sample(str(tmp_path / "MH_x.pik"), ch, km, {"DVR": np.ones(48), "R1": np.ones(48)}, roi_plot=3)
    assert [os.path.basename(f) for f in out] == ["MH_x_DVR_ROI3.png", "MH_x_DVR_ROI_all.png", "MH_x_R1_ROI3.png", "MH_x_R1_ROI_all.png"]
    assert all(os.path.getsize(f) > 1000 for f in out)


def test_shard_bounds_cover_everything():
    from pet_posterior_distribution_b200.distributed import shard_bounds, shard_sizes
    for n in (0, 1, 7, 100, 1048576):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(shard_sizes(n, w)) - min(shard_sizes(n, w)) <= 1


def test_chain_shards_cover_every_chain_once():
    """S < world: ranks r % S == t form TAC t's group and split its chains contiguously."""
    from pet_posterior_distribution_b200.distributed import chain_shards
    for S, C, W in ((1, 64, 8), (3, 1024, 8), (1, 4, 8), (2, 5, 3), (7, 16, 8)):
        plan = chain_shards(S, C, W)
        assert len(plan) == W
        for t in range(S):
            grp = [r for r in range(W) if plan[r]["tac"] == t]
            assert grp == plan[grp[0]]["group"] and all(plan[r]["owner"] == grp[0] for r in grp)
            cover = []
            for r in grp:
                cover += list(range(plan[r]["c_lo"], plan[r]["c_hi"]))
            assert cover == list(range(C))
            sizes = [plan[r]["c_hi"] - plan[r]["c_lo"] for r in grp]
            assert max(sizes) - min(sizes) <= 1


_GLOO = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
from pet_posterior_distribution_b200.distributed import shard_bounds, gather_summaries, gather_padded, chain_shards
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
n = 7                                     # 7 TACs over 2 ranks: ragged shards (4 + 3)
lo, hi = shard_bounds(n, world, rank)
full = torch.arange(n * 96 * 8, dtype=torch.float32).view(n, 96, 8)
got = gather_summaries(full[lo:hi].clone(), n)
assert torch.equal(got, full), "rank %%d: gathered summaries differ" %% rank
# chain-sharded regime: 1 TAC x 5 chains over 2 ranks (3 + 2): the owner reassembles every chain's draws in order
plan = chain_shards(1, 5, world)
me = plan[rank]
draws = torch.arange(5 * 4 * 96, dtype=torch.float32).view(5, 4, 96)
parts = gather_padded(draws[me["c_lo"]:me["c_hi"]].clone(), [p["c_hi"] - p["c_lo"] for p in plan])
assert torch.equal(torch.cat([parts[r] for r in me["group"]]), draws), "rank %%d: gathered draws differ" %% rank
dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_gloo_gather(tmp_path):
    """N > 1 path on CPU: world_size 2, gloo, ragged TAC shards, all-gather of summaries."""
    script = tmp_path / "g.py"
    script.write_text(_GLOO % ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_model_objects_pickle_without_their_handle(monkeypatch):
    """SURVEY.md 8 b: the reference's Op and its SRTM2 are pickled to PyMC's worker processes.  The mirrors pickle as their
    arrays + device and rebuild the library handle on load (the handle itself, a ctypes pointer, cannot travel)."""
    from pet_posterior_distribution_b200 import kinetic_model as km
    from pet_posterior_distribution_b200 import mcmc
    made = []

    class FakeSampler:
        def __init__(self, n_chains=4, max_tacs=1, max_draws=0, seed=0, device=0, tac_gid0=0):
            self.device = device
            made.append(self)

        def set_frames(self, t, dt):
            self.frames = (np.array(t), np.array(dt))

        def set_prior(self, *a):
            pass

    monkeypatch.setattr(km, "MHSampler", FakeSampler)
    t, dt, cr = np.linspace(1, 120, 54), np.full(54, 2.0), np.linspace(0.1, 5, 54)
    m = km.SRTM2(frame_time_list=t, frame_duration_list=dt, tac_reference=cr, device=3)
    op = mcmc.CreateTAC_SRTM2(m)
    op2 = pickle.loads(pickle.dumps(op))
    m2 = op2.k_srtm
    assert isinstance(m2, km.SRTM2) and m2 is not m and len(made) == 2 and made[1].device == 3
    assert np.array_equal(m2._tac_reference, cr) and np.array_equal(m2._frame_time_list, t) and m2._k2p is None
    assert np.array_equal(made[1].frames[0], t) and np.array_equal(made[1].frames[1], dt)
    s2 = pickle.loads(pickle.dumps(km.SRTM(frame_time_list=t, frame_duration_list=dt, device=1)))
    assert isinstance(s2, km.SRTM) and made[-1].device == 1 and np.array_equal(s2._frame_duration_list, dt)
