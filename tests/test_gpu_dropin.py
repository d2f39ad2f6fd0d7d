"""The drop-in entry points on a real GPU: mcmc.main() writes the reference's three outputs;
kinetic_model.SRTM2 / CreateTAC_SRTM2 mirror the operator seam; product-side generator."""
import os
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _write_test_set(root, dataset, prior, n=100):
    d = os.path.join(root, "sim_data", "nROI48", "26-01-01_00-00-00_test")
    os.makedirs(d)
    k = dataset["varDVR"].shape[0]
    idx = np.arange(n) % k
    ds = {"varDVR": list(dataset["varDVR"][idx]), "varR1": list(dataset["varR1"][idx]),
          "vark2p": [dataset["vark2p"][i] for i in idx], "vartacref": list(dataset["vartacref"][idx]),
          "tac_sampled": list(dataset["tac_sampled"][idx]), "tac_noisy_sampled": list(dataset["tac_noisy_sampled"][idx]),
          "mu_noise": dataset["mu_noise"], "sigma_noise": dataset["sigma_noise"], "mean_sigma_noise": 0.1,
          "flag_mahalanobis": True, "target_ROI_names": None, "time_vector": dataset["time_vector"], "dt": dataset["dt"]}
    pickle.dump(ds, open(os.path.join(d, "data_nROI48_n100_s1.0e-01.pik"), "wb"))
    pickle.dump({k2: prior[k2] for k2 in prior}, open(os.path.join(root, "prior_stats_nROI48.pik"), "wb"))
    return d


def test_mcmc_main_writes_reference_outputs(tmp_path, dataset, prior):
    from pet_posterior_distribution_b200 import mcmc
    d = _write_test_set(str(tmp_path), dataset, prior)
    mcmc.sample_range = range(0, 3)
    mcmc.iter_mcmc, mcmc.burn_mcmc, mcmc.chains = 60, 300, 4
    written = mcmc.main(data_dir=os.path.join(str(tmp_path), "sim_data"),
                        prior_path=os.path.join(str(tmp_path), "prior_stats_nROI48.pik"))
    assert len(written) == 3
    out = pickle.load(open(written[0], "rb"))
    assert set(out) == {"idata", "DVR_mcmc", "k2p_mcmc", "R1_mcmc", "iter", "burn", "y_obs", "km_obs", "chains", "elapsed_time"}
    assert out["DVR_mcmc"].shape == (4, 60, 48) and out["DVR_mcmc"].dtype == np.float64
    assert out["k2p_mcmc"].shape == (4, 60) and out["y_obs"].shape == (48, 54)
    assert os.path.basename(written[0]).startswith("MH_MCMC_nROI48_it6.0e+01_brn3.0e+02_km_obs-")
    csv = open(written[0].replace(".pik", "_summary.csv")).read().strip().split("\n")
    assert len(csv) == 98 and csv[1].startswith("var_DVR[0],")
    assert np.asarray(out["idata"].posterior["var_R1"]).shape == (4, 60, 48)
    # second call: everything exists -> skipped (mcmc.py:125-128)
    assert mcmc.main(data_dir=os.path.join(str(tmp_path), "sim_data"),
                     prior_path=os.path.join(str(tmp_path), "prior_stats_nROI48.pik")) == []
    assert os.path.isdir(os.path.join(d, "MCMC_s1.0e-01"))


def test_operator_seam(forward_golden):
    from pet_posterior_distribution_b200 import kinetic_model, mcmc
    g = forward_golden
    k = kinetic_model.SRTM2(frame_time_list=g["t"], frame_duration_list=g["dt"], tac_reference=g["c_r"][1])
    out = k.create_activity_curve(DVR=g["DVR"][1], R1=g["R1"][1], k2p=g["k2p"][1])
    assert out.shape == (54, 48) and np.abs(out / g["tac"][1] - 1).max() < 1e-5
    op = mcmc.CreateTAC_SRTM2(k)
    cell = [[None]]
    op.perform(None, [g["DVR"][1], g["R1"][1], g["k2p"][1]], cell)
    assert cell[0][0].shape == (48, 54) and cell[0][0].dtype == np.float64
    one = k.create_activity_curve(DVR=float(g["DVR"][1][3]), R1=float(g["R1"][1][3]), k2p=g["k2p"][1])
    assert one.shape == (54,) and np.abs(one / g["tac"][1][:, 3] - 1).max() < 1e-5


def test_srtm_k2_free_matches_reference_golden(forward_golden):
    """the other model of kinetic_model.py (SRTM, k2 free) vs the live reference's golden TACs."""
    from pet_posterior_distribution_b200 import kinetic_model
    g = forward_golden
    k = kinetic_model.SRTM(frame_time_list=g["t"], frame_duration_list=g["dt"])
    for c in range(3):
        out = k.forward_model(DVR=g["DVR"][c], k2=g["k2"][c], R1=g["R1"][c], tac_ref=g["c_r"][c])
        assert out.shape == (54, 48) and np.abs(out / g["tac_srtm"][c] - 1).max() < 1e-5


def test_product_generator_schema(prior):
    from pet_posterior_distribution_b200 import sample_sim_data as gen
    ds = gen.generate(prior, 6, 0.1, test_style=True, seed=5)
    assert len(ds["varDVR"]) == 6 and ds["tac_noisy_sampled"][0].shape == (48, 54)
    assert all((np.asarray(v) >= 0).all() for v in ds["tac_noisy_sampled"])
    from oracle import forward
    ref = forward.srtm2_tac(ds["time_vector"], ds["vartacref"][2], ds["varDVR"][2], ds["varR1"][2], float(prior["mu_k2p"])).T
    assert np.abs(ds["tac_sampled"][2] / ds["dt"][None, :] / ref - 1).max() < 1e-5


def test_generator_main_writes_reference_layout(tmp_path, monkeypatch, prior):
    """sample_sim_data.py as a script (:96-240): sim_data/nROI48/<ts>_{train,test}/data_*.pik with the reference's schema
    and args_*.txt, generated on the GPU; the test set obeys the Mahalanobis rule and feeds mcmc.find_test_file."""
    import json
    from scipy import stats
    from pet_posterior_distribution_b200 import mcmc
    from pet_posterior_distribution_b200 import sample_sim_data as gen
    monkeypatch.chdir(tmp_path)
    pickle.dump({k: prior[k] for k in prior}, open("prior_stats_nROI48.pik", "wb"))
    keys = {"varDVR", "varR1", "vark2p", "vartacref", "tac_sampled", "tac_noisy_sampled", "mu_noise", "sigma_noise",
            "mean_sigma_noise", "flag_mahalanobis", "target_ROI_names", "time_vector", "dt"}
    for flag, n in ((False, 24), (True, 100)):
        monkeypatch.setattr(gen, "n_samples", n)
        monkeypatch.setattr(gen, "flag_testing_data", flag)
        d = gen.main(seed=11)
        assert os.path.basename(d).endswith("_test" if flag else "_train") and os.path.dirname(d).endswith(os.path.join("sim_data", "nROI48"))
        ds = pickle.load(open(os.path.join(d, "data_nROI48_n%d_s1.0e-01.pik" % n), "rb"))
        assert set(ds) == keys and len(ds["varDVR"]) == n and ds["flag_mahalanobis"] is flag
        assert ds["tac_noisy_sampled"][0].shape == (48, 54) and ds["vartacref"][0].shape == (54,) and ds["sigma_noise"].shape == (48, 54)
        assert all((np.asarray(ds[k]) >= 0).all() for k in ("varDVR", "varR1", "vartacref", "tac_sampled", "tac_noisy_sampled"))
        a = json.load(open(os.path.join(d, "args_nROI48_n%d_s1.0e-01.txt" % n)))
        assert set(a) == {"mean_sigma_noise", "target_ROI_names", "MK_half_T", "MK_lambda", "n_samples", "n_ROI", "save_samples_dir",
                          "flag_mahalanobis"} and a["n_samples"] == n and a["flag_mahalanobis"] is flag
        if flag:
            dd = np.asarray(ds["varDVR"]) - prior["mu_DVR"]
            d2 = np.einsum("ni,ij,nj->n", dd, np.linalg.inv(prior["Cov_DVR"]), dd)
            assert (d2 < stats.chi2.ppf(0.8, 48) + 0.1).all()
            found_dir, found_name = mcmc.find_test_file(os.path.join(str(tmp_path), "sim_data"))
            assert os.path.samefile(found_dir, d) and found_name == "data_nROI48_n100_s1.0e-01.pik"


def test_frame_grid_rejected_when_pattern_differs(prior):
    from pet_posterior_distribution_b200 import MHSampler, PetmhError
    s = MHSampler()
    t = np.linspace(1, 120, 54)
    with pytest.raises(PetmhError) as e:
        s.set_frames(t, np.full(54, 120 / 54))
    assert e.value.code == -5
