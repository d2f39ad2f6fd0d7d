#!/usr/bin/env python
"""Run the LIVE reference generator script (/root/reference/sample_sim_data.py, unmodified source, exec'd from where it
lies) in the BUILD container and commit what it produces as small fixtures: the pin of oracle/generator.py (SURVEY.md
section 8 rows a11-a13) against outputs of the reference itself.

The script is a flat module that runs at import (no functions), configured by its module constants
(sample_sim_data.py:88-95).  It is executed with
  * its four configuration assignments rewritten IN MEMORY (n_samples, flag_testing_data, mean_sigma_noise_save,
    FLAG_PLOT = False) -- the same edit a user of the reference makes by hand;
  * the working directory set to a scratch directory that holds a link to prior_stats_nROI48.pik (the script reads
    ./prior_stats_nROI48.pik and writes ./sim_data/...);
  * environment shims only: `matplotlib` (not installed here; only used behind FLAG_PLOT) and `diffusion_model`
    (helper_func.py:7 imports NP_DTYPE from it; it needs tensorflow) are stub modules, and `np.Inf` (removed in numpy 2,
    used by helper_func.py:148) is aliased to `np.inf`;
  * `np.random.seed(seed)`: the reference draws from numpy's global RandomState.

  reference_generated_test_s0.1.npz   the pickle of a test-style run (Mahalanobis rule on), 6 samples, every key
  reference_generated_train_s0.1.npz  the pickle of a training-style run, 6 samples
  reference_generated_stats.npz       moments of larger runs (training-style 3000, test-style 1200 samples): per-coordinate
                                      mean / sd of the accepted DVR, R1, reference TAC draws, the per-ROI noise scale, and the
                                      moments / quantiles of the standardised noise the script added

Run:  python tools/make_reference_generated.py      (the GPU box never needs /root/reference)
"""
import glob
import os
import pickle
import sys
import tempfile
import types
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def run_reference_script(n_samples, test_style, noise, seed):
    """exec /root/reference/sample_sim_data.py with its constants set; returns the dict it pickled."""
    src = open(os.path.join(REF, "sample_sim_data.py")).read()
    edits = {"n_samples = 100000": "n_samples = %d" % n_samples,
             "flag_testing_data = False": "flag_testing_data = %s" % bool(test_style),
             "mean_sigma_noise_save = 1e-1": "mean_sigma_noise_save = %r" % float(noise),
             "FLAG_PLOT = True": "FLAG_PLOT = False"}
    for old, new in edits.items():
        assert src.count(old) == 1, old
        src = src.replace(old, new)
    if not hasattr(np, "Inf"):
        np.Inf = np.inf                                   # numpy >= 2 (the reference pins numpy < 1.28)
    mpl = types.ModuleType("matplotlib"); mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot"); mpl.pyplot = plt
    dm = types.ModuleType("diffusion_model"); dm.NP_DTYPE = np.float32      # diffusion_model.py:7,10
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "diffusion_model": dm})
    if REF not in sys.path:
        sys.path.insert(0, REF)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.symlink(os.path.join(REF, "prior_stats_nROI48.pik"), os.path.join(tmp, "prior_stats_nROI48.pik"))
        os.chdir(tmp)
        try:
            np.random.seed(seed)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")           # scipy's mahalanobis: sqrt of a negative form -> NaN (the rule rejects it)
                exec(compile(src, os.path.join(REF, "sample_sim_data.py"), "exec"), {"__name__": "__main__"})
            hits = glob.glob(os.path.join(tmp, "sim_data", "nROI48", "*", "data_nROI48_n%d_s*.pik" % n_samples))
            assert len(hits) == 1, hits
            d = pickle.load(open(hits[0], "rb"))
            d["_dir"] = os.path.basename(os.path.dirname(hits[0]))
            d["_file"] = os.path.basename(hits[0])
            d["_args_file"] = sorted(os.path.basename(p) for p in glob.glob(os.path.join(os.path.dirname(hits[0]), "args_*")))
        finally:
            os.chdir(cwd)
    return d


def as_arrays(d, seed):
    """The pickle's lists of per-sample arrays as stacked arrays (what np.asarray(load_test_dict[...]) gives, mcmc.py:73-88)."""
    out = {k: np.asarray(d[k], np.float64) for k in ("varDVR", "varR1", "vark2p", "vartacref", "tac_sampled", "tac_noisy_sampled",
                                                     "mu_noise", "sigma_noise", "mean_sigma_noise", "time_vector", "dt")}
    out["flag_mahalanobis"] = np.asarray(bool(d["flag_mahalanobis"]))
    out["target_ROI_names"] = np.asarray(d["target_ROI_names"]).astype(str)
    out["pickle_keys"] = np.asarray(sorted(k for k in d if not k.startswith("_")))
    out["pickle_types"] = np.asarray([type(d[k]).__name__ for k in sorted(d) if not k.startswith("_")])
    out["dir_suffix"] = np.asarray(d["_dir"].split("_")[-1])
    out["file_name"] = np.asarray(d["_file"])
    out["args_file"] = np.asarray(d["_args_file"])
    out["numpy_global_seed"] = np.asarray(seed)
    return out


def standardised_noise(d):
    """z = (noisy - clean) / (sqrt(clean) sigma) in concentration units (sample_sim_data.py:205-215): truncated standard normal,
    z >= -sqrt(clean)/sigma."""
    dt = np.asarray(d["dt"])[None, None, :]
    clean = np.asarray(d["tac_sampled"]) / dt
    noisy = np.asarray(d["tac_noisy_sampled"]) / dt
    sig = np.asarray(d["sigma_noise"])[None]
    with np.errstate(divide="ignore", invalid="ignore"):
        z = (noisy - clean) / (np.sqrt(clean) * sig)
        low = -np.sqrt(clean) / sig
    return z, low


def stats_of(d, tag):
    out = {}
    for k in ("varDVR", "varR1", "vartacref"):
        a = np.asarray(d[k], np.float64)
        out["%s_%s_mean" % (tag, k)] = a.mean(0)
        out["%s_%s_sd" % (tag, k)] = a.std(0, ddof=1)
    out["%s_n" % tag] = np.asarray(len(d["varDVR"]))
    z, low = standardised_noise(d)
    ok = np.isfinite(z)
    out["%s_noise_z_mean_by_frame" % tag] = np.array([z[:, :, f][ok[:, :, f]].mean() for f in range(z.shape[2])])
    out["%s_noise_z_sd_by_frame" % tag] = np.array([z[:, :, f][ok[:, :, f]].std() for f in range(z.shape[2])])
    # where the truncation is far away (low < -6) z is a plain standard normal: quantiles of those
    far = ok & (low < -6)
    out["%s_noise_z_far_quantiles" % tag] = np.quantile(z[far], [0.01, 0.1, 0.25, 0.5, 0.75, 0.9, 0.99])
    out["%s_noise_z_far_n" % tag] = np.asarray(int(far.sum()))
    # where it bites (low > -1): mean of z is that of a normal truncated at `low`; keep (low, z) pairs thinned
    near = ok & (low > -1.5)
    idx = np.flatnonzero(near.ravel())[:: max(1, int(near.sum()) // 4000)]
    out["%s_noise_near_low" % tag] = low.ravel()[idx]
    out["%s_noise_near_z" % tag] = z.ravel()[idx]
    # sigma_noise = sigma_roi / sqrt(dt exp(-lambda t)) (sample_sim_data.py:197-199): keep the table of this run
    out["%s_sigma_noise" % tag] = np.asarray(d["sigma_noise"], np.float64)
    out["%s_min_clean" % tag] = np.asarray(min(np.min(x) for x in d["tac_sampled"]))
    out["%s_min_noisy" % tag] = np.asarray(min(np.min(x) for x in d["tac_noisy_sampled"]))
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    for style, test in (("test", True), ("train", False)):
        seed = 20250710 + int(test)
        d = run_reference_script(6, test, 0.1, seed)
        np.savez_compressed(os.path.join(OUT, "reference_generated_%s_s0.1.npz" % style), **as_arrays(d, seed))
        print(style, "pickle keys:", sorted(k for k in d if not k.startswith("_")), d["_dir"], d["_file"], d["_args_file"])
    st = {}
    st.update(stats_of(run_reference_script(3000, False, 0.1, 31), "train"))
    st.update(stats_of(run_reference_script(1200, True, 0.1, 32), "test"))
    # sigma_roi ~ truncnorm(mean, 0.3 mean, low 0) (sample_sim_data.py:197): 48 values per run; pool a few runs for its moments
    rois = []
    for s in range(40):
        d = run_reference_script(1, False, 0.1, 1000 + s)
        t, dtv = np.asarray(d["time_vector"]), np.asarray(d["dt"])
        lam = np.log(2) / 109.8
        rois.append(np.asarray(d["sigma_noise"])[:, 0] * np.sqrt(dtv[0] * np.exp(-lam * t[0])))
    st["sigma_roi_pool"] = np.concatenate(rois)
    np.savez_compressed(os.path.join(OUT, "reference_generated_stats.npz"), **st)
    print("wrote reference_generated_*.npz;", {k: v.shape for k, v in st.items() if v.ndim})


if __name__ == "__main__":
    main()
