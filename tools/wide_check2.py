import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pet_posterior_distribution_b200 import MHSampler
g = "tests/golden/"
pr = np.load(g + "prior_stats_nROI48.npz"); ds = np.load(g + "dataset_s0.1.npz")
y = ds["tac_noisy_sampled"] / ds["dt"][None, None, :]
def run(C, nt, wide, draws=40, tune=230):
    os.environ["PETMH_WIDE"] = wide
    s = MHSampler(n_chains=C, max_tacs=nt, max_draws=draws, seed=11)
    s.set_frames(ds["time_vector"], ds["dt"]); s.set_prior(pr["mu_DVR"], pr["Cov_DVR"], pr["mu_R1"], pr["Cov_R1"])
    s.set_data(y[:nt], ds["vartacref"][:nt], ds["vark2p"][:nt], ds["sigma_noise"])
    s.run(draws=draws, tune=tune)
    dvr, r1 = s.chains(); s.close()
    return dvr
for nt in (1, 2):
    for tune in (0, 100, 230):
        a5n, a5w = run(5, nt, "0", tune=tune), run(5, nt, "1", tune=tune)
        a6n, a6w = run(6, 1, "0", tune=tune), run(6, 1, "1", tune=tune)
        print("nt", nt, "tune", tune, "n5==w5", np.array_equal(a5n, a5w), "| tac0: n5==n6[:5]", np.array_equal(a5n[0], a6n[0, :5]) if nt == 1 else "-",
              "w5==w6[:5]", np.array_equal(a5w[0], a6w[0, :5]) if nt == 1 else "-", "n6==w6", np.array_equal(a6n, a6w),
              "chains differing n5/w5:", sorted(set(np.argwhere(a5n != a5w)[:, 1].tolist())), "tacs", sorted(set(np.argwhere(a5n != a5w)[:, 0].tolist())))
