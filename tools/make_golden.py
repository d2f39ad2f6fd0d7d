#!/usr/bin/env python
"""Generate tests/golden/* in the BUILD container, where /root/reference exists.

  prior_stats_nROI48.npz  -- the reference's prior statistics (prior_stats_nROI48.pik),
                             re-saved as plain arrays (a data file, not source).
  forward_golden.npz      -- inputs/outputs of the LIVE reference kinetic_model.py
                             (SRTM2.create_activity_curve, estimate_continuous_convolution)
                             on seeded inputs: the pin for oracle/forward.py and the GPU path.
  dataset_s0.1.npz        -- a 4-TAC test-style synthetic dataset from oracle/generator.py
                             (restated sample_sim_data.py), seed recorded.

Run:  python tools/make_golden.py      (the GPU box never needs /root/reference)
"""
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    import kinetic_model as km          # the live reference
    from oracle import frames, generator
    os.makedirs(OUT, exist_ok=True)
    prior = pickle.load(open("/root/reference/prior_stats_nROI48.pik", "rb"))
    np.savez_compressed(os.path.join(OUT, "prior_stats_nROI48.npz"),
                        **{k: (np.asarray(v).astype(str) if k == "ROI_names" else np.asarray(v, np.float64))
                           for k, v in prior.items()})
    t, dt = frames.frame_grid()
    rng = np.random.default_rng(20260101)
    cases = {}
    n_case = 6
    c_r = np.empty((n_case, 54)); DVR = np.empty((n_case, 48)); R1 = np.empty((n_case, 48))
    tac = np.empty((n_case, 54, 48)); Mref = np.empty((n_case, 54, 54))
    k2 = np.empty((n_case, 48)); tac_srtm = np.empty((n_case, 54, 48))
    k2p = np.full(n_case, float(prior["mu_k2p"]))
    for c in range(n_case):
        spread = [0.02, 0.05, 0.1, 0.2, 0.3, 0.5][c]
        c_r[c] = np.abs(prior["mu_tac_ref"] * (1 + spread * rng.standard_normal(54)))
        DVR[c] = np.abs(prior["mu_DVR"] * (1 + spread * rng.standard_normal(48))) + 0.05
        R1[c] = np.abs(prior["mu_R1"] * (1 + spread * rng.standard_normal(48))) + 0.05
        if c == 5:
            k2p[c] = 0.03
        model = km.SRTM2(frame_time_list=t, frame_duration_list=dt, tac_reference=c_r[c])
        tac[c] = model.create_activity_curve(DVR=DVR[c], R1=R1[c], k2p=k2p[c])
        Mref[c] = km.estimate_continuous_convolution(t, c_r[c], np.eye(54))
        # SRTM with k2 free (kinetic_model.py:62-84), the other model of the file
        k2[c] = np.abs(0.0126 * R1[c] * (1 + spread * rng.standard_normal(48))) + 1e-3
        tac_srtm[c] = km.SRTM(frame_time_list=t, frame_duration_list=dt).forward_model(DVR=DVR[c], k2=k2[c], R1=R1[c], tac_ref=c_r[c])
    np.savez_compressed(os.path.join(OUT, "forward_golden.npz"), t=t, dt=dt, c_r=c_r, DVR=DVR, R1=R1,
                        k2p=k2p, tac=tac, M=Mref, k2=k2, tac_srtm=tac_srtm)
    ds = generator.generate(prior, 4, 0.1, test_style=True, seed=7, reject_negative=False)   # (the fixture predates the NaN rule)
    np.savez_compressed(os.path.join(OUT, "dataset_s0.1.npz"),
                        varDVR=np.array(ds["varDVR"]), varR1=np.array(ds["varR1"]),
                        vark2p=np.array(ds["vark2p"], np.float64), vartacref=np.array(ds["vartacref"]),
                        tac_sampled=np.array(ds["tac_sampled"]), tac_noisy_sampled=np.array(ds["tac_noisy_sampled"]),
                        mu_noise=ds["mu_noise"], sigma_noise=ds["sigma_noise"],
                        mean_sigma_noise=ds["mean_sigma_noise"], time_vector=ds["time_vector"], dt=ds["dt"],
                        seed=ds["seed"])
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
