#!/usr/bin/env python
"""Generate tests/golden/* in the BUILD container, where /root/reference exists.

  prior_stats_nROI48.npz  -- the reference's prior statistics (prior_stats_nROI48.pik),
                             re-saved as plain arrays (a data file, not source).
  forward_golden.npz      -- inputs/outputs of the LIVE reference kinetic_model.py
                             (SRTM2.create_activity_curve, estimate_continuous_convolution)
                             on seeded inputs: the pin for oracle/forward.py and the GPU path.
  dataset_s0.1.npz        -- a 4-TAC test-style synthetic dataset from oracle/generator.py
                             (restated sample_sim_data.py), seed recorded.

  kinetic_helpers_golden.npz -- inputs/outputs of the live reference's module-level helpers
                             (estimate_continuous_convolution, interp1d_linear_vec,
                             SRTM.make_time_exponential) on general grids: the pin for
                             petmh_conv.cuh (CPU harness oracle/c/conv_check.cpp and the GPU path).

Run:  python tools/make_golden.py            (everything; the GPU box never needs /root/reference)
      python tools/make_golden.py helpers    (only kinetic_helpers_golden.npz)
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


def helpers():
    import kinetic_model as km          # the live reference
    from oracle import frames
    t, _ = frames.frame_grid()
    prior = pickle.load(open("/root/reference/prior_stats_nROI48.pik", "rb"))
    rng = np.random.default_rng(20261019)
    out = {}
    # ---- estimate_continuous_convolution (kinetic_model.py:12-32): (x, y0, y1, num_points_resample or 0) ----
    xg = np.sort(rng.uniform(0.0, 10.0, 17))
    k2a = rng.uniform(0.004, 0.03, 48)
    conv_cases = [
        (t, np.abs(prior["mu_tac_ref"] * (1 + 0.1 * rng.standard_normal(54))), np.exp(-k2a[None, :] * t[:, None]), 0),
        (xg, rng.standard_normal(17), rng.standard_normal((17, 3)), 0),
        (xg, rng.standard_normal(17), rng.standard_normal((17, 5)), 64),
        (xg, rng.standard_normal(17), rng.standard_normal(17), 51),          # 1-D y1: np.convolve path, odd length allowed
        (t, np.abs(rng.standard_normal(54)), rng.standard_normal(54), 0),
        (np.array([1.0, 2.5]), np.array([0.3, -1.0]), np.array([[2.0, 1.0], [0.5, -3.0]]), 0),   # the smallest grid
    ]
    for k, (x, y0, y1, N) in enumerate(conv_cases):
        out["conv%d_x" % k], out["conv%d_y0" % k], out["conv%d_y1" % k], out["conv%d_N" % k] = x, y0, y1, N
        out["conv%d_out" % k] = km.estimate_continuous_convolution(x, y0, y1, num_points_resample=N or None)
    out["n_conv"] = len(conv_cases)
    # ---- interp1d_linear_vec (kinetic_model.py:35-57), incl. x < xp[0] and x == xp[0] (index -1 wraps), nodes, x == xp[-1] ----
    xp = np.sort(rng.uniform(-3.0, 7.0, 23))
    interp_cases = [
        (np.concatenate([[xp[0] - 2.0, xp[0] - 1e-9, xp[0], xp[4], xp[-1]], rng.uniform(xp[0], xp[-1], 40)]), xp, rng.standard_normal((23, 4))),
        (rng.uniform(xp[0], xp[-1], 9), xp, rng.standard_normal(23)),
        (np.linspace(t[0], t[-1], 108), t, rng.standard_normal((54, 48))),
    ]
    for k, (x, xp_, fp) in enumerate(interp_cases):
        out["interp%d_x" % k], out["interp%d_xp" % k], out["interp%d_fp" % k] = x, xp_, fp
        out["interp%d_out" % k] = km.interp1d_linear_vec(x, xp_, fp)
    out["n_interp"] = len(interp_cases)
    # ---- SRTM.make_time_exponential (kinetic_model.py:118-122) ----
    out["texp_param"], out["texp_t"] = -k2a, t
    out["texp_out"] = km.SRTM.make_time_exponential(-k2a, t)
    out["texp_scalar_out"] = km.SRTM2.make_time_exponential(-0.0123, t)
    np.savez_compressed(os.path.join(OUT, "kinetic_helpers_golden.npz"), **out)
    print("wrote kinetic_helpers_golden.npz:", {k: np.shape(v) for k, v in out.items() if k.endswith("_out")})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "helpers":
        helpers()
    else:
        main()
        helpers()
