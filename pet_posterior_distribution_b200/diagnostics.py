"""Host-side writers for the reference's side outputs (mcmc.py:162-194): the per-sample
pickle, the ArviZ-style ``_summary.csv`` and the ``rhat_less_than_102.txt`` log.

Every pm.summary column comes from the GPU: mean / sd / mcse_mean / ess_bulk / ess_tail / r_hat from
petmh_get_summary, hdi_3% / hdi_97% / mcse_sd from petmh_get_summary_ext (K3, stored draws).  The numpy ``hdi`` /
``mcse_sd`` below remain only for callers that pass no GPU columns (fewer than 8 stored draws).
"""
import io
import os

import numpy as np


def hdi(x, prob=0.94):
    """Narrowest interval containing `prob` of the pooled draws (arviz.hdi, unimodal)."""
    x = np.sort(np.asarray(x, np.float64).ravel())
    n = x.size
    k = int(np.floor(prob * n))
    if k < 1 or k >= n:
        return x[0], x[-1]
    w = x[k:] - x[:n - k]
    i = int(np.argmin(w))
    return x[i], x[i + k]


def mcse_sd(x, ess_sd):
    """arviz _mcse_sd: sd * sqrt(e * (1 - 1/ess)^(ess - 1) - 1) with ess = ess_sd."""
    sd = np.asarray(x, np.float64).std(ddof=1)
    ess_sd = max(float(ess_sd), 1.0 + 1e-9)
    fac = np.sqrt(np.exp(1) * (1 - 1 / ess_sd) ** (ess_sd - 1) - 1)
    return sd * fac


COLUMNS = ("mean", "sd", "hdi_3%", "hdi_97%", "mcse_mean", "mcse_sd", "ess_bulk", "ess_tail", "r_hat")
_DECIMALS = {c: (0 if c.startswith("ess") else 2 if c == "r_hat" else 3) for c in COLUMNS}   # az.summary's default rounding


def summary_table(dvr, r1, k2p, gpu_summary, gpu_ext=None):
    """(row labels, (97, 9) float64 array) of pm.summary(idata): rows var_DVR[i], var_R1[i], var_k2p; columns COLUMNS,
    unrounded.  dvr/r1: (chains, draws, 48); gpu_summary (96, 8) from MHSampler.summary(), gpu_ext (96, 4) = hdi_3%,
    hdi_97%, mcse_sd, ess_sd from MHSampler.summary_ext().  The var_k2p row is what ArviZ reports for a constant
    Deterministic (mcmc.py:150): sd and MCSE 0, ESS = chains * draws (a constant array's ESS is its size), r_hat NaN."""
    labels, rows = [], []
    for b, (name, arr) in enumerate((("var_DVR", dvr), ("var_R1", r1))):
        for i in range(arr.shape[-1]):
            g = gpu_summary[b * 48 + i]
            if gpu_ext is not None:
                lo, hi, msd = gpu_ext[b * 48 + i][:3]
            else:        # too few stored draws for the GPU path: numpy, with ess_sd ~ ess_bulk
                lo, hi = hdi(arr[..., i])
                msd = mcse_sd(arr[..., i], g[3])
            labels.append("%s[%d]" % (name, i))
            rows.append([g[0], g[1], lo, hi, g[2], msd, g[3], g[4], g[5]])
    k = float(np.asarray(k2p).reshape(-1)[0])
    n = float(np.prod(np.shape(dvr)[:2]))
    labels.append("var_k2p")
    rows.append([k, 0.0, k, k, 0.0, 0.0, n, n, np.nan])
    return labels, np.asarray(rows, np.float64)


def summary_csv(dvr, r1, k2p, gpu_summary, gpu_ext=None):
    """The text of pm.summary(idata).to_csv() (mcmc.py:180-181): ArviZ's default rounding (3 decimals, ESS to 0, r_hat to
    2) and pandas' own CSV formatting -- through pandas when it is importable, else the same shortest-repr floats."""
    labels, tab = summary_table(dvr, r1, k2p, gpu_summary, gpu_ext)
    try:
        import pandas as pd
        return pd.DataFrame(tab, index=labels, columns=list(COLUMNS)).round(_DECIMALS).to_csv()
    except ImportError:
        out = io.StringIO()
        out.write("," + ",".join(COLUMNS) + "\n")
        for lab, row in zip(labels, tab):
            cells = ["" if np.isnan(v) else repr(round(float(v), _DECIMALS[c]) + 0.0) for c, v in zip(COLUMNS, row)]
            out.write(lab + "," + ",".join(cells) + "\n")
        return out.getvalue()


def append_rhat_log(mcmc_dir, save_name, sample, gpu_summary, threshold=1.02):
    """mcmc.py:183-194: append a line when any DVR/R1 r_hat exceeds 1.02."""
    rmax = float(np.nanmax(gpu_summary[:, 5]))
    if rmax > threshold:
        with open(os.path.join(mcmc_dir, "rhat_less_than_102.txt"), "at") as f:
            f.write(save_name + " - sample {} - rhat_max = {:.4f}\n".format(sample, rmax))
    return rmax


class PosteriorStandIn(dict):
    """Stand-in for arviz.InferenceData when ArviZ is absent: idata.posterior['var_DVR'] etc.
    work the way mcmc.py:162-164 and main_script.py use them."""

    @property
    def posterior(self):
        return self["posterior"]

    @property
    def sample_stats(self):
        return self.get("sample_stats", {})


def make_idata(dvr, r1, k2p, sample_stats=None):
    post = {"var_DVR": np.asarray(dvr, np.float64), "var_R1": np.asarray(r1, np.float64),
            "var_k2p": np.full(dvr.shape[:2], float(np.asarray(k2p).reshape(-1)[0]))}
    try:                                   # a real InferenceData when ArviZ is importable
        import arviz as az                 # noqa: F401
        return az.from_dict(posterior=post, sample_stats=sample_stats or {})
    except Exception:
        return PosteriorStandIn(posterior=post, sample_stats=sample_stats or {})
