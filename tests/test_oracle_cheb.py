"""The Chebyshev-in-k2a operator of the sweep kernel, checked on the CPU against the
exact operator M of the pinned forward-model oracle: with the kernel's range and column counts the truncation
error of conv = M exp(-k2a t) stays below 1e-7 relative -- under the fp32 rounding of either form -- on every
reference TAC we hold (golden cases of the live reference + the golden dataset)."""
import os
import re

import numpy as np

from oracle import cheb, forward

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_constants_match_the_kernel_source():
    src = open(os.path.join(ROOT, "pet_posterior_distribution_b200", "csrc", "petmh_device.cuh")).read()
    lo = float(re.search(r"#define PETMH_CHEB_KT_LO ([0-9.]+)", src).group(1))
    hi = float(re.search(r"#define PETMH_CHEB_KT_HI ([0-9.]+)", src).group(1))
    n = tuple(int(x) for x in re.search(r"NCH0 = (\d+), NCH1 = (\d+), NCH2 = (\d+)", src).groups())
    assert (lo, hi) == (cheb.KT_LO, cheb.KT_HI) and n == cheb.NCOLS


def _crs(forward_golden, dataset):
    return [c for c in forward_golden["c_r"]] + [c for c in dataset["vartacref"]]


def test_truncation_error_below_fp32_rounding(forward_golden, dataset):
    t = dataset["time_vector"]
    lo, hi = cheb.k2a_range(t)
    ks = np.linspace(lo, hi, 801)
    worst = 0.0
    for c_r in _crs(forward_golden, dataset):
        exact = forward.build_M(t, c_r) @ np.exp(-ks[None, :] * t[:, None])
        approx = cheb.conv_cheb(t, c_r, ks)
        worst = max(worst, np.abs(approx / exact - 1).max())
    print("max relative truncation error of conv over the range: %.2e" % worst)
    assert worst < 1e-7


def _fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def test_fp32_evaluation_order(forward_golden, dataset, prior):
    """The kernel's arithmetic restated in numpy float32 (fp32 A, coef-scaled fp32 recurrence, FMA accumulation from
    zero in the order columns 2..n-1, 1, 0, then R1 c_r): the rms TAC error against the fp64 reference formula is
    below that of the exact-operator fp32 form of round 1 (4.7e-8), and far below the 1e-5 bar."""
    f32 = np.float32
    t = dataset["time_vector"]
    lo, hi = cheb.k2a_range(t)
    inv_h = f32(2.0 / (hi - lo))
    c0 = f32((cheb.KT_HI + cheb.KT_LO) / (cheb.KT_HI - cheb.KT_LO))
    hh = 1.0 / float(inv_h)
    kmid = float(c0) * hh
    rng = np.random.default_rng(5)
    sq, n, worst = 0.0, 0, 0.0
    for tac in range(dataset["varDVR"].shape[0]):
        c_r, k2p = dataset["vartacref"][tac], f32(dataset["vark2p"][tac])
        A = cheb.cheb_operator(t, c_r, kmid - hh, kmid + hh)
        for rep in range(3):
            d = (prior["mu_DVR"] * (1 + 0.03 * rep * rng.standard_normal(48))).astype(f32)
            a = (prior["mu_R1"] * (1 + 0.03 * rep * rng.standard_normal(48))).astype(f32)
            ref = forward.srtm2_tac(t, c_r, d.astype(np.float64), a.astype(np.float64), float(k2p))
            k2 = (k2p * a).astype(f32)
            k2a = (k2 / d).astype(f32)
            coef = _fma32(-a, k2a, k2)
            s = _fma32(k2a, np.full_like(k2a, inv_h), np.full_like(k2a, -c0))
            assert (np.abs(s) <= 1).all()
            ts, cs = (s + s).astype(f32), (coef * s).astype(f32)
            for b in range(3):
                nc = cheb.NCOLS[b]
                T = [coef, cs]
                for _ in range(2, nc):
                    T.append(_fma32(ts, T[-1], -T[-2]))
                A32 = A[b].astype(f32)
                acc = np.zeros((18, 48), f32)
                one = np.ones_like(acc)
                for c in list(range(2, nc)) + [1, 0]:
                    acc = _fma32(A32[:, c:c + 1] * one, T[c][None, :] * one, acc)
                acc = _fma32(c_r[18 * b:18 * b + 18].astype(f32)[:, None] * one, a[None, :] * one, acc)
                rel = acc / ref[18 * b:18 * b + 18] - 1
                sq += (rel ** 2).sum()
                n += rel.size
                worst = max(worst, np.abs(rel).max())
    print("fp32 Chebyshev TAC vs fp64 reference formula: rms %.2e max %.2e" % (np.sqrt(sq / n), worst))
    assert np.sqrt(sq / n) < 4.5e-8 and worst < 4e-7


def test_range_covers_the_reference_parameters(prior, dataset):
    """With the reference's grid and k2p the range is R1/DVR in [0, 3.97]: every golden-dataset truth and the
    prior mean are inside (chains spend their time near them; the rest falls back to the exact operator)."""
    t = dataset["time_vector"]
    lo, hi = cheb.k2a_range(t)
    k2p = float(dataset["vark2p"][0])
    for DVR, R1 in [(prior["mu_DVR"], prior["mu_R1"])] + list(zip(dataset["varDVR"], dataset["varR1"])):
        k2a = k2p * R1 / DVR
        assert (k2a > lo).all() and (k2a < hi).all()
