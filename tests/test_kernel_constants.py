"""The numerical constants compiled into the sweep kernel, read from the shipped source and checked on the CPU against
what DESIGN.md section 2 claims for them (the GPU suite checks the kernel's results; this suite runs without a GPU)."""
import os
import re

import numpy as np
from scipy import special

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "pet_posterior_distribution_b200", "csrc", "petmh_device.cuh")).read()


def _body(name):
    i = SRC.index(name)
    return SRC[i:SRC.index("\n}\n", i)]


def test_erfc_exponent_polynomial_meets_its_bound():
    """trunc_factor2: erfc(z)/2 = 2^R6(z) on [0, Z_CUT] (mcmc.py:153-155's TruncatedNormal normaliser 1 - erfc(z)/2).  With
    the seven coefficients of the source, Horner in fp32 like the kernel's FFMA2 chain and an exact exp2: |err| < 2.7e-7
    (the MUFU ex2.approx adds at most 2 ulp of a value <= 0.5: 1.2e-7); beyond Z_CUT the factor is dropped: erfc(3.5)/2 = 3.72e-7."""
    z_cut = float(re.search(r"constexpr float Z_CUT = ([0-9.]+)f", SRC).group(1))
    coef = [np.float32(c) for c in re.findall(r"pack2\((-?[0-9.]+e[-+][0-9]+)f, \1f\)", _body("u64 trunc_factor2("))]
    assert len(coef) == 7 and z_cut == 3.5
    z = np.linspace(0.0, z_cut, 200001).astype(np.float32)
    r = np.full_like(z, coef[0])
    for c in coef[1:]:
        r = (r.astype(np.float64) * z.astype(np.float64) + np.float64(c)).astype(np.float32)     # one fp32 FMA per step
    half_erfc = 0.5 * special.erfc(z.astype(np.float64))
    err = np.abs(np.exp2(r.astype(np.float64)) - half_erfc)
    assert err.max() < 2.7e-7, err.max()
    assert 0.5 * special.erfc(z_cut) < 3.72e-7
    # relative accuracy of the factor the likelihood uses, 1 - erfc(z)/2 in [0.5, 1]
    assert (err / (1 - half_erfc)).max() < 5.4e-7


def test_tune_table_is_pymc_s():
    """tune_factor(count over 100 sweeps) == pymc.step_methods.metropolis.tune on the acceptance rate count / 100
    (restated in oracle.mh.tune_factor)."""
    from oracle import mh
    body = _body("float tune_factor(int c)")
    rules = re.findall(r"if \(c ([<>]) (\d+)\) return ([0-9.]+)f;", body)
    assert len(rules) == 6 and "return 1.0f;" in body

    def kernel(c):
        for op, thr, fac in rules:
            if (c < int(thr)) if op == "<" else (c > int(thr)):
                return float(fac)
        return 1.0

    for c in range(0, 101):
        assert kernel(c) == mh.tune_factor(c), c
    # the published table itself
    assert [kernel(c) for c in (0, 1, 4, 5, 19, 20, 50, 51, 75, 76, 95, 96, 100)] == \
        [0.1, 0.5, 0.5, 0.9, 0.9, 1.0, 1.0, 1.1, 1.1, 2.0, 2.0, 10.0, 10.0]


def test_clamp_and_cut_constants():
    """mcmc.py:152 `switch(sn < 0, 1e-6, sn)`: the kernel's clamp value; Philox rounds as Random123's philox4x32-10."""
    assert re.search(r"1e-6f|1\.0e-6f|9\.99999997e-07f", SRC), "clamp value of mcmc.py:152"
    assert "0xD2511F53" in SRC.upper().replace("0XD2511F53", "0xD2511F53") and "0xCD9E8D57" in SRC.upper().replace("0XCD9E8D57", "0xCD9E8D57")
    assert "0x9E3779B9" in SRC.upper().replace("0X9E3779B9", "0x9E3779B9") and "0xBB67AE85" in SRC.upper().replace("0XBB67AE85", "0xBB67AE85")
