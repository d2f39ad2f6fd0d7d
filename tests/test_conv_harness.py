"""petmh_conv.cuh's per-element routines (the bodies of the general-grid interpolation / convolution kernels), compiled
for the host by oracle/c/conv_check.cpp, against golden vectors of the LIVE reference's kinetic_model.py
(tools/make_golden.py helpers): the index logic -- searchsorted, the wrap-around tap at x <= xp[0], np.interp's slope
form, the truncated causal convolution -- is checked here without a GPU; tests/test_gpu_kinetic_helpers.py checks the
kernels themselves through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def chk():
    lib = os.path.join(ROOT, "oracle", "_build", "libconv_check.so")
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle", "c")], check=True, capture_output=True)
    L = C.CDLL(lib)
    dp = C.POINTER(C.c_double)
    L.conv_check_interp.argtypes = [C.c_int, dp, C.c_int, dp, dp, C.c_int, dp]
    L.conv_check_convolution.argtypes = [C.c_int, dp, dp, dp, C.c_int, C.c_int, dp]
    return L


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "kinetic_helpers_golden.npz"))


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_interp_elements_match_reference(chk, gold):
    for k in range(int(gold["n_interp"])):
        x, xp, fp, ref = (np.ascontiguousarray(gold["interp%d_%s" % (k, n)], np.float64) for n in ("x", "xp", "fp", "out"))
        f2 = fp.reshape(xp.size, -1)
        out = np.empty((x.size, f2.shape[1]))
        chk.conv_check_interp(x.size, _d(x), xp.size, _d(xp), _d(f2), f2.shape[1], _d(out))
        assert np.abs(out.reshape(ref.shape) - ref).max() <= 1e-13 * np.abs(ref).max(), k


def test_convolution_elements_match_reference(chk, gold):
    for k in range(int(gold["n_conv"])):
        x, y0, y1, ref = (np.ascontiguousarray(gold["conv%d_%s" % (k, n)], np.float64) for n in ("x", "y0", "y1", "out"))
        N = int(gold["conv%d_N" % k]) or 2 * x.size
        y2 = y1.reshape(x.size, -1)
        out = np.empty_like(y2)
        chk.conv_check_convolution(x.size, _d(x), _d(y0), _d(y2), y2.shape[1], N, _d(out))
        assert np.abs(out.reshape(ref.shape) - ref).max() <= 1e-12 * np.abs(ref).max(), k


@pytest.mark.skipif(not os.path.isfile("/root/reference/kinetic_model.py"), reason="the live reference exists in the build container only")
def test_helper_golden_regenerates_from_the_live_reference():
    """Provenance: the LIVE /root/reference kinetic_model.py helpers, called on the fixture's inputs, return the fixture's
    outputs bit for bit (estimate_continuous_convolution, interp1d_linear_vec, make_time_exponential)."""
    import sys
    code = r"""
import sys, numpy as np
sys.path.insert(0, "/root/reference")
import kinetic_model as km
g = np.load(%r)
for k in range(int(g["n_conv"])):
    N = int(g["conv%%d_N" %% k])
    out = km.estimate_continuous_convolution(g["conv%%d_x" %% k], g["conv%%d_y0" %% k], g["conv%%d_y1" %% k], num_points_resample=N or None)
    assert np.array_equal(out, g["conv%%d_out" %% k]), ("conv", k)
for k in range(int(g["n_interp"])):
    out = km.interp1d_linear_vec(g["interp%%d_x" %% k], g["interp%%d_xp" %% k], g["interp%%d_fp" %% k])
    assert np.array_equal(out, g["interp%%d_out" %% k]), ("interp", k)
assert np.array_equal(km.SRTM.make_time_exponential(g["texp_param"], g["texp_t"]), g["texp_out"])
print("ok")
""" % os.path.join(ROOT, "tests", "golden", "kinetic_helpers_golden.npz")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (r.stdout[-300:], r.stderr[-800:])
