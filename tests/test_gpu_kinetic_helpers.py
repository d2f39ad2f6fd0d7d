"""The module-level helpers of the reference's kinetic_model.py on the GPU (csrc/petmh_conv.cuh through the C ABI and the
package's kinetic_model mirror) against golden vectors of the LIVE reference (tools/make_golden.py helpers):
estimate_continuous_convolution (kinetic_model.py:12-32), interp1d_linear_vec (:35-57), SRTM.make_time_exponential
(:118-122), SRTM.convolve (:124-128).  fp64 on both sides: 1e-12 relative to the largest value."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "kinetic_helpers_golden.npz"))


def test_continuous_convolution_matches_reference(gold):
    from pet_posterior_distribution_b200 import kinetic_model as km
    for k in range(int(gold["n_conv"])):
        x, y0, y1, ref = (gold["conv%d_%s" % (k, n)] for n in ("x", "y0", "y1", "out"))
        N = int(gold["conv%d_N" % k]) or None
        out = km.estimate_continuous_convolution(x, y0, y1, num_points_resample=N)
        assert out.shape == ref.shape and np.abs(out - ref).max() <= 1e-12 * np.abs(ref).max(), k
    # the classes' static method is the same function (kinetic_model.py:124-128, 198-201)
    x, y0, y1, ref = (gold["conv0_%s" % n] for n in ("x", "y0", "y1", "out"))
    assert np.array_equal(km.SRTM.convolve(x, y0, y1), km.estimate_continuous_convolution(x, y0, y1))
    assert np.array_equal(km.SRTM2.convolve(x, y0, y1), km.SRTM.convolve(x, y0, y1))
    # scipy's convolve1d has no origin -N//2 for an odd length: the reference's 2-D path raises ValueError there
    with pytest.raises(ValueError):
        km.estimate_continuous_convolution(gold["conv1_x"], gold["conv1_y0"], gold["conv1_y1"], num_points_resample=33)
    with pytest.raises(Exception):          # grid not increasing
        km.estimate_continuous_convolution(np.array([0.0, 2.0, 1.0]), np.ones(3), np.ones(3))


def test_general_convolution_equals_the_sampler_operator(gold, forward_golden):
    """On the 54-frame grid the general-grid routine and the sampler's device-built operator M (petmh_get_operator: what
    the hot path applies) are the same linear map: conv(t, c_r, E) == M E."""
    from pet_posterior_distribution_b200 import MHSampler
    from pet_posterior_distribution_b200 import kinetic_model as km
    g = forward_golden
    s = MHSampler(n_chains=1, max_tacs=1)
    s.set_frames(g["t"], g["dt"])
    s.set_prior(np.zeros(48), np.eye(48), np.zeros(48), np.eye(48))
    s.set_data(np.ones((1, 48, 54)), g["c_r"][2][None], np.array([0.0126]), np.ones((48, 54)))
    M = s.operator(0)
    E = np.exp(-np.linspace(0.004, 0.03, 48)[None, :] * g["t"][:, None])
    conv = km.estimate_continuous_convolution(g["t"], g["c_r"][2], E)
    assert np.abs(conv - M @ E).max() <= 1e-12 * np.abs(conv).max()
    assert np.abs(km.estimate_continuous_convolution(g["t"], g["c_r"][2], np.eye(54)) - g["M"][2]).max() <= 1e-12 * np.abs(g["M"][2]).max()
    s.close()


def test_interp1d_linear_vec_matches_reference(gold):
    from pet_posterior_distribution_b200 import kinetic_model as km
    for k in range(int(gold["n_interp"])):
        x, xp, fp, ref = (gold["interp%d_%s" % (k, n)] for n in ("x", "xp", "fp", "out"))
        out = km.interp1d_linear_vec(x, xp, fp)
        assert out.shape == ref.shape and np.abs(out - ref).max() <= 1e-13 * np.abs(ref).max(), k
    # dim = 1: the interpolated axis stays where it was (kinetic_model.py:51-57)
    x, xp, fp, ref = (gold["interp0_%s" % n] for n in ("x", "xp", "fp", "out"))
    out = km.interp1d_linear_vec(x, xp, np.ascontiguousarray(fp.T), dim=1)
    assert out.shape == (fp.shape[1], x.size) and np.abs(out.T - ref).max() <= 1e-13 * np.abs(ref).max()
    with pytest.raises(IndexError):         # beyond xp[-1] the reference indexes out of bounds
        km.interp1d_linear_vec(np.array([xp[-1] + 1.0]), xp, fp)


def test_make_time_exponential_matches_reference(gold):
    from pet_posterior_distribution_b200 import kinetic_model as km
    p, t, ref = gold["texp_param"], gold["texp_t"], gold["texp_out"]
    out = km.SRTM.make_time_exponential(p, t)
    assert out.shape == ref.shape and np.abs(out / ref - 1).max() < 1e-14
    one = km.SRTM2.make_time_exponential(-0.0123, t)
    assert one.shape == gold["texp_scalar_out"].shape and np.abs(one / gold["texp_scalar_out"] - 1).max() < 1e-14
    scaled = km.SRTM.make_time_exponential(p, t, time_scale=t)          # optional scale of make_time_func (:104-108)
    assert np.abs(scaled / (ref * t[:, None]) - 1).max() < 1e-14
