// petmh_conv.cuh -- the public helper functions of the reference's kinetic_model.py as general-size fp64 kernels:
//   interp1d_linear_vec(x, xp, fp)                          kinetic_model.py:35-57
//   estimate_continuous_convolution(x, y0, y1, num_points)  kinetic_model.py:12-32
//   SRTM.make_time_exponential(param, time_vector)          kinetic_model.py:118-122
// Any strictly increasing grid, any number of columns.  The sampler's hot path never runs these: it uses the same
// arithmetic in operator form (petmh_device.cuh: build_operators / eval3), specialised to the 54-frame grid.  These are the
// drop-ins for callers of the reference's module-level functions (kinetic_model.SRTM.convolve, custom models).
//
// The per-element routines are plain functions of their indices (PETMH_HD), so that the CPU test suite can compile this
// header with g++ (oracle/c/conv_check.cpp) and check the index logic against the live reference's golden vectors
// without a GPU; the kernels below only map threads to indices.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define PETMH_HD __host__ __device__ __forceinline__
#else
#define PETMH_HD inline
#endif

namespace petmh {

// numpy.searchsorted(xp, x, side='left'): the first index with xp[idx] >= x (np if none)
PETMH_HD int searchsorted_left(const double* xp, int np, double x) {
    int lo = 0, hi = np;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (xp[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// The two non-zero weights of row `x` of interp1d_linear_vec's weight matrix (kinetic_model.py:41-49): hi = searchsorted,
// lo = hi - 1 -- numpy's index -1 WRAPS to the last node when x <= xp[0] (:47-48) --, |xp[lo] - x| on hi and |xp[hi] - x|
// on lo, normalised by their sum.  The caller guarantees x <= xp[np - 1] (the reference raises IndexError beyond).
PETMH_HD void linear_taps(const double* xp, int np, double x, int& ia, double& wa, int& ib, double& wb) {
    const int hi = searchsorted_left(xp, np, x);
    int lo = hi - 1;
    if (lo < 0) lo += np;
    const double w_hi = fabs(xp[lo] - x), w_lo = fabs(xp[hi] - x);
    const double s = w_hi + w_lo;
    ia = lo; wa = w_lo / s;
    ib = hi; wb = w_hi / s;
}

// numpy.interp(x, xp, fp) for increasing xp (kinetic_model.py:21): slope form on [xp[j], xp[j+1]), end values outside
PETMH_HD double np_interp(const double* xp, const double* fp, int np, double x) {
    if (x <= xp[0]) return fp[0];
    if (x >= xp[np - 1]) return fp[np - 1];
    int lo = 0, hi = np - 1;                       // xp[lo] <= x < xp[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (xp[mid] <= x) lo = mid; else hi = mid;
    }
    const double slope = (fp[lo + 1] - fp[lo]) / (xp[lo + 1] - xp[lo]);
    return slope * (x - xp[lo]) + fp[lo];
}

// numpy.linspace(a, b, num)[i]: i * step + a, the last point exactly b (kinetic_model.py:17)
PETMH_HD double linspace_point(double a, double b, int num, int i) {
    if (i == num - 1) return b;
    return (double)i * ((b - a) / (double)(num - 1)) + a;
}

// interp1d_linear_vec(x, xp, fp)[i][c] for fp [np][m]
PETMH_HD double interp_elem(const double* xp, int np, const double* fp, int m, double x, int c) {
    int ia, ib;
    double wa, wb;
    linear_taps(xp, np, x, ia, wa, ib, wb);
    return wa * fp[(size_t)ia * m + c] + wb * fp[(size_t)ib * m + c];
}

// estimate_continuous_convolution, step 1 (kinetic_model.py:17-22): the resampled inputs at grid point i of
// x_rs = linspace(x[0], x[n-1], N): y0_rs[i] (np.interp) and y1_rs[i][c] (interp1d_linear_vec)
PETMH_HD double conv_resample_y0(const double* x, int n, const double* y0, int N, int i) {
    return np_interp(x, y0, n, linspace_point(x[0], x[n - 1], N, i));
}
PETMH_HD double conv_resample_y1(const double* x, int n, const double* y1, int m, int N, int i, int c) {
    return interp_elem(x, n, y1, m, linspace_point(x[0], x[n - 1], N, i), c);
}
// step 2 (:25-29): the causal discrete convolution truncated to N points, times the grid spacing --
// np.convolve(y0_rs, y1_rs)[:N] * dx, equally scipy convolve1d(y1_rs, y0_rs, mode='constant', origin=-N//2) * dx
PETMH_HD double conv_discrete(const double* y0_rs, const double* y1_rs, int m, double dx, int i, int c) {
    double s = 0.0;
    for (int k = 0; k <= i; k++) s += y0_rs[k] * y1_rs[(size_t)(i - k) * m + c];
    return s * dx;
}
// step 3 (:32): back onto the original grid, interp1d_linear_vec(x, x_rs, conv)[j][c]; x_rs is passed as an array
PETMH_HD double conv_back(const double* x_rs, int N, const double* conv, int m, double xj, int c) {
    return interp_elem(x_rs, N, conv, m, xj, c);
}

#ifdef __CUDACC__
// one CTA per output row, threads over the m columns
__global__ void interp_rows_kernel(const double* x, int nx, const double* xp, int np, const double* fp, int m, double* out) {
    const int i = blockIdx.x;
    for (int c = threadIdx.x; c < m; c += blockDim.x) out[(size_t)i * m + c] = interp_elem(xp, np, fp, m, x[i], c);
}
__global__ void conv_resample_kernel(const double* x, int n, const double* y0, const double* y1, int m, int N, double* x_rs,
                                     double* y0_rs, double* y1_rs) {
    const int i = blockIdx.x;
    if (threadIdx.x == 0) {
        x_rs[i] = linspace_point(x[0], x[n - 1], N, i);
        y0_rs[i] = conv_resample_y0(x, n, y0, N, i);
    }
    for (int c = threadIdx.x; c < m; c += blockDim.x) y1_rs[(size_t)i * m + c] = conv_resample_y1(x, n, y1, m, N, i, c);
}
__global__ void conv_discrete_kernel(const double* x_rs, const double* y0_rs, const double* y1_rs, int m, int N, double* conv) {
    const int i = blockIdx.x;
    const double dx = x_rs[1] - x_rs[0];                       // kinetic_model.py:18
    for (int c = threadIdx.x; c < m; c += blockDim.x) conv[(size_t)i * m + c] = conv_discrete(y0_rs, y1_rs, m, dx, i, c);
}
__global__ void conv_back_kernel(const double* x, int n, const double* x_rs, int N, const double* conv, int m, double* out) {
    const int j = blockIdx.x;
    for (int c = threadIdx.x; c < m; c += blockDim.x) out[(size_t)j * m + c] = conv_back(x_rs, N, conv, m, x[j], c);
}
// out[i][j] = exp(param[j] * t[i])
__global__ void time_exponential_kernel(const double* param, int np, const double* t, int nt, double* out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (size_t)nt * np) out[idx] = exp(param[idx % np] * t[idx / np]);
}
#endif

}  // namespace petmh
