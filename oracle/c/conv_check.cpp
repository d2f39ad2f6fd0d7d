// conv_check.cpp -- TEST INFRASTRUCTURE: compiles the per-element routines of the product header
// pet_posterior_distribution_b200/csrc/petmh_conv.cuh with g++ and drives them with plain loops, so that the CPU test
// suite can check their index logic (searchsorted / wrap-around taps / truncated causal convolution) against golden vectors
// of the live reference's kinetic_model.py (tests/golden/kinetic_helpers_golden.npz) without a GPU.  The CUDA kernels in that
// header map threads to exactly these calls.  Nothing in the product links or loads this file.
#include <vector>

#include "../../pet_posterior_distribution_b200/csrc/petmh_conv.cuh"

using namespace petmh;

extern "C" void conv_check_interp(int nx, const double* x, int np, const double* xp, const double* fp, int m, double* out) {
    for (int i = 0; i < nx; i++)
        for (int c = 0; c < m; c++) out[(size_t)i * m + c] = interp_elem(xp, np, fp, m, x[i], c);
}

extern "C" void conv_check_convolution(int n, const double* x, const double* y0, const double* y1, int m, int N, double* out) {
    std::vector<double> x_rs(N), y0_rs(N), y1_rs((size_t)N * m), conv((size_t)N * m);
    for (int i = 0; i < N; i++) {                       // conv_resample_kernel
        x_rs[i] = linspace_point(x[0], x[n - 1], N, i);
        y0_rs[i] = conv_resample_y0(x, n, y0, N, i);
        for (int c = 0; c < m; c++) y1_rs[(size_t)i * m + c] = conv_resample_y1(x, n, y1, m, N, i, c);
    }
    const double dx = x_rs[1] - x_rs[0];
    for (int i = 0; i < N; i++)                         // conv_discrete_kernel
        for (int c = 0; c < m; c++) conv[(size_t)i * m + c] = conv_discrete(y0_rs.data(), y1_rs.data(), m, dx, i, c);
    for (int j = 0; j < n; j++)                         // conv_back_kernel
        for (int c = 0; c < m; c++) out[(size_t)j * m + c] = conv_back(x_rs.data(), N, conv.data(), m, x[j], c);
}
