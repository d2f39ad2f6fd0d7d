#!/usr/bin/env python
"""Fit used by the kernel's truncation term (petmh_device.cuh trunc_factor2):
   0.5 * erfc(z) ~= 2^R(z),  R a degree-DEG polynomial in z on [0, Z_CUT]
(weighted so that the ABSOLUTE error of 2^R is minimised; beyond Z_CUT the factor is exactly 1 - 0).  Prints the
coefficients and the max abs error in fp64 and with the kernel's fp32 Horner + ex2.approx emulated."""
import numpy as np
from numpy.polynomial import chebyshev as Ch
from scipy.special import erfc

DEG, ZCUT = 6, 3.5
z = np.linspace(0, ZCUT, 40001)
h = 0.5 * erfc(z)
f = np.log2(h)
x = 2 * z / ZCUT - 1
w = h.copy()
c = Ch.chebfit(x, f, DEG, w=w)
for it in range(60):                                   # Lawson-style reweighting towards the minimax abs error of 2^R
    err = np.abs(2 ** Ch.chebval(x, c) - h)
    w = w * (1 + 2 * (err / err.max()) ** 2)
    w /= w.max()
    c = Ch.chebfit(x, f, DEG, w=w * h)
# power basis in z
px = Ch.cheb2poly(c)                                    # polynomial in x
pz = np.zeros(DEG + 1)
for k, a in enumerate(px):                              # x = (2/ZCUT) z - 1
    pz += a * np.pad(np.polynomial.polynomial.polypow([-1.0, 2.0 / ZCUT], k), (0, DEG - k))
print("max|err| fp64 = %.3g" % np.abs(2 ** np.polynomial.polynomial.polyval(z, pz) - h).max())
print("coef (z^0..z^%d) = {%s}" % (DEG, ", ".join("%.9ef" % v for v in pz)))
zf = z.astype(np.float32)
r = np.full_like(zf, np.float32(pz[-1]))
for k in range(DEG - 1, -1, -1):
    r = (r.astype(np.float64) * zf.astype(np.float64) + np.float64(np.float32(pz[k]))).astype(np.float32)   # FMA
e = np.exp2(r.astype(np.float64)) * (1 + 2e-7 * np.sign(np.sin(1e4 * z)))                                  # ex2.approx: ~2 ulp
print("max|err| fp32 emulation = %.3g" % np.abs(e - 0.5 * erfc(zf.astype(np.float64))).max())
