#!/usr/bin/env python
"""Least-squares fit used by the kernel's truncation term:
   0.5*erfc(z) ~= t * exp(-z^2) * Q(t),  t = 1/(1 + p z),  z in [0, 4.5], deg(Q) = 5.
Prints p, the coefficients and the max abs error (fp64 and emulated fp32 Horner)."""
import numpy as np
from scipy.special import erfc

DEG = 5
z = np.linspace(0, 4.5, 40001)
best = None
for p in np.linspace(0.30, 0.50, 201):
    t = 1 / (1 + p * z)
    w = np.exp(-z * z) * t
    A = np.vander(t, DEG + 1, increasing=True) * w[:, None]
    b = 0.5 * erfc(z)
    c, *_ = np.linalg.lstsq(A, b, rcond=None)
    err = np.abs(A @ c - b).max()
    if best is None or err < best[0]:
        best = (err, p, c)
err, p, c = best
print("p = %.17g  max|err| fp64 = %.3g" % (p, err))
print("coef (t^0..t^%d) = {%s}" % (DEG, ", ".join("%.9ef" % v for v in c)))
zf = z.astype(np.float32)
t = (np.float32(1) / (np.float32(1) + np.float32(p) * zf)).astype(np.float32)
q = np.full_like(t, np.float32(c[-1]))
for k in range(DEG - 1, -1, -1):
    q = (q * t + np.float32(c[k])).astype(np.float32)
ex = np.exp2((zf * zf * np.float32(-1.4426950408889634)).astype(np.float32)).astype(np.float32)
h = (q * t * ex).astype(np.float32)
print("max|err| fp32 emulation = %.3g" % np.abs(h.astype(np.float64) - 0.5 * erfc(zf.astype(np.float64))).max())
