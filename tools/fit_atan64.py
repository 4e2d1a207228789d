#!/usr/bin/env python
"""Coefficients of P(s) ~ atan(sqrt(s)) / sqrt(s) on s in [0, 1] for the piece builder's fp64 anchors
(csrc/vaw_pieces.cu): Chebyshev interpolation in extended precision, converted to monomials.
The builder needs ~1e-10 relative accuracy (4e-7 px at 4K); this gives < 1e-13.
python tools/fit_atan64.py [degree]  -> prints a C initialiser and the measured error."""
import sys
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P

deg = int(sys.argv[1]) if len(sys.argv) > 1 else 17
L = np.longdouble


def f(s):
    s = np.asarray(s, L)
    t = np.sqrt(s)
    out = np.ones_like(s)
    m = t > L(1e-6)
    out[m] = np.arctan(t[m]) / t[m]
    out[~m] = 1 - s[~m] / 3
    return out


n = deg + 1
k = np.arange(n, dtype=L)
z = np.cos(np.pi * (k + L(0.5)) / n)            # Chebyshev nodes on [-1, 1]
s = (z + 1) / 2
fz = f(s)
c = np.array([(L(2) / n) * np.sum(fz * np.cos(j * np.pi * (k + L(0.5)) / n)) for j in range(n)], L)
c[0] /= 2                                        # Chebyshev interpolation (n nodes, degree n - 1), no linear algebra
mono_z = C.cheb2poly(c)                          # in z = 2 s - 1
# substitute z = 2 s - 1
poly = np.zeros(1, L)
for a in mono_z[::-1]:
    poly = P.polyadd(P.polymul(poly, np.array([-1, 2], L)), np.array([a], L))
coef = np.array(poly, L)
test = np.linspace(0, 1, 2000001, dtype=L)
approx = np.zeros_like(test)
cd = coef.astype(np.float64)
td = test.astype(np.float64)
acc = np.zeros_like(td)
for a in cd[::-1]:
    acc = acc * td + a                           # double-precision Horner, as on the device
err = np.abs(acc.astype(L) - f(test)) / f(test)
print("// degree", deg, "max relative error of the double-precision Horner evaluation: %.3e" % float(err.max()))
print("static __device__ const double kAtanP[%d] = {" % len(cd))
print(",\n".join("    %.17e" % v for v in cd))
print("};")
