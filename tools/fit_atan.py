"""Fit the odd minimax polynomial used by vaw_atanf_pos (video_annotator_b200/csrc/vaw_coords.cuh).

atan(t) = t + t*s*P(s), s = t*t, t in [0, 1].  Remez exchange on P (relative error of atan),
then the coefficients are rounded to fp32.  Also searches two floats whose product is pi/2.
"""
import sys
import numpy as np
from numpy.polynomial import polynomial as Pn

deg = int(sys.argv[1]) if len(sys.argv) > 1 else 7   # degree of P in s

def target(s):
    t = np.sqrt(s)
    out = np.empty_like(s)
    small = s < 1e-8
    out[~small] = (np.arctan(t[~small]) - t[~small]) / (t[~small] * s[~small])
    out[small] = -1.0 / 3 + s[small] / 5
    return out

def weight(s):
    # error in atan = t*s*dP ; relative to atan(t)
    t = np.sqrt(s)
    return np.where(s > 0, t * s / np.maximum(np.arctan(t), 1e-300), 0.0)

n = deg + 2
xs = 0.5 - 0.5 * np.cos(np.pi * (np.arange(n) + 0.5) / n)      # Chebyshev nodes on [0,1]
grid = np.linspace(0, 1, 200001)
for it in range(60):
    A = np.zeros((n, n))
    for j in range(deg + 1):
        A[:, j] = xs ** j
    w = weight(xs)
    A[:, deg + 1] = [(-1) ** i / max(w[i], 1e-30) for i in range(n)]
    sol = np.linalg.solve(A, target(xs))
    coef, E = sol[:-1], sol[-1]
    err = (Pn.polyval(grid, coef) - target(grid)) * weight(grid)
    # new extrema: split at sign changes
    sign = np.sign(err)
    idx = np.where(np.diff(sign) != 0)[0]
    bounds = np.concatenate([[0], idx + 1, [len(grid)]])
    ext = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        k = a + np.argmax(np.abs(err[a:b]))
        ext.append(k)
    if len(ext) != n:
        # keep the n largest alternating
        ext = sorted(sorted(ext, key=lambda k: -abs(err[k]))[:n])
        if len(ext) < n:
            break
    new = grid[ext]
    if np.allclose(new, xs, atol=1e-9):
        break
    xs = new
print("max weighted (relative) error %.3e  (2^-24 = %.3e)" % (np.abs(err).max(), 2.0 ** -24))
c32 = coef.astype(np.float32)
print("coefficients of P(s), highest degree first:")
for c in c32[::-1]:
    print("    %.10ef," % c)
err32 = (Pn.polyval(grid, c32.astype(np.float64)) - target(grid)) * weight(grid)
print("after fp32 rounding of coefficients: %.3e" % np.abs(err32).max())

# pi/2 = a*b with a, b floats
best = None
half_pi = np.pi / 2
for a in np.float32(0.9) + np.arange(0, 200000, dtype=np.float32) * np.float32(2.0 ** -24):
    b = np.float32(half_pi / float(a))
    d = abs(float(a) * float(b) - half_pi)
    if best is None or d < best[0]:
        best = (d, a, b)
print("pi/2 ~= %.9ef * %.9ef  (error %.2e)" % (best[1], best[2], best[0]))
