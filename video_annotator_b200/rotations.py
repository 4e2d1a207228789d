"""Synthetic per-frame warp rotations of the BASELINE.json workloads.

Pure numpy and free of package-relative imports, so that bench.py's reference arm can load this
file stand-alone (without importing the package, hence without mapping libvaw.so) and feed the
CPU path exactly the rotations the GPU arm uses.

The sequences imitate what FrameSourceWarp::pull_frame feeds warp_frame
(FrameSourceWarp.cpp:441-442 accumulate by left-multiplication, :212/:471 Savitzky-Golay
smoothing with half-width 30 and order 2, :472-475 warp rotation = (smoothed * measured^-1)^-1).
"""
import numpy as np

ROT_SEED = 20260002


def _rodrigues(v):
    th = np.linalg.norm(v)
    if th < 1e-15:
        return np.eye(3)
    k = v / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def sg_weights(m):
    """Savitzky-Golay smoothing weights, window 2m+1, polynomial order 2 (centre point)."""
    i = np.arange(-m, m + 1, dtype=np.float64)
    return 3.0 * (3 * m * m + 3 * m - 1 - 5 * i * i) / ((2 * m + 3) * (2 * m + 1) * (2 * m - 1))


def make_rotations(n, sigma_deg, radius=30, seed=ROT_SEED):
    """(n, 3, 3) float64 warp rotations for a seeded gyro random walk (identity if sigma = 0)."""
    if sigma_deg == 0 or n == 0:
        return np.tile(np.eye(3), (n, 1, 1))
    rng = np.random.default_rng(seed)
    inc = rng.normal(0.0, np.deg2rad(sigma_deg), (n, 3))
    return rotations_from_increments(inc, radius)


def rotations_from_increments(inc, radius=30, start="zeros"):
    """Warp rotations for given per-frame axis-angle increments (n, 3): what FrameSourceWarp's
    consume_frame / pull_frame chain (:441-475) hands warp_frame for frames 1..n when the inter-frame
    rotation of frame i is rodrigues(inc[i-1]).  tests/test_host_shim.py feeds the same increments to the
    C++ shim's chain and compares.

    start: what the smoothing window holds before the first 2*radius+1 samples -- "zeros" (the library's
    constructor fills its buffer with zero matrices; host/FrameSourceWarp.hpp has the argument) or "first"
    (round 1's convention: the first sample replicated)."""
    inc = np.asarray(inc, np.float64)
    n = len(inc)
    measured = np.empty((n, 3, 3))
    acc = np.eye(3)
    for i in range(n):
        acc = _rodrigues(inc[i]) @ acc          # FrameSourceWarp.cpp:441-442
        measured[i] = acc
    w = sg_weights(radius)
    head = np.zeros((radius, 3, 3)) if start == "zeros" else np.repeat(measured[:1], radius, 0)
    pad = np.concatenate([head, measured, np.repeat(measured[-1:], radius, 0)])  # :456-461 pads with the last rotation
    out = np.empty_like(measured)
    for i in range(n):
        m = np.tensordot(w, pad[i:i + 2 * radius + 1], axes=(0, 0))
        u, _, vt = np.linalg.svd(m)              # back to SO(3)
        s = u @ vt
        if np.linalg.det(s) < 0:
            u[:, -1] *= -1
            s = u @ vt
        correction = s @ measured[i].T           # :472
        out[i] = correction.T                    # :475 (inverse of a rotation)
    return out
