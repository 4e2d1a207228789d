"""vaw_guess_rotation (host, csrc/vaw_camera.cpp) against guess_camera_rotation of the reference
(opencv/FrameSourceWarp.cpp:316-368) run on the real OpenCV functions (oracle/gftt_ref.guess_camera_rotation_cv2)
and against the motion the point pairs were synthesised from.  cv::solvePnPRansac draws random samples and the
reference randomises the depths with rand(), so the bar is a tolerance: 0.05 degrees between the two rotations."""
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN

TOL_DEG = 0.05


def _rot(rx, ry, rz):
    rx, ry, rz = np.deg2rad([rx, ry, rz])
    Rx = np.array([[1, 0, 0], [0, np.cos(rx), -np.sin(rx)], [0, np.sin(rx), np.cos(rx)]])
    Ry = np.array([[np.cos(ry), 0, np.sin(ry)], [0, 1, 0], [-np.sin(ry), 0, np.cos(ry)]])
    Rz = np.array([[np.cos(rz), -np.sin(rz), 0], [np.sin(rz), np.cos(rz), 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _angle_deg(Ra, Rb):
    c = (np.trace(Ra @ Rb.T) - 1) / 2
    return float(np.rad2deg(np.arccos(np.clip(c, -1, 1))))


def _fisheye_project(K, rays):
    # the equidistant model createMap.cl inverts (r = f theta), no distortion
    x, y, z = rays[:, 0], rays[:, 1], rays[:, 2]
    r = np.hypot(x, y)
    theta = np.arctan2(r, z)
    s = np.where(r > 1e-12, theta / np.maximum(r, 1e-12), 1.0)
    return np.stack([K[0, 0] * x * s + K[0, 2], K[1, 1] * y * s + K[1, 2]], axis=1)


def make_case(seed, n=180, outliers=0.2, noise=0.3, rot=(0.8, -1.1, 0.6), size=(1920, 1080)):
    import video_annotator_b200 as V
    cam = V.get_preset_camera(4, size[0], size[1])       # GOPRO_H4B_WIDE169_MEASURED
    out = V.get_output_camera(cam)
    rng = np.random.default_rng(seed)
    prev = np.stack([rng.uniform(40, size[0] - 40, n), rng.uniform(40, size[1] - 40, n)], axis=1)
    K = cam.K
    # pixels -> rays of the equidistant camera
    px, py = (prev[:, 0] - K[0, 2]) / K[0, 0], (prev[:, 1] - K[1, 2]) / K[1, 1]
    theta = np.hypot(px, py)
    s = np.where(theta > 1e-12, np.sin(theta) / np.maximum(theta, 1e-12), 1.0)
    rays = np.stack([px * s, py * s, np.cos(theta)], axis=1)
    R = _rot(*rot)
    cur = _fisheye_project(K, rays @ R.T) + rng.normal(0, noise, (n, 2))
    bad = rng.random(n) < outliers
    cur[bad] += rng.uniform(-80, 80, (int(bad.sum()), 2))
    return cam, out, prev.astype(np.float32), cur.astype(np.float32), R, int((~bad).sum())


CASES = [(1, dict()), (2, dict(rot=(0.0, 0.0, 0.0))), (3, dict(rot=(2.5, 1.5, -3.0), outliers=0.35)),
         (4, dict(n=60, outliers=0.1, rot=(-0.3, 0.2, 0.1))), (5, dict(noise=1.0, rot=(0.1, -2.0, 0.4), size=(3840, 2160)))]


@pytest.mark.parametrize("seed,kw", CASES)
def test_rotation_fit_recovers_the_motion_and_matches_the_cv2_route(seed, kw):
    cv2 = pytest.importorskip("cv2")
    import video_annotator_b200 as V
    from oracle import gftt_ref as G
    cam, out, prev, cur, R_true, n_good = make_case(seed, **kw)
    R, inl = V.guess_rotation(cam, out, prev, cur, seed=seed)
    assert abs(np.linalg.det(R) - 1) < 1e-9 and np.abs(R @ R.T - np.eye(3)).max() < 1e-9
    assert _angle_deg(R, R_true) < TOL_DEG
    assert inl >= 0.95 * n_good
    R_cv, inl_cv = G.guess_camera_rotation_cv2(cam.K, cam.distortion, out.K, prev, cur, seed=seed)
    assert _angle_deg(R, R_cv) < TOL_DEG, (_angle_deg(R, R_cv), _angle_deg(R_cv, R_true))
    assert abs(inl - inl_cv) <= 0.1 * len(prev)


def test_rotation_fit_degenerate_inputs():
    import video_annotator_b200 as V
    cam = V.get_preset_camera(4, 1920, 1080)
    out = V.get_output_camera(cam)
    R, inl = V.guess_rotation(cam, out, np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32))
    assert np.array_equal(R, np.eye(3)) and inl == 0
    # three pairs: cv::solvePnPRansac needs four -> identity, no inliers (the caller keeps the last rotation, :431-438)
    p = np.array([[100, 100], [500, 300], [900, 700]], np.float32)
    R, inl = V.guess_rotation(cam, out, p, p)
    assert np.array_equal(R, np.eye(3)) and inl == 0
    # all pairs identical: the camera did not move
    rng = np.random.default_rng(0)
    p = np.stack([rng.uniform(50, 1870, 100), rng.uniform(50, 1030, 100)], axis=1).astype(np.float32)
    R, inl = V.guess_rotation(cam, out, p, p)
    assert inl == 100 and _angle_deg(R, np.eye(3)) < 1e-6


def test_rotation_golden_fixture():
    """tests/golden/rotation_cases.npz: point pairs with the cv2 route's answers (make_golden.py --rotation)."""
    import video_annotator_b200 as V
    g = np.load(os.path.join(GOLDEN, "rotation_cases.npz"))
    for i in range(int(g["n_cases"])):
        cam = V.Camera.from_matrix(g[f"K_in_{i}"], int(g[f"size_{i}"][0]), int(g[f"size_{i}"][1]), model=1)
        out = V.Camera.from_matrix(g[f"K_out_{i}"], int(g[f"size_{i}"][0]), int(g[f"size_{i}"][1]))
        R, inl = V.guess_rotation(cam, out, g[f"prev_{i}"], g[f"cur_{i}"], seed=i)
        assert _angle_deg(R, g[f"cv_R_{i}"]) < TOL_DEG
        assert abs(inl - int(g[f"cv_inliers_{i}"])) <= 0.1 * len(g[f"prev_{i}"])
