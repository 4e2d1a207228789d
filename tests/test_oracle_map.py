"""Coordinate oracle (oracle/create_map_ref.c, restating opencv/createMap.cl:10-50).

Pinned to the reference's own kernel in tests/test_oracle_ref.py (oracle/_ref); here the
additional anchors: an independent implementation from the reference's own dependency
(cv2.fisheye.initUndistortRectifyMap) and frozen bit patterns."""
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN, rotation_xyz


def _k_from(oracle, arr):
    k = oracle.Intrinsics()
    for (name, _), v in zip(k._fields_, arr):
        setattr(k, name, float(v))
    return k


def test_map_frozen_bits_and_cv_fisheye(oracle):
    g = np.load(os.path.join(GOLDEN, "fisheye_map.npz"))
    k = _k_from(oracle, g["intrinsics"])
    mx, my = oracle.create_map(k, g["rot"].reshape(3, 3), 48, 64)
    assert np.array_equal(mx.view(np.uint32), g["oracle_x_bits"])
    assert np.array_equal(my.view(np.uint32), g["oracle_y_bits"])
    # independent projection code agrees to fp32 rounding level (tolerance 1e-3 px)
    assert np.abs(mx - g["cv_x"]).max() < 1e-3
    assert np.abs(my - g["cv_y"]).max() < 1e-3


def test_map_hand_checked_points(oracle):
    """Identity rotation, integer principal point: r = 0 -> NaN (createMap.cl:38-39)."""
    k = oracle.Intrinsics(100.0, 80.0, 50.0, 50.0, 8.0, 6.0, 25.0, 25.0)
    mx, my = oracle.create_map(k, np.eye(3), 12, 16)
    assert np.isnan(mx[6, 8]) and np.isnan(my[6, 8])
    # on the x axis: c = (u-8)/25, map_x = 100 + atan(c)*50 (k = atan(r)/r, c0*k = atan(c))
    u = np.arange(16)
    u = u[u != 8]
    expect = 100.0 + np.arctan((u - 8.0) / 25.0) * 50.0
    assert np.abs(mx[6, u] - expect).max() < 2e-5
    assert np.abs(my[6, u] - 80.0).max() < 1e-6


def test_map_vs_float64_formula(oracle):
    """fp32 transcription stays within 1e-3 px of the same formula in float64 at 4K."""
    w = h = None
    cam = oracle.get_preset_camera(4, 3840, 2160)
    K_out = np.array([[984.866 * 3840 / 3524, 0, 1919.5], [0, 984.866 * 3840 / 3524, 1079.5], [0, 0, 1]])
    rot = rotation_xyz(2.0, -3.0, 1.5)
    k = oracle.intrinsics(cam.K, K_out)
    rows, cols = 2160, 3840
    mx, my = oracle.create_map(k, rot, rows, cols, threads=8)
    r = oracle.rot32(rot).astype(np.float64).reshape(3, 3)
    u, v = np.meshgrid(np.arange(cols, dtype=np.float64), np.arange(rows, dtype=np.float64))
    x = (u - k.map_center_x) / k.map_focal_x
    y = (v - k.map_center_y) / k.map_focal_y
    q = [r[i, 0] * x + r[i, 1] * y + r[i, 2] for i in range(3)]
    c0, c1 = q[0] / q[2], q[1] / q[2]
    rad = np.sqrt(c0 * c0 + c1 * c1)
    kk = np.arctan(rad) / rad
    ex = k.src_center_x + c0 * kk * k.src_focal_x
    ey = k.src_center_y + c1 * kk * k.src_focal_y
    assert np.nanmax(np.abs(mx - ex)) < 1e-3
    assert np.nanmax(np.abs(my - ey)) < 1e-3


def test_map_live_cv_fisheye_full_frame(oracle):
    cv2 = pytest.importorskip("cv2")
    cam = oracle.get_preset_camera(4, 1920, 1080)
    outc = oracle.get_output_camera(cam, 1.0, False, 1.0)
    rot = rotation_xyz(-1.0, 2.5, 0.7)
    k = oracle.intrinsics(cam.K, outc.K)
    mx, my = oracle.create_map(k, rot, outc.height, outc.width, threads=4)
    r32 = oracle.rot32(rot).astype(np.float64).reshape(3, 3)
    Kin = np.array([[k.src_focal_x, 0, k.src_center_x], [0, k.src_focal_y, k.src_center_y], [0, 0, 1]], np.float64)
    Kout = np.array([[k.map_focal_x, 0, k.map_center_x], [0, k.map_focal_y, k.map_center_y], [0, 0, 1]], np.float64)
    cvx, cvy = cv2.fisheye.initUndistortRectifyMap(Kin, np.zeros(4), r32.T, Kout,
                                                   (outc.width, outc.height), cv2.CV_32FC1)
    assert np.nanmax(np.abs(mx - cvx)) < 1e-3
    assert np.nanmax(np.abs(my - cvy)) < 1e-3


@pytest.mark.parametrize("dist", [(0.02, -0.015, 0.006, -0.001), (-0.05, 0.01, 0.0, 0.0)])
def test_map_with_fisheye_distortion_vs_cv_fisheye(oracle, dist):
    """Extension: k1..k4 of the input camera.  cv2.fisheye.initUndistortRectifyMap is
    the pin: same model (theta_d = theta (1 + k1 theta^2 + ...)), evaluated in double by OpenCV."""
    cv2 = pytest.importorskip("cv2")
    cam = oracle.get_preset_camera(4, 1920, 1080)
    outc = oracle.get_output_camera(cam, 1.0, False, 1.0)
    rot = rotation_xyz(1.5, -2.0, 0.8)
    k = oracle.intrinsics(cam.K, outc.K, dist=dist)
    mx, my = oracle.create_map(k, rot, outc.height, outc.width, threads=4)
    r32 = oracle.rot32(rot).astype(np.float64).reshape(3, 3)
    Kin = np.array([[k.src_focal_x, 0, k.src_center_x], [0, k.src_focal_y, k.src_center_y], [0, 0, 1]], np.float64)
    Kout = np.array([[k.map_focal_x, 0, k.map_center_x], [0, k.map_focal_y, k.map_center_y], [0, 0, 1]], np.float64)
    D = np.array([np.float32(d) for d in dist], np.float64)
    cvx, cvy = cv2.fisheye.initUndistortRectifyMap(Kin, D, r32.T, Kout, (outc.width, outc.height), cv2.CV_32FC1)
    assert np.nanmax(np.abs(mx - cvx)) < 1e-3
    assert np.nanmax(np.abs(my - cvy)) < 1e-3
    # and it is not a no-op
    m0x, _ = oracle.create_map(oracle.intrinsics(cam.K, outc.K), rot, outc.height, outc.width, threads=4)
    assert np.nanmax(np.abs(mx - m0x)) > 1.0


def test_chroma_map_definition(oracle):
    rng = np.random.default_rng(5)
    mx = rng.uniform(0, 4000, (10, 12)).astype(np.float32)
    my = rng.uniform(0, 2000, (10, 12)).astype(np.float32)
    mx[2, 3] = np.nan
    cx, cy = oracle.chroma_map(mx, my)
    q = np.float32(0.25)
    h = np.float32(0.5)
    ex = (((mx[0::2, 0::2] + mx[0::2, 1::2]) + (mx[1::2, 0::2] + mx[1::2, 1::2])) * q - h) * h
    ey = (((my[0::2, 0::2] + my[0::2, 1::2]) + (my[1::2, 0::2] + my[1::2, 1::2])) * q - h) * h
    assert np.array_equal(cx.view(np.uint32), ex.astype(np.float32).view(np.uint32))
    assert np.array_equal(cy.view(np.uint32), ey.astype(np.float32).view(np.uint32))
    assert np.isnan(cx[1, 1])
