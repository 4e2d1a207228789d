"""Whole-path oracle (oracle/nv12_warp_ref.c) against cv2-generated fixtures."""
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN, rotation_xyz


def test_nv12_small_golden(oracle):
    g = np.load(os.path.join(GOLDEN, "nv12_small.npz"))
    k = oracle.intrinsics(g["K_in"], g["K_out"])
    for bi, border in enumerate(g["borders"]):
        got = oracle.warp_nv12(g["src"], 96, 64, 80, 48, k, g["rot"], border=tuple(border))
        assert np.array_equal(got, g[f"dst_b{bi}"]), bi


def test_nv12_live_cv2_1080p(oracle):
    cv2 = pytest.importorskip("cv2")
    sw, sh = 1920, 1080
    cam = oracle.get_preset_camera(4, sw, sh)
    outc = oracle.get_output_camera(cam, 1.0, False, 1.0)
    ow, oh = outc.width & ~1, outc.height & ~1
    assert (ow, oh) == (1758, 998)
    k = oracle.intrinsics(cam.K, outc.K)
    rot = rotation_xyz(1.0, -2.0, 0.5)
    src = oracle.synth_nv12(sw, sh, 1, white_noise=True)
    got = oracle.warp_nv12(src, sw, sh, ow, oh, k, rot, threads=8)
    mx, my = oracle.create_map(k, rot, oh, ow, threads=8)
    cx, cy = oracle.chroma_map(mx, my)
    ry = cv2.remap(src[:sh], mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0.0)
    ruv = cv2.remap(src[sh:].reshape(sh // 2, sw // 2, 2), cx, cy, cv2.INTER_LINEAR,
                    borderMode=cv2.BORDER_CONSTANT, borderValue=(128.0, 128.0))
    assert np.array_equal(got[:oh], ry)
    assert np.array_equal(got[oh:], ruv.reshape(oh // 2, ow))


def test_bgr_literal_reference_behaviour(oracle):
    """The reference warps BGR 8UC3 (FrameSourceWarp.cpp:401,445) at 1759x998 for 1080p input."""
    cv2 = pytest.importorskip("cv2")
    cam = oracle.get_preset_camera(4, 1920, 1080)
    outc = oracle.get_output_camera(cam, 1.0, False, 1.0)
    k = oracle.intrinsics(cam.K, outc.K)
    rng = np.random.default_rng(11)
    src = rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
    got = oracle.warp_bgr(src, outc.width, outc.height, k, np.eye(3), threads=8)
    mx, my = oracle.create_map(k, np.eye(3), outc.height, outc.width, threads=8)
    ref = cv2.remap(src, mx, my, cv2.INTER_LINEAR)  # defaults, as the reference calls it
    assert got.shape == (998, 1759, 3)
    assert np.array_equal(got, ref)


def test_synth_is_band_limited(oracle):
    f = oracle.synth_nv12(256, 128, 5).astype(np.int32)
    y = f[:128]
    assert np.abs(np.diff(y, axis=1)).max() <= 19 and np.abs(np.diff(y, axis=0)).max() <= 19
    uv = f[128:].reshape(64, 128, 2)
    assert np.abs(np.diff(uv, axis=1)).max() <= 21 and np.abs(np.diff(uv, axis=0)).max() <= 21
    w = oracle.synth_nv12(256, 128, 5, white_noise=True)
    assert 100 < w.mean() < 156 and w.std() > 60


def test_touched_bytes_c3_geometry(oracle):
    """the C3 geometry touches ~100% of the source luma plane."""
    cam = oracle.get_preset_camera(4, 960, 540)
    f = 984.866 * 3840 / 3524 / 4
    K_out = np.array([[f, 0, 479.5], [0, f, 269.5], [0, 0, 1]])
    k = oracle.intrinsics(cam.K, K_out)
    mx, my = oracle.create_map(k, np.eye(3), 540, 960, threads=4)
    assert oracle.touched_bytes(mx, my, 960, 540) > 0.99 * 960 * 540
