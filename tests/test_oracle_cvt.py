"""cv::cvtColor(COLOR_YUV2BGR_NV12) restated (oracle/cvt_ref.c) -- the conversion the reference runs
on every frame before the warp (opencv/FrameSourceWarp.cpp:399-401)."""
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN


def test_cvt_golden(oracle):
    g = np.load(os.path.join(GOLDEN, "cvt_nv12_bgr.npz"))
    got = oracle.nv12_to_bgr(g["nv12"], 64, 48)
    assert np.array_equal(got, g["bgr"])


def test_cvt_live_cv2_1080p(oracle):
    cv2 = pytest.importorskip("cv2")
    nv = oracle.synth_nv12(1920, 1080, 3, white_noise=True)
    assert np.array_equal(oracle.nv12_to_bgr(nv, 1920, 1080, threads=8), cv2.cvtColor(nv, cv2.COLOR_YUV2BGR_NV12))
    nv = oracle.synth_nv12(1920, 1080, 3)
    assert np.array_equal(oracle.nv12_to_bgr(nv, 1920, 1080, threads=8), cv2.cvtColor(nv, cv2.COLOR_YUV2BGR_NV12))
