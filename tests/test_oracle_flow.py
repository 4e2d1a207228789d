"""The optical-flow oracle (oracle/lk_ref.py) against the real OpenCV functions the reference calls
(cv::calcOpticalFlowPyrLK at opencv/FrameSourceWarp.cpp:255-262; cv::pyrDown inside it)."""
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN


def _textured_pair(h, w, seed, angle_deg=1.2, shift=(3.3, -2.1)):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(seed)
    base = cv2.GaussianBlur(rng.integers(0, 256, (h, w)).astype(np.uint8), (0, 0), 3)
    base = cv2.normalize(base, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
    M = cv2.getRotationMatrix2D((w / 2, h / 2), angle_deg, 1.0)
    M[:, 2] += shift
    nxt = cv2.warpAffine(base, M, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    return base, nxt


def test_pyr_down_and_scharr_equal_opencv():
    cv2 = pytest.importorskip("cv2")
    from oracle import lk_ref as L
    rng = np.random.default_rng(3)
    for shape in ((64, 96), (67, 101), (33, 40)):
        img = rng.integers(0, 256, shape).astype(np.uint8)
        assert np.array_equal(L.pyr_down(img), cv2.pyrDown(img))
        dx, dy = L.scharr_deriv(img)
        assert np.array_equal(dx, cv2.Scharr(img, cv2.CV_16S, 1, 0, borderType=cv2.BORDER_REFLECT_101))
        assert np.array_equal(dy, cv2.Scharr(img, cv2.CV_16S, 0, 1, borderType=cv2.BORDER_REFLECT_101))


@pytest.mark.parametrize("shape,seed", [((480, 640), 1), ((270, 480), 2)])
def test_lk_restatement_matches_cv2_live(shape, seed):
    cv2 = pytest.importorskip("cv2")
    from oracle import lk_ref as L
    prev, nxt = _textured_pair(shape[0], shape[1], seed)
    pts = cv2.goodFeaturesToTrack(prev, 120, 0.01, 30).reshape(-1, 2).astype(np.float32)
    # a few points at the border and outside the frame: status handling
    pts = np.concatenate([pts, np.array([[2.0, 3.0], [shape[1] - 2.5, shape[0] - 1.5], [-40.0, 10.0]], np.float32)])
    ref, st, _ = cv2.calcOpticalFlowPyrLK(prev, nxt, pts, None)
    ref, st = ref.reshape(-1, 2), st.reshape(-1).astype(bool)
    got, st2 = L.calc_optical_flow_pyr_lk(prev, nxt, pts)
    assert np.array_equal(st, st2)
    assert np.abs(got - ref)[st].max() < 1e-4
    assert np.abs(ref - pts)[st].mean() > 1.0  # the points really moved


def test_lk_golden_fixture():
    """tests/golden/lk_small.npz: frames, points and cv2's own answer (tests/golden/make_golden.py --lk)."""
    from oracle import lk_ref as L
    g = np.load(os.path.join(GOLDEN, "lk_small.npz"))
    got, st = L.calc_optical_flow_pyr_lk(g["prev"], g["next"], g["pts"])
    assert np.array_equal(st, g["cv_status"].astype(bool))
    assert np.abs(got - g["cv_next"])[st].max() < 1e-4
    assert np.array_equal(L.pyr_down(g["prev"]), g["cv_pyr1"])
