"""The corner-detection oracle (oracle/gftt_ref.py) against the real OpenCV functions the reference calls
(cv::goodFeaturesToTrack at opencv/FrameSourceWarp.cpp:230; cv::cornerMinEigenVal inside it)."""
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN
from tests.test_oracle_flow import _textured_pair

# OpenCV's own response differs in the last bit between its SIMD / IPP code paths (order of the fp32 sums);
# responses are <= ~0.1, so 3e-8 is a few ulp of the largest values
RESPONSE_TOL = 3e-8


@pytest.mark.parametrize("shape,seed", [((480, 640), 1), ((270, 480), 2), ((1080, 1920), 3), ((97, 131), 4)])
def test_response_and_corner_list_match_cv2_live(shape, seed):
    cv2 = pytest.importorskip("cv2")
    from oracle import gftt_ref as G
    img, _ = _textured_pair(shape[0], shape[1], seed)
    ref = cv2.cornerMinEigenVal(img, 3, ksize=3)
    got = G.corner_min_eigen_val(img)
    assert np.abs(got - ref).max() <= RESPONSE_TOL
    want = cv2.goodFeaturesToTrack(img, 200, 0.01, 30).reshape(-1, 2)
    # the selection logic alone, on OpenCV's own response: identical list, order included
    assert np.array_equal(G.select_corners(ref), want)
    # the whole restatement: the same corners (a last-bit difference could at most swap two near-equal responses)
    mine = G.good_features_to_track(img)
    assert len(mine) == len(want)
    assert len(set(map(tuple, mine.astype(int))) & set(map(tuple, want.astype(int)))) >= 0.98 * len(want)


def test_selection_parameters_match_cv2():
    cv2 = pytest.importorskip("cv2")
    from oracle import gftt_ref as G
    img, _ = _textured_pair(300, 400, 7)
    ref = cv2.cornerMinEigenVal(img, 3, ksize=3)
    for mc, q, md in ((50, 0.05, 10.0), (0, 0.2, 5.0), (1000, 0.01, 0.0), (30, 0.01, 45.5)):
        want = cv2.goodFeaturesToTrack(img, mc, q, md).reshape(-1, 2)
        assert np.array_equal(G.select_corners(ref, mc, q, md), want), (mc, q, md)


def test_corner_golden_fixture():
    """tests/golden/gftt_small.npz: a frame, cv2's response map and cv2's corner list (make_golden.py --gftt)."""
    from oracle import gftt_ref as G
    g = np.load(os.path.join(GOLDEN, "gftt_small.npz"))
    assert np.abs(G.corner_min_eigen_val(g["image"]) - g["cv_response"]).max() <= RESPONSE_TOL
    assert np.array_equal(G.select_corners(g["cv_response"]), g["cv_corners"])
    assert np.array_equal(G.good_features_to_track(g["image"]), g["cv_corners"])
