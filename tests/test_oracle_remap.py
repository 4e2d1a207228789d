"""Pin oracle/remap_ref.c to the real cv::remap (reference call site
opencv/FrameSourceWarp.cpp:306-312): committed cv2 outputs + live cv2 when present."""
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN


def test_remap_golden_vectors(oracle):
    g = np.load(os.path.join(GOLDEN, "remap_cases.npz"))
    mx, my = g["map_x"], g["map_y"]
    for cn in (1, 2, 3):
        src = g[f"src{cn}"]
        for bi, border in enumerate(g["borders"]):
            got = oracle.remap_u8(src, mx, my, border=border[:cn])
            assert np.array_equal(got, g[f"dst{cn}_b{bi}"]), (cn, bi)


def test_remap_documented_cases(oracle):
    """SURVEY 9.2 known answers: partial-border blends and specials."""
    src = np.full((16, 16), 221, np.uint8)
    mx = np.array([[15.5, -0.5, np.nan, np.inf, -np.inf, 1e9, -1e9, 3.0]], np.float32)
    my = np.full_like(mx, 4.0)
    out = oracle.remap_u8(src, mx, my, border=(0,))
    assert out[0, 0] == 111          # x = 15.5: tap 16 is outside, blends with border 0
    assert out[0, 1] == 111          # x = -0.5
    assert list(out[0, 2:7]) == [0] * 5
    assert out[0, 7] == 221
    src2 = np.full((8, 8, 2), 200, np.uint8)
    o2 = oracle.remap_u8(src2, np.array([[-0.5]], np.float32), np.array([[3.0]], np.float32), border=(128, 128))
    assert list(o2[0, 0]) == [164, 164]
    o3 = oracle.remap_u8(src2, np.array([[-0.5]], np.float32), np.array([[3.0]], np.float32), border=(0, 0))
    assert list(o3[0, 0]) == [100, 100]


def test_remap_33_levels(oracle):
    """A 0/255 step sampled at 257 sub-pixel offsets gives exactly 33 levels (1/32-px buckets)."""
    src = np.zeros((4, 8), np.uint8)
    src[:, 4:] = 255
    mx = (3.0 + np.arange(257) / 256.0).astype(np.float32)[None, :]
    my = np.full_like(mx, 1.0)
    out = oracle.remap_u8(src, mx, my)
    assert len(np.unique(out)) == 33


@pytest.mark.parametrize("cn", [1, 2, 3])
def test_remap_live_cv2(oracle, cn):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7 + cn)
    src = rng.integers(0, 256, (113, 75) if cn == 1 else (113, 75, cn), dtype=np.uint8)
    mx = rng.uniform(-4, 79, (200, 300)).astype(np.float32)
    my = rng.uniform(-4, 117, (200, 300)).astype(np.float32)
    # exact 1/64 ties everywhere in one block
    mx[:50] = np.round(mx[:50] * 64) / 64
    my[:50] = np.round(my[:50] * 64) / 64
    for border in [(0, 0, 0), (128, 128, 128), (255, 1, 77)]:
        bv = tuple(float(b) for b in border[:cn])
        ref = cv2.remap(src, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT,
                        borderValue=bv if cn > 1 else bv[0])
        got = oracle.remap_u8(src, mx, my, border=border[:cn])
        assert np.array_equal(ref, got)


def test_remap_threads_agree(oracle):
    rng = np.random.default_rng(3)
    src = rng.integers(0, 256, (64, 64), dtype=np.uint8)
    mx = rng.uniform(-2, 66, (33, 47)).astype(np.float32)
    my = rng.uniform(-2, 66, (33, 47)).astype(np.float32)
    assert np.array_equal(oracle.remap_u8(src, mx, my, threads=1), oracle.remap_u8(src, mx, my, threads=5))


@pytest.mark.parametrize("cn", [1, 2, 3])
def test_nearest_is_the_linear_filter_on_the_rounded_map(oracle, cn):
    """cv::remap(INTER_NEAREST) samples at (cvRound(x), cvRound(y)) -- saturate_cast<short> of the map,
    round-half-even -- and takes the border value outside.  The library implements it as the integer
    bilinear filter on a map rounded to whole pixels (all the weight on tap 00); this pins that
    identity on the real cv2.remap: random, half-integer, out-of-range and non-finite coordinates."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11 + cn)
    h, w = 37, 53
    src = rng.integers(0, 256, (h, w) if cn == 1 else (h, w, cn), dtype=np.uint8)
    mx = rng.uniform(-3, w + 3, (64, 80)).astype(np.float32)
    my = rng.uniform(-3, h + 3, (64, 80)).astype(np.float32)
    mx[0, :40] = np.arange(40, dtype=np.float32) - 2.5          # exact .5 ties, both parities
    my[0, :40] = 5.5
    mx[1, :6] = [np.nan, np.inf, -np.inf, 1e9, -1e9, w - 0.5]
    my[1, :6] = [3.0, 3.0, 3.0, 3.0, 3.0, h - 0.5]
    border = (7, 130, 250)[:cn]
    want = cv2.remap(src, mx, my, cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT, borderValue=border)
    got = oracle.remap_u8(src, np.rint(mx), np.rint(my), border=border)
    assert np.array_equal(got.reshape(want.shape), want)


@pytest.mark.parametrize("cn", [1, 2, 3])
def test_cubic_vs_live_cv2(oracle, cn):
    """oracle/remap_cubic_ref.c against the real cv2.remap(INTER_CUBIC and INTER_LANCZOS4, BORDER_CONSTANT): random maps
    that straddle the border, grid-aligned and non-finite coordinates, overshoot on a checkerboard."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(21 + cn)
    h, w = 41, 57
    src = rng.integers(0, 256, (h, w) if cn == 1 else (h, w, cn), dtype=np.uint8)
    src[::2, ::2] = 255
    src[1::2, 1::2] = 0
    mx = rng.uniform(-4, w + 4, (96, 120)).astype(np.float32)
    my = rng.uniform(-4, h + 4, (96, 120)).astype(np.float32)
    xs = np.array([-2.0, -1.5, -1.0, -0.5, 0.0, 0.25, 0.5, 1.0, w - 2.0, w - 1.5, w - 1.0, w - 0.5, w, w + 1.5,
                   np.nan, np.inf, -np.inf, 1e9, -1e9, 3.03125, 3.96875, 4.015625], np.float32)
    mx[0, :22] = xs
    my[0, :22] = 7.0
    mx[1, :22] = 9.0
    my[1, :22] = np.where(np.isfinite(xs), np.minimum(xs, h + 1.5), xs)
    for border in ((0, 0, 0), (77, 130, 255)):
        want = cv2.remap(src, mx, my, cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT, borderValue=border[:cn])
        got = oracle.remap_u8(src, mx, my, border=border[:cn], cubic=True, threads=2)
        assert np.array_equal(got.reshape(want.shape), want)
        want = cv2.remap(src, mx, my, cv2.INTER_LANCZOS4, borderMode=cv2.BORDER_CONSTANT, borderValue=border[:cn])
        got = oracle.remap_u8(src, mx, my, border=border[:cn], lanczos4=True, threads=2)
        assert np.array_equal(got.reshape(want.shape), want)
    tab = oracle.cubic_table()
    assert (tab.reshape(1024, 16).astype(np.int64).sum(1) == 32768).all()


def test_cubic_and_nearest_golden_vectors(oracle):
    """Committed cv2.remap(INTER_CUBIC / INTER_NEAREST) outputs (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLDEN, "remap_cases.npz"))
    c = np.load(os.path.join(GOLDEN, "remap_cubic.npz"))
    mx, my = g["map_x"], g["map_y"]
    for cn in (1, 2, 3):
        src = g[f"src{cn}"]
        for bi, border in enumerate(g["borders"]):
            got = oracle.remap_u8(src, mx, my, border=border[:cn], cubic=True)
            assert np.array_equal(got, c[f"cubic{cn}_b{bi}"].reshape(got.shape)), ("cubic", cn, bi)
            got = oracle.remap_u8(src, np.rint(mx), np.rint(my), border=border[:cn])
            assert np.array_equal(got, c[f"nearest{cn}_b{bi}"].reshape(got.shape)), ("nearest", cn, bi)
            got = oracle.remap_u8(src, mx, my, border=border[:cn], lanczos4=True)
            assert np.array_equal(got, c[f"lanczos{cn}_b{bi}"].reshape(got.shape)), ("lanczos", cn, bi)
