"""Motion measurement on the GPU (vaw_flow_*, csrc/vaw_flow.cu) against the optical-flow oracle
(oracle/lk_ref.py, pinned to cv2.calcOpticalFlowPyrLK): SURVEY 8 f4, the tracking step of
FrameSourceWarp::consume_frame (opencv/FrameSourceWarp.cpp:421-427, :242-270)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def V():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import video_annotator_b200 as V
    return V


def _pair(h, w, seed, angle=1.0, shift=(2.7, -1.9)):
    """A textured frame and the same texture rotated / shifted (numpy + the oracle only: no cv2 needed)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    def tex(x, y):
        v = np.zeros_like(x)
        for k in range(24):
            fx, fy, ph = rng2[k]
            v += np.sin(x * fx + y * fy + ph)
        return v
    rng2 = [(rng.uniform(0.03, 0.35) * rng.choice([-1, 1]), rng.uniform(0.03, 0.35) * rng.choice([-1, 1]), rng.uniform(0, 6.28)) for _ in range(24)]
    a = np.deg2rad(angle)
    cx, cy = w / 2, h / 2
    x2 = np.cos(a) * (xx - cx) - np.sin(a) * (yy - cy) + cx + shift[0]
    y2 = np.sin(a) * (xx - cx) + np.cos(a) * (yy - cy) + cy + shift[1]
    def to8(v):
        return np.clip(127.5 + v * 18.0, 0, 255).astype(np.uint8)
    return to8(tex(xx, yy)), to8(tex(x2, y2))


@pytest.mark.parametrize("shape", [(270, 480), (1080, 1920), (2160, 3840)])
def test_pyramid_and_derivatives_equal_the_oracle(V, shape):
    import torch
    from oracle import lk_ref as L
    h, w = shape
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (h, w)).astype(np.uint8)
    ft = V.FlowTracker(w, h)
    ft.push_frame(torch.from_numpy(img).cuda())
    pyr = L.build_pyramid(img)
    assert ft.levels == len(pyr) == 4
    for lvl, ref in enumerate(pyr):
        got, dx, dy = ft.level(1, lvl)
        assert np.array_equal(got, ref), lvl              # cv::pyrDown, bit for bit
        rdx, rdy = L.scharr_deriv(ref)
        assert np.array_equal(dx, rdx) and np.array_equal(dy, rdy), lvl
    ft.close()


@pytest.mark.parametrize("shape,seed,n", [((270, 480), 1, 80), ((1080, 1920), 2, 200), ((2160, 3840), 3, 200)])
def test_lk_tracking_equals_the_oracle(V, shape, seed, n):
    """Same fixed-point scheme, exact integer accumulation, OpenCV's fp32 operation order: the tracked positions
    equal the oracle's to 1e-5 px (bit-identical in practice) with identical status flags -- hence within 1e-4 px of
    cv2.calcOpticalFlowPyrLK, which the oracle is pinned to."""
    import torch
    from oracle import lk_ref as L
    h, w = shape
    prev, nxt = _pair(h, w, seed)
    rng = np.random.default_rng(seed + 100)
    pts = np.stack([rng.uniform(30, w - 30, n), rng.uniform(30, h - 30, n)], axis=1).astype(np.float32)
    pts = np.concatenate([pts, np.array([[1.5, 2.0], [w - 1.25, h - 2.0], [-35.0, 20.0], [w + 5.0, 10.0]], np.float32)])
    ft = V.FlowTracker(w, h)
    ft.push_frame(torch.from_numpy(prev).cuda())
    ft.push_frame(torch.from_numpy(nxt).cuda())
    got, st = ft.track(pts)
    want, wst = L.calc_optical_flow_pyr_lk(prev, nxt, pts)
    assert np.array_equal(st, wst)
    assert st[:n].mean() > 0.9
    d = np.abs(got - want)[st]
    assert d.max() <= 1e-5, d.max()
    # the flow is the known motion (rotation about the centre + shift) to within LK's own accuracy
    a = np.deg2rad(1.0)
    cx, cy = w / 2, h / 2
    # next(x2, y2) = prev(x, y) with (x2, y2) = R (x - c) + c + shift  ->  a prev point moves by the INVERSE map
    x, y = pts[:n, 0] - cx - 2.7, pts[:n, 1] - cy + 1.9
    ex = np.cos(a) * x + np.sin(a) * y + cx
    ey = -np.sin(a) * x + np.cos(a) * y + cy
    ok = st[:n]
    assert np.median(np.hypot(got[:n, 0] - ex, got[:n, 1] - ey)[ok]) < 0.2
    ft.close()


def test_lk_against_cv2_directly(V):
    """Where cv2 is importable: the GPU tracker against the real cv2.calcOpticalFlowPyrLK, no oracle in between."""
    cv2 = pytest.importorskip("cv2")
    import torch
    prev, nxt = _pair(720, 1280, 9, angle=-0.7, shift=(-3.1, 2.2))
    pts = cv2.goodFeaturesToTrack(prev, 200, 0.01, 30).reshape(-1, 2).astype(np.float32)
    ref, st, _ = cv2.calcOpticalFlowPyrLK(prev, nxt, pts, None)
    ft = V.FlowTracker(1280, 720)
    ft.push_frame(torch.from_numpy(prev).cuda())
    ft.push_frame(torch.from_numpy(nxt).cuda())
    got, gst = ft.track(pts)
    assert np.array_equal(gst, st.reshape(-1).astype(bool))
    assert np.abs(got - ref.reshape(-1, 2))[gst].max() < 1e-4
    ft.close()


@pytest.mark.parametrize("shape,seed", [((270, 480), 1), ((1080, 1920), 2), ((2160, 3840), 3), ((241, 333), 4)])
def test_corner_response_and_corners_equal_the_oracle(V, shape, seed):
    """cv::goodFeaturesToTrack(image, 200, 0.01, 30) (FrameSourceWarp.cpp:230): the device response map equals the
    oracle's bit for bit (same fp32 operations in the same order; the oracle is pinned to cv2.cornerMinEigenVal to
    3e-8), hence the same candidates, and the host tail is OpenCV's ordered greedy filter: identical corner lists."""
    import torch
    from oracle import gftt_ref as G
    h, w = shape
    img, _ = _pair(h, w, seed)
    ft = V.FlowTracker(w, h)
    ft.push_frame(torch.from_numpy(img).cuda())
    got = ft.corners(1)
    want_resp = G.corner_min_eigen_val(img)
    assert np.array_equal(ft.response(), want_resp)
    want = G.select_corners(want_resp)
    assert len(want) > 20
    assert np.array_equal(got, want)
    # other parameters of the selection: many corners, no distance filter, unlimited count
    for mc, q, md in ((1000, 0.05, 8.0), (0, 0.3, 0.0), (25, 0.01, 60.5)):
        assert np.array_equal(ft.corners(1, mc, q, md), G.select_corners(want_resp, mc, q, md)), (mc, q, md)
    ft.close()


def test_corners_on_white_noise_and_flat_frames(V):
    """Edge cases: a flat frame has no corners; white noise has a candidate at about one pixel in nine (the
    candidate list holds one in six)."""
    import torch
    from oracle import gftt_ref as G
    ft = V.FlowTracker(640, 360)
    ft.push_frame(torch.full((360, 640), 77, dtype=torch.uint8, device="cuda"))
    assert len(ft.corners(1)) == 0
    rng = np.random.default_rng(8)
    img = rng.integers(0, 256, (360, 640)).astype(np.uint8)
    ft.push_frame(torch.from_numpy(img).cuda())
    assert np.array_equal(ft.corners(1, 0, 0.01, 3.0), G.good_features_to_track(img, 0, 0.01, 3.0))
    # the previous frame is still there: which = 0
    assert len(ft.corners(0)) == 0
    ft.close()


def test_corners_against_cv2_directly(V):
    """Where cv2 is importable: the same corner list as the real cv2.goodFeaturesToTrack, no oracle in between."""
    cv2 = pytest.importorskip("cv2")
    import torch
    img, _ = _pair(720, 1280, 12)
    img = cv2.GaussianBlur(img, (0, 0), 1.5)
    want = cv2.goodFeaturesToTrack(img, 200, 0.01, 30).reshape(-1, 2)
    ft = V.FlowTracker(1280, 720)
    ft.push_frame(torch.from_numpy(img).cuda())
    got = ft.corners(1)
    assert len(got) == len(want)
    assert len(set(map(tuple, got.astype(int))) & set(map(tuple, want.astype(int)))) >= 0.98 * len(want)
    ft.close()


def _rot(rx, ry, rz):
    rx, ry, rz = np.deg2rad([rx, ry, rz])
    Rx = np.array([[1, 0, 0], [0, np.cos(rx), -np.sin(rx)], [0, np.sin(rx), np.cos(rx)]])
    Ry = np.array([[np.cos(ry), 0, np.sin(ry)], [0, 1, 0], [-np.sin(ry), 0, np.cos(ry)]])
    Rz = np.array([[np.cos(rz), -np.sin(rz), 0], [np.sin(rz), np.cos(rz), 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def test_measurement_chain_recovers_camera_rotations(V):
    """The whole measurement of consume_frame (FrameSourceWarp.cpp:403-438) on the GPU -- corners of the previous
    frame, pyramidal LK into the current one, rotation fit -- on frames with KNOWN camera rotations: a textured
    fisheye frame re-projected through the library's own fisheye -> fisheye warp for a sequence of camera poses.
    The measured inter-frame rotation must equal the pose change to 0.05 degrees (the warp's bilinear resampling
    and LK's own accuracy bound it), with most corners tracked and kept as inliers."""
    import torch
    w, h = 1920, 1080
    cam = V.get_preset_camera(4, w, h)
    out = V.get_output_camera(cam)
    fish_out = V.Camera.from_matrix(cam.K, w, h, model=1)      # the same fisheye camera on the output side
    ctx = V.WarpContext(cam, fish_out, out_size=(w, h), border=(0, 128, 128))
    luma, _ = _pair(h, w, 31)
    base = np.concatenate([luma, np.full((h // 2, w), 128, np.uint8)])
    src = torch.from_numpy(base).cuda().unsqueeze(0).contiguous()
    poses = [np.eye(3)]
    rng = np.random.default_rng(5)
    for _ in range(6):
        poses.append(_rot(*rng.normal(0, 0.5, 3)) @ poses[-1])
    ft = V.FlowTracker(w, h)
    frames = []
    for R in poses:
        dst = torch.empty_like(src)
        ctx.warp(src[0], dst[0], R)
        torch.cuda.synchronize()
        frames.append(dst[0, :h].contiguous())
    ft.push_frame(frames[0])
    worst = 0.0
    for k in range(1, len(poses)):
        ft.push_frame(frames[k])
        pts = ft.corners(0)                       # corners of the previous frame (:415-419)
        assert len(pts) >= 150
        nxt, st = ft.track(pts)
        assert st.mean() > 0.7     # points whose window leaves the re-projected frame are lost
        R, inl = V.guess_rotation(cam, out, pts[st], nxt[st], seed=k)
        assert inl >= 0.9 * st.sum() and inl >= 40
        # a pixel of view k with ray r shows the world ray poses[k] r, so rays move by poses[k]^-1 poses[k-1]
        want = poses[k].T @ poses[k - 1]
        err = np.rad2deg(np.arccos(np.clip((np.trace(R @ want.T) - 1) / 2, -1, 1)))
        worst = max(worst, err)
    assert worst < 0.05, worst
    ft.close()
    ctx.close()


def test_flow_errors(V):
    with pytest.raises(V.VawError):
        V.FlowTracker(8, 8)
    ft = V.FlowTracker(64, 64)
    with pytest.raises(V.VawError):
        ft.track(np.zeros((1, 2), np.float32))   # needs two frames
    ft.close()
