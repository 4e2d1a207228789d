"""CPU restatement of cv::goodFeaturesToTrack as the reference calls it
(/root/reference/opencv/FrameSourceWarp.cpp:228-230: goodFeaturesToTrack(image, corners, 200, 0.01, 30), i.e.
blockSize 3, gradientSize 3, minimum-eigenvalue response, no mask) and of the rotation fit of
guess_camera_rotation (:316-368).

TEST INFRASTRUCTURE ONLY (see oracle/vaw_oracle.h): only tests/ import this module.

The algorithms live in OpenCV's `imgproc` / `calib3d` modules (third-party, not under /root/reference;
meson.build:33 asks for opencv4 >= 4.5, the image carries opencv-python-headless 4.13.0).  They are restated
here and PINNED on the real functions (tests/test_oracle_corners.py, live, plus tests/golden/gftt_small.npz):

* corner response: cv::cornerMinEigenVal = Sobel derivatives scaled by 1 / (4 * 3 * 255), their products summed
  over 3 x 3 (BORDER_REFLECT_101), (a + c) - sqrt((a - c)^2 + b^2) with a = Sxx / 2, b = Sxy, c = Syy / 2.  fp32
  arithmetic; OpenCV's own result depends on its SIMD / IPP code path in the last bit (the order of the float
  sums), so the pin is to 3e-8 absolute (values reach ~0.1), not to the bit.
* selection: threshold at quality * max, 3 x 3 non-maximum suppression by equality with the dilated response,
  rows and columns 1 .. n-2 only, candidates in decreasing response (ties: higher address first), greedy
  minimum-distance filter on a grid of minDistance cells, stop at maxCorners.  Integer logic: equal to OpenCV's
  list whenever the responses order the same way.
"""
import numpy as np

F = np.float32


def _reflect(a, pad):
    h, w = a.shape
    yy = np.abs(np.arange(-pad, h + pad))
    yy = np.where(yy >= h, 2 * h - 2 - yy, yy)
    xx = np.abs(np.arange(-pad, w + pad))
    xx = np.where(xx >= w, 2 * w - 2 - xx, xx)
    return a[yy][:, xx]


def corner_min_eigen_val(img):
    """cv::cornerMinEigenVal(img, blockSize 3, ksize 3) for 8-bit input, fp32."""
    h, w = img.shape
    p = _reflect(img.astype(np.int32), 1)

    def s(dy, dx):
        return p[1 + dy:1 + dy + h, 1 + dx:1 + dx + w]
    # the Sobel sums are exact integers; one rounding when they are scaled
    scale = F(1.0 / (4.0 * 3.0 * 255.0))
    dx = ((s(-1, 1) - s(-1, -1)) + 2 * (s(0, 1) - s(0, -1)) + (s(1, 1) - s(1, -1))).astype(F) * scale
    dy = ((s(1, -1) - s(-1, -1)) + 2 * (s(1, 0) - s(-1, 0)) + (s(1, 1) - s(-1, 1))).astype(F) * scale

    def box(c):
        q = _reflect(c, 1)
        r = (q[:, :-2] + q[:, 1:-1]) + q[:, 2:]
        return (r[:-2] + r[1:-1]) + r[2:]
    a, b, c = box(dx * dx) * F(0.5), box(dx * dy), box(dy * dy) * F(0.5)
    return ((a + c) - np.sqrt((a - c) * (a - c) + b * b)).astype(F)


def select_corners(eig, max_corners=200, quality=0.01, min_distance=30.0):
    """The selection half of cv::goodFeaturesToTrack on a given response map -> (N, 2) float32 (x, y)."""
    h, w = eig.shape
    thr = F(np.float64(eig.max()) * quality)
    e = np.where(eig > thr, eig, F(0))           # THRESH_TOZERO
    q = np.full((h + 2, w + 2), -np.inf, F)      # cv::dilate: the border does not take part in the maximum
    q[1:-1, 1:-1] = e
    dil = np.max(np.stack([q[dy:dy + h, dx:dx + w] for dy in range(3) for dx in range(3)]), axis=0)
    cand = (e != 0) & (e == dil)
    cand[0, :] = cand[-1, :] = False
    cand[:, 0] = cand[:, -1] = False
    ys, xs = np.nonzero(cand)
    vals = e[ys, xs]
    order = np.lexsort((-(ys * w + xs), -vals.astype(np.float64)))  # value descending, then address descending
    cell = int(np.rint(min_distance))
    if min_distance < 1:
        pts = np.stack([xs[order], ys[order]], axis=1).astype(F)
        return pts[:max_corners] if max_corners > 0 else pts
    gw, gh = (w + cell - 1) // cell, (h + cell - 1) // cell
    grid = [[] for _ in range(gw * gh)]
    md2 = float(min_distance) * float(min_distance)
    out = []
    for i in order:
        x, y = int(xs[i]), int(ys[i])
        xc, yc = x // cell, y // cell
        good = True
        for yy in range(max(0, yc - 1), min(gh - 1, yc + 1) + 1):
            for xx in range(max(0, xc - 1), min(gw - 1, xc + 1) + 1):
                for (mx, my) in grid[yy * gw + xx]:
                    if (x - mx) * (x - mx) + (y - my) * (y - my) < md2:
                        good = False
                        break
                if not good:
                    break
            if not good:
                break
        if good:
            grid[yc * gw + xc].append((x, y))
            out.append((x, y))
            if max_corners > 0 and len(out) == max_corners:
                break
    return np.array(out, F).reshape(-1, 2)


def good_features_to_track(img, max_corners=200, quality=0.01, min_distance=30.0):
    return select_corners(corner_min_eigen_val(img), max_corners, quality, min_distance)


def guess_camera_rotation_cv2(K_in, D_in, K_out, pts_prev, pts_cur, seed=0):
    """guess_camera_rotation (FrameSourceWarp.cpp:316-368) step for step on the real OpenCV functions
    (needs cv2: used by tests/test_oracle_rotation.py and to generate tests/golden/rotation_cases.npz).
    The reference draws the depths from libc rand(); here from a seeded numpy generator.
    Returns (R (3, 3) float64, number of inliers)."""
    import cv2
    pts_prev = np.asarray(pts_prev, np.float32).reshape(-1, 1, 2)
    pts_cur = np.asarray(pts_cur, np.float32).reshape(-1, 1, 2)
    K_in = np.asarray(K_in, np.float64)
    K_out = np.asarray(K_out, np.float64)
    D_in = np.asarray(D_in, np.float64).reshape(4, 1)
    corners_output = cv2.fisheye.undistortPoints(pts_cur, K_in, D_in, R=np.eye(3), P=K_out)
    prev_identity = cv2.fisheye.undistortPoints(pts_prev, K_in, D_in).reshape(-1, 2)
    rng = np.random.default_rng(seed)
    scale = rng.random(len(prev_identity))
    obj = np.stack([prev_identity[:, 0] * scale, prev_identity[:, 1] * scale, scale], axis=1).astype(np.float64)
    try:
        ok, rvec, tvec, inliers = cv2.solvePnPRansac(obj, corners_output.reshape(-1, 2).astype(np.float64), K_out, np.zeros(4),
                                                     useExtrinsicGuess=False, iterationsCount=100, reprojectionError=8.0,
                                                     confidence=0.99)
    except cv2.error:
        return np.eye(3), 0
    if not ok or inliers is None:
        return np.eye(3), 0
    R, _ = cv2.Rodrigues(rvec)
    return R, len(inliers)
