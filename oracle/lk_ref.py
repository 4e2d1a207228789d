"""CPU restatement of cv::calcOpticalFlowPyrLK as the reference calls it
(/root/reference/opencv/FrameSourceWarp.cpp:242-270: defaults only -- winSize 21x21, maxLevel 3,
TermCriteria(COUNT + EPS, 30, 0.01), flags 0, minEigThreshold 1e-4), plus cv::pyrDown and the Scharr
derivative it builds its pyramids from.

TEST INFRASTRUCTURE ONLY (see oracle/vaw_oracle.h): only tests/ import this module.

The algorithm lives in OpenCV's `video` module (third-party, not under /root/reference; the reference's
meson.build:33 asks for opencv4 >= 4.5; the image carries opencv-python-headless 4.13.0).  It is restated
here from its published fixed-point scheme and PINNED on the real cv2.calcOpticalFlowPyrLK / cv2.pyrDown:
tests/test_oracle_flow.py (live, <= 1e-4 px, identical status) and tests/golden/lk_small.npz.

numpy for the window arithmetic; the loops over points, levels and iterations are plain Python (a few
hundred points at most)."""
import numpy as np

WIN = 21
W_BITS = 14


def reflect101(i, n):
    i = np.abs(i)
    return np.where(i >= n, 2 * n - 2 - i, i)


def pyr_down(img):
    """cv::pyrDown on 8-bit: 5x5 binomial kernel in integers, (sum + 128) >> 8, BORDER_REFLECT_101."""
    h, w = img.shape
    dh, dw = (h + 1) // 2, (w + 1) // 2
    k = np.array([1, 4, 6, 4, 1], np.int64)
    ys = reflect101(2 * np.arange(dh)[:, None] + np.arange(-2, 3)[None, :], h)   # (dh, 5)
    xs = reflect101(2 * np.arange(dw)[:, None] + np.arange(-2, 3)[None, :], w)   # (dw, 5)
    a = img.astype(np.int64)
    rows = (a[:, xs] * k).sum(axis=2)            # (h, dw): horizontal pass
    out = (rows[ys, :] * k[None, :, None]).sum(axis=1)
    return ((out + 128) >> 8).astype(np.uint8)


def scharr_deriv(img):
    """calcSharrDeriv: int16 (dx, dy), the image extended by BORDER_REFLECT_101."""
    h, w = img.shape
    a = img.astype(np.int32)
    yy = reflect101(np.arange(-1, h + 1), h)
    xx = reflect101(np.arange(-1, w + 1), w)
    p = a[yy][:, xx]

    def s(dy, dx):
        return p[1 + dy:1 + dy + h, 1 + dx:1 + dx + w]
    dx = 3 * (s(-1, 1) - s(-1, -1)) + 10 * (s(0, 1) - s(0, -1)) + 3 * (s(1, 1) - s(1, -1))
    dy = 3 * (s(1, -1) - s(-1, -1)) + 10 * (s(1, 0) - s(-1, 0)) + 3 * (s(1, 1) - s(-1, 1))
    return dx.astype(np.int16), dy.astype(np.int16)


def build_pyramid(img, max_level=3):
    """cv::buildOpticalFlowPyramid: a level is added while the next size stays above the window."""
    pyr = [np.ascontiguousarray(img)]
    for _ in range(max_level):
        h, w = pyr[-1].shape
        if (w + 1) // 2 <= WIN or (h + 1) // 2 <= WIN:
            break
        pyr.append(pyr_down(pyr[-1]))
    return pyr


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _pad(a, pad, reflect):
    h, w = a.shape
    if reflect:
        yy = reflect101(np.arange(-pad, h + pad), h)
        xx = reflect101(np.arange(-pad, w + pad), w)
        return a[yy][:, xx].astype(np.int64)
    out = np.zeros((h + 2 * pad, w + 2 * pad), np.int64)
    out[pad:pad + h, pad:pad + w] = a
    return out


def _weights(a, b):
    f = np.float32
    s = f(1 << W_BITS)
    w00 = int(np.rint((f(1) - a) * (f(1) - b) * s))
    w01 = int(np.rint(a * (f(1) - b) * s))
    w10 = int(np.rint((f(1) - a) * b * s))
    return w00, w01, w10, (1 << W_BITS) - w00 - w01 - w10


def calc_optical_flow_pyr_lk(prev, nxt, pts, max_level=3, iters=30, eps=0.01, min_eig_threshold=1e-4):
    """prev, nxt: (H, W) uint8; pts: (N, 2) float32 (x, y).  Returns (next_pts (N, 2) float32, status (N,) bool)."""
    f = np.float32
    pts = np.asarray(pts, np.float32).reshape(-1, 2)
    pyr_p, pyr_n = build_pyramid(prev, max_level), build_pyramid(nxt, max_level)
    top = len(pyr_p) - 1
    half = f((WIN - 1) * 0.5)
    status = np.ones(len(pts), bool)
    nextpts = np.zeros_like(pts)
    scale = f(1.0 / (1 << 20))
    for level in range(top, -1, -1):
        I, J = pyr_p[level], pyr_n[level]
        h, w = I.shape
        dIx, dIy = scharr_deriv(I)
        pad = WIN + 1
        Ip, Jp = _pad(I, pad, True), _pad(J, pad, True)
        dxp, dyp = _pad(dIx, pad, False), _pad(dIy, pad, False)

        def interp(A, x0, y0, wt, n):
            w00, w01, w10, w11 = wt
            return _descale(A[y0:y0 + WIN, x0:x0 + WIN] * w00 + A[y0:y0 + WIN, x0 + 1:x0 + WIN + 1] * w01 +
                            A[y0 + 1:y0 + WIN + 1, x0:x0 + WIN] * w10 + A[y0 + 1:y0 + WIN + 1, x0 + 1:x0 + WIN + 1] * w11, n)

        for i, pt in enumerate(pts):
            prev_pt = pt * f(1.0 / (1 << level))
            next_pt = prev_pt.copy() if level == top else nextpts[i] * f(2.0)
            nextpts[i] = next_pt
            pp = prev_pt - half
            ip = np.floor(pp).astype(np.int64)
            if ip[0] < -WIN or ip[0] >= w or ip[1] < -WIN or ip[1] >= h:
                if level == 0:
                    status[i] = False
                continue
            wt = _weights(f(pp[0] - f(ip[0])), f(pp[1] - f(ip[1])))
            x0, y0 = int(ip[0]) + pad, int(ip[1]) + pad
            Iw, Ix, Iy = interp(Ip, x0, y0, wt, W_BITS - 5), interp(dxp, x0, y0, wt, W_BITS), interp(dyp, x0, y0, wt, W_BITS)
            # the sums are exact integers (OpenCV accumulates the same integers in float lanes)
            A11, A12, A22 = f(int((Ix * Ix).sum())) * scale, f(int((Ix * Iy).sum())) * scale, f(int((Iy * Iy).sum())) * scale
            D = A11 * A22 - A12 * A12
            dd = A11 - A22
            min_eig = (A22 + A11 - np.sqrt(dd * dd + f(4.0) * A12 * A12)) / f(2 * WIN * WIN)
            if min_eig < f(min_eig_threshold) or D < np.finfo(np.float32).eps:
                if level == 0:
                    status[i] = False
                continue
            D = f(1.0) / D
            npt = next_pt - half
            prev_delta = np.zeros(2, np.float32)
            for j in range(iters):
                inp = np.floor(npt).astype(np.int64)
                if inp[0] < -WIN or inp[0] >= w or inp[1] < -WIN or inp[1] >= h:
                    if level == 0:
                        status[i] = False
                    break
                wj = _weights(f(npt[0] - f(inp[0])), f(npt[1] - f(inp[1])))
                diff = interp(Jp, int(inp[0]) + pad, int(inp[1]) + pad, wj, W_BITS - 5) - Iw
                b1, b2 = f(int((diff * Ix).sum())) * scale, f(int((diff * Iy).sum())) * scale
                delta = np.array([(A12 * b2 - A22 * b1) * D, (A12 * b1 - A11 * b2) * D], np.float32)
                npt = npt + delta
                nextpts[i] = npt + half
                if delta[0] * delta[0] + delta[1] * delta[1] <= f(eps * eps):
                    break
                if j > 0 and abs(delta[0] + prev_delta[0]) < 0.01 and abs(delta[1] + prev_delta[1]) < 0.01:
                    nextpts[i] = nextpts[i] - delta * f(0.5)
                    break
                prev_delta = delta
    return nextpts, status
