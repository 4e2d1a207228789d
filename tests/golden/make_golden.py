"""Regenerate the golden fixtures under tests/golden/.

Run HERE (the authoring container), where cv2 (opencv-python-headless 4.13.0) is
importable:  python tests/golden/make_golden.py

What pins what:
  remap_cases.npz   outputs of cv2.remap (the real cv::remap the reference calls at
                    opencv/FrameSourceWarp.cpp:306-312) on random + adversarial maps.
                    PINS oracle/remap_ref.c.
  remap_cubic.npz   cv2.remap(INTER_CUBIC / INTER_LANCZOS4 / INTER_NEAREST) on the maps and sources of remap_cases.npz.
                    PINS oracle/remap_cubic_ref.c and the nearest = linear-on-rounded-map identity.
  fisheye_map.npz   cv2.fisheye.initUndistortRectifyMap(K_in, D=0, R=rot^T, P=K_out)
                    -- an independent implementation of createMap.cl's projection
                    from the reference's own dependency -- next to the oracle's map
                    bit patterns for the same inputs.  The reference has no golden
                    vectors for createMap.cl (parity unpinned); this is the anchor.
  camera_table.npz  output cameras computed by a Python transcription of
                    FrameSourceWarp.cpp:27-165 using cv2.fisheye.undistortPoints.
  cvt_nv12_bgr.npz  cv2.cvtColor(COLOR_YUV2BGR_NV12) on random + extreme samples.  PINS oracle/cvt_ref.c.
  nv12_small.npz    a small NV12 warp computed with cv2.remap on the oracle's luma
                    map and a numpy chroma map.  PINS oracle/nv12_warp_ref.c.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402
from tests.conftest import rotation_xyz  # noqa: E402


def cv_remap(src, mx, my, border):
    cn = 1 if src.ndim == 2 else src.shape[2]
    bv = tuple(float(b) for b in border[:cn])
    return cv2.remap(src, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT,
                     borderValue=bv if cn > 1 else bv[0])


def remap_cases():
    rng = np.random.default_rng(20260003)
    out = {}
    mx = rng.uniform(-3, 134, (60, 80)).astype(np.float32)
    my = rng.uniform(-3, 100, (60, 80)).astype(np.float32)
    mx[0, :8] = [np.nan, np.inf, -np.inf, 1e9, -1e9, 15.5, -0.5, 130.5]
    my[1, :8] = [np.nan, np.inf, -np.inf, 1e9, -1e9, 15.5, -0.5, 96.5]
    mx[2, :] = np.arange(80) * 0.5 - 2          # exact integers and halves
    my[2, :] = 5
    mx[3, :] = 10 + (np.arange(80) + 0.5) / 64  # ties at odd multiples of 1/64
    my[3, :] = 7 + (np.arange(80) + 0.5) / 64
    mx[4, :] = 130 + np.arange(80) / 32.0 - 1    # right edge walk
    my[4, :] = 96 + np.arange(80) / 32.0 - 1     # bottom edge walk
    mx[5, :] = -1.5 + np.arange(80) / 32.0       # left edge walk
    my[5, :] = -1.5 + np.arange(80) / 32.0
    out["map_x"], out["map_y"] = mx, my
    for cn in (1, 2, 3):
        src = rng.integers(0, 256, (97, 131) if cn == 1 else (97, 131, cn), dtype=np.uint8)
        out[f"src{cn}"] = src
        for bi, border in enumerate([(0, 0, 0), (80, 80, 80), (10, 128, 250)]):
            out[f"dst{cn}_b{bi}"] = cv_remap(src, mx, my, border)
    out["borders"] = np.array([(0, 0, 0), (80, 80, 80), (10, 128, 250)], np.uint8)
    np.savez_compressed(os.path.join(HERE, "remap_cases.npz"), **out)


def fisheye_map():
    w, h = 1920, 1080
    cam = O.get_preset_camera(4, w, h)
    outc = O.get_output_camera(cam, 1.0, False, 1.0)
    rows, cols = 48, 64
    # a window of the full map: offset the output principal point so the 64x48
    # patch sits well off-axis in the real geometry
    K_out = outc.K.copy()
    K_out[0, 2] -= 300.0
    K_out[1, 2] -= 200.0
    rot = rotation_xyz(2.0, -3.0, 1.5)
    k = O.intrinsics(cam.K, K_out)
    mx, my = O.create_map(k, rot, rows, cols)
    r32 = O.rot32(rot).astype(np.float64).reshape(3, 3)
    Kin32 = np.array([[k.src_focal_x, 0, k.src_center_x], [0, k.src_focal_y, k.src_center_y], [0, 0, 1]], np.float64)
    Kout32 = np.array([[k.map_focal_x, 0, k.map_center_x], [0, k.map_focal_y, k.map_center_y], [0, 0, 1]], np.float64)
    cvx, cvy = cv2.fisheye.initUndistortRectifyMap(Kin32, np.zeros(4), r32.T, Kout32, (cols, rows), cv2.CV_32FC1)
    np.savez_compressed(os.path.join(HERE, "fisheye_map.npz"),
                        intrinsics=np.array([getattr(k, f[0]) for f in k._fields_], np.float32),
                        rot=O.rot32(rot), oracle_x_bits=mx.view(np.uint32), oracle_y_bits=my.view(np.uint32),
                        cv_x=cvx, cv_y=cvy)


def py_output_camera(K, w, h, scale, crop, zoom):
    """Python transcription of FrameSourceWarp.cpp:88-165 on cv2.fisheye.undistortPoints."""
    pts = np.array([[0, 0], [0, h - 1], [w - 1, 0], [w - 1, h - 1],
                    [K[0, 2], 0], [w - 1, K[1, 2]], [K[0, 2], h - 1], [0, K[1, 2]]], np.float64)
    ext = cv2.fisheye.undistortPoints(pts.reshape(-1, 1, 2), K, np.zeros(4)).reshape(-1, 2)
    sel = ext[4:] if crop else ext
    max_x, min_x, max_y, min_y = sel[:, 0].max(), sel[:, 0].min(), sel[:, 1].max(), sel[:, 1].min()
    idx, idy = round(w - 1), round(h - 1)
    d = ext[3] - ext[0]
    odx, ody = int(np.rint(d[0])), int(np.rint(d[1]))
    scale = scale * np.sqrt(1. * idx * idx + idy * idy) / np.sqrt(1. * odx * odx + ody * ody)
    return np.array([scale, scale * -min_x / zoom, scale * -min_y / zoom,
                     int(scale * (max_x - min_x) / zoom), int(scale * (max_y - min_y) / zoom)])


def camera_table():
    rows = []
    for preset, (w, h), scale, crop, zoom in [
            (1, (1920, 1440), 0.5, False, 1.0), (4, (1920, 1080), 1.0, False, 1.0),
            (4, (2704, 1520), 1.0, False, 1.0), (4, (3840, 2160), 1.0, False, 1.0),
            (4, (5312, 2988), 1.0, False, 1.0), (0, (1920, 1440), 1.0, True, 1.0),
            (3, (1920, 1080), 1.0, False, 1.2), (2, (1920, 1440), 0.75, True, 1.1),
            (5, (2704, 1520), 1.0, False, 1.0)]:
        cam = O.get_preset_camera(preset, w, h)  # presets are plain arithmetic; K checked in test
        rows.append(np.concatenate([[preset, w, h, scale, crop, zoom],
                                    py_output_camera(cam.K, w, h, scale, crop, zoom)]))
    np.savez_compressed(os.path.join(HERE, "camera_table.npz"), table=np.array(rows))


def nv12_small():
    sw, sh, ow, oh = 96, 64, 80, 48
    src = O.synth_nv12(sw, sh, frame_index=3, white_noise=True)
    K_in = np.array([[48.0, 0, 47.3], [0, 47.5, 31.6], [0, 0, 1]])
    K_out = np.array([[30.0, 0, 39.5], [0, 30.0, 23.5], [0, 0, 1]])
    rot = rotation_xyz(4.0, -6.0, 3.0)
    k = O.intrinsics(K_in, K_out)
    mx, my = O.create_map(k, rot, oh, ow)
    # numpy restatement of the chroma map definition (nv12_warp_ref.c header)
    sx = ((mx[0::2, 0::2] + mx[0::2, 1::2]) + (mx[1::2, 0::2] + mx[1::2, 1::2])) * np.float32(0.25)
    sy = ((my[0::2, 0::2] + my[0::2, 1::2]) + (my[1::2, 0::2] + my[1::2, 1::2])) * np.float32(0.25)
    cx = ((sx - np.float32(0.5)) * np.float32(0.5)).astype(np.float32)
    cy = ((sy - np.float32(0.5)) * np.float32(0.5)).astype(np.float32)
    Y = src[:sh]
    UV = src[sh:].reshape(sh // 2, sw // 2, 2)
    out = {}
    for bi, border in enumerate([(0, 128, 128), (0, 0, 0), (33, 77, 201)]):
        dy = cv_remap(Y, mx, my, (border[0],))
        duv = cv_remap(UV, cx, cy, border[1:])
        out[f"dst_b{bi}"] = np.concatenate([dy, duv.reshape(oh // 2, ow)], axis=0)
    np.savez_compressed(os.path.join(HERE, "nv12_small.npz"), src=src, K_in=K_in, K_out=K_out, rot=rot,
                        borders=np.array([(0, 128, 128), (0, 0, 0), (33, 77, 201)], np.uint8), **out)


def cvt_nv12_bgr():
    """cv2.cvtColor(COLOR_YUV2BGR_NV12) -- the call at opencv/FrameSourceWarp.cpp:399-401 -- on random
    NV12 plus the extreme (Y, U, V) combinations.  PINS oracle/cvt_ref.c."""
    rng = np.random.default_rng(20260004)
    w, h = 64, 48
    nv = rng.integers(0, 256, (h * 3 // 2, w), dtype=np.uint8)
    ext = np.array([0, 1, 15, 16, 17, 127, 128, 129, 234, 235, 236, 254, 255], np.uint8)
    k = 0
    for yv in ext:                       # first rows: extreme combinations
        for uv in ext[::3]:
            if k >= w * 8:
                break
            nv[(k // w), k % w] = yv
            nv[h + (k // w) // 2, (k % w) & ~1] = uv
            nv[h + (k // w) // 2, ((k % w) & ~1) + 1] = ext[(k * 7) % len(ext)]
            k += 1
    np.savez_compressed(os.path.join(HERE, "cvt_nv12_bgr.npz"), nv12=nv,
                        bgr=cv2.cvtColor(nv, cv2.COLOR_YUV2BGR_NV12))


def remap_cubic():
    """INTER_CUBIC and INTER_NEAREST outputs of the real cv2.remap for the maps / sources of remap_cases.npz."""
    g = np.load(os.path.join(HERE, "remap_cases.npz"))
    out = {}
    for cn in (1, 2, 3):
        src = g[f"src{cn}"]
        for bi, border in enumerate(g["borders"]):
            bv = tuple(float(b) for b in border[:cn])
            for name, flag in (("cubic", cv2.INTER_CUBIC), ("nearest", cv2.INTER_NEAREST), ("lanczos", cv2.INTER_LANCZOS4)):
                out[f"{name}{cn}_b{bi}"] = cv2.remap(src, g["map_x"], g["map_y"], flag, borderMode=cv2.BORDER_CONSTANT,
                                                     borderValue=bv if cn > 1 else bv[0])
    np.savez_compressed(os.path.join(HERE, "remap_cubic.npz"), **out)


if __name__ == "__main__" and not any(a.startswith("--") for a in sys.argv[1:]):
    cvt_nv12_bgr()
    remap_cases()
    remap_cubic()
    fisheye_map()
    camera_table()
    nv12_small()
    print("golden fixtures written to", HERE)


def make_lk_fixture():
    """tests/golden/lk_small.npz: a 160x120 textured frame pair, 26 points (two of them at / outside the frame edge)
    and the real cv2.calcOpticalFlowPyrLK / cv2.pyrDown answers (run: python tests/golden/make_golden.py --lk)."""
    import cv2
    from tests.test_oracle_flow import _textured_pair
    prev, nxt = _textured_pair(120, 160, 11, angle_deg=0.8, shift=(1.7, -1.2))
    pts = cv2.goodFeaturesToTrack(prev, 24, 0.01, 12).reshape(-1, 2).astype(np.float32)
    pts = np.concatenate([pts, np.array([[1.0, 1.0], [-30.0, 5.0]], np.float32)])
    ref, st, _ = cv2.calcOpticalFlowPyrLK(prev, nxt, pts, None)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lk_small.npz"), prev=prev, next=nxt, pts=pts,
                        cv_next=ref.reshape(-1, 2), cv_status=st.reshape(-1), cv_pyr1=cv2.pyrDown(prev),
                        cv_version=np.array(cv2.__version__))


def make_gftt_fixture():
    """tests/golden/gftt_small.npz: a 200x150 textured frame with the real cv2.cornerMinEigenVal map and the real
    cv2.goodFeaturesToTrack(image, 200, 0.01, 30) list (run: python tests/golden/make_golden.py --gftt)."""
    import cv2
    from tests.test_oracle_flow import _textured_pair
    img, _ = _textured_pair(150, 200, 21)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gftt_small.npz"), image=img,
                        cv_response=cv2.cornerMinEigenVal(img, 3, ksize=3),
                        cv_corners=cv2.goodFeaturesToTrack(img, 200, 0.01, 30).reshape(-1, 2),
                        cv_version=np.array(cv2.__version__))


def make_rotation_fixture():
    """tests/golden/rotation_cases.npz: synthetic point pairs (tests/test_oracle_rotation.make_case) and what the
    reference's guess_camera_rotation computes for them on the real cv2.fisheye.undistortPoints / cv2.solvePnPRansac
    (oracle/gftt_ref.guess_camera_rotation_cv2); run: python tests/golden/make_golden.py --rotation."""
    from oracle import gftt_ref as G
    from tests.test_oracle_rotation import CASES, make_case
    out = {"n_cases": np.array(len(CASES))}
    for i, (seed, kw) in enumerate(CASES):
        cam, oc, prev, cur, R_true, _ = make_case(seed, **kw)
        R, inl = G.guess_camera_rotation_cv2(cam.K, cam.distortion, oc.K, prev, cur, seed=seed)
        out.update({f"K_in_{i}": cam.K, f"K_out_{i}": oc.K, f"size_{i}": np.array(cam.size), f"prev_{i}": prev, f"cur_{i}": cur,
                    f"cv_R_{i}": R, f"cv_inliers_{i}": np.array(inl), f"true_R_{i}": R_true})
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "rotation_cases.npz"), **out)


if __name__ == "__main__" and "--rotation" in sys.argv:
    make_rotation_fixture()
if __name__ == "__main__" and "--lk" in sys.argv:
    make_lk_fixture()
if __name__ == "__main__" and "--gftt" in sys.argv:
    make_gftt_fixture()
