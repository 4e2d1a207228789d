"""The C-ABI library (include/vaw.h -> video_annotator_b200/libvaw.so) on a CPU-only host:
it builds, loads, exports every declared symbol, its host-only entry points work, and the
compute entry points fail loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests.conftest import GOLDEN, ROOT


def _declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            text = open(os.path.join(ROOT, "include", fn)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            names |= set(re.findall(r"\b(vaw_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_exports_every_declared_symbol():
    from video_annotator_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 17
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/*.h but not exported"
    # the ctypes mirror binds exactly the declared surface
    assert set(_lib.SIGNATURES) == declared


def test_params_struct_layout_matches_header():
    from video_annotator_b200 import _lib
    # 8 doubles + 6 int32 + 4 bytes + int32 + 4 floats + 3 int32 = 124, padded to the 8-byte alignment;
    # vaw_api.cu static_asserts the same numbers on the C side
    assert C.sizeof(_lib.VawParams) == 128
    assert C.sizeof(_lib.VawCamera) == 4 * 4 + 9 * 8 + 4 * 8 == 120


def test_strerror_and_frame_bytes():
    from video_annotator_b200 import _lib
    lib = _lib.load()
    assert lib.vaw_strerror(0) == b"ok"
    assert b"CUDA" in lib.vaw_strerror(-3)
    assert lib.vaw_frame_bytes(0, 3840, 2160, 3840) == 3840 * 2160 * 3 // 2  # NV12
    assert lib.vaw_frame_bytes(1, 1759, 998, 1759 * 3) == 1759 * 998 * 3
    assert lib.vaw_frame_bytes(2, 100, 50, 128) == 128 * 50


def test_camera_producers_match_golden_table_and_oracle(oracle):
    """vaw_get_preset_camera / vaw_get_output_camera (FrameSourceWarp.cpp:27-165)."""
    import video_annotator_b200 as V
    t = np.load(os.path.join(GOLDEN, "camera_table.npz"))["table"]
    for row in t:
        preset, w, h, scale, crop, zoom, f, cx, cy, ow, oh = row
        cam = V.get_preset_camera(int(preset), int(w), int(h))
        ocam = oracle.get_preset_camera(int(preset), int(w), int(h))
        assert np.array_equal(cam.K, ocam.K)
        out = V.get_output_camera(cam, scale, bool(crop), zoom)
        assert out.K[0, 0] == pytest.approx(f, rel=1e-12)
        assert out.K[0, 2] == pytest.approx(cx, rel=1e-10)
        assert out.K[1, 2] == pytest.approx(cy, rel=1e-10)
        assert out.size == (int(ow), int(oh))
        oout = oracle.get_output_camera(ocam, scale, bool(crop), zoom)
        assert np.allclose(out.K, oout.K, rtol=1e-13, atol=0)
        assert out.size == (oout.width, oout.height)


def test_camera_producers_reject_bad_arguments():
    from video_annotator_b200 import _lib
    lib = _lib.load()
    cam = _lib.VawCamera()
    assert lib.vaw_get_preset_camera(99, 1920, 1080, C.byref(cam)) == -2
    assert lib.vaw_get_preset_camera(0, 0, 1080, C.byref(cam)) == -2


def test_create_validates_before_touching_cuda():
    import video_annotator_b200 as V
    cam = V.get_preset_camera(4, 1920, 1080)
    out = V.get_output_camera(cam)
    with pytest.raises(V.VawError) as e:  # cv::INTER_AREA (FrameSourceWarp.hpp:90): not a remap filter, not implemented
        V.WarpContext(cam, out, interpolation=3)
    assert e.value.code == -4
    with pytest.raises(V.VawError) as e:  # no table filter on variant POLY (GATHER and TILED carry them)
        V.WarpContext(cam, out, interpolation=V.INTER_CUBIC, variant=2)
    assert e.value.code == -4
    with pytest.raises(V.VawError) as e:  # NV12 needs even sizes
        V.WarpContext(cam, out, out_size=(1759, 998))
    assert e.value.code == -2
    with pytest.raises(V.VawError) as e:  # `short` indices (createMap.cl:10-11)
        V.WarpContext(cam, out, out_size=(40000, 998))
    assert e.value.code == -2


def test_cubic_table_equals_the_oracles(oracle):
    """The library builds cv::remap's INTER_CUBIC weight table itself (csrc/vaw_cubic.cuh, host code); it
    must equal the oracle's restatement (pinned on cv2.remap) entry for entry."""
    from video_annotator_b200 import _lib
    tab = np.zeros(32 * 32 * 16, np.int16)
    assert _lib.load().vaw_cubic_table(tab.ctypes.data_as(C.POINTER(C.c_int16))) == 0
    assert np.array_equal(tab.reshape(32, 32, 4, 4), oracle.cubic_table())
    tab8 = np.zeros(32 * 32 * 64, np.int16)
    assert _lib.load().vaw_lanczos4_table(tab8.ctypes.data_as(C.POINTER(C.c_int16))) == 0
    assert np.array_equal(tab8.reshape(32, 32, 8, 8), oracle.lanczos4_table())


def test_no_cpu_fallback_without_device():
    """The product path must fail loudly when there is no GPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import video_annotator_b200 as V
    cam = V.get_preset_camera(4, 1920, 1080)
    out = V.get_output_camera(cam)
    with pytest.raises(V.VawError) as e:
        V.WarpContext(cam, out)
    assert e.value.code == -3
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing in the package or include/ may reference it."""
    pkg = os.path.join(ROOT, "video_annotator_b200")
    for base, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(base, fn)).read()
                assert "import oracle" not in text and "from oracle" not in text, fn
                assert "liboracle" not in text and "vaw_oracle_" not in text, fn


def test_output_camera_honours_distortion_like_cv_fisheye():
    """get_output_camera passes distortion_coefficients to fisheye::undistortPoints
    (FrameSourceWarp.cpp:93-110); the library's Newton inversion must agree with the real
    cv2.fisheye.undistortPoints on the 8 probe points, through the derived intrinsics."""
    cv2 = pytest.importorskip("cv2")
    import video_annotator_b200 as V
    dist = np.array([0.05, -0.02, 0.01, -0.003])
    cam0 = V.get_preset_camera(1, 1920, 1440)
    cam = V.Camera.from_matrix(cam0.K, 1920, 1440, model=1, distortion=dist)
    K = cam0.K
    w1, h1 = 1919.0, 1439.0
    probes = np.array([[0, 0], [0, h1], [w1, 0], [w1, h1], [K[0, 2], 0], [w1, K[1, 2]], [K[0, 2], h1], [0, K[1, 2]]],
                      np.float64).reshape(-1, 1, 2)
    e = cv2.fisheye.undistortPoints(probes, K, dist).reshape(-1, 2)
    for crop in (False, True):
        pts = e[4:] if crop else e
        mn, mx = pts.min(0), pts.max(0)
        od = np.rint(e[3] - e[0])
        f = 0.5 * np.hypot(np.rint(w1), np.rint(h1)) / np.hypot(od[0], od[1])
        out = V.get_output_camera(cam, 0.5, crop, 1.0)
        assert out.K[0, 0] == pytest.approx(f, rel=1e-12)
        assert out.K[0, 2] == pytest.approx(f * -mn[0], rel=1e-9)
        assert out.K[1, 2] == pytest.approx(f * -mn[1], rel=1e-9)
        assert out.size == (int(f * (mx[0] - mn[0])), int(f * (mx[1] - mn[1])))
    # and it differs from the zero-distortion camera (the coefficients are not ignored)
    assert V.get_output_camera(cam, 0.5).size != V.get_output_camera(cam0, 0.5).size


def test_camera_models_select_the_projection_pair():
    """CameraModel (FrameSourceWarp.hpp:23-26) is carried into vaw_params::projection; createMap.cl's pair
    (FISHEYE in, RECTILINEAR out) is 0.  The reference's get_output_camera is fisheye-only."""
    import video_annotator_b200 as V
    from video_annotator_b200 import _lib
    lib = _lib.load()
    cam = V.get_preset_camera(4, 1920, 1080)
    rect_in = V.Camera.from_matrix(cam.K, 1920, 1080, model=0)
    out = _lib.VawCamera()
    assert lib.vaw_get_output_camera(C.byref(rect_in._c), 1.0, 0, 1.0, C.byref(out)) == -4
    good_out = V.get_output_camera(cam)
    fish_out = V.Camera.from_matrix(good_out.K, 1758, 998, model=1)
    p = _lib.VawParams()
    for cin, cout, want in ((cam, good_out, 0), (rect_in, good_out, 1), (cam, fish_out, 2), (rect_in, fish_out, 3)):
        assert lib.vaw_params_from_cameras(C.byref(cin._c), C.byref(cout._c), 0, C.byref(p)) == 0
        assert p.projection == want
    bad = V.Camera.from_matrix(cam.K, 1920, 1080, model=7)
    assert lib.vaw_params_from_cameras(C.byref(bad._c), C.byref(good_out._c), 0, C.byref(p)) == -2


def test_params_from_cameras_sets_the_reference_defaults():
    """A zero-initialised vaw_params must come back with INTER_LINEAR (FrameSourceWarp.hpp:90), not 0 = NEAREST."""
    import video_annotator_b200 as V
    from video_annotator_b200 import _lib
    cam = V.get_preset_camera(4, 1920, 1080)
    out = V.get_output_camera(cam)
    p = _lib.VawParams()
    assert _lib.load().vaw_params_from_cameras(C.byref(cam._c), C.byref(out._c), 0, C.byref(p)) == 0
    assert p.interpolation == V.INTER_LINEAR and p.variant == 0
    assert (p.out_width, p.out_height) == (1758, 998)
