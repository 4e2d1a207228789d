"""Camera producers (oracle/camera_ref.c restating opencv/FrameSourceWarp.cpp:27-165)."""
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN


def test_output_camera_golden_table(oracle):
    """Table generated with cv2.fisheye.undistortPoints (tests/golden/make_golden.py)."""
    t = np.load(os.path.join(GOLDEN, "camera_table.npz"))["table"]
    for row in t:
        preset, w, h, scale, crop, zoom, f, cx, cy, ow, oh = row
        cam = oracle.get_preset_camera(int(preset), int(w), int(h))
        out = oracle.get_output_camera(cam, scale, bool(crop), zoom)
        assert out.K[0, 0] == pytest.approx(f, rel=1e-12)
        assert out.K[1, 1] == pytest.approx(f, rel=1e-12)
        assert out.K[0, 2] == pytest.approx(cx, rel=1e-10)
        assert out.K[1, 2] == pytest.approx(cy, rel=1e-10)
        assert (out.width, out.height) == (int(ow), int(oh))


def test_output_camera_survey_values(oracle):
    """SURVEY 8 a4 derived outputs."""
    expect = {
        (1, 1920, 1440, 0.5): (187.299, 507.55, 362.08, 988, 744),
        (4, 1920, 1080, 1.0): (492.281, 896.27, 485.84, 1759, 998),
        (4, 2704, 1520, 1.0): (693.310, 1264.61, 684.87, 2483, 1408),
        (4, 3840, 2160, 1.0): (984.866, 1793.86, 971.98, 3524, 1999),
        (4, 5312, 2988, 1.0): (1362.514, 2482.01, 1344.69, 4877, 2766),
    }
    for (preset, w, h, scale), (f, cx, cy, ow, oh) in expect.items():
        out = oracle.get_output_camera(oracle.get_preset_camera(preset, w, h), scale, False, 1.0)
        assert out.K[0, 0] == pytest.approx(f, abs=1e-3)
        assert out.K[0, 2] == pytest.approx(cx, abs=1e-2)
        assert out.K[1, 2] == pytest.approx(cy, abs=1e-2)
        assert (out.width, out.height) == (ow, oh)


def test_preset_quirks(oracle):
    """Published FOVs truncate to int (FrameSourceWarp.cpp:22-25); MEASURED fx scales by height (:54)."""
    cam = oracle.get_preset_camera(0, 1920, 1440)
    assert cam.K[0, 0] == pytest.approx(1920 / (122 * np.pi / 180))
    assert cam.K[1, 1] == pytest.approx(1440 / (94 * np.pi / 180))
    assert cam.K[0, 2] == (1920 - 1.0) / 2 and cam.K[1, 2] == (1440 - 1.0) / 2
    cam = oracle.get_preset_camera(4, 3840, 2160)
    assert cam.K[0, 0] == pytest.approx(1392.49 * 2160 / 1520)
    assert cam.K[0, 2] == pytest.approx(1361.80 * 3840 / 2704)
    assert cam.model == 1 and list(cam.dist) == [0, 0, 0, 0]
