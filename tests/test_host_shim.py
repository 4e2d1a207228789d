"""The C++ FrameSource / FrameSourceWarp shim (video_annotator_b200/host), which mirrors
/root/reference/opencv/FrameSource.hpp and FrameSourceWarp.hpp over the C-ABI."""
import os
import re
import subprocess
import zlib

import numpy as np
import pytest

from tests.conftest import ROOT

HOST_TEST = os.path.join(ROOT, "tests", "host", "test_state_machine")


def _build_state_machine_test():
    from video_annotator_b200 import _build
    _build.build()
    pkg = os.path.join(ROOT, "video_annotator_b200")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-Werror", "-o", HOST_TEST,
                           os.path.join(ROOT, "tests", "host", "test_state_machine.cpp"),
                           os.path.join(pkg, "host", "FrameSourceWarp.cpp"),
                           "-L" + pkg, "-l:libvaw.so", "-Wl,-rpath," + pkg, "-pthread"])


def test_state_machine_on_cpu():
    """Frame-0 drop, look-ahead latency, EOF drain, peek == pull, rotation smoothing, cameras
    (opencv/FrameSourceWarp.cpp:397-480) -- warp_frame overridden, so no GPU is touched."""
    _build_state_machine_test()
    out = subprocess.run([HOST_TEST], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().startswith("OK:")


def test_shim_and_demo_build_with_werror():
    from video_annotator_b200 import _build
    demo = _build.build_host_shim(force=True)
    assert os.access(demo, os.X_OK)
    # the shim is plain C++ over the C-ABI: no CUDA or OpenCV headers
    for fn in ("FrameSource.hpp", "FrameSourceWarp.hpp", "FrameSourceWarp.cpp", "vaw_demo.cpp"):
        text = open(os.path.join(ROOT, "video_annotator_b200", "host", fn)).read()
        assert "cuda_runtime" not in text and "#include <opencv" not in text and "#include <cuda" not in text


@pytest.mark.gpu
def test_demo_chain_on_gpu(oracle):
    """DisplayImage.cpp's loop over the shim: every emitted frame equals the Python API's warp of
    the same synthetic frame with the rotation the shim reports, and the frames come out in order
    with frame 0 dropped."""
    import torch
    import video_annotator_b200 as V
    from video_annotator_b200 import _build
    demo = _build.build_host_shim()
    w, h, n, radius = 1920, 1080, 12, 3
    out = subprocess.run([demo, str(w), str(h), str(n), str(radius), "0.5"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    ow, oh = [int(v) for v in lines[0].split()[1:]]
    assert (ow, oh) == (1758, 998)
    frames = [ln.split() for ln in lines[1:]]
    assert [int(f[1]) for f in frames] == list(range(1, n))           # frame 0 is never emitted
    cam = V.get_preset_camera(V.warp.GOPRO_H4B_WIDE169_MEASURED, w, h)
    ctx = V.WarpContext(cam, V.get_output_camera(cam), out_size=(ow, oh), border=(0, 128, 128))
    src = torch.empty(ctx.frame_shape("src"), dtype=torch.uint8, device="cuda")
    dst = torch.empty(ctx.frame_shape("dst"), dtype=torch.uint8, device="cuda")
    for f in frames:
        idx, crc = int(f[1]), int(f[3], 16)
        rot = np.array([float(v) for v in f[5:14]]).reshape(3, 3)
        assert np.allclose(rot @ rot.T, np.eye(3), atol=1e-9)
        V.synth_nv12(src, w, h, 1, first_index=idx)
        ctx.warp(src, dst, rot)
        torch.cuda.synchronize()
        assert zlib.crc32(dst.cpu().numpy().tobytes()) == crc, idx
    assert re.search(r"\d+ frames", out.stderr)
    ctx.close()
