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


def test_shim_rotation_chain_equals_the_python_workload_generator(tmp_path):
    """The two restatements of FrameSourceWarp.cpp:441-475 -- the C++ shim's accumulate / Savitzky-Golay /
    SO(3)-projection / inverse chain (batched look-ahead on) and video_annotator_b200/rotations.py, which
    generates the benchmark's rotations -- agree on the same gyro increments to 1e-12."""
    from video_annotator_b200 import rotations as R
    _build_state_machine_test()
    rng = np.random.default_rng(7)
    for n, radius in ((75, 30), (12, 3)):
        inc = rng.normal(0.0, np.deg2rad(0.6), (n, 3))
        path = tmp_path / f"inc_{n}.txt"
        np.savetxt(path, inc, fmt="%.17g")
        env = dict(os.environ, VAW_ROT_INCREMENTS=str(path), VAW_ROT_RADIUS=str(radius))
        out = subprocess.run([HOST_TEST], capture_output=True, text=True, timeout=120, env=env)
        assert out.returncode == 0, out.stdout + out.stderr
        got = np.array([[float(v) for v in ln.split()[1:]] for ln in out.stdout.splitlines() if ln.startswith("ROT")])
        want = R.rotations_from_increments(inc, radius).reshape(n, 9)
        assert got.shape == want.shape
        assert np.abs(got - want).max() < 1e-12


def test_shim_and_demo_build_with_werror():
    from video_annotator_b200 import _build
    demo = _build.build_host_shim(force=True)
    assert os.access(demo, os.X_OK)
    # the shim is plain C++ over the C-ABI: no CUDA or OpenCV headers
    for fn in ("FrameSource.hpp", "FrameSourceWarp.hpp", "FrameSourceWarp.cpp", "vaw_demo.cpp"):
        text = open(os.path.join(ROOT, "video_annotator_b200", "host", fn)).read()
        assert "cuda_runtime" not in text and "#include <opencv" not in text and "#include <cuda" not in text


@pytest.mark.gpu
@pytest.mark.parametrize("warp_batch", [1, 5])
def test_demo_chain_on_gpu(oracle, warp_batch):
    """DisplayImage.cpp's loop over the shim: every emitted frame equals the Python API's warp of
    the same synthetic frame with the rotation the shim reports, and the frames come out in order
    with frame 0 dropped."""
    import torch
    import video_annotator_b200 as V
    from video_annotator_b200 import _build
    demo = _build.build_host_shim()
    w, h, n, radius = 1920, 1080, 12, 3
    # warp_batch 5: look-ahead frames are warped five per launch from the pooled slab (vaw_warp_batch +
    # vaw_bind_clip); frames, order, rotations and bytes must not change
    out = subprocess.run([demo, str(w), str(h), str(n), str(radius), "0.5", str(warp_batch)], capture_output=True,
                         text=True, timeout=300)
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


@pytest.mark.gpu
def test_demo_bench_mode_reports_both_pull_patterns():
    from video_annotator_b200 import _build
    demo = _build.build_host_shim()
    out = subprocess.run([demo, "--bench", "300", "16", "1920", "1080"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    import json
    d = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][0])
    assert d["frames"] == 299 and d["warp_batch"] == 16
    assert d["fps_batched"] > 0 and d["fps_one_warp_per_call"] > 0
    assert d["batched_launches"] <= 299 // 16 + 8   # frames really went through batched launches (a few runs split at the slab's wrap)


@pytest.mark.gpu
def test_demo_chain_with_the_optical_flow_measurement():
    """vaw_demo --flow: FrameSourceWarp fed by OpticalFlowRotationSource (the reference's own measurement,
    FrameSourceWarp.cpp:403-438: corners every 20 frames or below 150 points, pyramidal LK, rotation fit, the
    40-inlier rule) on a synthetic camera whose rotations are known.  Every frame after the first is fitted,
    to 0.05 degrees, and the chain emits every frame but the first (which the reference drops, :403-406)."""
    import json
    from video_annotator_b200 import _build
    demo = _build.build_host_shim()
    n = 30
    out = subprocess.run([demo, "--flow", str(n), "1920", "1080", "0.5"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().split("\n")
    d = json.loads(lines[-1])
    assert d["emitted"] == n - 1
    assert d["fits"] == n - 1
    assert d["min_inliers"] >= 100
    assert d["worst_error_deg"] < 0.05, lines
