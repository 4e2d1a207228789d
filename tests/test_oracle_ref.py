"""Pins the coordinate oracle to the reference's own kernel.

oracle/_ref/libcreatemap_ref.so is /root/reference/opencv/createMap.cl compiled UNMODIFIED with
gcc behind oracle/ref_build/cl_shim.h, run over the NDRange and argument binding of
FrameSourceWarp.cpp:275-304.  The transcription (oracle/create_map_ref.c), which is what
travels everywhere, must equal it bit for bit:
  - against the committed fixture tests/golden/createmap_ref.npz (generated from _ref by
    tests/golden/make_golden_ref.py) -- runs anywhere;
  - live against _ref on whole C1 / C2 / C3 / C5 maps when _ref is built (authoring container;
    on the GPU box the prebuilt .so travels with the snapshot)."""
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN, rotation_xyz
from tests.gpu_util import bits_equal

NCPU = os.cpu_count() or 1


def test_transcription_equals_reference_kernel_fixture(oracle):
    g = np.load(os.path.join(GOLDEN, "createmap_ref.npz"))
    names = [str(n) for n in g["names"]]
    assert len(names) >= 40
    for name in names:
        k = oracle.Intrinsics(*[float(v) for v in g[name + "_k"]])
        rows, cols, y0, x0, h, w = [int(v) for v in g[name + "_shape"]]
        mx, my = oracle.create_map(k, g[name + "_rot"], rows, cols, threads=NCPU)
        assert np.array_equal(mx[y0:y0 + h, x0:x0 + w].view(np.uint32), g[name + "_x"]), name
        assert np.array_equal(my[y0:y0 + h, x0:x0 + w].view(np.uint32), g[name + "_y"]), name
        sums = [int(mx.view(np.uint32).astype(np.uint64).sum()), int(my.view(np.uint32).astype(np.uint64).sum())]
        assert sums == [int(v) for v in g[name + "_sum"]], name  # the whole map, not only the window
    # the NaN at r == 0 is the reference kernel's own output (createMap.cl:38-39)
    nan_x = g["axis_nan_x"].view(np.float32)
    assert np.isnan(nan_x[6, 8]) and np.isnan(nan_x).sum() == 1


def test_fixture_matches_the_mounted_reference_source(oracle):
    """When the reference tree is mounted, the fixture must come from exactly that source file."""
    if not os.path.exists(oracle.REF_SOURCE):
        pytest.skip("reference tree not mounted")
    import hashlib
    g = np.load(os.path.join(GOLDEN, "createmap_ref.npz"))
    assert hashlib.sha256(open(oracle.REF_SOURCE, "rb").read()).hexdigest() == str(g["source_sha256"])


@pytest.mark.parametrize("geom", ["C1", "C2", "C3", "C5"])
def test_transcription_equals_reference_kernel_live(oracle, geom):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built and /root/reference not mounted")
    from video_annotator_b200 import configs  # host-only: camera scalars of the BASELINE workloads
    w = configs.workload(geom)
    k = oracle.intrinsics(w.input_camera.K, w.output_camera.K)
    (ow, oh) = w.out_size
    for R in (np.eye(3), rotation_xyz(2.0, -3.0, 1.5), rotation_xyz(10.0, -15.0, 20.0), rotation_xyz(-75.0, 40.0, 3.0)):
        ax, ay = oracle.create_map(k, R, oh, ow, threads=NCPU)
        bx, by = oracle.ref_create_map(k, R, oh, ow, threads=NCPU, sentinel=-12345.0)
        assert bits_equal(ax, bx) and bits_equal(ay, by)
        assert not np.any(bx == -12345.0)  # the NDRange covered every pixel


def test_reference_kernel_bounds_check_and_nan(oracle):
    """createMap.cl:13 keeps work-items of the rounded-up NDRange from writing; :38-39 NaN at r == 0."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built and /root/reference not mounted")
    k = oracle.Intrinsics(100.0, 80.0, 50.0, 50.0, 8.0, 6.0, 25.0, 25.0)
    rx, ry = oracle.ref_create_map(k, np.eye(3), 11, 13)  # odd sizes: NDRange rounds up to 16 x 12
    tx, ty = oracle.create_map(k, np.eye(3), 11, 13)
    assert bits_equal(rx, tx) and bits_equal(ry, ty)
    assert np.isnan(rx[6, 8]) and np.isnan(ry[6, 8]) and np.isnan(rx).sum() == 1


def test_reference_create_map_prefers_the_reference(oracle):
    k = oracle.Intrinsics(100.0, 80.0, 50.0, 50.0, 8.5, 6.5, 25.0, 25.0)
    mx, my, kind = oracle.reference_create_map(k, rotation_xyz(1, 2, 3), 12, 16)
    tx, ty = oracle.create_map(k, rotation_xyz(1, 2, 3), 12, 16)
    assert bits_equal(mx, tx) and bits_equal(my, ty)
    assert kind.startswith("reference") == oracle.ref_available()
    k.dist[0] = 0.01  # the distortion extension is not in createMap.cl: falls back to the port
    assert oracle.reference_create_map(k, np.eye(3), 4, 4)[2].startswith("port")
