"""The BASELINE.json workloads: geometry and synthetic per-frame rotations.

Host-only parameter producers for the benchmark and the parity tests.  Geometry comes from
the library's own get_preset_camera / get_output_camera (FrameSourceWarp.cpp:27-165); the
rotation sequences imitate what FrameSourceWarp::pull_frame feeds warp_frame
(FrameSourceWarp.cpp:441-442 accumulate by left-multiplication, :212/:471 Savitzky-Golay
smoothing with half-width 30 and order 2, :472-475 warp rotation = (smoothed * measured^-1)^-1).
The smoothing library (gram_savitzky_golay, un-vendored) is upstream of the hot path; any
smooth rotation sequence is a valid input, so the plain least-squares SG weights are used.
"""
from dataclasses import dataclass

import numpy as np

from . import warp as W

CONTENT_SEED = 20260001


@dataclass
class Workload:
    name: str
    description: str
    src_size: tuple      # (w, h) luma
    out_size: tuple      # (w, h) luma, even
    input_camera: W.Camera
    output_camera: W.Camera
    n_frames: int        # clip length the config names
    sigma_deg: float     # per-frame gyro increment, 0 -> identity rotation

    @property
    def src_frame_bytes(self):
        return self.src_size[0] * self.src_size[1] * 3 // 2

    @property
    def out_frame_bytes(self):
        return self.out_size[0] * self.out_size[1] * 3 // 2

    @property
    def algorithmic_bytes_per_frame(self):
        """unique source bytes touched (~100 % of the frame for these geometries)
        + output bytes + the 36-byte rotation.  No map bytes: the map is never materialised."""
        return self.src_frame_bytes + self.out_frame_bytes + 36

    def rotations(self, n=None, first=0, total=None):
        """Warp rotations of frames [first, first + n) of a clip of `total` frames (default: the clip
        ends at first + n).  A rank's slice of a longer clip must pass the clip's `total`, so that
        the smoothing window near the end of its range sees the following frames."""
        n = self.n_frames if n is None else n
        total = first + n if total is None else total
        return make_rotations(total, self.sigma_deg)[first:first + n]


from .rotations import ROT_SEED, make_rotations, sg_weights  # noqa: E402,F401  (pure numpy, also loaded stand-alone by bench.py's reference arm)


def _centred(f, size):
    return W.Camera.from_matrix([[f, 0, (size[0] - 1) / 2.0], [0, f, (size[1] - 1) / 2.0], [0, 0, 1]],
                                size[0], size[1])


def workload(name):
    """C1..C5 of BASELINE.json `configs` as SURVEY 8d defines them."""
    P = W.GOPRO_H4B_WIDE169_MEASURED
    if name == "C1":
        cam = W.get_preset_camera(P, 1920, 1080)
        out = W.get_output_camera(cam)
        size = (out.size[0] & ~1, out.size[1] & ~1)  # 1758 x 998
        return Workload("C1", "1920x1080 NV12 fisheye->rectilinear, identity rotation", (1920, 1080), size,
                        cam, out, 1, 0.0)
    if name == "C2":
        cam = W.get_preset_camera(P, 2704, 1520)
        out = W.get_output_camera(cam)
        size = (out.size[0] & ~1, out.size[1] & ~1)  # 2482 x 1408
        return Workload("C2", "2704x1520 GoPro fisheye->rectilinear, 120 frames, smoothed gyro rotations",
                        (2704, 1520), size, cam, out, 120, 0.4)
    if name in ("C3", "C4"):
        cam = W.get_preset_camera(P, 3840, 2160)
        ref = W.get_output_camera(cam)               # f 984.866, 3524 x 1999
        out = _centred(ref.K[0, 0] * 3840.0 / ref.size[0], (3840, 2160))
        n = 64 if name == "C3" else 600
        return Workload(name, "3840x2160 NV12 fisheye->rectilinear 3840x2160, per-frame rotation",
                        (3840, 2160), (3840, 2160), cam, out, n, 0.4)
    if name == "C5":
        cam = W.get_preset_camera(P, 5312, 2988)
        ref = W.get_output_camera(cam)               # f 1362.514, 4877 x 2766
        out = _centred(ref.K[0, 0] * 3840.0 / ref.size[0], (3840, 2160))
        return Workload("C5", "5312x2988 fisheye -> 3840x2160 wide-FOV rectilinear, large gather footprint",
                        (5312, 2988), (3840, 2160), cam, out, 32, 1.0)
    raise KeyError(name)
