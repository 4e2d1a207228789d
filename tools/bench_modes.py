#!/usr/bin/env python
"""Frames/s and fraction of the HBM roofline of every format / interpolation the library offers, on the C3 geometry
(3840x2160 -> 3840x2160, 16 frames per launch, per-frame rotations), plus the motion-measurement kernels at 4K against
the OpenCV calls they replace.  GPU: PYTHONPATH=. python tools/bench_modes.py > gpurun_out/modes.json"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import video_annotator_b200 as V  # noqa: E402
from video_annotator_b200 import configs  # noqa: E402

PEAK = 6553.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    w = configs.workload("C3")
    n = 16
    rots = w.rotations(n, first=10, total=64)
    out = []
    modes = [("NV12 linear (tile kernel, default)", V.FORMAT_NV12, V.INTER_LINEAR, 0),
             ("NV12 linear, variant POLY (L1 gathers)", V.FORMAT_NV12, V.INTER_LINEAR, 2),
             ("NV12 linear, variant GATHER (op-for-op map)", V.FORMAT_NV12, V.INTER_LINEAR, 1),
             ("NV12 nearest (staged-tile kernel, default)", V.FORMAT_NV12, V.INTER_NEAREST, 0),
             ("NV12 nearest, variant GATHER", V.FORMAT_NV12, V.INTER_NEAREST, 1),
             ("BGR24 nearest (staged-tile kernel, default)", V.FORMAT_BGR24, V.INTER_NEAREST, 0),
             ("NV12 cubic (staged-tile kernel, default)", V.FORMAT_NV12, V.INTER_CUBIC, 0),
             ("NV12 cubic, variant GATHER (per-pixel taps)", V.FORMAT_NV12, V.INTER_CUBIC, 1),
             ("NV12 lanczos4 (staged-tile kernel, default)", V.FORMAT_NV12, V.INTER_LANCZOS4, 0),
             ("NV12 lanczos4, variant GATHER (per-pixel taps)", V.FORMAT_NV12, V.INTER_LANCZOS4, 1),
             ("BGR24 cubic (staged-tile kernel, default)", V.FORMAT_BGR24, V.INTER_CUBIC, 0),
             ("BGR24 cubic, variant GATHER (per-pixel taps)", V.FORMAT_BGR24, V.INTER_CUBIC, 1),
             ("BGR24 lanczos4 (staged-tile kernel, default)", V.FORMAT_BGR24, V.INTER_LANCZOS4, 0),
             ("BGR24 lanczos4, variant GATHER (per-pixel taps)", V.FORMAT_BGR24, V.INTER_LANCZOS4, 1),
             ("GRAY8 cubic (staged-tile kernel, default)", V.FORMAT_GRAY8, V.INTER_CUBIC, 0),
             ("GRAY8 cubic, variant GATHER (per-pixel taps)", V.FORMAT_GRAY8, V.INTER_CUBIC, 1),
             ("GRAY8 lanczos4 (staged-tile kernel, default)", V.FORMAT_GRAY8, V.INTER_LANCZOS4, 0),
             ("NV12 in -> BGR24 out, one launch (cvtColor + 3-channel remap fused, variant POLY)", V.FORMAT_NV12_TO_BGR24, V.INTER_LINEAR, 2),
             ("NV12 in -> BGR24 out, cvtColor into an L2-resident scratch + staged BGR kernel (variant TILED)", V.FORMAT_NV12_TO_BGR24, V.INTER_LINEAR, 3),
             ("BGR24 linear (the reference's literal format; staged-tile kernel, default)", V.FORMAT_BGR24, V.INTER_LINEAR, 0),
             ("BGR24 linear, variant GATHER (per-pixel taps)", V.FORMAT_BGR24, V.INTER_LINEAR, 1),
             ("GRAY8 linear (staged-tile kernel, default)", V.FORMAT_GRAY8, V.INTER_LINEAR, 0),
             ("GRAY8 linear, variant GATHER", V.FORMAT_GRAY8, V.INTER_LINEAR, 1)]
    only = os.environ.get("VAW_BENCH_MODES")  # substring filter, e.g. "cubic": just those modes, no motion measurement
    if only:
        modes = [m for m in modes if any(k in m[0] for k in only.split(","))]
    for name, fmt, interp, variant in modes:
        try:
            ctx = V.WarpContext(w.input_camera, w.output_camera, fmt=fmt, out_size=w.out_size, variant=variant, interpolation=interp)
        except V.VawError as e:
            out.append({"mode": name, "error": str(e)})
            continue
        src = torch.randint(0, 256, (n,) + tuple(ctx.frame_shape("src")), dtype=torch.uint8, device="cuda")
        dst = torch.empty((n,) + tuple(ctx.frame_shape("dst")), dtype=torch.uint8, device="cuda")
        rdev = torch.empty(n * 9, dtype=torch.float32, device="cuda")
        ctx.upload_rotations(rots, rdev)
        ms = timed(lambda: ctx.warp_batch(src, dst, rdev, n), 20)
        nbytes = src[0].numel() + dst[0].numel()
        out.append({"mode": name, "frames_per_s": round(n / ms * 1e3), "ms_per_16_frames": round(ms, 4),
                    "algorithmic_bytes_per_frame": nbytes, "roofline_frac": round(nbytes * n / ms / 1e6 / PEAK, 4)})
        ctx.close()
        del src, dst
    if only:
        print(json.dumps({"workload": "C3 geometry, 16 frames per launch", "peak_gbs": PEAK, "modes": out}, indent=1))
        return
    # motion measurement at 4K: pyramid + derivatives of a new frame, corners of the previous frame, 200 points tracked
    h, wd = 2160, 3840
    rng = np.random.default_rng(1)
    yy, xx = np.mgrid[0:h, 0:wd].astype(np.float32)
    img = np.zeros((h, wd), np.float32)
    for _ in range(16):
        img += np.sin(xx * rng.uniform(0.03, 0.3) + yy * rng.uniform(0.03, 0.3) + rng.uniform(0, 6.28))
    a = np.clip(127.5 + 20 * img, 0, 255).astype(np.uint8)
    b = np.roll(a, (3, -2), axis=(0, 1))
    ft = V.FlowTracker(wd, h)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    ft.push_frame(ta)
    ft.push_frame(tb)
    push_ms = timed(lambda: ft.push_frame(tb), 20)

    def wall(fn, reps=10):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps * 1e3
    pts = ft.corners(0)
    corners_ms = wall(lambda: ft.corners(0))
    track_ms = wall(lambda: ft.track(pts))
    flow = {"frame": [wd, h], "push_frame_ms (pyramid + Scharr, device time)": round(push_ms, 3),
            "corners_ms (wall, incl. host selection)": round(corners_ms, 3), "track_200_points_ms (wall)": round(track_ms, 3),
            "corners_found": int(len(pts))}
    try:
        import cv2
        cv2.setNumThreads(os.cpu_count() or 1)
        flow["cv2_goodFeaturesToTrack_ms"] = round(wall(lambda: cv2.goodFeaturesToTrack(a, 200, 0.01, 30), 3), 2)
        flow["cv2_calcOpticalFlowPyrLK_ms"] = round(wall(lambda: cv2.calcOpticalFlowPyrLK(a, b, pts.reshape(-1, 1, 2), None), 3), 2)
        flow["cv2_threads"] = cv2.getNumThreads()
    except ImportError:
        pass
    print(json.dumps({"workload": "C3 geometry, 16 frames per launch", "peak_gbs": PEAK, "modes": out, "motion_measurement": flow}, indent=1))


if __name__ == "__main__":
    main()
