#!/usr/bin/env python
"""BGR24 / GRAY8 through the staged-tile kernel on a BASELINE geometry: frames/s, fraction of the HBM roofline for the
format's bytes, piece / tile statistics.  GPU: PYTHONPATH=. python tools/bench_packed.py [C3] [frames] [bgr|gray]"""
import json
import os
import sys

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


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C3"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    fmts = sys.argv[3:] or ["bgr", "gray"]
    w = configs.workload(name)
    rots = w.rotations(n, first=10, total=max(64, n + 10))
    for f in fmts:
        fmt = V.FORMAT_BGR24 if f == "bgr" else V.FORMAT_GRAY8
        ctx = V.WarpContext(w.input_camera, w.output_camera, fmt=fmt, out_size=w.out_size)
        src = torch.randint(0, 256, (n,) + tuple(ctx.frame_shape("src")), dtype=torch.uint8, device="cuda")
        dst = torch.empty((n,) + tuple(ctx.frame_shape("dst")), dtype=torch.uint8, device="cuda")
        rdev = torch.empty(n * 9, dtype=torch.float32, device="cuda")
        ctx.upload_rotations(rots, rdev)
        ctx.set_option("time_kernels", 1)
        for _ in range(3):
            ctx.warp_batch(src, dst, rdev, n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            ctx.warp_batch(src, dst, rdev, n)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        bt, wt = ctx.kernel_times(reps)
        nbytes = src[0].numel() + dst[0].numel()
        st = ctx.piece_stats(rots[n // 2])
        print(json.dumps({"workload": name, "format": f, "frames_per_launch": n, "variant": ctx.variant, "frames_per_s": round(n / ms * 1e3),
                          "step_ms": round(ms, 4), "sampler_ms": round(float(np.mean(wt)), 4), "builder_ms": round(float(np.mean(bt)), 4),
                          "algorithmic_bytes_per_frame": nbytes, "roofline_frac_step": round(nbytes * n / ms / 1e6 / PEAK, 4),
                          "roofline_frac_sampler": round(nbytes * n / float(np.mean(wt)) / 1e6 / PEAK, 4), "pieces": st,
                          "ph_env": os.environ.get("VAW_EXPERIMENT_PH")}))
        ctx.close()
        del src, dst


if __name__ == "__main__":
    main()
