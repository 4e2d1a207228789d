#!/usr/bin/env python
"""A few launches of one interpolation mode on the C3 geometry (16 frames per launch), for ncu:
    ncu --set full -k regex:warp_nv12_quad -s 2 -c 1 -o gpurun_out/prof_cubic python tools/run_mode.py cubic"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import video_annotator_b200 as V  # noqa: E402
from video_annotator_b200 import configs  # noqa: E402

interp = {"linear": V.INTER_LINEAR, "nearest": V.INTER_NEAREST, "cubic": V.INTER_CUBIC, "lanczos4": V.INTER_LANCZOS4}[sys.argv[1]]
w = configs.workload("C3")
n = 16
rots = w.rotations(n, first=10, total=64)
ctx = V.WarpContext(w.input_camera, w.output_camera, out_size=w.out_size, interpolation=interp)
src = torch.randint(0, 256, (n,) + tuple(ctx.frame_shape("src")), dtype=torch.uint8, device="cuda")
dst = torch.empty((n,) + tuple(ctx.frame_shape("dst")), dtype=torch.uint8, device="cuda")
rdev = torch.empty(n * 9, dtype=torch.float32, device="cuda")
ctx.upload_rotations(rots, rdev)
for _ in range(4):
    ctx.warp_batch(src, dst, rdev, n)
torch.cuda.synchronize()
print("ok", sys.argv[1], ctx.variant)
ctx.close()
