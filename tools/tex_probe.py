#!/usr/bin/env python
"""Variant TEX against variant TILED on one workload: byte differences (histogram) and kernel times.
python tools/tex_probe.py [C3] [frames] [white]"""
import json, sys
import numpy as np
import torch
sys.path.insert(0, ".")
import video_annotator_b200 as V
from video_annotator_b200 import configs

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
white = int(sys.argv[3]) if len(sys.argv) > 3 else 0
wl = configs.workload(name)
dev = torch.device("cuda", 0)
sw, sh = wl.src_size
rots = wl.rotations(n, first=100, total=100 + n)
out = {}
res = {}
for variant in (3, 5):
    ctx = V.WarpContext(wl.input_camera, wl.output_camera, out_size=wl.out_size, border=(0, 128, 128), variant=variant)
    src = torch.empty((n,) + ctx.frame_shape("src"), dtype=torch.uint8, device=dev)
    dst = torch.zeros((n,) + ctx.frame_shape("dst"), dtype=torch.uint8, device=dev)
    V.synth_nv12(src, sw, sh, n, first_index=0, device=0, white_noise=bool(white))
    rdev = torch.empty(n * 9, dtype=torch.float32, device=dev)
    ctx.upload_rotations(rots, rdev)
    for _ in range(3):
        ctx.warp_batch(src, dst, rdev, n)
    ctx.set_option("time_kernels", 1)
    for _ in range(20):
        ctx.warp_batch(src, dst, rdev, n)
    torch.cuda.synchronize()
    b, w = ctx.kernel_times(20)
    a, t = ctx.kernel_times_split(20)
    res[variant] = dict(builder_ms=float(b.mean()), warp_ms=float(w.mean()), tex_ms=float(a.mean()), tile_ms=float(t.mean()),
                        pieces=ctx.piece_stats(rots[n // 2]))
    out[variant] = dst.clone()
    ctx.close()
    del src
d = (out[5].to(torch.int16) - out[3].to(torch.int16)).flatten()
hist = torch.bincount((d + 255).to(torch.int64), minlength=511).cpu().numpy()
nz = {int(i - 255): int(c) for i, c in enumerate(hist) if c and i != 255}
res["diff_hist_tex_minus_tiled"] = nz
res["samples"] = int(d.numel())
res["workload"] = name; res["frames"] = n; res["white"] = white
print(json.dumps(res))
