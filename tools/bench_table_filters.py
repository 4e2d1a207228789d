#!/usr/bin/env python
"""INTER_CUBIC / INTER_LANCZOS4 on the staged-tile kernel (C3 geometry, 16 frames per launch): frames/s against the
shared-memory / L1 split and the resident CTA count (analysis switches VAW_EXPERIMENT_FULL_SMEM, VAW_EXPERIMENT_MAX_CTAS).
GPU: PYTHONPATH=. python tools/bench_table_filters.py > gpurun_out/table_filters.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import video_annotator_b200 as V  # noqa: E402
from video_annotator_b200 import configs  # noqa: E402
from tools.bench_modes import timed  # noqa: E402


def main():
    w = configs.workload("C3")
    n = 16
    rots = w.rotations(n, first=10, total=64)
    src = torch.randint(0, 256, (n, 3240, 3840), dtype=torch.uint8, device="cuda")
    dst = torch.empty_like(src)
    rdev = torch.empty(n * 9, dtype=torch.float32, device="cuda")
    for interp, name in ((V.INTER_CUBIC, "cubic"), (V.INTER_LANCZOS4, "lanczos4")):
        for env in ({}, {"VAW_EXPERIMENT_PITCH64": "1"}, {"VAW_EXPERIMENT_FULL_SMEM": "1"}, {"VAW_EXPERIMENT_MAX_CTAS": "5"},
                    {"VAW_EXPERIMENT_MAX_CTAS": "3"}, {"VAW_EXPERIMENT_MAX_CTAS": "2"}, {"VAW_EXPERIMENT_TABLE_PAD": "110"},
                    {"VAW_EXPERIMENT_TABLE_PAD": "160"}):
            for k in ("VAW_EXPERIMENT_FULL_SMEM", "VAW_EXPERIMENT_MAX_CTAS", "VAW_EXPERIMENT_TABLE_PAD", "VAW_EXPERIMENT_PITCH64"):
                os.environ.pop(k, None)
            os.environ.update(env)
            ctx = V.WarpContext(w.input_camera, w.output_camera, out_size=w.out_size, interpolation=interp)
            ctx.upload_rotations(rots, rdev)
            ms = timed(lambda: ctx.warp_batch(src, dst, rdev, n), 10)
            st = ctx.piece_stats(rots[0])
            print(f"{name:9s} {str(env):40s} {ms:8.4f} ms / 16 frames  {n / ms * 1e3:9.0f} frames/s  tile_cap {st.get('tile_cap')}  "
                  f"max_tile {st.get('max_tile_bytes')} over_cap {st.get('over_cap')}", flush=True)
            ctx.close()


if __name__ == "__main__":
    main()
