#!/usr/bin/env python
"""End-to-end frames/s of the IN-LIBRARY frame-parallel scheduler (vaw_clip_warp_host, csrc/vaw_clip.cpp):
ONE process, one host thread + context + stream set per GPU, contiguous frame ranges, host buffers in and
out.  This is the code path BASELINE.json's north_star item 3 names; bench.py's torchrun ranks measure the
same kernels with one process per GPU instead.

    python tools/bench_clip.py --gpus N [--frames-per-gpu 64] [--steps 4] [--shards-local]

--shards-local: every device's shard of the host clip is allocated (pinned) by a thread bound to that
device's NUMA node, so its staging traffic stays on the local socket; without it the whole clip is one
pinned allocation made by the main thread.  Prints one JSON line."""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--frames-per-gpu", type=int, default=64)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--check", action="store_true", help="compare the result with a single-context run")
    a = ap.parse_args()
    import numpy as np
    import torch
    import video_annotator_b200 as V
    from video_annotator_b200 import _lib, configs
    lib = _lib.load()
    n_dev = a.gpus
    assert torch.cuda.device_count() >= n_dev, f"{torch.cuda.device_count()} GPUs visible"
    wl = configs.workload(a.workload)
    n = a.frames_per_gpu * n_dev
    ctx0 = V.WarpContext(wl.input_camera, wl.output_camera, out_size=wl.out_size, border=(0, 128, 128), device=0)
    params = ctx0.params
    sshape, dshape = ctx0.frame_shape("src"), ctx0.frame_shape("dst")
    rots = wl.rotations(n, first=100)
    nodes = [lib.vaw_device_numa_node(d) for d in range(n_dev)]
    # one pinned clip for the whole run, the way a caller that knows nothing about the topology holds it
    src = torch.empty((n,) + sshape, dtype=torch.uint8).pin_memory()
    dst = torch.empty((n,) + dshape, dtype=torch.uint8).pin_memory()
    dev_src = torch.empty((a.frames_per_gpu,) + sshape, dtype=torch.uint8, device="cuda:0")
    for g in range(n_dev):
        V.synth_nv12(dev_src, wl.src_size[0], wl.src_size[1], a.frames_per_gpu, first_index=g * a.frames_per_gpu, device=0)
        src[g * a.frames_per_gpu:(g + 1) * a.frames_per_gpu].copy_(dev_src)
    torch.cuda.synchronize()
    clip = V.ClipWarper(params, list(range(n_dev)))
    clip.warp_host(src, dst, rots)  # warm-up: contexts, staging rings
    clip.warp_host(src, dst, rots)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        clip.warp_host(src, dst, rots)
    dt = time.perf_counter() - t0
    line = {"scheduler": "vaw_clip_warp_host (one process, one host thread per GPU)", "n_gpus": n_dev, "workload": a.workload,
            "frames_per_step": n, "steps": a.steps, "value": n * a.steps / dt, "unit": "frames/s",
            "ms_per_step": dt / a.steps * 1e3,
            "gbs_per_direction_total": 0.5 * n * (wl.src_frame_bytes + wl.out_frame_bytes) * a.steps / dt / 1e9,
            "device_numa_nodes": nodes, "host_clip": "one pinned allocation (main thread)"}
    if a.check:
        want = torch.empty_like(dst[:a.frames_per_gpu])
        ctx0.warp_batch_host(src[:a.frames_per_gpu], want, rots[:a.frames_per_gpu])
        line["first_shard_equals_single_context"] = bool(torch.equal(want, dst[:a.frames_per_gpu]))
        last = n - a.frames_per_gpu
        ctx0.warp_batch_host(src[last:], want, rots[last:])
        line["last_shard_equals_single_context"] = bool(torch.equal(want, dst[last:]))
    print(json.dumps(line))
    clip.close()
    ctx0.close()


if __name__ == "__main__":
    main()
