#!/usr/bin/env python
"""bench.py -- warped frames/s of the fused fisheye->rectilinear rotate-and-warp path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic NV12 frames: ONE launch of
the fused map+remap kernel over `--batch` frames (default 64) of workload C3
(BASELINE.json configs[2]: 3840x2160 NV12 fisheye->rectilinear with per-frame rotation).
With N GPUs each rank owns a contiguous range of `batch` frames of an N*batch-frame clip with
their rotations (frame-parallel, no collective on the data path; torch.distributed is only
the barrier and the max-over-ranks of the timing) -> weak scaling.

  value     frames/s with the frames resident in HBM (CUDA events on the launch stream)
  e2e       the same through vaw_warp_batch_host: pinned HOST buffers in, host buffers out,
            host<->device copies inside the timed region
  roofline  algorithmic bytes per launch / average launch duration, against the measured
            HBM copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline  the reference's CPU path (createMap transcription + cv::remap, all host
            threads) on a bounded sample of the same workload (rank 0, N = 1 only)

`--impl reference` times only that CPU path (oracle/ is test infrastructure: this file's
cpu_baseline / reference legs are the only place outside tests/ that execute it).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "4K NV12 undistort+rotate warp frames/s at 1/2/4/8 B200; % of HBM roofline"
UNIT = "frames/s"
L2_BYTES = 126 << 20


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3", choices=["C1", "C2", "C3", "C5"])
    ap.add_argument("--batch", type=int, default=64, help="frames per launch per GPU")
    ap.add_argument("--variant", type=int, default=0, help="VAW_VARIANT_* (0 = auto)")
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample budget")
    return ap.parse_args()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic(workload, batch):
    """DRAM bytes per launch from the committed ncu --set full capture (profiles/traffic.json)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        e = t.get(f"{workload}_batch{batch}")
        return float(e["dram_bytes_per_launch"]) if e else None
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and clock-event reasons of one GPU through NVML while a region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join()
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- the reference's CPU path --------------------------------------------------------------------
class CpuPath:
    """createMap.cl transcription (oracle/create_map_ref.c, rows over all threads) + cv::remap
    INTER_LINEAR/BORDER_CONSTANT on the Y plane and the 2-channel UV plane (cv2 = the real
    cv::remap when importable, with cv::setNumThreads(all cores); else the oracle's port)."""

    def __init__(self, wl):
        from oracle import oracle as O
        O.build()
        self.O, self.wl = O, wl
        self.threads = os.cpu_count() or 1
        try:
            import cv2
            cv2.setNumThreads(self.threads)
            self.cv2 = cv2
            self.remap_kind = f"cv2.remap {cv2.__version__} ({cv2.getNumThreads()} threads)"
        except Exception:
            self.cv2 = None
            self.remap_kind = "oracle/remap_ref.c (pthreads)"
        (sw, sh) = wl.src_size
        self.k = O.intrinsics(wl.input_camera.K, wl.output_camera.K)
        self.frames = [O.synth_nv12(sw, sh, i) for i in range(2)]

    def warp(self, frame, rot):
        O, wl = self.O, self.wl
        (sw, sh), (ow, oh) = wl.src_size, wl.out_size
        mx, my = O.create_map(self.k, rot, oh, ow, threads=self.threads)
        cx, cy = O.chroma_map(mx, my, threads=self.threads)
        if self.cv2 is not None:
            cv2 = self.cv2
            y = cv2.remap(frame[:sh], mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0.0)
            uv = cv2.remap(frame[sh:].reshape(sh // 2, sw // 2, 2), cx, cy, cv2.INTER_LINEAR,
                           borderMode=cv2.BORDER_CONSTANT, borderValue=(128.0, 128.0))
        else:
            y = O.remap_u8(frame[:sh], mx, my, border=(0,), threads=self.threads)
            uv = O.remap_u8(frame[sh:].reshape(sh // 2, sw // 2, 2), cx, cy, border=(128, 128), threads=self.threads)
        return y, uv

    def run(self, rots):
        t0 = time.perf_counter()
        for i, r in enumerate(rots):
            self.warp(self.frames[i & 1], r)
        return time.perf_counter() - t0

    def describe(self, sample):
        return {"cores": self.threads, "kind": "port",
                "sample": f"{sample}; createMap.cl transcription (oracle/create_map_ref.c, pthreads) + {self.remap_kind}"}


def cpu_baseline(wl, budget_s):
    cpu = CpuPath(wl)
    rots = wl.rotations(64, first=100)
    t1 = cpu.run(rots[:1])                      # warm-up / calibration frame
    n = max(2, min(64, int(budget_s / max(t1, 1e-3))))
    dt = cpu.run(rots[:n])
    d = cpu.describe(f"{n} frames of {wl.name} ({wl.src_size[0]}x{wl.src_size[1]}->{wl.out_size[0]}x{wl.out_size[1]})"
                     f" in {dt:.2f} s")
    d.update(value=n / dt, unit=UNIT)
    return d


def config_dict(wl, args, world, extra=None):
    c = {"workload": f"{wl.name}: {wl.description}", "src": list(wl.src_size), "out": list(wl.out_size),
         "format": "NV12", "frames_per_launch_per_gpu": args.batch, "clip_frames": args.batch * world,
         "sharding": f"frame-parallel, contiguous ranges of {args.batch} frames per GPU, no collective",
         "rotations": f"seeded gyro random walk sigma {wl.sigma_deg} deg/frame, SG-smoothed (radius 30, order 2)",
         "content": "integer triangle waves + hash noise (vaw_synth_nv12)"}
    if extra:
        c.update(extra)
    return c


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import video_annotator_b200 as V  # host-only camera producers (no GPU needed)
    from video_annotator_b200 import configs
    wl = configs.workload(args.workload)
    cpu = CpuPath(wl)
    per_step = 2
    rots = wl.rotations(per_step * (args.steps + args.warmup), first=100)
    for s in range(args.warmup):
        cpu.run(rots[s * per_step:(s + 1) * per_step])
    t0 = time.perf_counter()
    for s in range(args.warmup, args.warmup + args.steps):
        cpu.run(rots[s * per_step:(s + 1) * per_step])
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    base = cpu.describe(f"{per_step} frames per step, {args.steps} steps")
    base.update(value=value, unit=UNIT)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": config_dict(wl, args, 1, {"frames_per_step": per_step}),
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


KERNEL_NAMES = {1: "warp_nv12_gather_kernel", 2: "warp_nv12_poly_kernel", 3: "warp_nv12_tile_kernel",
                4: "warp_nv12_pipe_kernel", 5: "warp_nv12_tex_kernel + warp_nv12_tile_kernel"}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the warp path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # (NCCL prints its version banner on stdout at init: main() has moved fd 1 to stderr, see emit())
        dist.init_process_group("nccl", device_id=dev)

    import video_annotator_b200 as V
    from video_annotator_b200 import configs

    wl = configs.workload(args.workload)
    n = args.batch
    (sw, sh) = wl.src_size
    ctx = V.WarpContext(wl.input_camera, wl.output_camera, out_size=wl.out_size, border=(0, 128, 128),
                        device=local, variant=args.variant)
    # this rank's contiguous frame range of the clip, with its rotations
    first = rank * n
    rots = wl.rotations(n, first=100 + first, total=100 + world * n)
    src = torch.empty((n,) + ctx.frame_shape("src"), dtype=torch.uint8, device=dev)
    dst = torch.empty((n,) + ctx.frame_shape("dst"), dtype=torch.uint8, device=dev)
    V.synth_nv12(src, sw, sh, n, first_index=first, device=local)
    rdev = torch.empty(n * 9, dtype=torch.float32, device=dev)
    ctx.upload_rotations(rots, rdev)
    torch.cuda.synchronize()
    pieces = ctx.piece_stats(rots[n // 2])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.current_stream()
    for _ in range(args.warmup):
        ctx.warp_batch(src, dst, rdev, n)
    timed_kernels = ctx.fmt == V.FORMAT_NV12 and args.variant != 1
    if timed_kernels:
        ctx.set_option("time_kernels", 1)   # CUDA-event stamps around each kernel, on the launch stream
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches0 = ctx.launch_count
    ev[0].record(stream)
    for s in range(args.steps):
        ctx.warp_batch(src, dst, rdev, n)
        ev[s + 1].record(stream)
    torch.cuda.synchronize()
    launches = ctx.launch_count - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    if timed_kernels:
        builder_ms, warp_ms = ctx.kernel_times(min(args.steps, 512))
        ctx.set_option("time_kernels", 0)
    else:
        builder_ms, warp_ms = np.zeros(1, np.float32), np.array(per_launch_ms, np.float32)
    barrier()
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    ms_per_step = total_ms_max / args.steps
    value = n * world * args.steps / (total_ms_max * 1e-3)

    # ---- end to end: pinned host buffers through the C-ABI's host entry point ----------------
    e2e = None
    if not args.no_e2e:
        src_h = torch.empty((n,) + ctx.frame_shape("src"), dtype=torch.uint8).pin_memory()
        dst_h = torch.empty((n,) + ctx.frame_shape("dst"), dtype=torch.uint8).pin_memory()
        src_h.copy_(src)
        ctx.warp_batch_host(src_h, dst_h, rots)          # warm-up (allocates the staging ring)
        ctx.warp_batch_host(src_h, dst_h, rots)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            ctx.warp_batch_host(src_h, dst_h, rots)      # returns when dst_h is complete
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        same = bool(torch.equal(dst_h.to(dev), dst))
        e2e = {"value": n * world * args.e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": n * wl.src_frame_bytes + n * 36,
               "d2h_bytes_per_step": n * wl.out_frame_bytes,
               "steps": args.e2e_steps, "ms_per_step": dt / args.e2e_steps * 1e3,
               "api": "vaw_warp_batch_host (pinned host src/dst, chunked H2D -> warp -> D2H pipeline)",
               "output_equals_device_path": same}
        del src_h, dst_h
    clocks = sampler.stop()

    if rank == 0:
        peak, peak_src = measured_peak()
        alg = wl.algorithmic_bytes_per_frame * n
        avg_step_ms = total_ms / args.steps
        avg_launch_ms = float(np.mean(warp_ms))          # the dominant kernel alone
        achieved = alg / (avg_launch_ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": config_dict(wl, args, world, {
                    "l2": f"inputs {n * wl.src_frame_bytes >> 20} MiB per step > {L2_BYTES >> 20} MiB L2 (no flush needed)",
                    "variant": args.variant, "variant_resolved": ctx.variant, "pieces_128x32": pieces}),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": recorded_traffic(wl.name, n),
                             "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                             "kernel": KERNEL_NAMES.get(ctx.variant, "warp_nv12_tile_kernel")
                                       + " (fused map + remap, luma + chroma, one launch per step)",
                             "launch_ms": {"avg": avg_launch_ms, "median": float(np.median(warp_ms)),
                                           "best": float(np.min(warp_ms)), "launches_timed": int(len(warp_ms))},
                             "other_kernels_ms": {"build_pieces_kernel": float(np.mean(builder_ms))},
                             "step_ms": {"avg": avg_step_ms, "median": statistics.median(per_launch_ms),
                                         "best": min(per_launch_ms)},
                             "whole_step_frac": alg / (avg_step_ms * 1e-3) / 1e9 / peak,
                             "frac_of_8TBps": achieved / 8000.0},
                "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl, args.cpu_seconds)
        emit(line)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_RESULT_FD = None


def emit(line):
    """The one JSON line of the contract, on the process's ORIGINAL stdout.  Everything else any library
    writes to fd 1 (NCCL's version banner, torchrun notices) goes to stderr: fd 1 is redirected at start-up."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


if __name__ == "__main__":
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)   # from here on fd 1 is stderr, for C libraries as well as for print()
    a = parse_args()
    sys.exit(run_reference(a) if a.impl == "reference" else run_ours(a))
