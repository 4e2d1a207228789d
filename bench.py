#!/usr/bin/env python
"""bench.py -- warped frames/s of the fused fisheye->rectilinear rotate-and-warp path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C3|C4|...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over this rank's frames, repeated `launches_per_step`
times back to back (so that K steps fill a timed region of >= ~0.6 s; every launch streams
its whole input from HBM because one batch is larger than L2):

  C3 (default, BASELINE.json configs[2]): a batch of `--batch` (64) 3840x2160 NV12 frames per GPU,
      each with its own rotation, ONE library call per pass.  With N GPUs every rank owns a
      contiguous range of `batch` frames of an N*batch-frame clip -> weak scaling.
  C4 (configs[3]): a 600-frame 4K clip resident in HBM, sharded frame-parallel: rank r owns the
      contiguous range vaw_shard_range(600, N, r) with its rotations -> strong scaling.
  C1 / C2 / C5: the other BASELINE geometries (parity-test cases; benchable for reference).

No collective on the data path; torch.distributed is only the barrier and the max-over-ranks
of the timing.

  value     frames/s with the frames resident in HBM (CUDA events on the launch stream)
  e2e       the same through vaw_warp_batch_host: pinned HOST buffers in, host buffers out,
            host<->device copies inside the timed region; next to it the measured pinned
            bidirectional copy rate of the same buffers (its own roofline)
  roofline  algorithmic bytes per launch / average duration of the sampler kernel(s), against the
            measured HBM copy bandwidth of MEASURED_PEAKS.json
  parity    this workload's middle frame against the oracle (coordinates vs createMap.cl,
            pixels vs cv::remap's integer filter on the same map, and vs the oracle's whole path)
  cpu_baseline  the reference's CPU path (createMap.cl + cv::remap, all host threads) on a
            bounded sample of the same workload (rank 0, N = 1 only)

`--impl reference` times only that CPU path; it never imports the package (libvaw.so is not
mapped).  oracle/ is test infrastructure: the cpu_baseline / reference / parity legs of this
file are the only place outside tests/ that execute it, and only as the checker / baseline.
"""
import argparse
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "4K NV12 undistort+rotate warp frames/s at 1/2/4/8 B200; % of HBM roofline"
UNIT = "frames/s"
L2_BYTES = 126 << 20
C4_FRAMES = 600

# name -> (description, src (w, h), sigma of the per-frame gyro increment in degrees)
WORKLOADS = {
    "C1": ("1920x1080 NV12 fisheye->rectilinear, identity rotation", (1920, 1080), 0.0),
    "C2": ("2704x1520 GoPro fisheye->rectilinear, 120 frames, smoothed gyro rotations", (2704, 1520), 0.4),
    "C3": ("3840x2160 NV12 fisheye->rectilinear 3840x2160, per-frame rotation", (3840, 2160), 0.4),
    "C4": ("3840x2160 NV12 fisheye->rectilinear 3840x2160, 600-frame clip sharded frame-parallel", (3840, 2160), 0.4),
    "C5": ("5312x2988 fisheye -> 3840x2160 wide-FOV rectilinear, large gather footprint", (5312, 2988), 1.0),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=64, help="frames per launch per GPU (C4: ignored, the rank's range)")
    ap.add_argument("--variant", type=int, default=0, help="VAW_VARIANT_* (0 = auto)")
    ap.add_argument("--inner", type=int, default=0, help="launches per step (0: chosen so that the timed region is >= 0.6 s)")
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-shim", action="store_true")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the rank to its GPU's NUMA node")
    ap.add_argument("--split-builder", action="store_true", help="build all but the head frames' piece table on a side stream (measured neutral)")
    ap.add_argument("--host-stages", type=int, default=0, help="chunks in flight of the host-buffer pipeline (0 = library default)")
    ap.add_argument("--host-chunk-mb", type=int, default=0, help="MiB of source frames per chunk of the host-buffer pipeline (0 = default)")
    ap.add_argument("--fused-bgr", action="store_true",
                    help="NV12 in, BGR24 out (cvtColor + 3-channel remap, SURVEY 8 f2; --variant 2 = in one launch) instead of NV12 -> NV12")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample budget")
    return ap.parse_args()


def frames_of_rank(args, world, rank):
    """(first, count, clip_total) of this rank's contiguous frame range."""
    if args.workload == "C4":
        base, rem = divmod(C4_FRAMES, world)  # = vaw_shard_range (tests/test_shard.py pins the rule)
        return rank * base + min(rank, rem), base + (1 if rank < rem else 0), C4_FRAMES
    return rank * args.batch, args.batch, world * args.batch


def config_dict(args, world):
    """Names the workload.  Both arms build it from the command line alone, so the dicts are identical."""
    desc, src, sigma = WORKLOADS[args.workload]
    c4 = args.workload == "C4"
    return {"workload": f"{args.workload}: {desc}", "src": list(src),
            "format": "NV12 -> BGR24 (cvtColor + 3-channel remap)" if getattr(args, "fused_bgr", False) else "NV12",
            "frames_per_launch_per_gpu": (C4_FRAMES // world) if c4 else args.batch,
            "clip_frames": C4_FRAMES if c4 else args.batch * world,
            "sharding": ("frame-parallel, contiguous ranges of 600/N frames per GPU, no collective" if c4 else
                         f"frame-parallel, contiguous ranges of {args.batch} frames per GPU, no collective"),
            "rotations": (f"seeded gyro random walk sigma {sigma} deg/frame, SG-smoothed (radius 30, order 2)"
                          if sigma else "identity"),
            "content": "integer triangle waves + hash noise (vaw_synth_nv12 / oracle synth_ref.c)",
            "l2": "one launch's input is larger than the 126 MiB L2 (no flush needed)"}


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


# ---- NUMA placement ---------------------------------------------------------------------------------
def _cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def numa_bind(local, enable=True):
    """Pin this process to the CPUs of its GPU's NUMA node BEFORE any pinned allocation: first-touch
    then places the staging buffers next to the GPU's PCIe root.  Returns what was done."""
    info = {"bound": False}
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bus = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        info["pci"] = bus
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        info["node"] = node
        info["nodes_online"] = open("/sys/devices/system/node/online").read().strip()
        if node < 0 or not enable:
            return info
        cpus = _cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if target:
            os.sched_setaffinity(0, target)
            info.update(bound=True, cpus=len(target))
    except Exception as exc:  # no sysfs / not Linux / old torch: run unbound and say so
        info["error"] = repr(exc)[:120]
    return info


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
def load_rotations_module():
    """video_annotator_b200/rotations.py loaded stand-alone (pure numpy): the package is NOT imported."""
    spec = importlib.util.spec_from_file_location("vaw_rotations", os.path.join(ROOT, "video_annotator_b200", "rotations.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class RefGeometry:
    """The BASELINE geometries from the oracle's camera producers (oracle/camera_ref.c restating
    FrameSourceWarp.cpp:27-165) -- the same numbers video_annotator_b200.configs derives from the
    library's (tests/test_abi.py compares the two producers)."""

    def __init__(self, name, O):
        import numpy as np
        desc, (sw, sh), sigma = WORKLOADS[name]
        cam = O.get_preset_camera(4, sw, sh)  # GOPRO_H4B_WIDE169_MEASURED
        ref = O.get_output_camera(cam, 1.0, False, 1.0)
        if name in ("C1", "C2"):
            K_out, out = ref.K, (ref.width & ~1, ref.height & ~1)
        else:
            f = ref.K[0, 0] * 3840.0 / ref.width
            out = (3840, 2160)
            K_out = np.array([[f, 0, (out[0] - 1) / 2.0], [0, f, (out[1] - 1) / 2.0], [0, 0, 1]])
        self.name, self.src_size, self.out_size, self.sigma = name, (sw, sh), out, sigma
        self.k = O.intrinsics(cam.K, K_out)


class CpuPath:
    """The reference's path on the host: its own createMap kernel (oracle/_ref = createMap.cl compiled
    unmodified, rows over all threads; the transcription oracle/create_map_ref.c where _ref is absent)
    + cv::remap INTER_LINEAR/BORDER_CONSTANT on the Y plane and the 2-channel UV plane (cv2 = the real
    cv::remap, with cv::setNumThreads(all cores); else the oracle's port)."""

    def __init__(self, name):
        from oracle import oracle as O
        O.build()
        self.O = O
        self.geo = RefGeometry(name, O)
        self.threads = os.cpu_count() or 1
        try:
            import cv2
            cv2.setNumThreads(self.threads)
            self.cv2 = cv2
            self.remap_kind = f"cv2.remap {cv2.__version__} ({cv2.getNumThreads()} threads)"
        except Exception:
            self.cv2 = None
            self.remap_kind = "oracle/remap_ref.c (pthreads)"
        self.map_is_ref = O.ref_available()
        (sw, sh) = self.geo.src_size
        self.frames = [O.synth_nv12(sw, sh, i) for i in range(2)]

    def maps(self, rot):
        O, g = self.O, self.geo
        (ow, oh) = g.out_size
        mx, my, _ = O.reference_create_map(g.k, rot, oh, ow, threads=self.threads)
        cx, cy = O.chroma_map(mx, my, threads=self.threads)
        return mx, my, cx, cy

    def remap(self, frame, mx, my, cx, cy):
        O = self.O
        (sw, sh) = self.geo.src_size
        if self.cv2 is not None:
            cv2 = self.cv2
            y = cv2.remap(frame[:sh], mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0.0)
            uv = cv2.remap(frame[sh:].reshape(sh // 2, sw // 2, 2), cx, cy, cv2.INTER_LINEAR,
                           borderMode=cv2.BORDER_CONSTANT, borderValue=(128.0, 128.0))
        else:
            y = O.remap_u8(frame[:sh], mx, my, border=(0,), threads=self.threads)
            uv = O.remap_u8(frame[sh:].reshape(sh // 2, sw // 2, 2), cx, cy, border=(128, 128), threads=self.threads)
        return y, uv

    def warp(self, frame, rot):
        return self.remap(frame, *self.maps(rot))

    def run(self, rots):
        t0 = time.perf_counter()
        for i, r in enumerate(rots):
            self.warp(self.frames[i & 1], r)
        return time.perf_counter() - t0

    def describe(self, sample):
        stage1 = ("createMap.cl compiled unmodified (oracle/_ref, pthreads over rows)" if self.map_is_ref
                  else "createMap.cl transcription (oracle/create_map_ref.c, pthreads)")
        return {"cores": self.threads, "kind": "reference" if self.map_is_ref and self.cv2 is not None else "port",
                "sample": f"{sample}; {stage1} + {self.remap_kind}"}


def workload_rotations(name, n, first, total):
    R = load_rotations_module()
    return R.make_rotations(100 + total, WORKLOADS[name][2])[100 + first:100 + first + n]


def cpu_baseline(name, budget_s):
    cpu = CpuPath(name)
    rots = workload_rotations(name, 64, 0, 64)
    t1 = cpu.run(rots[:1])                      # warm-up / calibration frame
    n = max(2, min(64, int(budget_s / max(t1, 1e-3))))
    dt = cpu.run(rots[:n])
    g = cpu.geo
    d = cpu.describe(f"{n} frames of {name} ({g.src_size[0]}x{g.src_size[1]}->{g.out_size[0]}x{g.out_size[1]}) in {dt:.2f} s")
    d.update(value=n / dt, unit=UNIT)
    return d


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cpu = CpuPath(args.workload)
    per_step = 2
    rots = workload_rotations(args.workload, per_step * (args.steps + args.warmup), 0, per_step * (args.steps + args.warmup))
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
            "higher_is_better": True, "scaling": "strong" if args.workload == "C4" else "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config_dict(args, world),
            "frames_per_step": per_step, "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# ---- parity of the benched workload --------------------------------------------------------------------
def parity_block(ctx, wl, src, dst, rots, index):
    """The bench workload's own numbers for the three bars of BASELINE.json (rank 0): frame `index` of
    the batch the timed region just warped, checked on the host against the oracle."""
    import numpy as np
    from oracle import oracle as O
    from tests import gpu_util as G
    O.build()
    threads = os.cpu_count() or 1
    (sw, sh), (ow, oh) = wl.src_size, wl.out_size
    rot = rots[index]
    mx, my = [t.cpu().numpy() for t in ctx.dump_coords(rot, 0)]
    cx, cy = [t.cpu().numpy() for t in ctx.dump_coords(rot, 1)]
    got = dst[index].cpu().numpy()
    frame = src[index].cpu().numpy()
    k = O.intrinsics(wl.input_camera.K, wl.output_camera.K)
    ox, oy, kind = O.reference_create_map(k, rot, oh, ow, threads=threads)
    err = max(float(np.nanmax(np.abs(mx - ox))), float(np.nanmax(np.abs(my - oy))))

    def warp_on(ax, ay, bx, by):
        y = O.remap_u8(frame[:sh], ax, ay, border=(0,), threads=threads)
        uv = O.remap_u8(frame[sh:].reshape(sh // 2, sw // 2, 2), bx, by, border=(128, 128), threads=threads)
        return np.concatenate([y, uv.reshape(oh // 2, ow)], axis=0)

    same = G.diff_stats(got, warp_on(mx, my, cx, cy))
    ocx, ocy = O.chroma_map(ox, oy, threads=threads)
    whole = G.diff_stats(got, warp_on(ox, oy, ocx, ocy))
    return {"frame": int(index), "coordinate_oracle": kind, "coord_max_err_px": err, "coord_bar_px": 1e-3,
            "nan_pattern_equal": bool(np.array_equal(np.isnan(mx), np.isnan(ox))),
            "same_map": {"max_lsb": same["max"], "differ": same["differ"],
                         "what": "output vs cv::remap's integer filter (oracle/remap_ref.c, pinned to cv2.remap) on the map the kernel used"},
            "vs_reference_path": {"max_lsb": whole["max"], "differ": whole["differ"], "gt1": whole["gt1"],
                                  "hist": whole["hist"], "psnr_db": whole["psnr"],
                                  "what": "output vs createMap.cl's own map + cv::remap: differences are 1/32-px bucket flips "
                                          "where the two maps differ in the last bits (<= 1e-3 px)"}}


# ---- pinned copy rate: the roofline of the end-to-end number ---------------------------------------------
def copy_roofline(torch, dev, src_h, dst_h, src_d, dst_d, chunk_bytes=32 << 20, reps=3):
    """Plain pinned cudaMemcpyAsync of the e2e buffers in the library's chunk size: H2D alone, D2H alone,
    and both at once on two streams (what a perfect pipeline could sustain).  GB/s per direction."""
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    sh, dh = src_h.view(-1), dst_h.view(-1)
    sd, dd = src_d.view(-1), dst_d.view(-1)

    def h2d():
        with torch.cuda.stream(s1):
            for o in range(0, sh.numel(), chunk_bytes):
                sd[o:o + chunk_bytes].copy_(sh[o:o + chunk_bytes], non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            for o in range(0, dh.numel(), chunk_bytes):
                dh[o:o + chunk_bytes].copy_(dd[o:o + chunk_bytes], non_blocking=True)

    def timed(fns):
        best = float("inf")
        for _ in range(reps + 1):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for f in fns:
                f()
            torch.cuda.synchronize(dev)
            best = min(best, time.perf_counter() - t0)
        return best

    t_h, t_d, t_b = timed([h2d]), timed([d2h]), timed([h2d, d2h])
    return {"h2d_alone_gbs": sh.numel() / t_h / 1e9, "d2h_alone_gbs": dh.numel() / t_d / 1e9,
            "bidirectional_gbs_per_direction": 0.5 * (sh.numel() + dh.numel()) / t_b / 1e9,
            "chunk_bytes": chunk_bytes}


def shim_bench():
    """frames/s through the C++ drop-in (host/FrameSourceWarp: pull_frame on a 4K source): vaw_demo --bench."""
    exe = os.path.join(ROOT, "video_annotator_b200", "host", "vaw_demo")
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe, "--bench"], capture_output=True, text=True, timeout=300)
        for line in out.stdout.splitlines():
            if line.startswith("{"):
                return json.loads(line)
        return {"error": (out.stderr or out.stdout)[-200:]}
    except Exception as exc:
        return {"error": repr(exc)[:200]}


KERNEL_NAMES = {1: "warp_nv12_gather_kernel", 2: "warp_nv12_poly_kernel", 3: "warp_nv12_quad_kernel",
                5: "warp_nv12_tex_kernel + warp_nv12_quad_kernel"}


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
    numa = numa_bind(local, enable=not args.no_numa)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # (NCCL prints its version banner on stdout at init: main() has moved fd 1 to stderr, see emit())
        dist.init_process_group("nccl", device_id=dev)

    import video_annotator_b200 as V
    from video_annotator_b200 import configs

    wl = configs.workload(args.workload)
    first, n, clip_total = frames_of_rank(args, world, rank)
    (sw, sh) = wl.src_size
    ctx = V.WarpContext(wl.input_camera, wl.output_camera, out_size=wl.out_size,
                        border=(0, 0, 0) if args.fused_bgr else (0, 128, 128), device=local, variant=args.variant,
                        fmt=V.FORMAT_NV12_TO_BGR24 if args.fused_bgr else V.FORMAT_NV12)
    out_frame_bytes = wl.out_size[0] * wl.out_size[1] * 3 if args.fused_bgr else wl.out_frame_bytes
    if args.split_builder:
        ctx.set_option("split_builder", 1)
    if args.host_stages:
        ctx.set_option("host_stages", args.host_stages)
    if args.host_chunk_mb:
        ctx.set_option("host_chunk_mb", args.host_chunk_mb)
    # this rank's contiguous frame range of the clip, with its rotations
    rots = wl.rotations(n, first=100 + first, total=100 + clip_total)
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
    # launches per step: enough that K steps last >= 0.6 s (estimated from two untimed passes)
    ctx.warp_batch(src, dst, rdev, n)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.warp_batch(src, dst, rdev, n)
    ctx.warp_batch(src, dst, rdev, n)
    torch.cuda.synchronize()
    est_ms = (time.perf_counter() - t0) * 500.0
    inner = args.inner if args.inner > 0 else max(1, int(np.ceil(600.0 / (max(est_ms, 1e-3) * args.steps))))
    if world > 1:  # every rank must run the same count
        t = torch.tensor([inner], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        inner = int(t.item())
    for _ in range(args.warmup):
        for _ in range(inner):
            ctx.warp_batch(src, dst, rdev, n)
    timed_kernels = ctx.fmt in (V.FORMAT_NV12, V.FORMAT_NV12_TO_BGR24) and args.variant != 1
    if timed_kernels:
        ctx.set_option("time_kernels", 1)   # CUDA-event stamps around each kernel, on the launch stream
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches0 = ctx.launch_count
    ev[0].record(stream)
    for s in range(args.steps):
        for _ in range(inner):
            ctx.warp_batch(src, dst, rdev, n)
        ev[s + 1].record(stream)
    torch.cuda.synchronize()
    launches = ctx.launch_count - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    per_step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    if timed_kernels:
        builder_ms, warp_ms = ctx.kernel_times(min(args.steps * inner, 512))
        ctx.set_option("time_kernels", 0)
    else:
        builder_ms, warp_ms = np.zeros(1, np.float32), np.array(per_step_ms, np.float32) / inner
    barrier()
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    ms_per_step = total_ms_max / args.steps
    frames_all = clip_total if args.workload == "C4" else n * world
    value = frames_all * inner * args.steps / (total_ms_max * 1e-3)

    # ---- end to end: pinned host buffers through the C-ABI's host entry point ----------------
    e2e = None
    if not args.no_e2e:
        ne = min(n, 128)  # C4 at N = 1, 2 would need tens of GB of pinned memory: its first 128 frames per rank
        src_h = torch.empty((ne,) + ctx.frame_shape("src"), dtype=torch.uint8).pin_memory()
        dst_h = torch.empty((ne,) + ctx.frame_shape("dst"), dtype=torch.uint8).pin_memory()
        src_h.copy_(src[:ne])
        ctx.warp_batch_host(src_h, dst_h, rots[:ne])          # warm-up (allocates the staging ring)
        ctx.warp_batch_host(src_h, dst_h, rots[:ne])
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            ctx.warp_batch_host(src_h, dst_h, rots[:ne])      # returns when dst_h is complete
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        same = bool(torch.equal(dst_h.to(dev), dst[:ne]))
        copy = copy_roofline(torch, dev, src_h, dst_h, src[:ne], torch.empty_like(dst[:ne]))
        barrier()
        # all ranks copy at once in the micro-benchmark too (same barrier-bracketed phase): the per-rank rate
        # under contention is what bounds the aggregate; rank 0 reports the slowest rank's figure
        t = torch.tensor([copy["bidirectional_gbs_per_direction"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        copy["bidirectional_gbs_per_direction_min_over_ranks"] = float(t.item())
        h2d = ne * wl.src_frame_bytes + ne * 36
        d2h = ne * out_frame_bytes
        per_dir = 0.5 * (h2d + d2h) * args.e2e_steps / dt / 1e9
        e2e = {"value": ne * world * args.e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": args.e2e_steps, "ms_per_step": dt / args.e2e_steps * 1e3, "frames_per_step_per_gpu": ne,
               "api": "vaw_warp_batch_host (pinned host src/dst, chunked H2D -> warp -> D2H pipeline)",
               "output_equals_device_path": same,
               "gbs_per_direction_per_gpu": per_dir, "copy_peak": copy,
               "copy_peak_gbs": copy["bidirectional_gbs_per_direction_min_over_ranks"],
               "roofline_frac": per_dir / copy["bidirectional_gbs_per_direction_min_over_ranks"],
               "roofline_note": "bytes moved per direction per second / the plain pinned bidirectional cudaMemcpyAsync "
                                "rate of the same buffers (PCIe + host memory), all ranks copying at once"}
        del src_h, dst_h
    clocks = sampler.stop()

    if rank == 0:
        peak, peak_src = measured_peak()
        alg = (wl.src_frame_bytes + out_frame_bytes + 36) * n
        avg_pass_ms = total_ms / (args.steps * inner)
        avg_launch_ms = float(np.mean(warp_ms))          # the sampler kernel(s) of one pass
        achieved = alg / (avg_launch_ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong" if args.workload == "C4" else "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": config_dict(args, world),
                "launches_per_step": inner, "frames_per_step": frames_all * inner,
                "details": {"out": list(wl.out_size), "variant": args.variant, "variant_resolved": ctx.variant,
                            "pieces_128x32": pieces, "frames_this_rank": n, "numa": numa,
                            "split_builder": bool(args.split_builder)},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": recorded_traffic(wl.name, n),
                             "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                             "kernel": (("nv12_to_bgr_kernel + warp_packed_tile_kernel<3> per L2-sized chunk (cvtColor, then fused map + 3-channel remap on staged tiles)"
                                         if ctx.variant == 3 else "warp_nv12_to_bgr_kernel (fused map + cvtColor + 3-channel remap)") if args.fused_bgr else
                                        KERNEL_NAMES.get(ctx.variant, "warp_nv12_quad_kernel") + " (fused map + remap, luma + chroma)"),
                             "launch_ms": {"avg": avg_launch_ms, "median": float(np.median(warp_ms)),
                                           "best": float(np.min(warp_ms)), "launches_timed": int(len(warp_ms))},
                             "other_kernels_ms": {"build_pieces_kernel": float(np.mean(builder_ms))},
                             "pass_ms": {"avg": avg_pass_ms, "median": statistics.median(per_step_ms) / inner,
                                         "best": min(per_step_ms) / inner},
                             "whole_step_frac": alg / (avg_pass_ms * 1e-3) / 1e9 / peak,
                             "frac_of_8TBps": achieved / 8000.0},
                "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        if world == 1 and not args.no_parity and not args.fused_bgr:
            line["parity"] = parity_block(ctx, wl, src, dst, rots, n // 2)
        if world == 1 and not args.no_shim:
            line["shim"] = shim_bench()
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload, args.cpu_seconds)
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
