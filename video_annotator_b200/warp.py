"""Python mirror of the reference's warp interface, over the C-ABI.

Mirrors /root/reference/opencv/FrameSourceWarp.hpp:14-34 (CameraPreset, Camera) and the
private FrameSourceWarp::warp_frame (FrameSourceWarp.cpp:272-314) as WarpContext.warp*.
Device buffers are torch uint8 CUDA tensors (plumbing only); the arithmetic is all in
libvaw.so.
"""
import ctypes as C

import numpy as np

from . import _lib

FORMAT_NV12, FORMAT_BGR24, FORMAT_GRAY8, FORMAT_NV12_TO_BGR24 = 0, 1, 2, 3
INTER_NEAREST = 0
INTER_LINEAR = 1
INTER_CUBIC = 2
INTER_LANCZOS4 = 4

# CameraPreset, FrameSourceWarp.hpp:14-21
GOPRO_H4B_WIDE43_PUBLISHED = 0
GOPRO_H4B_WIDE43_MEASURED = 1
GOPRO_H4B_WIDE43_MEASURED_STABILISATION = 2
GOPRO_H4B_WIDE169_PUBLISHED = 3
GOPRO_H4B_WIDE169_MEASURED = 4
GOPRO_H4B_WIDE169_MEASURED_STABILISATION = 5


class VawError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"vaw error {code}: {msg}")
        self.code = code


class Camera:
    """Camera (FrameSourceWarp.hpp:28-34): model, 3x3 matrix, distortion, size."""

    def __init__(self, c):
        self._c = c

    @property
    def K(self):
        return np.array(self._c.matrix[:], np.float64).reshape(3, 3)

    @property
    def size(self):
        return self._c.width, self._c.height

    @property
    def model(self):
        return self._c.model

    @property
    def distortion(self):
        return np.array(self._c.distortion[:], np.float64)

    @staticmethod
    def from_matrix(K, width, height, model=0, distortion=None):
        c = _lib.VawCamera()
        c.model, c.width, c.height = model, width, height
        for i, v in enumerate(np.asarray(K, np.float64).reshape(9)):
            c.matrix[i] = v
        if distortion is not None:
            for i in range(4):
                c.distortion[i] = float(distortion[i])
        return Camera(c)


def _check(rc, ctx=None):
    if rc != 0:
        msg = _lib.load().vaw_last_error(ctx)
        raise VawError(rc, (msg or b"").decode() or _lib.load().vaw_strerror(rc).decode())


def get_preset_camera(preset, width, height):
    """get_preset_camera, FrameSourceWarp.cpp:27-86."""
    c = _lib.VawCamera()
    _check(_lib.load().vaw_get_preset_camera(preset, width, height, C.byref(c)))
    return Camera(c)


def get_output_camera(cam, scale=1.0, crop_borders=False, zoom=1.0):
    """get_output_camera, FrameSourceWarp.cpp:88-165."""
    c = _lib.VawCamera()
    _check(_lib.load().vaw_get_output_camera(C.byref(cam._c), scale, int(crop_borders), zoom, C.byref(c)))
    return Camera(c)


def guess_rotation(input_camera, output_camera, prev_pts, cur_pts, seed=1):
    """guess_camera_rotation, FrameSourceWarp.cpp:316-368 (vaw_guess_rotation) -> (R (3, 3) float64, inliers)."""
    p = np.ascontiguousarray(prev_pts, np.float32).reshape(-1, 2)
    c = np.ascontiguousarray(cur_pts, np.float32).reshape(-1, 2)
    assert len(p) == len(c)
    R = np.zeros(9, np.float64)
    n = C.c_int(0)
    _check(_lib.load().vaw_guess_rotation(C.byref(input_camera._c), C.byref(output_camera._c), p.ctypes.data_as(_lib.f32p),
                                          c.ctypes.data_as(_lib.f32p), len(p), int(seed), R.ctypes.data_as(_lib.f64p), C.byref(n)))
    return R.reshape(3, 3), n.value


def _rot_arg(rot):
    r = np.ascontiguousarray(np.asarray(rot, np.float64).reshape(-1))
    return r, r.ctypes.data_as(_lib.f64p)


def _stream_handle(stream):
    if stream is None:
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)
    return C.c_void_p(int(stream))


class WarpContext:
    """Owns a vaw_ctx: the state FrameSourceWarp's constructor builds (FrameSourceWarp.cpp:199-226)."""

    def __init__(self, input_camera, output_camera, fmt=FORMAT_NV12, border=(0, 128, 128, 0),
                 device=0, variant=0, out_size=None, interpolation=INTER_LINEAR):
        lib = _lib.load()
        p = _lib.VawParams()
        _check(lib.vaw_params_from_cameras(C.byref(input_camera._c), C.byref(output_camera._c), fmt, C.byref(p)))
        if out_size is not None:
            p.out_width, p.out_height = out_size
        p.interpolation = interpolation
        p.variant = variant
        b = list(border) + [0] * (4 - len(border))
        for i in range(4):
            p.border[i] = int(b[i])
        self.params = p
        self.fmt = fmt
        self.device = device
        self.channels = 3 if fmt == FORMAT_BGR24 else 1              # of a source row
        self.dst_channels = 3 if fmt in (FORMAT_BGR24, FORMAT_NV12_TO_BGR24) else 1
        h = C.c_void_p()
        _check(lib.vaw_create(C.byref(p), device, C.byref(h)))
        self._h = h
        self._lib = lib

    # sizes ---------------------------------------------------------------------------
    @property
    def src_size(self):
        return self.params.src_width, self.params.src_height

    @property
    def out_size(self):
        return self.params.out_width, self.params.out_height

    def _side(self, which):
        # (format, channels) of the source or the output side (they differ for FORMAT_NV12_TO_BGR24)
        if self.fmt == FORMAT_NV12_TO_BGR24:
            return (FORMAT_NV12, 1) if which == "src" else (FORMAT_BGR24, 3)
        return self.fmt, self.channels

    def frame_bytes(self, which):
        w, h = self.src_size if which == "src" else self.out_size
        fmt, cn = self._side(which)
        return self._lib.vaw_frame_bytes(fmt, w, h, w * cn)

    def frame_shape(self, which):
        w, h = self.src_size if which == "src" else self.out_size
        fmt, _ = self._side(which)
        if fmt == FORMAT_NV12:
            return (h * 3 // 2, w)
        return (h, w, 3) if fmt == FORMAT_BGR24 else (h, w)

    @property
    def variant(self):
        """The kernel variant the context resolved to (VAW_VARIANT_*, never AUTO)."""
        return int(self._lib.vaw_get_variant(self._h))

    @property
    def launch_count(self):
        return int(self._lib.vaw_launch_count(self._h))

    def set_option(self, name, value):
        _check(self._lib.vaw_set_option(self._h, name.encode(), int(value)), self._h)

    # the warp ------------------------------------------------------------------------
    def warp(self, src, dst, rotation, stream=None, src_pitch=None, dst_pitch=None):
        """warp_frame(input, rotation), FrameSourceWarp.cpp:272-314.  src/dst: CUDA uint8 tensors."""
        _, rp = _rot_arg(rotation)
        sp = src_pitch or self.src_size[0] * self.channels
        dp = dst_pitch or self.out_size[0] * self.dst_channels
        _check(self._lib.vaw_warp(self._h, src.data_ptr(), sp, dst.data_ptr(), dp, rp,
                                  _stream_handle(stream)), self._h)
        return dst

    def upload_rotations(self, rotations, out, stream=None):
        """rotations: (n, 3, 3) float64 host -> out: CUDA float32 tensor (n*9)."""
        r, rp = _rot_arg(rotations)
        _check(self._lib.vaw_upload_rotations(self._h, rp, r.size // 9, out.data_ptr(),
                                              _stream_handle(stream)), self._h)
        return out

    def warp_batch(self, src, dst, rotations_dev, n_frames, stream=None, src_pitch=None,
                   dst_pitch=None, src_stride=None, dst_stride=None):
        sp = src_pitch or self.src_size[0] * self.channels
        dp = dst_pitch or self.out_size[0] * self.dst_channels
        ss = src_stride if src_stride is not None else self.frame_bytes("src")
        ds = dst_stride if dst_stride is not None else self.frame_bytes("dst")
        _check(self._lib.vaw_warp_batch(self._h, src.data_ptr(), sp, ss, dst.data_ptr(), dp, ds,
                                        rotations_dev.data_ptr(), n_frames, _stream_handle(stream)), self._h)
        return dst

    def warp_batch_host(self, src_host, dst_host, rotations):
        """Host buffers (numpy arrays or pinned torch CPU tensors), tightly packed frames."""
        r, rp = _rot_arg(rotations)
        n = r.size // 9

        def ptr(a):
            return a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data
        _check(self._lib.vaw_warp_batch_host(self._h, ptr(src_host), ptr(dst_host), rp, n), self._h)
        return dst_host

    def dump_coords(self, rotation, plane=0, stream=None):
        """The map createMap.cl would write (createMap.cl:42-49); returns two CUDA float32 tensors."""
        import torch
        w, h = self.out_size
        if plane == 1:
            w, h = w // 2, h // 2
        dev = torch.device("cuda", self.device)
        mx = torch.empty((h, w), dtype=torch.float32, device=dev)
        my = torch.empty((h, w), dtype=torch.float32, device=dev)
        _, rp = _rot_arg(rotation)
        _check(self._lib.vaw_dump_coords(self._h, rp, plane, mx.data_ptr(), my.data_ptr(), w,
                                         _stream_handle(stream)), self._h)
        return mx, my

    def kernel_times(self, max_launches=512):
        """(builder_ms, warp_ms) numpy arrays of the most recent launches (option time_kernels)."""
        b = np.zeros(max_launches, np.float32)
        w = np.zeros(max_launches, np.float32)
        n = C.c_int(0)
        _check(self._lib.vaw_kernel_times(self._h, max_launches, b.ctypes.data_as(_lib.f32p),
                                          w.ctypes.data_as(_lib.f32p), C.byref(n)), self._h)
        return b[:n.value], w[:n.value]

    def kernel_times_split(self, max_launches=512):
        """(tex_ms, tile_ms): the warp time of the same launches split by kernel (vaw_kernel_times_split)."""
        a = np.zeros(max_launches, np.float32)
        w = np.zeros(max_launches, np.float32)
        n = C.c_int(0)
        _check(self._lib.vaw_kernel_times_split(self._h, max_launches, a.ctypes.data_as(_lib.f32p),
                                                w.ctypes.data_as(_lib.f32p), C.byref(n)), self._h)
        return a[:n.value], w[:n.value]

    def piece_stats(self, rotation, stream=None):
        """Piece classification and tile sizing for this rotation (vaw_piece_stats)."""
        out = (C.c_uint32 * 8)()
        _, rp = _rot_arg(rotation)
        _check(self._lib.vaw_piece_stats(self._h, rp, out, _stream_handle(stream)), self._h)
        return dict(zip(("pieces", "poly", "interior", "outside", "max_tile_bytes", "tile_cap", "over_cap"),
                        [int(v) for v in out]))

    def piece_tiles(self, rotation, stream=None):
        # (pieces, 4) uint32: flags, tile bytes needed, tile pitch, luma rows | chroma rows << 16 (vaw_piece_tiles)
        cap = ((self.out_size[0] + 127) // 128) * ((self.out_size[1] + 7) // 8)
        out = np.zeros(cap * 4, np.uint32)
        _, rp = _rot_arg(rotation)
        _check(self._lib.vaw_piece_tiles(self._h, rp, out.ctypes.data_as(C.POINTER(C.c_uint32)), cap, _stream_handle(stream)), self._h)
        return out.reshape(cap, 4)

    def piece_flags(self, rotation, stream=None):
        # (pieces_y, pieces_x) uint32 flags of the 128 x piece_h pieces for this rotation, and piece_h (vaw_piece_flags)
        cap = ((self.out_size[0] + 127) // 128) * ((self.out_size[1] + 7) // 8)
        out = np.zeros(cap, np.uint32)
        nx, ny, ph = C.c_int(0), C.c_int(0), C.c_int(0)
        _, rp = _rot_arg(rotation)
        _check(self._lib.vaw_piece_flags(self._h, rp, out.ctypes.data_as(C.POINTER(C.c_uint32)), cap, C.byref(nx), C.byref(ny),
                                         C.byref(ph), _stream_handle(stream)), self._h)
        return out[:nx.value * ny.value].reshape(ny.value, nx.value), ph.value

    def close(self):
        if getattr(self, "_h", None):
            self._lib.vaw_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shard_range(n_frames, n_parts, part):
    """Contiguous frame range (first, count) of one rank / device (vaw_shard_range)."""
    first, count = C.c_int(0), C.c_int(0)
    _check(_lib.load().vaw_shard_range(n_frames, n_parts, part, C.byref(first), C.byref(count)))
    return first.value, count.value


class ClipWarper:
    """Frame-parallel warp of a host-resident clip over several GPUs of one box (vaw_clip_*)."""

    def __init__(self, ctx_params, devices):
        lib = _lib.load()
        h = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        rc = lib.vaw_clip_create(C.byref(ctx_params), len(devices), arr, C.byref(h))
        if rc != 0:
            raise VawError(rc, (lib.vaw_clip_last_error(None) or b"").decode())
        self._h, self._lib = h, lib

    def warp_host(self, src_host, dst_host, rotations):
        r, rp = _rot_arg(rotations)

        def ptr(a):
            return a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data
        rc = self._lib.vaw_clip_warp_host(self._h, ptr(src_host), ptr(dst_host), rp, r.size // 9)
        if rc != 0:
            raise VawError(rc, (self._lib.vaw_clip_last_error(self._h) or b"").decode())
        return dst_host

    def close(self):
        if getattr(self, "_h", None):
            self._lib.vaw_clip_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FlowTracker:
    """cv::calcOpticalFlowPyrLK (defaults) between consecutive frames, on the GPU (vaw_flow_*): the tracking
    step of FrameSourceWarp::consume_frame, FrameSourceWarp.cpp:421-427 / :242-270."""

    def __init__(self, width, height, device=0):
        lib = _lib.load()
        h = C.c_void_p()
        rc = lib.vaw_flow_create(width, height, device, C.byref(h))
        if rc != 0:
            raise VawError(rc, (lib.vaw_flow_last_error(None) or b"").decode())
        self._h, self._lib, self.size = h, lib, (width, height)

    def _check(self, rc):
        if rc != 0:
            raise VawError(rc, (self._lib.vaw_flow_last_error(self._h) or b"").decode())

    @property
    def levels(self):
        return int(self._lib.vaw_flow_levels(self._h))

    def push_frame(self, luma, pitch=None, stream=None):
        """luma: CUDA uint8 tensor holding the plane (an NV12 frame works: its first `height` rows)."""
        self._check(self._lib.vaw_flow_push_frame(self._h, luma.data_ptr(), pitch or self.size[0], _stream_handle(stream)))

    def track(self, prev_pts, stream=None):
        """prev_pts: (N, 2) float32 host array -> (next_pts (N, 2) float32, status (N,) bool)."""
        pts = np.ascontiguousarray(prev_pts, np.float32).reshape(-1, 2)
        nxt = np.zeros_like(pts)
        st = np.zeros(len(pts), np.uint8)
        self._check(self._lib.vaw_flow_track(self._h, pts.ctypes.data_as(_lib.f32p), len(pts), nxt.ctypes.data_as(_lib.f32p),
                                             st.ctypes.data_as(_lib.u8p), _stream_handle(stream)))
        return nxt, st.astype(bool)

    def corners(self, which=1, max_corners=200, quality=0.01, min_distance=30.0, stream=None):
        """cv::goodFeaturesToTrack on a pushed frame (which: 0 previous, 1 current) -> (N, 2) float32 (x, y)."""
        cap = max_corners if max_corners > 0 else 65536
        out = np.zeros((cap, 2), np.float32)
        n = C.c_int(0)
        self._check(self._lib.vaw_flow_corners(self._h, which, max_corners, quality, min_distance, out.ctypes.data_as(_lib.f32p),
                                               cap, C.byref(n), _stream_handle(stream)))
        return out[:n.value].copy()

    def response(self):
        """The cv::cornerMinEigenVal map of the last corners() call, (h, w) float32."""
        out = np.zeros((self.size[1], self.size[0]), np.float32)
        self._check(self._lib.vaw_flow_get_response(self._h, out.ctypes.data_as(_lib.f32p)))
        return out

    def level(self, which, level):
        """(image (h, w) uint8, dx (h, w) int16, dy (h, w) int16) of a pyramid level; which: 0 previous, 1 current."""
        w, h = C.c_int(0), C.c_int(0)
        self._check(self._lib.vaw_flow_get_level(self._h, which, level, None, None, C.byref(w), C.byref(h)))
        img = np.zeros((h.value, w.value), np.uint8)
        der = np.zeros((h.value, w.value, 2), np.int16)
        self._check(self._lib.vaw_flow_get_level(self._h, which, level, img.ctypes.data_as(_lib.u8p),
                                                 der.ctypes.data_as(C.POINTER(C.c_int16)), C.byref(w), C.byref(h)))
        return img, der[:, :, 0].copy(), der[:, :, 1].copy()

    def close(self):
        if getattr(self, "_h", None):
            self._lib.vaw_flow_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def synth_nv12(dst, width, height, n_frames, first_index=0, seed=20260001, white_noise=False,
               device=0, stream=None):
    """Fill a CUDA uint8 tensor with n_frames tightly packed synthetic NV12 frames."""
    lib = _lib.load()
    fb = width * height * 3 // 2
    _check(lib.vaw_synth_nv12(dst.data_ptr(), width, height, width, fb, first_index, n_frames,
                              seed, int(white_noise), device, _stream_handle(stream)))
    return dst


def remap_u8(src, map_x, map_y, border=(0, 0, 0), stream=None):
    """cv::remap(INTER_LINEAR, BORDER_CONSTANT) as called at FrameSourceWarp.cpp:306-312.
    src: CUDA uint8 (H, W) or (H, W, cn<=3); map_x/map_y: CUDA float32 (rows, cols)."""
    import torch
    cn = 1 if src.dim() == 2 else src.shape[2]
    h, w = src.shape[:2]
    rows, cols = map_x.shape
    dst = torch.empty((rows, cols) if src.dim() == 2 else (rows, cols, cn), dtype=torch.uint8, device=src.device)
    b = (C.c_uint8 * 4)(*([int(v) for v in border] + [0] * (4 - len(border))))
    _check(_lib.load().vaw_remap_u8(src.data_ptr(), w, h, src.stride(0), cn, map_x.data_ptr(), map_y.data_ptr(),
                                    rows, cols, map_x.stride(0), dst.data_ptr(), dst.stride(0), b,
                                    src.device.index or 0, _stream_handle(stream)))
    return dst


def nv12_to_bgr(src, width, height, n_frames=1, stream=None):
    """cv::cvtColor(COLOR_YUV2BGR_NV12), FrameSourceWarp.cpp:399-401.  src: CUDA uint8, n_frames tightly
    packed NV12 frames; returns a CUDA uint8 tensor (n_frames, height, width, 3)."""
    import torch
    dst = torch.empty((n_frames, height, width, 3), dtype=torch.uint8, device=src.device)
    _check(_lib.load().vaw_nv12_to_bgr(src.data_ptr(), width, height, width, width * height * 3 // 2, dst.data_ptr(),
                                       width * 3, width * height * 3, n_frames, src.device.index or 0,
                                       _stream_handle(stream)))
    return dst


def selftest_math(device=0, seed=1, n_per_thread=1000):
    out = (C.c_uint64 * 4)()
    _check(_lib.load().vaw_selftest_math(device, seed, n_per_thread, out))
    return dict(zip(("rcp", "div", "sqrt", "k"), [int(v) for v in out]))
