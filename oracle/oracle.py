"""ctypes loader for the CPU oracle (oracle/liboracle_vaw.so).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product
(video_annotator_b200, libvaw.so) never does and has no CPU fallback.

The C sources restate the reference's algorithm (citations in vaw_oracle.h):
createMap.cl:10-50, FrameSourceWarp.cpp:27-165 and :272-314, and cv::remap.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_vaw.so")


def build(force=False):
    """Compile the oracle with gcc (oracle/Makefile)."""
    srcs = [f for f in os.listdir(_HERE) if f.endswith((".c", ".h")) or f == "Makefile"]
    newest = max(os.path.getmtime(os.path.join(_HERE, f)) for f in srcs)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class Intrinsics(C.Structure):
    """The 8 floats of FrameSourceWarp.cpp:283-290 (+ the fisheye distortion extension, zeros = createMap.cl)."""
    _fields_ = [(n, C.c_float) for n in (
        "src_center_x", "src_center_y", "src_focal_x", "src_focal_y",
        "map_center_x", "map_center_y", "map_focal_x", "map_focal_y")] + [("dist", C.c_float * 4)]


class Camera(C.Structure):
    """FrameSourceWarp.hpp:28-34."""
    _fields_ = [("model", C.c_int), ("matrix", C.c_double * 9), ("dist", C.c_double * 4),
                ("width", C.c_int), ("height", C.c_int)]

    @property
    def K(self):
        return np.array(self.matrix[:], dtype=np.float64).reshape(3, 3)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        fp = C.POINTER(C.c_float)
        u8 = C.POINTER(C.c_uint8)
        ip = C.POINTER(Intrinsics)
        L.vaw_oracle_create_map.argtypes = [fp, fp, C.c_int, C.c_int, C.c_int, ip, fp, C.c_int]
        L.vaw_oracle_create_map.restype = None
        L.vaw_oracle_chroma_map.argtypes = [fp, fp, C.c_int, C.c_int, C.c_int, fp, fp, C.c_int, C.c_int]
        L.vaw_oracle_chroma_map.restype = None
        L.vaw_oracle_remap_u8.argtypes = [u8, C.c_int, C.c_int, C.c_int, C.c_int, fp, fp, C.c_int,
                                          C.c_int, C.c_int, u8, C.c_int, u8, C.c_int]
        L.vaw_oracle_remap_u8.restype = None
        L.vaw_oracle_remap_cubic_u8.argtypes = L.vaw_oracle_remap_u8.argtypes
        L.vaw_oracle_remap_cubic_u8.restype = None
        L.vaw_oracle_cubic_table.argtypes = []
        L.vaw_oracle_cubic_table.restype = C.POINTER(C.c_short)
        L.vaw_oracle_remap_lanczos4_u8.argtypes = L.vaw_oracle_remap_u8.argtypes
        L.vaw_oracle_remap_lanczos4_u8.restype = None
        L.vaw_oracle_lanczos4_table.argtypes = []
        L.vaw_oracle_lanczos4_table.restype = C.POINTER(C.c_short)
        L.vaw_oracle_warp_nv12.argtypes = [u8, C.c_int, C.c_int, C.c_int, u8, C.c_int, C.c_int,
                                           C.c_int, ip, fp, C.c_int, C.c_int, C.c_int, C.c_int]
        L.vaw_oracle_warp_nv12.restype = None
        L.vaw_oracle_warp_bgr.argtypes = [u8, C.c_int, C.c_int, C.c_int, u8, C.c_int, C.c_int,
                                          C.c_int, ip, fp, u8, C.c_int]
        L.vaw_oracle_warp_bgr.restype = None
        L.vaw_oracle_touched_bytes.argtypes = [fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.vaw_oracle_touched_bytes.restype = C.c_int64
        L.vaw_oracle_get_preset_camera.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(Camera)]
        L.vaw_oracle_get_preset_camera.restype = None
        L.vaw_oracle_get_output_camera.argtypes = [C.POINTER(Camera), C.c_double, C.c_int, C.c_double,
                                                   C.POINTER(Camera)]
        L.vaw_oracle_get_output_camera.restype = None
        L.vaw_oracle_nv12_to_bgr.argtypes = [u8, C.c_int, C.c_int, C.c_int, u8, C.c_int, C.c_int]
        L.vaw_oracle_nv12_to_bgr.restype = None
        L.vaw_oracle_synth_nv12.argtypes = [u8, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int]
        L.vaw_oracle_synth_nv12.restype = None
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


# ---- oracle/_ref: the reference's own createMap.cl, compiled unmodified by oracle/ref_build ------
_REF_DIR = os.path.join(_HERE, "_ref")
_REF_SO = os.path.join(_REF_DIR, "libcreatemap_ref.so")
_REF_BUILD = os.path.join(_HERE, "ref_build")
REF_SOURCE = "/root/reference/opencv/createMap.cl"
_ref_lib = None


def build_ref(force=False):
    """gcc -x c -include cl_shim.h /root/reference/opencv/createMap.cl -> oracle/_ref/libcreatemap_ref.so.
    Only where the reference tree is mounted (the authoring container); the GPU box uses the prebuilt
    file that travels with the snapshot.  Returns the path, or None when neither exists."""
    if os.path.exists(REF_SOURCE):
        deps = [REF_SOURCE] + [os.path.join(_REF_BUILD, f) for f in ("cl_shim.h", "createmap_driver.c", "Makefile")]
        if force or not os.path.exists(_REF_SO) or os.path.getmtime(_REF_SO) < max(os.path.getmtime(d) for d in deps):
            subprocess.check_call(["make", "-C", _REF_BUILD, "-s"])
    return _REF_SO if os.path.exists(_REF_SO) else None


def ref_available():
    return build_ref() is not None


def ref_create_map(k, rot, rows, cols, threads=1, sentinel=None):
    """The map createMap.cl writes (the reference's own kernel source, run on the host over the
    NDRange {cols, rows} with the argument binding of FrameSourceWarp.cpp:275-300).  Ignores k.dist:
    the reference kernel has no distortion term."""
    global _ref_lib
    if _ref_lib is None:
        so = build_ref()
        if so is None:
            raise RuntimeError("oracle/_ref is not built and /root/reference is not mounted")
        L = C.CDLL(so)
        fp = C.POINTER(C.c_float)
        L.vaw_ref_create_map.argtypes = [fp, fp, C.c_int, C.c_int, C.c_int, fp, fp, C.c_int]
        L.vaw_ref_create_map.restype = None
        _ref_lib = L
    if sentinel is None:
        mx = np.empty((rows, cols), np.float32)
        my = np.empty((rows, cols), np.float32)
    else:  # tests: pre-fill so that a work-item that never ran is visible
        mx = np.full((rows, cols), sentinel, np.float32)
        my = np.full((rows, cols), sentinel, np.float32)
    kk = np.array([k.src_center_x, k.src_center_y, k.src_focal_x, k.src_focal_y,
                   k.map_center_x, k.map_center_y, k.map_focal_x, k.map_focal_y], np.float32)
    r = rot32(rot)
    _ref_lib.vaw_ref_create_map(_fp(mx), _fp(my), rows, cols, cols * 4, _fp(kk), _fp(r), threads)
    return mx, my


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def intrinsics(K_in, K_out, dist=None):
    """double -> float cast exactly where the reference does it (FrameSourceWarp.cpp:283-290).
    dist: optional k1..k4 of the input camera's cv::fisheye distortion (extension; None = createMap.cl)."""
    K_in = np.asarray(K_in, dtype=np.float64)
    K_out = np.asarray(K_out, dtype=np.float64)
    k = Intrinsics(np.float32(K_in[0, 2]), np.float32(K_in[1, 2]), np.float32(K_in[0, 0]),
                   np.float32(K_in[1, 1]), np.float32(K_out[0, 2]), np.float32(K_out[1, 2]),
                   np.float32(K_out[0, 0]), np.float32(K_out[1, 1]))
    if dist is not None:
        for i in range(4):
            k.dist[i] = np.float32(dist[i])
    return k


def rot32(rot):
    """3x3 double -> 9 floats row-major (FrameSourceWarp.cpp:291-299)."""
    return np.ascontiguousarray(np.asarray(rot, dtype=np.float64).reshape(9).astype(np.float32))


def create_map(k, rot, rows, cols, threads=1):
    mx = np.empty((rows, cols), np.float32)
    my = np.empty((rows, cols), np.float32)
    r = rot32(rot)
    lib().vaw_oracle_create_map(_fp(mx), _fp(my), rows, cols, cols, C.byref(k), _fp(r), threads)
    return mx, my


def chroma_map(mx, my, threads=1):
    rows, cols = mx.shape
    cx = np.empty((rows // 2, cols // 2), np.float32)
    cy = np.empty_like(cx)
    lib().vaw_oracle_chroma_map(_fp(mx), _fp(my), rows, cols, mx.strides[0] // 4,
                                _fp(cx), _fp(cy), cols // 2, threads)
    return cx, cy


def cubic_table():
    """cv::remap's INTER_CUBIC weights: (32, 32, 4, 4) int16, [fraction y][fraction x][tap row][tap column]."""
    p = lib().vaw_oracle_cubic_table()
    return np.ctypeslib.as_array(p, shape=(32 * 32 * 16,)).reshape(32, 32, 4, 4).copy()


def lanczos4_table():
    """cv::remap's INTER_LANCZOS4 weights: (32, 32, 8, 8) int16."""
    p = lib().vaw_oracle_lanczos4_table()
    return np.ctypeslib.as_array(p, shape=(32 * 32 * 64,)).reshape(32, 32, 8, 8).copy()


def remap_u8(src, mx, my, border=None, threads=1, cubic=False, lanczos4=False):
    """src: (H, W) or (H, W, cn) uint8; returns (rows, cols[, cn]).  cubic / lanczos4: INTER_CUBIC / INTER_LANCZOS4
    instead of INTER_LINEAR."""
    src = np.ascontiguousarray(src)
    cn = 1 if src.ndim == 2 else src.shape[2]
    h, w = src.shape[:2]
    mx = np.ascontiguousarray(mx, dtype=np.float32)
    my = np.ascontiguousarray(my, dtype=np.float32)
    rows, cols = mx.shape
    dst = np.empty((rows, cols) if src.ndim == 2 else (rows, cols, cn), np.uint8)
    b = np.zeros(4, np.uint8)
    if border is not None:
        b[:cn] = np.asarray(border, dtype=np.uint8).reshape(-1)[:cn]
    fn = lib().vaw_oracle_remap_lanczos4_u8 if lanczos4 else (lib().vaw_oracle_remap_cubic_u8 if cubic else lib().vaw_oracle_remap_u8)
    fn(_u8(src), w, h, src.strides[0], cn, _fp(mx), _fp(my), rows, cols,
       cols, _u8(dst), dst.strides[0], _u8(b), threads)
    return dst


def warp_nv12(src, src_w, src_h, out_w, out_h, k, rot, border=(0, 128, 128), threads=1):
    """src: (3*src_h/2, src_w) uint8 NV12 buffer -> (3*out_h/2, out_w)."""
    src = np.ascontiguousarray(src)
    assert src.shape == (src_h * 3 // 2, src_w)
    dst = np.empty((out_h * 3 // 2, out_w), np.uint8)
    r = rot32(rot)
    lib().vaw_oracle_warp_nv12(_u8(src), src_w, src_h, src.strides[0], _u8(dst), out_w, out_h,
                               dst.strides[0], C.byref(k), _fp(r), int(border[0]), int(border[1]),
                               int(border[2]), threads)
    return dst


def warp_bgr(src, out_w, out_h, k, rot, border=(0, 0, 0), threads=1):
    src = np.ascontiguousarray(src)
    h, w, _ = src.shape
    dst = np.empty((out_h, out_w, 3), np.uint8)
    r = rot32(rot)
    b = np.asarray(border, dtype=np.uint8)
    lib().vaw_oracle_warp_bgr(_u8(src), w, h, src.strides[0], _u8(dst), out_w, out_h,
                              dst.strides[0], C.byref(k), _fp(r), _u8(b), threads)
    return dst


def nv12_to_bgr(src, w, h, threads=1):
    """cv::cvtColor(COLOR_YUV2BGR_NV12): (3h/2, w) uint8 -> (h, w, 3)."""
    src = np.ascontiguousarray(src)
    dst = np.empty((h, w, 3), np.uint8)
    lib().vaw_oracle_nv12_to_bgr(_u8(src), w, h, src.strides[0], _u8(dst), dst.strides[0], threads)
    return dst


def touched_bytes(mx, my, src_w, src_h, cn=1):
    mx = np.ascontiguousarray(mx, dtype=np.float32)
    my = np.ascontiguousarray(my, dtype=np.float32)
    return int(lib().vaw_oracle_touched_bytes(_fp(mx), _fp(my), mx.shape[0], mx.shape[1],
                                              mx.shape[1], src_w, src_h, cn))


def get_preset_camera(preset, width, height):
    cam = Camera()
    lib().vaw_oracle_get_preset_camera(int(preset), width, height, C.byref(cam))
    return cam


def get_output_camera(cam, scale=1.0, crop_borders=False, zoom=1.0):
    out = Camera()
    lib().vaw_oracle_get_output_camera(C.byref(cam), scale, int(crop_borders), zoom, C.byref(out))
    return out


def synth_nv12(w, h, frame_index=0, seed=20260001, white_noise=False):
    dst = np.empty((h * 3 // 2, w), np.uint8)
    lib().vaw_oracle_synth_nv12(_u8(dst), w, h, w, frame_index, seed, int(white_noise))
    return dst


def reference_create_map(k, rot, rows, cols, threads=1):
    """The coordinate oracle the parity tests check against: the reference's own kernel
    (oracle/_ref, compiled from createMap.cl) when it is built -- else, or when the fisheye
    distortion extension is in use (createMap.cl has no such term), the transcription, which
    tests/test_oracle_ref.py pins to oracle/_ref bit for bit.  Returns (map_x, map_y, kind)."""
    if not any(k.dist[:]) and ref_available():
        mx, my = ref_create_map(k, rot, rows, cols, threads)
        return mx, my, "reference (oracle/_ref: createMap.cl compiled unmodified)"
    mx, my = create_map(k, rot, rows, cols, threads)
    return mx, my, "port (oracle/create_map_ref.c)"
