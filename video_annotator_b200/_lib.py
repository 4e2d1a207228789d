"""ctypes binding of include/vaw.h (libvaw.so).  No CPU fallback: import fails loudly
when the CUDA library has not been built."""
import ctypes as C
import os

from . import _build

u8p = C.POINTER(C.c_uint8)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)


class VawParams(C.Structure):
    """vaw_params (include/vaw.h)."""
    _fields_ = [
        ("src_center_x", C.c_double), ("src_center_y", C.c_double),
        ("src_focal_x", C.c_double), ("src_focal_y", C.c_double),
        ("map_center_x", C.c_double), ("map_center_y", C.c_double),
        ("map_focal_x", C.c_double), ("map_focal_y", C.c_double),
        ("src_width", C.c_int32), ("src_height", C.c_int32),
        ("out_width", C.c_int32), ("out_height", C.c_int32),
        ("format", C.c_int32), ("interpolation", C.c_int32),
        ("border", C.c_uint8 * 4), ("variant", C.c_int32), ("src_distortion", C.c_float * 4),
        ("projection", C.c_int32), ("reserved", C.c_int32 * 2),
    ]


class VawCamera(C.Structure):
    """vaw_camera (include/vaw.h)."""
    _fields_ = [("model", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("reserved", C.c_int32), ("matrix", C.c_double * 9), ("distortion", C.c_double * 4)]


# every symbol include/vaw.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "vaw_get_preset_camera": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(VawCamera)]),
    "vaw_get_output_camera": (C.c_int, [C.POINTER(VawCamera), C.c_double, C.c_int, C.c_double,
                                        C.POINTER(VawCamera)]),
    "vaw_params_from_cameras": (C.c_int, [C.POINTER(VawCamera), C.POINTER(VawCamera), C.c_int,
                                          C.POINTER(VawParams)]),
    "vaw_create": (C.c_int, [C.POINTER(VawParams), C.c_int, C.POINTER(C.c_void_p)]),
    "vaw_destroy": (None, [C.c_void_p]),
    "vaw_last_error": (C.c_char_p, [C.c_void_p]),
    "vaw_strerror": (C.c_char_p, [C.c_int]),
    "vaw_frame_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "vaw_launch_count": (C.c_uint64, [C.c_void_p]),
    "vaw_get_variant": (C.c_int, [C.c_void_p]),
    "vaw_cubic_table": (C.c_int, [C.POINTER(C.c_int16)]),
    "vaw_lanczos4_table": (C.c_int, [C.POINTER(C.c_int16)]),
    "vaw_warp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, f64p, C.c_void_p]),
    "vaw_warp_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_int,
                                 C.c_size_t, C.c_void_p, C.c_int, C.c_void_p]),
    "vaw_bind_clip": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int]),
    "vaw_upload_rotations": (C.c_int, [C.c_void_p, f64p, C.c_int, C.c_void_p, C.c_void_p]),
    "vaw_warp_batch_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, f64p, C.c_int]),
    "vaw_dump_coords": (C.c_int, [C.c_void_p, f64p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "vaw_remap_u8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, u8p, C.c_int, C.c_void_p]),
    "vaw_nv12_to_bgr": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_size_t,
                                  C.c_int, C.c_int, C.c_void_p]),
    "vaw_synth_nv12": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int,
                                 C.c_uint32, C.c_int, C.c_int, C.c_void_p]),
    "vaw_malloc": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]),
    "vaw_free": (C.c_int, [C.c_int, C.c_void_p]),
    "vaw_memcpy": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "vaw_sync": (C.c_int, [C.c_int, C.c_void_p]),
    "vaw_kernel_times": (C.c_int, [C.c_void_p, C.c_int, f32p, f32p, C.POINTER(C.c_int)]),
    "vaw_kernel_times_split": (C.c_int, [C.c_void_p, C.c_int, f32p, f32p, C.POINTER(C.c_int)]),
    "vaw_shard_range": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vaw_device_numa_node": (C.c_int, [C.c_int]),
    "vaw_bind_thread_to_device": (C.c_int, [C.c_int]),
    "vaw_clip_create": (C.c_int, [C.POINTER(VawParams), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]),
    "vaw_clip_destroy": (None, [C.c_void_p]),
    "vaw_clip_warp_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, f64p, C.c_int]),
    "vaw_clip_last_error": (C.c_char_p, [C.c_void_p]),
    "vaw_flow_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "vaw_flow_destroy": (None, [C.c_void_p]),
    "vaw_flow_last_error": (C.c_char_p, [C.c_void_p]),
    "vaw_flow_levels": (C.c_int, [C.c_void_p]),
    "vaw_flow_push_frame": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "vaw_flow_track": (C.c_int, [C.c_void_p, f32p, C.c_int, f32p, u8p, C.c_void_p]),
    "vaw_flow_get_level": (C.c_int, [C.c_void_p, C.c_int, C.c_int, u8p, C.POINTER(C.c_int16), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vaw_flow_corners": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, f32p, C.c_int, C.POINTER(C.c_int), C.c_void_p]),
    "vaw_flow_get_response": (C.c_int, [C.c_void_p, f32p]),
    "vaw_guess_rotation": (C.c_int, [C.c_void_p, C.c_void_p, f32p, f32p, C.c_int, C.c_uint32, f64p, C.POINTER(C.c_int)]),
    "vaw_piece_stats": (C.c_int, [C.c_void_p, f64p, C.POINTER(C.c_uint32), C.c_void_p]),
    "vaw_piece_flags": (C.c_int, [C.c_void_p, f64p, C.POINTER(C.c_uint32), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.POINTER(C.c_int), C.c_void_p]),
    "vaw_piece_tiles": (C.c_int, [C.c_void_p, f64p, C.POINTER(C.c_uint32), C.c_int, C.c_void_p]),
    "vaw_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "vaw_debug_oob_count": (C.c_longlong, [C.c_int]),
    "vaw_selftest_math": (C.c_int, [C.c_int, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint64)]),
}

_lib = None


def load():
    """Load libvaw.so (building it with nvcc if sources are newer)."""
    global _lib
    if _lib is None:
        path = os.environ.get("VAW_LIBRARY")  # A/B builds of tools/build_variants.py
        if not path:
            path = _build.LIB
            if not os.path.exists(path) or _build.needs_build():
                path = _build.build()
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
