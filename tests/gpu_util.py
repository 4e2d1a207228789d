"""Helpers for the -m gpu parity tests: device buffers, the host restatement of the device
coordinate function, mismatch statistics."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_DEVATAN_SRC = os.path.join(HERE, "helpers", "devatan_map.c")
_DEVATAN_SO = os.path.join(HERE, "helpers", "libdevatan.so")
_devatan = None


def devatan_lib():
    """tests/helpers/devatan_map.c: the oracle's createMap transcription with the device's
    atan polynomial substituted -- must equal the GPU's coordinates bit for bit."""
    global _devatan
    if _devatan is None:
        hdr = os.path.join(os.path.dirname(HERE), "video_annotator_b200", "csrc", "vaw_atan_poly.h")
        newest = max(os.path.getmtime(_DEVATAN_SRC), os.path.getmtime(hdr))
        if not os.path.exists(_DEVATAN_SO) or os.path.getmtime(_DEVATAN_SO) < newest:
            subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-fno-fast-math", "-shared",
                                   "-fPIC", "-o", _DEVATAN_SO, _DEVATAN_SRC, "-lm"])
        _devatan = C.CDLL(_DEVATAN_SO)
        fp = C.POINTER(C.c_float)
        _devatan.devatan_create_map.argtypes = [fp, fp, C.c_int, C.c_int, fp, fp]
        _devatan.devatan_create_map.restype = None
    return _devatan


def host_device_map(k, rot, rows, cols):
    """Luma map as the DEVICE computes it, evaluated on the host (k: oracle Intrinsics)."""
    kk = np.array([getattr(k, f[0]) for f in k._fields_[:8]] + list(k.dist[:]), np.float32)
    r = np.ascontiguousarray(np.asarray(rot, np.float64).reshape(9).astype(np.float32))
    mx = np.empty((rows, cols), np.float32)
    my = np.empty((rows, cols), np.float32)
    fp = C.POINTER(C.c_float)
    devatan_lib().devatan_create_map(mx.ctypes.data_as(fp), my.ctypes.data_as(fp), rows, cols,
                                     kk.ctypes.data_as(fp), r.ctypes.data_as(fp))
    return mx, my


def to_dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def bits_equal(a, b):
    """Bit equality of two fp32 arrays, any NaN == any NaN."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb])


def diff_stats(a, b):
    """Mismatch histogram and PSNR between two uint8 arrays."""
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    hist = np.bincount(d.ravel(), minlength=9)
    mse = float(np.mean(d.astype(np.float64) ** 2))
    return {"n": int(d.size), "max": int(d.max()), "differ": float((d > 0).mean()),
            "gt1": float((d > 1).mean()), "hist": [int(v) for v in hist[:9]] + [int(hist[9:].sum())],
            "psnr": float("inf") if mse == 0 else float(10 * np.log10(255.0 ** 2 / mse))}


def oracle_k(oracle, ctx_or_cams):
    """oracle Intrinsics from (input_camera, output_camera) of the library."""
    cin, cout = ctx_or_cams
    return oracle.intrinsics(cin.K, cout.K)
