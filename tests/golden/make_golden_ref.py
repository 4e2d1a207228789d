#!/usr/bin/env python
"""Generates tests/golden/createmap_ref.npz from oracle/_ref, i.e. from the reference's OWN
kernel source (/root/reference/opencv/createMap.cl compiled unmodified by oracle/ref_build).

Run in the authoring container (the reference tree is not present on the GPU box):
    python tests/golden/make_golden_ref.py
The fixture stores fp32 bit patterns of map windows for the BASELINE geometries under several
rotations (incl. the NaN pixel at r == 0, createMap.cl:38-39, and rays behind the camera), so
that the transcription oracle/create_map_ref.c stays pinned to the reference kernel wherever
the tests run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests.conftest import rotation_xyz  # noqa: E402


def cases():
    """(name, intrinsics[8], rot 3x3, rows, cols, window (y0, x0, h, w))"""
    out = []
    geoms = {
        "C1": (4, 1920, 1080, None),
        "C2": (4, 2704, 1520, None),
        "C3": (4, 3840, 2160, (984.866 * 3840 / 3524, 3840, 2160)),
        "C5": (4, 5312, 2988, (1362.514 * 3840 / 4877, 3840, 2160)),
    }
    rots = {"identity": np.eye(3), "small": rotation_xyz(2.0, -3.0, 1.5), "large": rotation_xyz(10.0, -15.0, 20.0),
            "behind": rotation_xyz(80.0, 10.0, 0.0)}
    for gname, (preset, w, h, explicit) in geoms.items():
        cam = O.get_preset_camera(preset, w, h)
        if explicit is None:
            outc = O.get_output_camera(cam, 1.0, False, 1.0)
            Kout, ow, oh = outc.K, outc.width & ~1, outc.height & ~1
        else:
            f, ow, oh = explicit
            Kout = np.array([[f, 0, (ow - 1) / 2], [0, f, (oh - 1) / 2], [0, 0, 1]])
        k = O.intrinsics(cam.K, Kout)
        kk = np.array([getattr(k, n) for n, _ in k._fields_[:8]], np.float32)
        for rname, R in rots.items():
            # three windows: top-left corner, centre, bottom-right corner
            for wname, (y0, x0) in {"tl": (0, 0), "mid": (oh // 2 - 8, ow // 2 - 16), "br": (oh - 16, ow - 32)}.items():
                out.append((f"{gname}_{rname}_{wname}", kk, R, oh, ow, (y0, x0, 16, 32)))
    # integer principal point + identity: the optical axis hits a pixel centre -> NaN (createMap.cl:38-39)
    kk = np.array([100.0, 80.0, 50.0, 50.0, 8.0, 6.0, 25.0, 25.0], np.float32)
    out.append(("axis_nan", kk, np.eye(3), 12, 16, (0, 0, 12, 16)))
    return out


def main():
    assert O.ref_available(), "oracle/_ref could not be built (is /root/reference mounted?)"
    data = {}
    names = []
    for name, kk, R, rows, cols, (y0, x0, h, w) in cases():
        k = O.Intrinsics(*[float(v) for v in kk])
        mx, my = O.ref_create_map(k, R, rows, cols, threads=os.cpu_count() or 1)
        data[name + "_k"] = kk
        data[name + "_rot"] = np.asarray(R, np.float64)
        data[name + "_shape"] = np.array([rows, cols, y0, x0, h, w], np.int32)
        data[name + "_x"] = mx[y0:y0 + h, x0:x0 + w].view(np.uint32).copy()
        data[name + "_y"] = my[y0:y0 + h, x0:x0 + w].view(np.uint32).copy()
        # a checksum of the WHOLE map as well (order-independent sum of the bit patterns)
        data[name + "_sum"] = np.array([mx.view(np.uint32).astype(np.uint64).sum(),
                                        my.view(np.uint32).astype(np.uint64).sum()], np.uint64)
        names.append(name)
    data["names"] = np.array(names)
    data["source_sha256"] = np.array(open(os.path.join(ROOT, "oracle", "_ref", "SOURCE.sha256")).read().split()[0])
    path = os.path.join(ROOT, "tests", "golden", "createmap_ref.npz")
    np.savez_compressed(path, **data)
    print(path, len(names), "cases", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
