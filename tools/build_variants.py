#!/usr/bin/env python
"""Build extra copies of libvaw.so with different compile-time switches, for A/B runs on the GPU box:

    python tools/build_variants.py name1="-DVAW_X=1 -DVAW_Y" name2="..."

Objects of the translation units that do not see the switches are compiled once (build/obj); per variant
only the listed units (default: vaw_tile.cu) are recompiled.  Output: build/variants/libvaw_<name>.so; select
one at run time with VAW_LIBRARY=<path> (video_annotator_b200/_lib.py)."""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_annotator_b200 import _build  # noqa: E402

OBJ = os.path.join(ROOT, "build", "obj")
OUT = os.path.join(ROOT, "build", "variants")
UNITS = os.environ.get("VAW_VARIANT_UNITS", "vaw_tile.cu").split()
CFLAGS = [f for f in _build.NVCC_FLAGS if f not in ("--shared",)]


def compile_one(src, obj, defs):
    cmd = [_build.nvcc(), *CFLAGS, "-ccbin", shutil.which("g++") or "g++", *defs, "-c", os.path.join(_build.CSRC, src), "-o", obj]
    subprocess.check_call(cmd)
    return obj


def main():
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(OUT, exist_ok=True)
    variants = dict(a.split("=", 1) for a in sys.argv[1:])
    common = [s for s in _build.SOURCES if s not in UNITS]
    newest = max(os.path.getmtime(os.path.join(_build.CSRC, f)) for f in os.listdir(_build.CSRC))
    jobs = []
    with ThreadPoolExecutor(max_workers=8) as ex:
        for s in common:
            obj = os.path.join(OBJ, s + ".o")
            if not os.path.exists(obj) or os.path.getmtime(obj) < newest:
                jobs.append(ex.submit(compile_one, s, obj, []))
        for name, defs in variants.items():
            for s in UNITS:
                jobs.append(ex.submit(compile_one, s, os.path.join(OBJ, f"{s}.{name}.o"), defs.split()))
        for j in jobs:
            j.result()
    for name in variants:
        objs = [os.path.join(OBJ, s + ".o") for s in common] + [os.path.join(OBJ, f"{s}.{name}.o") for s in UNITS]
        lib = os.path.join(OUT, f"libvaw_{name}.so")
        subprocess.check_call([_build.nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-cudart", "static",
                               "-Xcompiler", "-fPIC,-pthread", "-o", lib, *objs])
        print(lib)


if __name__ == "__main__":
    main()
