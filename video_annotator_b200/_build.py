"""Build recipe for libvaw.so (hand-written CUDA for sm_100a behind the C-ABI of include/vaw.h).

nvcc cross-compiles without a GPU.  The library is built IN-TREE
(video_annotator_b200/libvaw.so) so that it travels with the repository snapshot.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvaw.so")

SOURCES = ["vaw_kernels.cu", "vaw_pieces.cu", "vaw_poly.cu", "vaw_tile.cu", "vaw_packed_tile.cu", "vaw_bgr.cu", "vaw_tex.cu", "vaw_flow.cu", "vaw_corners.cu", "vaw_flow_api.cu", "vaw_api.cu", "vaw_camera.cpp", "vaw_clip.cpp"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # the coordinate code is written with explicit _rn intrinsics; this is belt and braces
    "-fmad=false",
    "-Xcompiler", "-fPIC,-O2,-Wall,-pthread",
    "--shared", "-cudart", "static",
    "--threads", "0",  # the translation units in parallel
]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libvaw.so cannot be built (there is no CPU fallback)")
    return exe


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "vaw.h"))
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    env = dict(os.environ)
    # use the system host compiler regardless of CC/CXX wrappers
    cmd = [nvcc(), *NVCC_FLAGS, "-ccbin", shutil.which("g++") or "g++", "-o", LIB,
           *[os.path.join(CSRC, s) for s in SOURCES]]
    for macro in os.environ.get("VAW_DEFINES", "").split():
        cmd.insert(1, "-D" + macro)
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd, env=env)
    return LIB


HOST = os.path.join(HERE, "host")
DEMO = os.path.join(HOST, "vaw_demo")


def build_host_shim(force=False):
    """g++ build of the C++ FrameSource / FrameSourceWarp shim + demo (pure C++ over the C-ABI:
    no CUDA headers, links libvaw.so only).  -Werror like the reference's meson.build:11."""
    srcs = [os.path.join(HOST, "vaw_demo.cpp"), os.path.join(HOST, "FrameSourceWarp.cpp")]
    deps = srcs + [os.path.join(HOST, "FrameSourceWarp.hpp"), os.path.join(HOST, "FrameSource.hpp"), LIB]
    if not force and os.path.exists(DEMO) and all(os.path.getmtime(d) <= os.path.getmtime(DEMO) for d in deps):
        return DEMO
    subprocess.check_call([shutil.which("g++") or "g++", "-std=c++14", "-O2", "-Wall", "-Werror", "-o", DEMO, *srcs,
                           "-L" + HERE, "-l:libvaw.so", "-Wl,-rpath," + HERE, "-pthread"])
    return DEMO


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
