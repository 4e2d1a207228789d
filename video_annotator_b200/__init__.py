"""video_annotator_b200 -- B200-native warp path of hedgepigdaniel/video-annotator.

The product is libvaw.so (hand-written CUDA for sm_100a, C-ABI in include/vaw.h) plus the
C++ FrameSource/FrameSourceWarp shim under host/.  This Python package is the thin
ctypes mirror the tests and bench.py call through; torch only provides device memory
and streams.  Nothing here falls back to the CPU.
"""
from ._lib import VawCamera, VawParams, load  # noqa: F401
from .warp import (FORMAT_BGR24, FORMAT_GRAY8, FORMAT_NV12, FORMAT_NV12_TO_BGR24, INTER_CUBIC, INTER_LANCZOS4, INTER_LINEAR, INTER_NEAREST, Camera, ClipWarper, FlowTracker, VawError,  # noqa: F401
                   WarpContext,
                   get_output_camera, get_preset_camera, guess_rotation, nv12_to_bgr, remap_u8, selftest_math, shard_range, synth_nv12)
