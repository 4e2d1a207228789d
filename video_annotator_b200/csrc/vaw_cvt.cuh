// vaw_cvt.cuh -- one pixel of cv::cvtColor(COLOR_YUV2BGR_NV12), the conversion the reference runs on every
// frame before buffering and warping it (/root/reference/opencv/FrameSourceWarp.cpp:399-401): OpenCV's 20-bit
// fixed-point BT.601 (oracle/cvt_ref.c, pinned to cv2.cvtColor by tests/golden/cvt_nv12_bgr.npz).
#pragma once
#include <cuda_runtime.h>

namespace vaw {

__device__ __forceinline__ unsigned sat8(int v) { return (unsigned)min(max(v, 0), 255); }

// u, v already centred (sample - 128)
__device__ __forceinline__ void yuv_pixel(int y, int u, int v, unsigned& b, unsigned& g, unsigned& r)
{
    const int yy = max(y - 16, 0) * 1220542 + (1 << 19);
    b = sat8((yy + 2116026 * u) >> 20);
    g = sat8((yy - 852492 * v - 409993 * u) >> 20);
    r = sat8((yy + 1673527 * v) >> 20);
}

}  // namespace vaw
