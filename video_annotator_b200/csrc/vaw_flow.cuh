// vaw_flow.cuh -- pyramids for the Lucas-Kanade tracker (vaw_flow.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vaw {

constexpr int kFlowMaxLevels = 4;  // maxLevel 3 (cv::calcOpticalFlowPyrLK's default) = 4 images

struct FlowLevel {
    const uint8_t* image;  // w x h, pitch bytes
    const short2* deriv;   // Scharr (dx, dy), dpitch pairs per row
    int w, h, pitch, dpitch;
};

struct FlowPyramid {
    FlowLevel level[kFlowMaxLevels];
    int levels;
};

cudaError_t launch_pyr_down(const uint8_t* src, int sw, int sh, int spitch, uint8_t* dst, int dpitch, cudaStream_t st);
cudaError_t launch_scharr(const uint8_t* src, int w, int h, int pitch, short2* deriv, int dpitch, cudaStream_t st);
cudaError_t launch_lk_track(const FlowPyramid& prev, const FlowPyramid& next, const float2* prev_pts, int n_pts,
                            float2* next_pts, uint8_t* status, int max_iters, float eps, float min_eig, cudaStream_t st);

}  // namespace vaw
