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

// Corner detection (vaw_corners.cu): cv::cornerMinEigenVal response + its maximum, then the candidate list
// (response bits, y * w + x) of cv::goodFeaturesToTrack; `count` may exceed `capacity` (the list is then truncated).
cudaError_t launch_corner_response(const uint8_t* img, int w, int h, int pitch, float* eig, unsigned* max_bits, cudaStream_t st);
cudaError_t launch_corner_candidates(const float* eig, int w, int h, const unsigned* max_bits, double quality, uint2* list,
                                     unsigned capacity, unsigned* count, cudaStream_t st);

}  // namespace vaw
