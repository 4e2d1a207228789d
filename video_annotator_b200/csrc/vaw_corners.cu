// vaw_corners.cu -- corner detection for the motion measurement: the response map and the candidate list of
// cv::goodFeaturesToTrack.
//
// Replaces find_corners (/root/reference/opencv/FrameSourceWarp.cpp:228-240): goodFeaturesToTrack(image, corners,
// 200, 0.01, 30) on the luma plane, i.e. cv::cornerMinEigenVal with blockSize 3 and a 3x3 Sobel, threshold at
// quality * max, 3x3 non-maximum suppression, then a greedy minimum-distance selection in decreasing response.
// OpenCV's `imgproc` module is third-party (not under /root/reference); the algorithm is restated in
// oracle/gftt_ref.py, pinned to cv2.cornerMinEigenVal (<= 3e-8: OpenCV's own last bit depends on its SIMD / IPP
// path) and to cv2.goodFeaturesToTrack (identical corner lists).  The kernels below do the oracle's fp32 operations
// in the oracle's order (integer Sobel sums scaled once by 1 / 3060; products; 3x3 sums as ((r0 + r1) + r2) by rows,
// then by columns; a = Sxx / 2, c = Syy / 2; (a + c) - sqrt((a - c)^2 + b^2)), so the response equals the oracle's
// bit for bit.  The sequential part -- sorting the few thousand candidates and the greedy distance filter -- runs on
// the host (vaw_flow_api.cu), as it does inside OpenCV.
#include <cuda_runtime.h>
#include <stdint.h>
#include "vaw_flow.cuh"

namespace vaw {

namespace {

__device__ __forceinline__ int reflect101c(int i, int n)
{
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

constexpr int kTx = 32, kTy = 8;  // outputs per CTA

// cv::cornerMinEigenVal(img, 3, 3) for 8-bit input + the maximum of the map (as the bits of a non-negative float)
__global__ void __launch_bounds__(kTx * kTy)
corner_response_kernel(const uint8_t* __restrict__ img, int w, int h, int pitch, float* __restrict__ eig, unsigned* __restrict__ max_bits)
{
    __shared__ float cxx[kTy + 2][kTx + 2], cxy[kTy + 2][kTx + 2], cyy[kTy + 2][kTx + 2];
    __shared__ unsigned wmax[kTx * kTy / 32];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * kTx + tx;
    const int x0 = blockIdx.x * kTx, y0 = blockIdx.y * kTy;
    const float scale = 1.0f / 3060.0f;  // 1 / (2^(3-1) * blockSize 3 * 255)
    // derivative products on the halo tile; a position outside the image takes the products of its mirror image
    // (the box filter reflects the product image, BORDER_REFLECT_101), whose own Sobel reflects the pixels
    for (int i = tid; i < (kTx + 2) * (kTy + 2); i += kTx * kTy) {
        const int ly = i / (kTx + 2), lx = i - ly * (kTx + 2);
        const int gx = reflect101c(min(x0 + lx - 1, w), w), gy = reflect101c(min(y0 + ly - 1, h), h);
        const int xm = reflect101c(gx - 1, w), xp = reflect101c(gx + 1, w);
        const uint8_t* r0 = img + (size_t)reflect101c(gy - 1, h) * pitch;
        const uint8_t* r1 = img + (size_t)gy * pitch;
        const uint8_t* r2 = img + (size_t)reflect101c(gy + 1, h) * pitch;
        const int a00 = __ldg(r0 + xm), a01 = __ldg(r0 + gx), a02 = __ldg(r0 + xp);
        const int a10 = __ldg(r1 + xm), a12 = __ldg(r1 + xp);
        const int a20 = __ldg(r2 + xm), a21 = __ldg(r2 + gx), a22 = __ldg(r2 + xp);
        const float dx = __fmul_rn((float)((a02 - a00) + 2 * (a12 - a10) + (a22 - a20)), scale);
        const float dy = __fmul_rn((float)((a20 - a00) + 2 * (a21 - a01) + (a22 - a02)), scale);
        cxx[ly][lx] = __fmul_rn(dx, dx); cxy[ly][lx] = __fmul_rn(dx, dy); cyy[ly][lx] = __fmul_rn(dy, dy);
    }
    __syncthreads();
    const int x = x0 + tx, y = y0 + ty;
    float e = 0.f;
    if (x < w && y < h) {
        auto box = [&](const float (&c)[kTy + 2][kTx + 2]) {
            const float q0 = __fadd_rn(__fadd_rn(c[ty][tx], c[ty][tx + 1]), c[ty][tx + 2]);
            const float q1 = __fadd_rn(__fadd_rn(c[ty + 1][tx], c[ty + 1][tx + 1]), c[ty + 1][tx + 2]);
            const float q2 = __fadd_rn(__fadd_rn(c[ty + 2][tx], c[ty + 2][tx + 1]), c[ty + 2][tx + 2]);
            return __fadd_rn(__fadd_rn(q0, q1), q2);
        };
        const float a = __fmul_rn(box(cxx), 0.5f), b = box(cxy), c = __fmul_rn(box(cyy), 0.5f);
        const float d = __fsub_rn(a, c);
        e = __fsub_rn(__fadd_rn(a, c), __fsqrt_rn(__fadd_rn(__fmul_rn(d, d), __fmul_rn(b, b))));
        eig[(size_t)y * w + x] = e;
    }
    // maximum of the map: non-negative floats order like their bit patterns
    unsigned bits = e > 0.f ? __float_as_uint(e) : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bits = max(bits, __shfl_xor_sync(0xffffffffu, bits, o));
    if ((tid & 31) == 0) wmax[tid >> 5] = bits;
    __syncthreads();
    if (tid == 0) {
        unsigned m = 0;
        for (int i = 0; i < kTx * kTy / 32; ++i) m = max(m, wmax[i]);
        if (m) atomicMax(max_bits, m);
    }
}

// Candidates of goodFeaturesToTrack: response above the threshold and not smaller than any of its eight
// neighbours (= equal to the 3x3 dilation of the thresholded map), rows and columns 1 .. n-2 only.
__global__ void __launch_bounds__(256)
corner_candidates_kernel(const float* __restrict__ eig, int w, int h, const unsigned* __restrict__ max_bits, double quality,
                         uint2* __restrict__ list, unsigned capacity, unsigned* __restrict__ count)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < 1 || y < 1 || x >= w - 1 || y >= h - 1) return;
    const float thr = (float)((double)__uint_as_float(*max_bits) * quality);
    const float v = __ldg(eig + (size_t)y * w + x);
    if (!(v > thr)) return;
    bool is_max = true;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx)
            if (dx | dy) is_max = is_max && v >= __ldg(eig + (size_t)(y + dy) * w + (x + dx));
    if (!is_max) return;
    const unsigned slot = atomicAdd(count, 1u);
    if (slot < capacity) list[slot] = make_uint2(__float_as_uint(v), (unsigned)(y * w + x));
}

}  // namespace

cudaError_t launch_corner_response(const uint8_t* img, int w, int h, int pitch, float* eig, unsigned* max_bits, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(max_bits, 0, sizeof(unsigned), st);
    if (e != cudaSuccess) return e;
    dim3 block(kTx, kTy), grid((w + kTx - 1) / kTx, (h + kTy - 1) / kTy);
    corner_response_kernel<<<grid, block, 0, st>>>(img, w, h, pitch, eig, max_bits);
    return cudaGetLastError();
}

cudaError_t launch_corner_candidates(const float* eig, int w, int h, const unsigned* max_bits, double quality, uint2* list,
                                     unsigned capacity, unsigned* count, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(unsigned), st);
    if (e != cudaSuccess) return e;
    dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
    corner_candidates_kernel<<<grid, block, 0, st>>>(eig, w, h, max_bits, quality, list, capacity, count);
    return cudaGetLastError();
}

}  // namespace vaw
