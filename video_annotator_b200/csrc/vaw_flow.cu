// vaw_flow.cu -- the measurement half of the stabiliser on the GPU: image pyramids, Scharr derivatives and the
// pyramidal Lucas-Kanade tracker that follows corners from one frame to the next.
//
// Replaces find_point_pairs_with_optical_flow (/root/reference/opencv/FrameSourceWarp.cpp:242-270), i.e.
// cv::calcOpticalFlowPyrLK with its defaults (winSize 21x21, maxLevel 3, 30 iterations / eps 0.01, no flags,
// minEigThreshold 1e-4) on the luma plane of consecutive frames.  OpenCV's `video` module is third-party (not
// under /root/reference; the image carries opencv-python-headless 4.13.0): its algorithm is restated in
// oracle/lk_ref.py, which tests/test_oracle_flow.py pins to the real cv2.calcOpticalFlowPyrLK (<= 1e-4 px).
// This file follows the same fixed-point scheme so that the results agree with that oracle bit for bit:
//   - pyramid levels by cv::pyrDown's 5x5 binomial kernel in integers, (sum + 128) >> 8, BORDER_REFLECT_101;
//   - derivatives by the 3x3 Scharr operator into int16 (calcSharrDeriv), zero outside the image;
//   - patches interpolated with 14-bit bilinear weights (the fourth weight is the remainder), image values kept
//     with 5 extra bits, the 2x2 gradient matrix and the mismatch vector accumulated EXACTLY (64-bit integers;
//     OpenCV accumulates the same integers in float lanes) and scaled by 2^-20; every fp32 operation after that
//     in OpenCV's order (the library is compiled with -fmad=false);
//   - per level: up to 30 Newton steps, stop at |delta|^2 <= 1e-4 or when two consecutive steps cancel.
// One warp tracks one point through all levels; its 21x21 template (value, Ix, Iy) lives in shared memory.
#include <cuda_runtime.h>
#include <stdint.h>
#include "vaw_flow.cuh"

namespace vaw {

namespace {

__device__ __forceinline__ int reflect101(int i, int n)
{
    // BORDER_REFLECT_101 for |offset| < n: -1 -> 1, n -> n - 2
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

// cv::pyrDown on 8-bit: dst(y, x) = (sum_ij w_i w_j src(2y + i, 2x + j) + 128) >> 8, w = 1 4 6 4 1
__global__ void __launch_bounds__(256)
pyr_down_kernel(const uint8_t* __restrict__ src, int sw, int sh, int spitch, uint8_t* __restrict__ dst, int dw, int dh,
                int dpitch)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const int w5[5] = {1, 4, 6, 4, 1};
    int acc = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const uint8_t* row = src + (size_t)reflect101(2 * y + i - 2, sh) * spitch;
        int r = 0;
#pragma unroll
        for (int j = 0; j < 5; ++j) r += w5[j] * (int)__ldg(row + reflect101(2 * x + j - 2, sw));
        acc += w5[i] * r;
    }
    dst[(size_t)y * dpitch + x] = (uint8_t)((acc + 128) >> 8);
}

// calcSharrDeriv: (dx, dy) as int16 pairs; the image itself is extended by BORDER_REFLECT_101
__global__ void __launch_bounds__(256)
scharr_kernel(const uint8_t* __restrict__ src, int w, int h, int pitch, short2* __restrict__ deriv, int dpitch /* pairs */)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
    const uint8_t* r0 = src + (size_t)reflect101(y - 1, h) * pitch;
    const uint8_t* r1 = src + (size_t)y * pitch;
    const uint8_t* r2 = src + (size_t)reflect101(y + 1, h) * pitch;
    const int a00 = __ldg(r0 + xm), a01 = __ldg(r0 + x), a02 = __ldg(r0 + xp);
    const int a10 = __ldg(r1 + xm), a12 = __ldg(r1 + xp);
    const int a20 = __ldg(r2 + xm), a21 = __ldg(r2 + x), a22 = __ldg(r2 + xp);
    const int dx = 3 * (a02 - a00) + 10 * (a12 - a10) + 3 * (a22 - a20);
    const int dy = 3 * (a20 - a00) + 10 * (a21 - a01) + 3 * (a22 - a02);
    deriv[(size_t)y * dpitch + x] = make_short2((short)dx, (short)dy);
}

constexpr int kWin = 21, kWinArea = kWin * kWin, kWBits = 14;
constexpr int kWarpsPerBlock = 4;

__device__ __forceinline__ int descale(int v, int n) { return (v + (1 << (n - 1))) >> n; }

struct Weights { int w00, w01, w10, w11; };

__device__ __forceinline__ Weights bilinear_weights(float a, float b)
{
    // cvRound of the three products, the fourth weight is what is left of 2^14
    Weights w;
    const float s = (float)(1 << kWBits);
    w.w00 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.0f, a), __fsub_rn(1.0f, b)), s));
    w.w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, __fsub_rn(1.0f, b)), s));
    w.w10 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.0f, a), b), s));
    w.w11 = (1 << kWBits) - w.w00 - w.w01 - w.w10;
    return w;
}

__device__ __forceinline__ int image_at(const FlowLevel& L, const uint8_t* img, int x, int y)
{
    return (int)__ldg(img + (size_t)reflect101(y, L.h) * L.pitch + reflect101(x, L.w));
}

__device__ __forceinline__ long long warp_sum(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(32 * kWarpsPerBlock)
lk_track_kernel(const FlowPyramid prev, const FlowPyramid next, const float2* __restrict__ prev_pts, int n_pts,
                float2* __restrict__ next_pts, uint8_t* __restrict__ status, int max_iters, float eps2, float min_eig_threshold)
{
    __shared__ short s_patch[kWarpsPerBlock][kWinArea][3];  // value (x32), Ix, Iy of the template window
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int pt = blockIdx.x * kWarpsPerBlock + wib;
    if (pt >= n_pts) return;  // warp-uniform
    short (*patch)[3] = s_patch[wib];
    const float2 p0 = prev_pts[pt];
    const float half = (float)(kWin - 1) * 0.5f;
    float2 nxt = make_float2(0.f, 0.f);
    bool ok = true;
    for (int level = prev.levels - 1; level >= 0; --level) {
        const FlowLevel L = prev.level[level];
        const uint8_t* J = next.level[level].image;
        const float inv = 1.0f / (float)(1 << level);
        const float2 prevPt = make_float2(__fmul_rn(p0.x, inv), __fmul_rn(p0.y, inv));
        if (level == prev.levels - 1) nxt = prevPt;
        else nxt = make_float2(__fmul_rn(nxt.x, 2.0f), __fmul_rn(nxt.y, 2.0f));
        const float ppx = __fsub_rn(prevPt.x, half), ppy = __fsub_rn(prevPt.y, half);
        const int ix = (int)floorf(ppx), iy = (int)floorf(ppy);
        if (ix < -kWin || ix >= L.w || iy < -kWin || iy >= L.h) {
            if (level == 0) ok = false;
            continue;
        }
        const Weights wt = bilinear_weights(__fsub_rn(ppx, (float)ix), __fsub_rn(ppy, (float)iy));
        long long a11 = 0, a12 = 0, a22 = 0;
        __syncwarp();
        for (int k = lane; k < kWinArea; k += 32) {
            const int wy = k / kWin, wx = k - wy * kWin;
            const int x = ix + wx, y = iy + wy;
            const int v = descale(image_at(L, L.image, x, y) * wt.w00 + image_at(L, L.image, x + 1, y) * wt.w01 +
                                  image_at(L, L.image, x, y + 1) * wt.w10 + image_at(L, L.image, x + 1, y + 1) * wt.w11, kWBits - 5);
            // derivatives are zero outside the image (BORDER_CONSTANT padding of the derivative buffer)
            int gx = 0, gy = 0;
            {
                int sx = 0, sy = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int xx = x + (c & 1), yy = y + (c >> 1);
                    const int wgt = c == 0 ? wt.w00 : (c == 1 ? wt.w01 : (c == 2 ? wt.w10 : wt.w11));
                    if ((unsigned)xx < (unsigned)L.w && (unsigned)yy < (unsigned)L.h) {
                        const short2 d = __ldg(L.deriv + (size_t)yy * L.dpitch + xx);
                        sx += d.x * wgt;
                        sy += d.y * wgt;
                    }
                }
                gx = descale(sx, kWBits);
                gy = descale(sy, kWBits);
            }
            patch[k][0] = (short)v; patch[k][1] = (short)gx; patch[k][2] = (short)gy;
            a11 += (long long)gx * gx; a12 += (long long)gx * gy; a22 += (long long)gy * gy;
        }
        __syncwarp();
        const float scale = 1.0f / (float)(1 << 20);
        const float A11 = __fmul_rn((float)warp_sum(a11), scale), A12 = __fmul_rn((float)warp_sum(a12), scale),
                    A22 = __fmul_rn((float)warp_sum(a22), scale);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dd = __fsub_rn(A11, A22);
        const float min_eig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11),
                                                  __fsqrt_rn(__fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.0f, A12), A12)))),
                                        (float)(2 * kWinArea));
        if (min_eig < min_eig_threshold || D < 1.1920929e-07f) {
            if (level == 0) ok = false;
            continue;
        }
        D = __fdiv_rn(1.0f, D);
        float npx = __fsub_rn(nxt.x, half), npy = __fsub_rn(nxt.y, half);
        float pdx = 0.f, pdy = 0.f;
        for (int j = 0; j < max_iters; ++j) {
            const int jx = (int)floorf(npx), jy = (int)floorf(npy);
            if (jx < -kWin || jx >= L.w || jy < -kWin || jy >= L.h) {
                if (level == 0) ok = false;
                break;
            }
            const Weights wj = bilinear_weights(__fsub_rn(npx, (float)jx), __fsub_rn(npy, (float)jy));
            long long b1 = 0, b2 = 0;
            for (int k = lane; k < kWinArea; k += 32) {
                const int wy = k / kWin, wx = k - wy * kWin;
                const int x = jx + wx, y = jy + wy;
                const int v = descale(image_at(L, J, x, y) * wj.w00 + image_at(L, J, x + 1, y) * wj.w01 +
                                      image_at(L, J, x, y + 1) * wj.w10 + image_at(L, J, x + 1, y + 1) * wj.w11, kWBits - 5);
                const int diff = v - patch[k][0];
                b1 += (long long)diff * patch[k][1];
                b2 += (long long)diff * patch[k][2];
            }
            const float B1 = __fmul_rn((float)warp_sum(b1), scale), B2 = __fmul_rn((float)warp_sum(b2), scale);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, B2), __fmul_rn(A22, B1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, B1), __fmul_rn(A11, B2)), D);
            npx = __fadd_rn(npx, dx); npy = __fadd_rn(npy, dy);
            nxt = make_float2(__fadd_rn(npx, half), __fadd_rn(npy, half));
            if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) <= eps2) break;
            if (j > 0 && fabsf(__fadd_rn(dx, pdx)) < 0.01f && fabsf(__fadd_rn(dy, pdy)) < 0.01f) {
                nxt = make_float2(__fsub_rn(nxt.x, __fmul_rn(dx, 0.5f)), __fsub_rn(nxt.y, __fmul_rn(dy, 0.5f)));
                break;
            }
            pdx = dx; pdy = dy;
        }
    }
    if (lane == 0) {
        next_pts[pt] = nxt;
        status[pt] = ok ? 1 : 0;
    }
}

}  // namespace

cudaError_t launch_pyr_down(const uint8_t* src, int sw, int sh, int spitch, uint8_t* dst, int dpitch, cudaStream_t st)
{
    const int dw = (sw + 1) / 2, dh = (sh + 1) / 2;
    dim3 block(32, 8), grid((dw + 31) / 32, (dh + 7) / 8);
    pyr_down_kernel<<<grid, block, 0, st>>>(src, sw, sh, spitch, dst, dw, dh, dpitch);
    return cudaGetLastError();
}

cudaError_t launch_scharr(const uint8_t* src, int w, int h, int pitch, short2* deriv, int dpitch, cudaStream_t st)
{
    dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
    scharr_kernel<<<grid, block, 0, st>>>(src, w, h, pitch, deriv, dpitch);
    return cudaGetLastError();
}

cudaError_t launch_lk_track(const FlowPyramid& prev, const FlowPyramid& next, const float2* prev_pts, int n_pts,
                            float2* next_pts, uint8_t* status, int max_iters, float eps, float min_eig, cudaStream_t st)
{
    if (n_pts <= 0) return cudaSuccess;
    lk_track_kernel<<<(n_pts + kWarpsPerBlock - 1) / kWarpsPerBlock, 32 * kWarpsPerBlock, 0, st>>>(
        prev, next, prev_pts, n_pts, next_pts, status, max_iters, eps * eps, min_eig);
    return cudaGetLastError();
}

}  // namespace vaw
