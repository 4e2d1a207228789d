// vaw_sample.cuh -- the integer bilinear filter of cv::remap, per tap from global memory.
//
// Replaces cv::remap(INTER_LINEAR, BORDER_CONSTANT) on 8-bit data as the reference
// calls it at /root/reference/opencv/FrameSourceWarp.cpp:306-312.  OpenCV (imgproc,
// third-party) does NOT filter in floating point: it rounds the coordinate to 1/32 px
// (round-half-even), takes the 2x2 neighbourhood with each out-of-image tap replaced
// by the border value, and blends with integer weights:
//     out = (w00*t00 + w01*t01 + w10*t10 + w11*t11 + 512) >> 10,  w = (32-ax|ax)(32-ay|ay)
// The hardware texture filter does not reproduce this bit for bit (variant TEX, vaw_tex.cu,
// measures it: 1 LSB off on 5 % of white-noise samples, and slower), so the blend is done in
// integer ALU on every default path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vaw {

// cvRound(m * 32) with the SSE semantics OpenCV relies on: round-half-even, and the
// "integer indefinite" INT_MIN for NaN / +-inf / out of int range (so the pixel lands
// far outside the image and takes the border value).
__device__ __forceinline__ int fix5(float m)
{
    float t = __fmul_rn(m, 32.0f);
    int s = __float2int_rn(t);  // cvt.rni.s32.f32: round-half-even, saturating, NaN -> 0
    return (fabsf(t) < 2147483648.0f) ? s : (int)0x80000000;
}

// cv::remap(INTER_NEAREST) takes the sample at (cvRound(x), cvRound(y)) (saturate_cast<short> of the map
// value: round-half-even), border value when that falls outside.  That is the integer filter above on a
// coordinate rounded to a whole pixel first: ax = ay = 0, all the weight on tap 00.  rintf keeps NaN / inf,
// which fix5 then sends far outside.
__device__ __forceinline__ float nearest_coord(float m) { return rintf(m); }

// Two-stage form of the 4-weight blend (algebraically identical integer arithmetic).
__device__ __forceinline__ int blend(int t00, int t01, int t10, int t11, int ax, int ay)
{
    int top = t00 * (32 - ax) + t01 * ax;
    int bot = t10 * (32 - ax) + t11 * ax;
    return (top * (32 - ay) + bot * ay + 512) >> 10;
}

// One luma (1-channel) sample.  `plane`: H rows of `pitch` bytes, W valid columns.
__device__ __forceinline__ int sample_c1(const uint8_t* __restrict__ plane, int pitch, int w, int h,
                                         float mx, float my, int border)
{
    int sx = fix5(mx), sy = fix5(my);
    int ix = sx >> 5, iy = sy >> 5, ax = sx & 31, ay = sy & 31;
    int t00, t01, t10, t11;
    if ((unsigned)ix < (unsigned)(w - 1) && (unsigned)iy < (unsigned)(h - 1)) {
        const uint8_t* p = plane + (size_t)iy * pitch + ix;
        t00 = __ldg(p);
        t01 = __ldg(p + 1);
        t10 = __ldg(p + pitch);
        t11 = __ldg(p + pitch + 1);
    } else {
        bool x0 = (unsigned)ix < (unsigned)w, x1 = (unsigned)(ix + 1) < (unsigned)w;
        bool y0 = (unsigned)iy < (unsigned)h, y1 = (unsigned)(iy + 1) < (unsigned)h;
        const uint8_t* p = plane + (ptrdiff_t)iy * pitch + ix;
        t00 = (x0 && y0) ? __ldg(p) : border;
        t01 = (x1 && y0) ? __ldg(p + 1) : border;
        t10 = (x0 && y1) ? __ldg(p + pitch) : border;
        t11 = (x1 && y1) ? __ldg(p + pitch + 1) : border;
    }
    return blend(t00, t01, t10, t11, ax, ay);
}

// One chroma (2-channel interleaved) sample; returns U | V << 8.
// `plane`: H rows of `pitch` bytes holding W (U,V) pairs; 2-byte aligned rows.
__device__ __forceinline__ unsigned sample_c2(const uint8_t* __restrict__ plane, int pitch, int w,
                                              int h, float mx, float my, unsigned border_uv)
{
    int sx = fix5(mx), sy = fix5(my);
    int ix = sx >> 5, iy = sy >> 5, ax = sx & 31, ay = sy & 31;
    unsigned t00, t01, t10, t11;  // each U | V << 8
    if ((unsigned)ix < (unsigned)(w - 1) && (unsigned)iy < (unsigned)(h - 1)) {
        const uint16_t* p = reinterpret_cast<const uint16_t*>(plane + (size_t)iy * pitch) + ix;
        const uint16_t* q = reinterpret_cast<const uint16_t*>(plane + (size_t)(iy + 1) * pitch) + ix;
        t00 = __ldg(p);
        t01 = __ldg(p + 1);
        t10 = __ldg(q);
        t11 = __ldg(q + 1);
    } else {
        bool x0 = (unsigned)ix < (unsigned)w, x1 = (unsigned)(ix + 1) < (unsigned)w;
        bool y0 = (unsigned)iy < (unsigned)h, y1 = (unsigned)(iy + 1) < (unsigned)h;
        const uint16_t* p = reinterpret_cast<const uint16_t*>(plane + (ptrdiff_t)iy * pitch) + ix;
        const uint16_t* q = reinterpret_cast<const uint16_t*>(plane + (ptrdiff_t)(iy + 1) * pitch) + ix;
        t00 = (x0 && y0) ? __ldg(p) : border_uv;
        t01 = (x1 && y0) ? __ldg(p + 1) : border_uv;
        t10 = (x0 && y1) ? __ldg(q) : border_uv;
        t11 = (x1 && y1) ? __ldg(q + 1) : border_uv;
    }
    int u = blend(t00 & 255, t01 & 255, t10 & 255, t11 & 255, ax, ay);
    int v = blend(t00 >> 8, t01 >> 8, t10 >> 8, t11 >> 8, ax, ay);
    return (unsigned)u | ((unsigned)v << 8);
}

// One 3-channel interleaved sample (the reference's literal BGR case); returns B | G<<8 | R<<16.
__device__ __forceinline__ unsigned sample_c3(const uint8_t* __restrict__ plane, int pitch, int w,
                                              int h, float mx, float my, unsigned border_bgr)
{
    int sx = fix5(mx), sy = fix5(my);
    int ix = sx >> 5, iy = sy >> 5, ax = sx & 31, ay = sy & 31;
    bool x0 = (unsigned)ix < (unsigned)w, x1 = (unsigned)(ix + 1) < (unsigned)w;
    bool y0 = (unsigned)iy < (unsigned)h, y1 = (unsigned)(iy + 1) < (unsigned)h;
    const uint8_t* p = plane + (ptrdiff_t)iy * pitch + (ptrdiff_t)ix * 3;
    unsigned out = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int b = (border_bgr >> (8 * c)) & 255;
        int t00 = (x0 && y0) ? __ldg(p + c) : b;
        int t01 = (x1 && y0) ? __ldg(p + 3 + c) : b;
        int t10 = (x0 && y1) ? __ldg(p + pitch + c) : b;
        int t11 = (x1 && y1) ? __ldg(p + pitch + 3 + c) : b;
        out |= (unsigned)blend(t00, t01, t10, t11, ax, ay) << (8 * c);
    }
    return out;
}

}  // namespace vaw
