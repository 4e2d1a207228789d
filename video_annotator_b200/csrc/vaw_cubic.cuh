// vaw_cubic.cuh -- cv::remap(INTER_CUBIC, BORDER_CONSTANT) on 8-bit data, per tap from global memory.
//
// FrameSourceWarp's constructor takes a cv::InterpolationFlags
// (/root/reference/opencv/FrameSourceWarp.hpp:90) and hands it to cv::remap
// (/root/reference/opencv/FrameSourceWarp.cpp:306-312).  Like the bilinear case, OpenCV's bicubic
// remap is a fixed-point filter: the coordinate is rounded to 1/32 px, the 4 x 4 weights come from a
// table of shorts scaled by 2^15 (Keys kernel, A = -0.75, float outer product, saturate_cast<short>,
// the block sum forced to 2^15 through one of its four central entries), a tap outside the image
// contributes the border value, and the result is saturate_cast<uchar>((sum + 2^14) >> 15).
// INTER_LANCZOS4 is the same scheme with 8 x 8 taps starting three samples up and left and OpenCV's
// interpolateLanczos4 weights (sin / cos in double, normalised in float).
// build_cubic_table() / build_lanczos4_table() are this library's own host-side constructions of those
// tables; tests compare them entry for entry with the oracle's and the outputs bit for bit with cv2.remap.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "vaw_sample.cuh"

namespace vaw {

constexpr int kCubicTabEntries = 32 * 32 * 16;

constexpr int kLanczosTabEntries = 32 * 32 * 64;

namespace detail {

inline void keys_coeffs(float x, float* c)
{
    const float A = -0.75f;
    c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
    c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
    c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
    c[3] = 1.f - c[0] - c[1] - c[2];
}

inline void lanczos4_coeffs(float x, float* c)
{
    const double s45 = 0.70710678118654752440084436210485, pi = 3.1415926535897932384626433832795;
    const double cs[8][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
    float sum = 0;
    const double y0 = -(x + 3) * pi * 0.25, s0 = sin(y0), c0 = cos(y0);
    for (int i = 0; i < 8; ++i) {
        const float d = (x + 3 - i);
        if (fabsf(d) >= 1e-6f) {
            const double y = -d * pi * 0.25;
            c[i] = (float)((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
        } else {
            c[i] = 1e30f;  // the tap on the sample itself takes all the weight
        }
        sum += c[i];
    }
    sum = 1.f / sum;
    for (int i = 0; i < 8; ++i) c[i] *= sum;
}

inline int to_short(float v)
{
    // saturate_cast<short>(float): round-half-even, then clamp (|v| <= 2^15 here)
    const long q0 = (long)v;
    const float r = v - (float)q0;
    long q = q0;
    if (r > 0.5f || (r == 0.5f && (q0 & 1))) ++q;
    else if (r < -0.5f || (r == -0.5f && (q0 & 1))) --q;
    return (int)(q < -32768 ? -32768 : (q > 32767 ? 32767 : q));
}

// ks = 4 (cubic) or 8 (Lanczos4)
inline void build_table(int16_t* tab, int ks)
{
    float t1[32][8];
    for (int i = 0; i < 32; ++i) {
        if (ks == 4) keys_coeffs(i * (1.f / 32), t1[i]);
        else lanczos4_coeffs(i * (1.f / 32), t1[i]);
    }
    const int h = ks / 2;
    for (int i = 0; i < 32; ++i)
        for (int j = 0; j < 32; ++j) {
            int16_t* it = tab + (i * 32 + j) * ks * ks;
            int isum = 0;
            for (int a = 0; a < ks; ++a)
                for (int b = 0; b < ks; ++b) {
                    const float v = t1[i][a] * t1[j][b];
                    it[a * ks + b] = (int16_t)to_short(v * 32768.f);
                    isum += it[a * ks + b];
                }
            if (isum != 32768) {  // the correction goes to the smallest / largest of the central 2 x 2 entries
                const int diff = isum - 32768;
                int Ma = h, Mb = h, ma = h, mb = h;
                for (int a = h; a < h + 2; ++a)
                    for (int b = h; b < h + 2; ++b) {
                        if (it[a * ks + b] < it[ma * ks + mb]) { ma = a; mb = b; }
                        else if (it[a * ks + b] > it[Ma * ks + Mb]) { Ma = a; Mb = b; }
                    }
                if (diff < 0) it[Ma * ks + Mb] = (int16_t)(it[Ma * ks + Mb] - diff);
                else it[ma * ks + mb] = (int16_t)(it[ma * ks + mb] - diff);
            }
        }
}

}  // namespace detail

inline void build_cubic_table(int16_t* tab) { detail::build_table(tab, 4); }
inline void build_lanczos4_table(int16_t* tab) { detail::build_table(tab, 8); }

#ifdef __CUDACC__
// kCn interleaved channels per sample (1 luma / gray, 2 NV12 chroma, 3 BGR); returns the channels packed
// into bytes 0 .. kCn-1.  `border` likewise packed.
// kKs x kKs taps: 4 = INTER_CUBIC, 8 = INTER_LANCZOS4.
template <int kCn, int kKs>
__device__ __forceinline__ unsigned sample_taps(const uint8_t* __restrict__ plane, int pitch, int w, int h,
                                                float mx, float my, unsigned border, const int16_t* __restrict__ tab)
{
    const int sx = fix5(mx), sy = fix5(my);
    // saturate_cast<short> of the integer part, then the block starts kKs / 2 - 1 samples up and left
    const int ix = max(-32768, min(32767, sx >> 5)) - (kKs / 2 - 1), iy = max(-32768, min(32767, sy >> 5)) - (kKs / 2 - 1);
    const int16_t* wt = tab + (((sy & 31) << 5) | (sx & 31)) * (kKs * kKs);
    int sum[kCn];
#pragma unroll
    for (int c = 0; c < kCn; ++c) sum[c] = (int)((border >> (8 * c)) & 255u) << 15;
#pragma unroll 4
    for (int a = 0; a < kKs; ++a) {
        const int yy = iy + a;
        if ((unsigned)yy >= (unsigned)h) continue;
        const uint8_t* row = plane + (ptrdiff_t)yy * pitch;
#pragma unroll
        for (int b = 0; b < kKs; ++b) {
            const int xx = ix + b;
            if ((unsigned)xx >= (unsigned)w) continue;
            const int wv = __ldg(wt + a * kKs + b);
#pragma unroll
            for (int c = 0; c < kCn; ++c)
                sum[c] += ((int)__ldg(row + xx * kCn + c) - (int)((border >> (8 * c)) & 255u)) * wv;
        }
    }
    unsigned out = 0;
#pragma unroll
    for (int c = 0; c < kCn; ++c) out |= (unsigned)min(255, max(0, (sum[c] + (1 << 14)) >> 15)) << (8 * c);
    return out;
}

// the filter the context was created for (Geom::cubic_tab / Geom::tab_ks)
template <int kCn>
__device__ __forceinline__ unsigned sample_hi(const uint8_t* __restrict__ plane, int pitch, int w, int h, float mx,
                                              float my, unsigned border, const int16_t* __restrict__ tab, int ks)
{
    return ks == 8 ? sample_taps<kCn, 8>(plane, pitch, w, h, mx, my, border, tab)
                   : sample_taps<kCn, 4>(plane, pitch, w, h, mx, my, border, tab);
}
#endif

}  // namespace vaw
