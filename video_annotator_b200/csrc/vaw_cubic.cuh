// vaw_cubic.cuh -- cv::remap(INTER_CUBIC, BORDER_CONSTANT) on 8-bit data, per tap from global memory.
//
// FrameSourceWarp's constructor takes a cv::InterpolationFlags
// (/root/reference/opencv/FrameSourceWarp.hpp:90) and hands it to cv::remap
// (/root/reference/opencv/FrameSourceWarp.cpp:306-312).  Like the bilinear case, OpenCV's bicubic
// remap is a fixed-point filter: the coordinate is rounded to 1/32 px, the 4 x 4 weights come from a
// table of shorts scaled by 2^15 (Keys kernel, A = -0.75, float outer product, saturate_cast<short>,
// the block sum forced to 2^15 through one of its four central entries), a tap outside the image
// contributes the border value, and the result is saturate_cast<uchar>((sum + 2^14) >> 15).
// build_cubic_table() is this library's own host-side construction of that table;
// tests compare it entry for entry with the oracle's and the outputs bit for bit with cv2.remap.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "vaw_sample.cuh"

namespace vaw {

constexpr int kCubicTabEntries = 32 * 32 * 16;

inline void build_cubic_table(int16_t* tab)
{
    auto keys = [](float x, float* c) {
        const float A = -0.75f;
        c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
        c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
        c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
        c[3] = 1.f - c[0] - c[1] - c[2];
    };
    auto to_short = [](float v) -> int {
        // saturate_cast<short>(float): round-half-even, then clamp
        float r = v - (float)(long)v;            // |v| < 2^15 here: exact
        long q = (long)v;
        if (r > 0.5f || (r == 0.5f && (q & 1))) ++q;
        else if (r < -0.5f || (r == -0.5f && (q & 1))) --q;
        return (int)(q < -32768 ? -32768 : (q > 32767 ? 32767 : q));
    };
    float t1[32][4];
    for (int i = 0; i < 32; ++i) keys(i * (1.f / 32), t1[i]);
    for (int i = 0; i < 32; ++i)
        for (int j = 0; j < 32; ++j) {
            int16_t* it = tab + (i * 32 + j) * 16;
            int isum = 0;
            for (int a = 0; a < 4; ++a)
                for (int b = 0; b < 4; ++b) {
                    const float v = t1[i][a] * t1[j][b];
                    it[a * 4 + b] = (int16_t)to_short(v * 32768.f);
                    isum += it[a * 4 + b];
                }
            if (isum != 32768) {  // the correction goes to the smallest / largest of the central 2 x 2 entries
                const int diff = isum - 32768;
                int Ma = 2, Mb = 2, ma = 2, mb = 2;
                for (int a = 2; a < 4; ++a)
                    for (int b = 2; b < 4; ++b) {
                        if (it[a * 4 + b] < it[ma * 4 + mb]) { ma = a; mb = b; }
                        else if (it[a * 4 + b] > it[Ma * 4 + Mb]) { Ma = a; Mb = b; }
                    }
                if (diff < 0) it[Ma * 4 + Mb] = (int16_t)(it[Ma * 4 + Mb] - diff);
                else it[ma * 4 + mb] = (int16_t)(it[ma * 4 + mb] - diff);
            }
        }
}

#ifdef __CUDACC__
// kCn interleaved channels per sample (1 luma / gray, 2 NV12 chroma, 3 BGR); returns the channels packed
// into bytes 0 .. kCn-1.  `border` likewise packed.
template <int kCn>
__device__ __forceinline__ unsigned sample_cubic(const uint8_t* __restrict__ plane, int pitch, int w, int h,
                                                 float mx, float my, unsigned border, const int16_t* __restrict__ tab)
{
    const int sx = fix5(mx), sy = fix5(my);
    // saturate_cast<short> of the integer part, then the 4 x 4 block starts one sample up and left
    const int ix = max(-32768, min(32767, sx >> 5)) - 1, iy = max(-32768, min(32767, sy >> 5)) - 1;
    const int16_t* wt = tab + (((sy & 31) << 5) | (sx & 31)) * 16;
    int sum[kCn];
#pragma unroll
    for (int c = 0; c < kCn; ++c) sum[c] = (int)((border >> (8 * c)) & 255u) << 15;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int yy = iy + a;
        if ((unsigned)yy >= (unsigned)h) continue;
        const uint8_t* row = plane + (ptrdiff_t)yy * pitch;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int xx = ix + b;
            if ((unsigned)xx >= (unsigned)w) continue;
            const int wv = __ldg(wt + a * 4 + b);
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
#endif

}  // namespace vaw
