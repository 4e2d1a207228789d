// vaw_synth.cuh -- integer synthetic NV12 content, generated directly in device memory.
//
// No reference counterpart: decode (VAAPI/libav, /root/reference/opencv/AvFrameSource*.cpp)
// is out of scope and BASELINE.json's north_star asks for frames synthesised on the device.
// Integer-only so the host mirror used by the tests produces identical bytes:
// two triangle waves drifting with the frame index (neighbouring samples differ by
// <= 19) plus +-8 of hash noise, or pure white noise.
#pragma once
#include <stdint.h>

namespace vaw {

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t h)
{
    h ^= h >> 16; h *= 0x7FEB352Du;
    h ^= h >> 15; h *= 0x846CA68Bu;
    h ^= h >> 16;
    return h;
}

__host__ __device__ __forceinline__ uint32_t hash32(uint32_t seed, uint32_t n, uint32_t plane,
                                                    uint32_t y, uint32_t x)
{
    uint32_t h = seed;
    h = mix32(h ^ (n * 0x9E3779B1u));
    h = mix32(h ^ (plane * 0x85EBCA77u + y * 0xC2B2AE3Du));
    h = mix32(h ^ (x * 0x27D4EB2Fu));
    return h;
}

__host__ __device__ __forceinline__ int tri_wave(int t, int period)  // 0 .. period/2
{
    int ph = t % period;
    if (ph < 0) ph += period;
    return ph < period / 2 ? ph : period - ph;
}

// plane 0: luma byte (y, xb); plane 1: byte xb of UV row y (channel = xb & 1)
__host__ __device__ __forceinline__ uint8_t synth_byte(int plane, int y, int xb, int n,
                                                       uint32_t seed, int white)
{
    uint32_t h = hash32(seed, (uint32_t)n, (uint32_t)plane, (uint32_t)y, (uint32_t)xb);
    if (white) return (uint8_t)(h & 255u);
    int X = xb, Y = y;
    if (plane) {
        X = 2 * (xb >> 1) + 37 * (1 + (xb & 1));
        Y = 2 * y;
    }
    int v = 128 + (tri_wave(X + 3 * n, 148) * 120) / 74 - 60 + (tri_wave(Y - 2 * n, 92) * 100) / 46 - 50 +
            (int)((h & 31u) >> 1) - 8;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

}  // namespace vaw
