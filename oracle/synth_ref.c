/*
 * synth_ref.c -- host mirror of the synthetic NV12 content generator.
 * TEST INFRASTRUCTURE ONLY (see vaw_oracle.h).
 *
 * No reference counterpart (decode is out of scope; frames are synthesised,
 * BASELINE.json north_star).  The recipe is integer-only so that this file and
 * the device generator (video_annotator_b200/csrc/vaw_synth.cuh) agree bit for
 * bit: two triangle waves that drift with the frame index (band-limited:
 * neighbouring samples differ by <= 19, so a 1/32-px bucket flip moves an output
 * sample by < 1 LSB) plus +-8 of hash noise; or pure white noise (worst case).
 */
#include "vaw_oracle.h"

static inline uint32_t mix32(uint32_t h)
{
    h ^= h >> 16; h *= 0x7FEB352Du;
    h ^= h >> 15; h *= 0x846CA68Bu;
    h ^= h >> 16;
    return h;
}

uint32_t vaw_oracle_hash32(uint32_t seed, uint32_t n, uint32_t plane, uint32_t y, uint32_t x)
{
    uint32_t h = seed;
    h = mix32(h ^ (n * 0x9E3779B1u));
    h = mix32(h ^ (plane * 0x85EBCA77u + y * 0xC2B2AE3Du));
    h = mix32(h ^ (x * 0x27D4EB2Fu));
    return h;
}

static inline int tri(int t, int period) /* 0 .. period/2 */
{
    int ph = t % period;
    if (ph < 0) ph += period;
    return ph < period / 2 ? ph : period - ph;
}

static inline uint8_t synth_byte(int plane, int y, int xb, int n, uint32_t seed, int white)
{
    uint32_t h = vaw_oracle_hash32(seed, (uint32_t)n, (uint32_t)plane, (uint32_t)y, (uint32_t)xb);
    if (white) return (uint8_t)(h & 255u);
    int X = xb, Y = y;
    if (plane) { /* UV: luma-scale position, per-channel phase */
        X = 2 * (xb >> 1) + 37 * (1 + (xb & 1));
        Y = 2 * y;
    }
    int v = 128 + (tri(X + 3 * n, 148) * 120) / 74 - 60
                + (tri(Y - 2 * n, 92) * 100) / 46 - 50
                + (int)((h & 31u) >> 1) - 8;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

void vaw_oracle_synth_nv12(uint8_t *dst, int w, int h, int pitch, int frame_index,
                           uint32_t seed, int white_noise)
{
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            dst[(long)y * pitch + x] = synth_byte(0, y, x, frame_index, seed, white_noise);
    for (int y = 0; y < h / 2; ++y)
        for (int x = 0; x < w; ++x)
            dst[(long)(h + y) * pitch + x] = synth_byte(1, y, x, frame_index, seed, white_noise);
}
