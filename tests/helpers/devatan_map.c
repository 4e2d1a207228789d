/*
 * devatan_map.c -- TEST HELPER: the oracle's createMap transcription
 * (oracle/create_map_ref.c, /root/reference/opencv/createMap.cl:10-50) with ONE
 * substitution: atan is the device polynomial (video_annotator_b200/csrc/vaw_atan_poly.h)
 * instead of libm's atanf.  Every other operation is IEEE-defined, so the GPU's dumped
 * coordinates must equal this map BIT FOR BIT; that isolates atan (the one step the
 * reference does not define to the bit) from division, sqrt and operation order.
 * Build: gcc -O2 -mfma -ffp-contract=off -shared -fPIC (fmaf must be a true FMA).
 */
#include <math.h>
#include "../../video_annotator_b200/csrc/vaw_atan_poly.h"
#define VAW_FMA(a, b, c) fmaf((a), (b), (c))
#define VAW_MUL(a, b) ((a) * (b))

static float dev_atanf_pos(float r)
{
    int big = r > 1.0f;
    float t = big ? 1.0f / r : r;
    float p;
    VAW_ATAN_REDUCED(t, big, p);
    return p;
}

/* k[12] = scx, scy, sfx, sfy, mcx, mcy, mfx, mfy, k1..k4 (cv::fisheye distortion extension; zeros = createMap.cl) */
void devatan_create_map(float *map_x, float *map_y, int rows, int cols, const float *k, const float *rot)
{
    for (int v = 0; v < rows; ++v)
        for (int u = 0; u < cols; ++u) {
            float x = ((float)(short)u - k[4]) / k[6];
            float y = ((float)(short)v - k[5]) / k[7];
            float q0 = (rot[0] * x + rot[1] * y) + rot[2];
            float q1 = (rot[3] * x + rot[4] * y) + rot[5];
            float q2 = (rot[6] * x + rot[7] * y) + rot[8];
            float c0 = q0 / q2, c1 = q1 / q2;
            float rad = sqrtf(c0 * c0 + c1 * c1);
            float theta = dev_atanf_pos(rad);
            if (k[8] != 0.0f || k[9] != 0.0f || k[10] != 0.0f || k[11] != 0.0f) {
                float t2 = theta * theta;
                theta = theta * (1.0f + t2 * (k[8] + t2 * (k[9] + t2 * (k[10] + t2 * k[11]))));
            }
            float kk = theta / rad;
            map_x[(long)v * cols + u] = k[0] + c0 * kk * k[2];
            map_y[(long)v * cols + u] = k[1] + c1 * kk * k[3];
        }
}
