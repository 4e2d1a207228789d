/*
 * remap_cubic_ref.c -- integer restatement of cv::remap(INTER_CUBIC, BORDER_CONSTANT) for
 * 8-bit images and CV_32FC1 maps.  TEST INFRASTRUCTURE ONLY (see vaw_oracle.h).
 *
 * Why it exists: FrameSourceWarp's constructor takes a cv::InterpolationFlags
 * (/root/reference/opencv/FrameSourceWarp.hpp:90, used at FrameSourceWarp.cpp:306-312);
 * SURVEY 8 f3 lists INTER_CUBIC among the next rows.  The algorithm lives in OpenCV imgproc
 * (third party, not vendored; the reference pins opencv4 >= 4.5, this image carries 4.13.0).
 * Its published algorithm, restated:
 *   1. coordinates to fixed point with 5 fractional bits exactly as for INTER_LINEAR:
 *      s = cvRound(m * 32), integer part s >> 5 saturated to int16, fraction s & 31;
 *   2. a table of 32 x 32 sub-pixel positions x (4 x 4) weights: the 1-D Keys kernel with
 *      A = -0.75 evaluated in float at k/32 (interpolateCubic), the outer product of the row and
 *      column weights in float, each scaled by 2^15 and converted with saturate_cast<short>
 *      (round-half-even); when the 16 integers do not sum to 2^15 the difference is taken from the
 *      smallest (sum too large) or added to the largest (sum too small) of the four central
 *      entries (rows / columns 2..3 of the 4 x 4 block as OpenCV scans them, first hit wins);
 *   3. the 4 x 4 taps start at (iy - 1, ix - 1); a tap outside the image contributes the border
 *      value; sum = border * 2^15 + sum_inside (tap - border) * w; result =
 *      saturate_cast<uchar>((sum + 2^14) >> 15).
 * PINNED: bit-exact against cv2.remap(INTER_CUBIC) in tests/test_oracle_remap.py (random maps,
 * grid-aligned, border-straddling and non-finite coordinates, overshoot on a checkerboard,
 * 1-3 channels) and tests/golden/remap_cubic.npz.
 */
#include <limits.h>
#include <math.h>
#include <pthread.h>
#include "vaw_oracle.h"
#include "par_rows.h"

static short g_tab[32 * 32 * 16];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void cubic_coeffs(float x, float *c)
{
    const float A = -0.75f;
    c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
    c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
    c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
    c[3] = 1.f - c[0] - c[1] - c[2];
}

static short sat_short_f(float v)
{
    long r = lrintf(v); /* round-half-even in the default rounding mode */
    return (short)(r < -32768 ? -32768 : (r > 32767 ? 32767 : r));
}

static void build_table(void)
{
    float t1[32][4];
    const float scale = 1.f / 32;
    for (int i = 0; i < 32; ++i) cubic_coeffs(i * scale, t1[i]);
    for (int i = 0; i < 32; ++i)
        for (int j = 0; j < 32; ++j) {
            short *it = g_tab + (i * 32 + j) * 16;
            int isum = 0;
            for (int k1 = 0; k1 < 4; ++k1)
                for (int k2 = 0; k2 < 4; ++k2) {
                    float v = t1[i][k1] * t1[j][k2];
                    it[k1 * 4 + k2] = sat_short_f(v * 32768.f);
                    isum += it[k1 * 4 + k2];
                }
            if (isum != 32768) {
                int diff = isum - 32768, Mk1 = 2, Mk2 = 2, mk1 = 2, mk2 = 2;
                for (int k1 = 2; k1 < 4; ++k1)
                    for (int k2 = 2; k2 < 4; ++k2) {
                        if (it[k1 * 4 + k2] < it[mk1 * 4 + mk2]) { mk1 = k1; mk2 = k2; }
                        else if (it[k1 * 4 + k2] > it[Mk1 * 4 + Mk2]) { Mk1 = k1; Mk2 = k2; }
                    }
                if (diff < 0) it[Mk1 * 4 + Mk2] = (short)(it[Mk1 * 4 + Mk2] - diff);
                else it[mk1 * 4 + mk2] = (short)(it[mk1 * 4 + mk2] - diff);
            }
        }
}

const short *vaw_oracle_cubic_table(void)
{
    pthread_once(&g_once, build_table);
    return g_tab;
}

static inline int cv_round_sse(float v)
{
    if (!(v >= -2147483648.0f && v < 2147483648.0f)) return INT_MIN;
    return (int)nearbyintf(v);
}
static inline int saturate_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

typedef struct {
    const uint8_t *src; int src_w, src_h, src_pitch, cn;
    const float *map_x, *map_y; int cols, map_step;
    uint8_t *dst; int dst_pitch; const uint8_t *border;
} cubic_job;

static void cubic_rows(int y0, int y1, void *p)
{
    cubic_job *j = (cubic_job *)p;
    const short *tab = vaw_oracle_cubic_table();
    for (int y = y0; y < y1; ++y) {
        const float *mx = j->map_x + (long)y * j->map_step, *my = j->map_y + (long)y * j->map_step;
        uint8_t *d = j->dst + (long)y * j->dst_pitch;
        for (int x = 0; x < j->cols; ++x) {
            int sx = cv_round_sse(mx[x] * 32.0f), sy = cv_round_sse(my[x] * 32.0f);
            const short *w = tab + ((sy & 31) * 32 + (sx & 31)) * 16;
            int ix = saturate_short(sx >> 5) - 1, iy = saturate_short(sy >> 5) - 1;
            for (int c = 0; c < j->cn; ++c) {
                int cv = j->border[c];
                long sum = (long)cv * 32768;
                for (int a = 0; a < 4; ++a) {
                    int yy = iy + a;
                    if (yy < 0 || yy >= j->src_h) continue;
                    for (int b = 0; b < 4; ++b) {
                        int xx = ix + b;
                        if (xx < 0 || xx >= j->src_w) continue;
                        sum += (long)((int)j->src[(long)yy * j->src_pitch + (long)xx * j->cn + c] - cv) * w[a * 4 + b];
                    }
                }
                long v = (sum + (1 << 14)) >> 15;
                d[x * j->cn + c] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
            }
        }
    }
}

void vaw_oracle_remap_cubic_u8(const uint8_t *src, int src_w, int src_h, int src_pitch, int cn,
                               const float *map_x, const float *map_y, int rows, int cols, int map_step,
                               uint8_t *dst, int dst_pitch, const uint8_t *border, int n_threads)
{
    cubic_job j = {src, src_w, src_h, src_pitch, cn, map_x, map_y, cols, map_step, dst, dst_pitch, border};
    vaw_oracle_cubic_table();
    vaw_par_rows(rows, n_threads, cubic_rows, &j);
}
