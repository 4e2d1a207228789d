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
 * INTER_LANCZOS4 is the same scheme with 8 x 8 taps starting at (iy - 3, ix - 3); its 1-D weights are
 * OpenCV's interpolateLanczos4: sin / cos of -(x + 3) pi / 4 in double, the eight taps obtained by
 * rotating that pair in steps of 45 degrees and dividing by y^2, a tap at distance < 1e-6 replaced by
 * 1e30 (so that it takes all the weight), normalised in float; the sum correction looks at rows /
 * columns 4..5 of the block.
 * PINNED: bit-exact against cv2.remap(INTER_CUBIC / INTER_LANCZOS4) in tests/test_oracle_remap.py
 * (random maps, grid-aligned, border-straddling and non-finite coordinates, overshoot on a
 * checkerboard, 1-3 channels) and tests/golden/remap_cubic.npz.
 */
#include <limits.h>
#include <math.h>
#include <pthread.h>
#include "vaw_oracle.h"
#include "par_rows.h"

static short g_tab[32 * 32 * 16];
static short g_tab8[32 * 32 * 64];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static pthread_once_t g_once8 = PTHREAD_ONCE_INIT;

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

static void lanczos4_coeffs(float x, float *c)
{
    static const double s45 = 0.70710678118654752440084436210485;
    static const double cs[8][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
    float sum = 0;
    double y0 = -(x + 3) * 3.1415926535897932384626433832795 * 0.25, s0 = sin(y0), c0 = cos(y0);
    for (int i = 0; i < 8; ++i) {
        float y0_ = (x + 3 - i);
        if (fabsf(y0_) >= 1e-6f) {
            double y = -y0_ * 3.1415926535897932384626433832795 * 0.25;
            c[i] = (float)((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
        } else {
            c[i] = 1e30f;
        }
        sum += c[i];
    }
    sum = 1.f / sum;
    for (int i = 0; i < 8; ++i) c[i] *= sum;
}

/* ks = 4 (cubic) or 8 (Lanczos4): the fixed-point 2-D table from the 1-D float weights */
static void build_table_k(short *tab, int ks)
{
    float t1[32][8];
    const float scale = 1.f / 32;
    for (int i = 0; i < 32; ++i) {
        if (ks == 4) cubic_coeffs(i * scale, t1[i]);
        else lanczos4_coeffs(i * scale, t1[i]);
    }
    const int h = ks / 2;
    for (int i = 0; i < 32; ++i)
        for (int j = 0; j < 32; ++j) {
            short *it = tab + (i * 32 + j) * ks * ks;
            int isum = 0;
            for (int k1 = 0; k1 < ks; ++k1)
                for (int k2 = 0; k2 < ks; ++k2) {
                    float v = t1[i][k1] * t1[j][k2];
                    it[k1 * ks + k2] = sat_short_f(v * 32768.f);
                    isum += it[k1 * ks + k2];
                }
            if (isum != 32768) {
                int diff = isum - 32768, Mk1 = h, Mk2 = h, mk1 = h, mk2 = h;
                for (int k1 = h; k1 < h + 2; ++k1)
                    for (int k2 = h; k2 < h + 2; ++k2) {
                        if (it[k1 * ks + k2] < it[mk1 * ks + mk2]) { mk1 = k1; mk2 = k2; }
                        else if (it[k1 * ks + k2] > it[Mk1 * ks + Mk2]) { Mk1 = k1; Mk2 = k2; }
                    }
                if (diff < 0) it[Mk1 * ks + Mk2] = (short)(it[Mk1 * ks + Mk2] - diff);
                else it[mk1 * ks + mk2] = (short)(it[mk1 * ks + mk2] - diff);
            }
        }
}

static void build_table(void) { build_table_k(g_tab, 4); }
static void build_table8(void) { build_table_k(g_tab8, 8); }

const short *vaw_oracle_cubic_table(void)
{
    pthread_once(&g_once, build_table);
    return g_tab;
}

const short *vaw_oracle_lanczos4_table(void)
{
    pthread_once(&g_once8, build_table8);
    return g_tab8;
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
    uint8_t *dst; int dst_pitch; const uint8_t *border; int ks;
} cubic_job;

static void cubic_rows(int y0, int y1, void *p)
{
    cubic_job *j = (cubic_job *)p;
    const int ks = j->ks, ks2 = ks * ks, off = ks / 2 - 1;
    const short *tab = ks == 4 ? vaw_oracle_cubic_table() : vaw_oracle_lanczos4_table();
    for (int y = y0; y < y1; ++y) {
        const float *mx = j->map_x + (long)y * j->map_step, *my = j->map_y + (long)y * j->map_step;
        uint8_t *d = j->dst + (long)y * j->dst_pitch;
        for (int x = 0; x < j->cols; ++x) {
            int sx = cv_round_sse(mx[x] * 32.0f), sy = cv_round_sse(my[x] * 32.0f);
            const short *w = tab + ((sy & 31) * 32 + (sx & 31)) * ks2;
            int ix = saturate_short(sx >> 5) - off, iy = saturate_short(sy >> 5) - off;
            for (int c = 0; c < j->cn; ++c) {
                int cv = j->border[c];
                long sum = (long)cv * 32768;
                for (int a = 0; a < ks; ++a) {
                    int yy = iy + a;
                    if (yy < 0 || yy >= j->src_h) continue;
                    for (int b = 0; b < ks; ++b) {
                        int xx = ix + b;
                        if (xx < 0 || xx >= j->src_w) continue;
                        sum += (long)((int)j->src[(long)yy * j->src_pitch + (long)xx * j->cn + c] - cv) * w[a * ks + b];
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
    cubic_job j = {src, src_w, src_h, src_pitch, cn, map_x, map_y, cols, map_step, dst, dst_pitch, border, 4};
    vaw_oracle_cubic_table();
    vaw_par_rows(rows, n_threads, cubic_rows, &j);
}

void vaw_oracle_remap_lanczos4_u8(const uint8_t *src, int src_w, int src_h, int src_pitch, int cn,
                                  const float *map_x, const float *map_y, int rows, int cols, int map_step,
                                  uint8_t *dst, int dst_pitch, const uint8_t *border, int n_threads)
{
    cubic_job j = {src, src_w, src_h, src_pitch, cn, map_x, map_y, cols, map_step, dst, dst_pitch, border, 8};
    vaw_oracle_lanczos4_table();
    vaw_par_rows(rows, n_threads, cubic_rows, &j);
}
