/*
 * remap_ref.c -- integer restatement of cv::remap(INTER_LINEAR, BORDER_CONSTANT)
 * for 8-bit images and CV_32FC1 maps.  TEST INFRASTRUCTURE ONLY (see vaw_oracle.h).
 *
 * Reference call site: /root/reference/opencv/FrameSourceWarp.cpp:306-312
 * (border mode/value left at the OpenCV defaults BORDER_CONSTANT / Scalar()).
 * The algorithm lives in a third-party dependency that is not vendored in the
 * reference: OpenCV imgproc (reference pins opencv4 >= 4.5, opencv/meson.build:33;
 * this image carries 4.13.0).  Its published algorithm, restated:
 *   1. each map coordinate is converted to fixed point with 5 fractional bits:
 *      s = cvRound(m * 32)  (round-half-even; NaN / out-of-int-range -> INT_MIN,
 *      the SSE "integer indefinite"); the integer part s >> 5 is stored
 *      saturated to int16, the fraction s & 31 selects the weight row;
 *   2. the four taps (iy,ix) (iy,ix+1) (iy+1,ix) (iy+1,ix+1) are fetched, each one
 *      independently replaced by borderValue when outside the image;
 *   3. the weights are (1-fx)(1-fy), fx(1-fy), (1-fx)fy, fx*fy scaled to 2^15 --
 *      exact multiples of 32 because fx, fy are multiples of 1/32 -- and the
 *      result is (sum + 2^14) >> 15, i.e. (sum_1024 + 512) >> 10.
 * PINNED: bit-exact against cv2.remap in tests/test_oracle_remap.py (random and
 * adversarial maps, 1-3 channels) and tests/golden/remap_*.npz.
 */
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "vaw_oracle.h"
#include "par_rows.h"

/* cvRound(float) as OpenCV's SSE path computes it (cvtss2si / cvtps2dq). */
static inline int cv_round_sse(float v)
{
    if (!(v >= -2147483648.0f && v < 2147483648.0f)) /* NaN, +-inf, overflow */
        return INT_MIN;
    return (int)nearbyintf(v); /* default rounding mode = round-half-even */
}

static inline int saturate_short(int v)
{
    return v < -32768 ? -32768 : (v > 32767 ? 32767 : v);
}

static void remap_row(const uint8_t *src, int src_w, int src_h, int src_pitch, int cn,
                      const float *mx, const float *my, int cols, uint8_t *dst,
                      const uint8_t *border)
{
    for (int x = 0; x < cols; ++x) {
        int sx = cv_round_sse(mx[x] * 32.0f);
        int sy = cv_round_sse(my[x] * 32.0f);
        int ax = sx & 31, ay = sy & 31;
        int ix = saturate_short(sx >> 5); /* arithmetic shift = floor */
        int iy = saturate_short(sy >> 5);
        int w00 = (32 - ax) * (32 - ay), w01 = ax * (32 - ay);
        int w10 = (32 - ax) * ay, w11 = ax * ay;
        int in_x0 = ix >= 0 && ix < src_w, in_x1 = ix + 1 >= 0 && ix + 1 < src_w;
        int in_y0 = iy >= 0 && iy < src_h, in_y1 = iy + 1 >= 0 && iy + 1 < src_h;
        /* tap offsets are only formed for in-range taps */
        long o00 = (long)iy * src_pitch + (long)ix * cn;
        long o01 = o00 + cn, o10 = o00 + src_pitch, o11 = o10 + cn;
        for (int c = 0; c < cn; ++c) {
            int t00 = (in_x0 && in_y0) ? src[o00 + c] : border[c];
            int t01 = (in_x1 && in_y0) ? src[o01 + c] : border[c];
            int t10 = (in_x0 && in_y1) ? src[o10 + c] : border[c];
            int t11 = (in_x1 && in_y1) ? src[o11 + c] : border[c];
            dst[x * cn + c] = (uint8_t)((w00 * t00 + w01 * t01 + w10 * t10 + w11 * t11 + 512) >> 10);
        }
    }
}

typedef struct {
    const uint8_t *src; int src_w, src_h, src_pitch, cn;
    const float *map_x, *map_y; int cols, map_step;
    uint8_t *dst; int dst_pitch; const uint8_t *border;
} remap_job;

static void remap_rows(int y0, int y1, void *p)
{
    remap_job *j = (remap_job *)p;
    for (int y = y0; y < y1; ++y)
        remap_row(j->src, j->src_w, j->src_h, j->src_pitch, j->cn,
                  j->map_x + (long)y * j->map_step, j->map_y + (long)y * j->map_step,
                  j->cols, j->dst + (long)y * j->dst_pitch, j->border);
}

void vaw_oracle_remap_u8(const uint8_t *src, int src_w, int src_h, int src_pitch, int cn,
                         const float *map_x, const float *map_y, int rows, int cols, int map_step,
                         uint8_t *dst, int dst_pitch, const uint8_t *border, int n_threads)
{
    remap_job j = {src, src_w, src_h, src_pitch, cn, map_x, map_y, cols, map_step,
                   dst, dst_pitch, border};
    vaw_par_rows(rows, n_threads, remap_rows, &j);
}

int64_t vaw_oracle_touched_bytes(const float *map_x, const float *map_y, int rows, int cols,
                                 int map_step, int src_w, int src_h, int cn)
{
    uint8_t *mark = (uint8_t *)calloc((size_t)src_w * src_h, 1);
    if (!mark) return -1;
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            int sx = cv_round_sse(map_x[(long)y * map_step + x] * 32.0f);
            int sy = cv_round_sse(map_y[(long)y * map_step + x] * 32.0f);
            int ix = saturate_short(sx >> 5), iy = saturate_short(sy >> 5);
            for (int dy = 0; dy < 2; ++dy)
                for (int dx = 0; dx < 2; ++dx) {
                    int px = ix + dx, py = iy + dy;
                    if (px >= 0 && px < src_w && py >= 0 && py < src_h)
                        mark[(long)py * src_w + px] = 1;
                }
        }
    int64_t n = 0;
    for (long i = 0; i < (long)src_w * src_h; ++i) n += mark[i];
    free(mark);
    return n * cn;
}
