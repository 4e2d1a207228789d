/*
 * cvt_ref.c -- cv::cvtColor(COLOR_YUV2BGR_NV12) restated in integers.
 * TEST INFRASTRUCTURE ONLY (see vaw_oracle.h).
 *
 * Reference call site: /root/reference/opencv/FrameSourceWarp.cpp:399-401 (the NV12 buffer is
 * converted to BGR before it is buffered and warped).  The algorithm lives in OpenCV imgproc
 * (third-party, not vendored; opencv4 >= 4.5 per opencv/meson.build:33; 4.13.0 in this image).
 * Its published fixed-point form (ITU-R BT.601, 20 fractional bits), restated:
 *     y = max(0, Y - 16) * 1220542
 *     R = (y + 1673527 * (V - 128)                        + 2^19) >> 20
 *     G = (y -  852492 * (V - 128) - 409993 * (U - 128)   + 2^19) >> 20
 *     B = (y + 2116026 * (U - 128)                        + 2^19) >> 20
 * saturated to [0, 255]; each (U, V) pair serves its 2x2 block of luma samples.
 * PINNED: bit-exact against cv2.cvtColor in tests/test_oracle_cvt.py and
 * tests/golden/cvt_nv12_bgr.npz.
 */
#include "vaw_oracle.h"
#include "par_rows.h"

static inline uint8_t sat8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

typedef struct { const uint8_t *src; int w, h, src_pitch; uint8_t *dst; int dst_pitch; } cvt_job;

static void cvt_rows(int r0, int r1, void *p)
{
    cvt_job *j = (cvt_job *)p;
    for (int y = r0; y < r1; ++y) {
        const uint8_t *yrow = j->src + (long)y * j->src_pitch;
        const uint8_t *uvrow = j->src + (long)(j->h + y / 2) * j->src_pitch;
        uint8_t *out = j->dst + (long)y * j->dst_pitch;
        for (int x = 0; x < j->w; ++x) {
            const int u = uvrow[x & ~1] - 128, v = uvrow[(x & ~1) + 1] - 128;
            int yy = yrow[x] - 16;
            if (yy < 0) yy = 0;
            yy *= 1220542;
            out[3 * x + 0] = sat8((yy + 2116026 * u + (1 << 19)) >> 20);
            out[3 * x + 1] = sat8((yy - 852492 * v - 409993 * u + (1 << 19)) >> 20);
            out[3 * x + 2] = sat8((yy + 1673527 * v + (1 << 19)) >> 20);
        }
    }
}

void vaw_oracle_nv12_to_bgr(const uint8_t *src, int w, int h, int src_pitch, uint8_t *dst, int dst_pitch, int n_threads)
{
    cvt_job j = {src, w, h, src_pitch, dst, dst_pitch};
    vaw_par_rows(h, n_threads, cvt_rows, &j);
}
