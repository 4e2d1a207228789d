/*
 * nv12_warp_ref.c -- the whole warp path on the CPU: map, chroma map, remap.
 * TEST INFRASTRUCTURE ONLY (see vaw_oracle.h).
 *
 * Reference path: FrameSourceWarp::warp_frame,
 * /root/reference/opencv/FrameSourceWarp.cpp:272-314 (createMap kernel, then
 * cv::remap with default border).  The reference applies it to a BGR 8UC3 frame
 * (FrameSourceWarp.cpp:401,445): vaw_oracle_warp_bgr is that literal behaviour.
 *
 * NV12-plane semantics are NEW (north_star asks for NV12; no reference behaviour
 * exists).  The NV12 buffer layout is the reference's: one 8-bit plane of
 * W x 3H/2, Y rows 0..H-1, interleaved U,V rows H..3H/2-1
 * (/root/reference/opencv/FrameSourceFfmpegOpenCl.cpp:58,75-85; consumer
 * FrameSourceWarp.cpp:217,399).  Definitions fixed here:
 *  (i)   luma: createMap coordinates, remap with 1 channel, border_y.
 *  (ii)  chroma sample (cx,cy) is centre-sited over the luma quad
 *        (2cx..2cx+1, 2cy..2cy+1).  Its luma-space source position is the mean of
 *        the four createMap outputs of that quad, summed as
 *        ((m00 + m01) + (m10 + m11)) * 0.25 in fp32; the chroma-plane coordinate
 *        is ((s - 0.5) * 0.5).  NaN in any of the four -> NaN -> border.
 *  (iii) chroma: remap with 2 channels on the (H/2) x (W/2) UV plane with border
 *        (border_u, border_v).
 *  (iv)  all widths/heights are even.
 */
#include <stdlib.h>
#include "vaw_oracle.h"
#include "par_rows.h"

typedef struct {
    const float *map_x, *map_y; int cols, step; float *cmap_x, *cmap_y; int cstep;
} cmap_job;

static void cmap_rows(int r0, int r1, void *p)
{
    cmap_job *j = (cmap_job *)p;
    for (int cy = r0; cy < r1; ++cy) {
        const float *x0 = j->map_x + (long)(2 * cy) * j->step, *x1 = x0 + j->step;
        const float *y0 = j->map_y + (long)(2 * cy) * j->step, *y1 = y0 + j->step;
        for (int cx = 0; cx < j->cols / 2; ++cx) {
            float sx = ((x0[2 * cx] + x0[2 * cx + 1]) + (x1[2 * cx] + x1[2 * cx + 1])) * 0.25f;
            float sy = ((y0[2 * cx] + y0[2 * cx + 1]) + (y1[2 * cx] + y1[2 * cx + 1])) * 0.25f;
            j->cmap_x[(long)cy * j->cstep + cx] = (sx - 0.5f) * 0.5f;
            j->cmap_y[(long)cy * j->cstep + cx] = (sy - 0.5f) * 0.5f;
        }
    }
}

void vaw_oracle_chroma_map(const float *map_x, const float *map_y, int rows, int cols, int step,
                           float *cmap_x, float *cmap_y, int cstep, int n_threads)
{
    cmap_job j = {map_x, map_y, cols, step, cmap_x, cmap_y, cstep};
    vaw_par_rows(rows / 2, n_threads, cmap_rows, &j);
}

void vaw_oracle_warp_nv12(const uint8_t *src, int src_w, int src_h, int src_pitch,
                          uint8_t *dst, int out_w, int out_h, int dst_pitch,
                          const vaw_oracle_intrinsics *k, const float rot[9],
                          int border_y, int border_u, int border_v, int n_threads)
{
    size_t n = (size_t)out_w * out_h;
    float *mx = (float *)malloc(n * sizeof(float)), *my = (float *)malloc(n * sizeof(float));
    float *cx = (float *)malloc(n / 4 * sizeof(float)), *cy = (float *)malloc(n / 4 * sizeof(float));
    uint8_t by[1] = {(uint8_t)border_y}, buv[2] = {(uint8_t)border_u, (uint8_t)border_v};

    vaw_oracle_create_map(mx, my, out_h, out_w, out_w, k, rot, n_threads);
    vaw_oracle_remap_u8(src, src_w, src_h, src_pitch, 1, mx, my, out_h, out_w, out_w,
                        dst, dst_pitch, by, n_threads);
    vaw_oracle_chroma_map(mx, my, out_h, out_w, out_w, cx, cy, out_w / 2, n_threads);
    vaw_oracle_remap_u8(src + (size_t)src_pitch * src_h, src_w / 2, src_h / 2, src_pitch, 2,
                        cx, cy, out_h / 2, out_w / 2, out_w / 2,
                        dst + (size_t)dst_pitch * out_h, dst_pitch, buv, n_threads);
    free(mx); free(my); free(cx); free(cy);
}

void vaw_oracle_warp_bgr(const uint8_t *src, int src_w, int src_h, int src_pitch,
                         uint8_t *dst, int out_w, int out_h, int dst_pitch,
                         const vaw_oracle_intrinsics *k, const float rot[9],
                         const uint8_t border[3], int n_threads)
{
    size_t n = (size_t)out_w * out_h;
    float *mx = (float *)malloc(n * sizeof(float)), *my = (float *)malloc(n * sizeof(float));
    vaw_oracle_create_map(mx, my, out_h, out_w, out_w, k, rot, n_threads);
    vaw_oracle_remap_u8(src, src_w, src_h, src_pitch, 3, mx, my, out_h, out_w, out_w,
                        dst, dst_pitch, border, n_threads);
    free(mx); free(my);
}
