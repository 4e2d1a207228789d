/*
 * create_map_ref.c -- scalar fp32 restatement of the reference's map generator.
 * TEST INFRASTRUCTURE ONLY (see vaw_oracle.h).
 *
 * Follows /root/reference/opencv/createMap.cl:10-50 statement by statement,
 * with the argument order of /root/reference/opencv/FrameSourceWarp.cpp:280-300.
 * PARITY PINNED: the reference has no tests or golden vectors, but its kernel source
 * compiles unmodified with gcc behind a shim header (oracle/ref_build -> oracle/_ref,
 * built from /root/reference/opencv/createMap.cl where it lies).  This transcription
 * equals that library bit for bit on whole C1/C2/C3/C5 maps (identity to 80 degree
 * rotations, the NaN pixel at r == 0 included) -- tests/test_oracle_ref.py, live and
 * through the committed fixture tests/golden/createmap_ref.npz -- and is additionally
 * cross-checked against cv2.fisheye.initUndistortRectifyMap (tests/test_oracle_map.py).
 * It exists because /root/reference does not travel to the GPU box and because the
 * fisheye-distortion extension below is not in the reference kernel.
 *
 * Build with -O2 -ffp-contract=off (no FMA contraction, no x87): every
 * operation below rounds once to fp32, in the order written.  Where OpenCL leaves
 * the order open (dot(), length()) the order is fixed here and documented.
 */
#include <math.h>
#include "vaw_oracle.h"
#include "par_rows.h"

void vaw_oracle_create_map_point(int gid_x, int gid_y, const vaw_oracle_intrinsics *k,
                                 const float rot[9], float *out_x, float *out_y)
{
    /* createMap.cl:10-11 -- the indices are `short` */
    short map_x = (short)gid_x;
    short map_y = (short)gid_y;

    /* createMap.cl:15-19 -- location vector of the mapped pixel in the output camera */
    float vi0 = ((float)map_x - k->map_center_x) / k->map_focal_x;
    float vi1 = ((float)map_y - k->map_center_y) / k->map_focal_y;
    float vi2 = 1.0f;

    /* createMap.cl:22-30 -- dot(row, v) fixed as ((r0*x + r1*y) + r2*z), no FMA */
    float vr0 = (rot[0] * vi0 + rot[1] * vi1) + rot[2] * vi2;
    float vr1 = (rot[3] * vi0 + rot[4] * vi1) + rot[5] * vi2;
    float vr2 = (rot[6] * vi0 + rot[7] * vi1) + rot[8] * vi2;

    /* createMap.cl:32-35 -- perspective divide, no guard for vr2 <= 0 */
    float c0 = vr0 / vr2;
    float c1 = vr1 / vr2;

    /* createMap.cl:38-39 -- length() fixed as sqrt(x*x + y*y); NaN at radius 0 */
    float radius_identity = sqrtf(c0 * c0 + c1 * c1);
    float theta = atanf(radius_identity);
    if (k->dist[0] != 0.0f || k->dist[1] != 0.0f || k->dist[2] != 0.0f || k->dist[3] != 0.0f) {
        /* extension, not in createMap.cl: the cv::fisheye distortion polynomial (Horner, no FMA) */
        float t2 = theta * theta;
        float poly = 1.0f + t2 * (k->dist[0] + t2 * (k->dist[1] + t2 * (k->dist[2] + t2 * k->dist[3])));
        theta = theta * poly;
    }
    float fisheye_correction = theta / radius_identity;

    /* createMap.cl:48-49 -- center + ((c * k) * focal) */
    *out_x = k->src_center_x + c0 * fisheye_correction * k->src_focal_x;
    *out_y = k->src_center_y + c1 * fisheye_correction * k->src_focal_y;
}

typedef struct {
    float *map_x, *map_y; int cols, step; const vaw_oracle_intrinsics *k; const float *rot;
} map_job;

static void map_rows(int y0, int y1, void *p)
{
    map_job *j = (map_job *)p;
    /* global size {cols, rows}, bounds check createMap.cl:13 */
    for (int y = y0; y < y1; ++y) {
        float *rx = j->map_x + (long)y * j->step;
        float *ry = j->map_y + (long)y * j->step;
        for (int x = 0; x < j->cols; ++x)
            vaw_oracle_create_map_point(x, y, j->k, j->rot, &rx[x], &ry[x]);
    }
}

void vaw_oracle_create_map(float *map_x, float *map_y, int rows, int cols, int step,
                           const vaw_oracle_intrinsics *k, const float rot[9], int n_threads)
{
    map_job j = {map_x, map_y, cols, step, k, rot};
    vaw_par_rows(rows, n_threads, map_rows, &j);
}
