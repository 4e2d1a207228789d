/*
 * createmap_driver.c -- runs the reference's own createMap kernel (compiled from
 * /root/reference/opencv/createMap.cl by oracle/ref_build/Makefile) over a 2-D NDRange on the
 * host.  TEST INFRASTRUCTURE ONLY: it pins oracle/create_map_ref.c (the transcription that
 * travels to the GPU box) to the reference's source, and times the reference's map stage.
 *
 * Arguments are bound exactly as FrameSourceWarp::warp_frame binds them
 * (/root/reference/opencv/FrameSourceWarp.cpp:275-300): KernelArg::WriteOnly(map_x) expands to
 * (ptr, step, offset, rows, cols), WriteOnlyNoSize(map_y) to (ptr, step, offset); then the input
 * camera's cx, cy, fx, fy, the output camera's cx, cy, fx, fy and the rotation row-major, each
 * already cast to float by the caller (the (cl_float) casts of :283-299).  The NDRange is
 * {cols, rows} (:278), rounded up here to a multiple of 16 in x so that the kernel's own bounds
 * check (createMap.cl:13) is exercised as it is on a device.
 */
#include <pthread.h>
#include <stddef.h>

#include "cl_shim.h"

__thread size_t vaw_cl_global_id[3];

/* The reference kernel: its own source file, included from where it lies under /root/reference
 * (VAW_REF_CL is that path, set by the Makefile), so that gcc can inline the work-item function
 * into the NDRange loop below -- the CPU baseline then does not pay a 25-argument call per pixel.
 * The text of the kernel is not touched; with -ffp-contract=off and no fast-math, inlining cannot
 * change a single rounding (tests/test_oracle_ref.py compares whole maps bit for bit). */
#include VAW_REF_CL

typedef struct {
    float *map_x, *map_y;
    int rows, cols, step_bytes;
    const float *k;   /* src cx, cy, fx, fy, map cx, cy, fx, fy */
    const float *rot; /* 9, row-major */
    int y0, y1;
} job_t;

static void *run_rows(void *p)
{
    const job_t *j = (const job_t *)p;
    const int gx = (j->cols + 15) & ~15;
    for (int y = j->y0; y < j->y1; ++y)
        for (int x = 0; x < gx; ++x) {
            vaw_cl_global_id[0] = (size_t)x;
            vaw_cl_global_id[1] = (size_t)y;
            vaw_cl_global_id[2] = 0;
            createMap(j->map_x, j->step_bytes, 0, j->rows, j->cols, j->map_y, j->step_bytes, 0,
                      j->k[0], j->k[1], j->k[2], j->k[3], j->k[4], j->k[5], j->k[6], j->k[7],
                      j->rot[0], j->rot[1], j->rot[2], j->rot[3], j->rot[4], j->rot[5],
                      j->rot[6], j->rot[7], j->rot[8]);
        }
    return NULL;
}

/* map_x / map_y: rows x cols floats, row pitch step_bytes.  gy is rounded up to a multiple of 2. */
void vaw_ref_create_map(float *map_x, float *map_y, int rows, int cols, int step_bytes,
                        const float k[8], const float rot[9], int n_threads)
{
    const int gy = (rows + 1) & ~1;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    if (n_threads > gy) n_threads = gy;
    pthread_t th[256];
    job_t jobs[256];
    for (int t = 0; t < n_threads; ++t) {
        job_t j = {map_x, map_y, rows, cols, step_bytes, k, rot,
                   (int)((long)gy * t / n_threads), (int)((long)gy * (t + 1) / n_threads)};
        jobs[t] = j;
    }
    for (int t = 1; t < n_threads; ++t) pthread_create(&th[t], NULL, run_rows, &jobs[t]);
    run_rows(&jobs[0]);
    for (int t = 1; t < n_threads; ++t) pthread_join(th[t], NULL);
}

const char *vaw_ref_source(void) { return VAW_REF_SOURCE; }
