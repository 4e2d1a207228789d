/*
 * par_rows.h -- split a row loop over pthreads.  TEST INFRASTRUCTURE ONLY.
 * (No reference counterpart: the reference is single-threaded; the all-cores CPU
 * baseline is this build's measurement set-up)
 */
#ifndef VAW_PAR_ROWS_H
#define VAW_PAR_ROWS_H
#include <pthread.h>

typedef void (*vaw_row_fn)(int y0, int y1, void *ctx);

typedef struct { vaw_row_fn fn; void *ctx; int y0, y1; } vaw_row_job;

static void *vaw_row_tramp(void *p)
{
    vaw_row_job *j = (vaw_row_job *)p;
    j->fn(j->y0, j->y1, j->ctx);
    return 0;
}

static inline void vaw_par_rows(int rows, int n_threads, vaw_row_fn fn, void *ctx)
{
    if (n_threads > 256) n_threads = 256;
    if (n_threads > rows) n_threads = rows;
    if (n_threads <= 1) { fn(0, rows, ctx); return; }
    pthread_t th[256];
    vaw_row_job job[256];
    for (int t = 0; t < n_threads; ++t) {
        job[t].fn = fn; job[t].ctx = ctx;
        job[t].y0 = (int)((long)rows * t / n_threads);
        job[t].y1 = (int)((long)rows * (t + 1) / n_threads);
        if (t + 1 == n_threads) fn(job[t].y0, job[t].y1, ctx);
        else pthread_create(&th[t], 0, vaw_row_tramp, &job[t]);
    }
    for (int t = 0; t + 1 < n_threads; ++t) pthread_join(th[t], 0);
}
#endif
