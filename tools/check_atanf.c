/*
 * check_atanf.c -- exhaustive error measurement of the device atan used by the
 * fused map generator (vaw_atanf_pos, video_annotator_b200/csrc/vaw_coords.cuh).
 * The device function uses only IEEE-defined operations (rcp.rn, mul.rn, fma.rn,
 * sub.rn), so this host build evaluates bit-identical results.
 *
 *   gcc -O2 -mfma -ffp-contract=off -pthread check_atanf.c -lm -o check_atanf && ./check_atanf
 *
 * Reports, over every non-negative finite float: max error in ulp against a
 * correctly rounded atan, how often the result differs from the correctly rounded
 * value and from this libc's atanf (the oracle's atan, oracle/create_map_ref.c).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../video_annotator_b200/csrc/vaw_atan_poly.h"
#define VAW_FMA(a, b, c) fmaf((a), (b), (c))
#define VAW_MUL(a, b) ((a) * (b))

static inline float vaw_atanf_pos(float r)
{
    int big = r > 1.0f;
    float t = big ? 1.0f / r : r; /* == __frcp_rn(r) */
    float p;
    VAW_ATAN_REDUCED(t, big, p);
    return p;
}

typedef struct { uint32_t lo, hi; double max_ulp; uint64_t n, ne_cr, ne_libm, libm_ne_cr; float worst; } job_t;

static void *run(void *arg)
{
    job_t *j = (job_t *)arg;
    for (uint32_t b = j->lo; b < j->hi; ++b) {
        float x; memcpy(&x, &b, 4);
        float got = vaw_atanf_pos(x);
        double ref = atan((double)x);
        float cr = (float)ref;
        float lm = atanf(x);
        double ulp = ref > 0 ? ldexp(1.0, ilogb(ref) - 23) : 1.0;
        double e = fabs((double)got - ref) / ulp;
        if (ref > 0 && e > j->max_ulp) { j->max_ulp = e; j->worst = x; }
        j->n++;
        j->ne_cr += got != cr;
        j->ne_libm += got != lm;
        j->libm_ne_cr += lm != cr;
    }
    return 0;
}

int main(void)
{
    enum { T = 8 };
    /* normal range: 2^-126 .. max finite; plus denormals trivially */
    uint32_t lo = 0x00000000u, hi = 0x7f800000u;
    pthread_t th[T]; job_t jobs[T];
    for (int t = 0; t < T; ++t) {
        memset(&jobs[t], 0, sizeof(job_t));
        jobs[t].lo = lo + (uint32_t)(((uint64_t)(hi - lo) * t) / T);
        jobs[t].hi = lo + (uint32_t)(((uint64_t)(hi - lo) * (t + 1)) / T);
        pthread_create(&th[t], 0, run, &jobs[t]);
    }
    job_t tot; memset(&tot, 0, sizeof tot);
    for (int t = 0; t < T; ++t) {
        pthread_join(th[t], 0);
        if (jobs[t].max_ulp > tot.max_ulp) { tot.max_ulp = jobs[t].max_ulp; tot.worst = jobs[t].worst; }
        tot.n += jobs[t].n; tot.ne_cr += jobs[t].ne_cr; tot.ne_libm += jobs[t].ne_libm;
        tot.libm_ne_cr += jobs[t].libm_ne_cr;
    }
    printf("inputs=%llu max_err=%.4f ulp at x=%a (%g)\n", (unsigned long long)tot.n, tot.max_ulp, tot.worst, tot.worst);
    printf("differs from correctly rounded: %.4f%%   differs from libc atanf: %.4f%%   (libc atanf vs CR: %.4f%%)\n",
           100.0 * tot.ne_cr / tot.n, 100.0 * tot.ne_libm / tot.n, 100.0 * tot.libm_ne_cr / tot.n);
    return 0;
}
