/* vaw_atan_poly.h -- the atan polynomial of the fused map generator, in one place.
 *
 * atan(t) = t + t*s*P(s), s = t*t, for t in [0,1]; callers reduce r > 1 with t = 1/r
 * and reflect with pi/2 - p (pi/2 = VAW_ATAN_PIO2_A * VAW_ATAN_PIO2_B inside one FMA).
 * P has degree 8, fitted by tools/fit_atan.py.  The expression uses only IEEE-defined
 * operations, so the device build (vaw_coords.cuh: __fmaf_rn/__fmul_rn) and a host
 * build (fmaf, no contraction) produce identical bits; tools/check_atanf.c measures
 * the error exhaustively on the host (max 1.42 ulp, r <= 1: 0.98 ulp).
 * Replaces the `atan` builtin of /root/reference/opencv/createMap.cl:39.
 */
#ifndef VAW_ATAN_POLY_H
#define VAW_ATAN_POLY_H

#define VAW_ATAN_C8 (-1.7936229706e-03f)
#define VAW_ATAN_C7 (1.0914611630e-02f)
#define VAW_ATAN_C6 (-3.1177856028e-02f)
#define VAW_ATAN_C5 (5.7957604527e-02f)
#define VAW_ATAN_C4 (-8.4034509957e-02f)
#define VAW_ATAN_C3 (1.0952185839e-01f)
#define VAW_ATAN_C2 (-1.4264242351e-01f)
#define VAW_ATAN_C1 (1.9998548925e-01f)
#define VAW_ATAN_C0 (-3.3333298564e-01f)
#define VAW_ATAN_PIO2_A (9.045259356e-01f)
#define VAW_ATAN_PIO2_B (1.736596227e+00f)

/* VAW_FMA(a,b,c) and VAW_MUL(a,b) must round once (fmaf / __fmaf_rn, a*b / __fmul_rn). */
#define VAW_ATAN_REDUCED(t, big, p)                         \
    do {                                                    \
        const float s_ = VAW_MUL((t), (t));                 \
        (p) = VAW_ATAN_C8;                                  \
        (p) = VAW_FMA((p), s_, VAW_ATAN_C7);                \
        (p) = VAW_FMA((p), s_, VAW_ATAN_C6);                \
        (p) = VAW_FMA((p), s_, VAW_ATAN_C5);                \
        (p) = VAW_FMA((p), s_, VAW_ATAN_C4);                \
        (p) = VAW_FMA((p), s_, VAW_ATAN_C3);                \
        (p) = VAW_FMA((p), s_, VAW_ATAN_C2);                \
        (p) = VAW_FMA((p), s_, VAW_ATAN_C1);                \
        (p) = VAW_FMA((p), s_, VAW_ATAN_C0);                \
        (p) = VAW_MUL((p), s_);                             \
        (p) = VAW_FMA((p), (t), (t));                       \
        if (big) (p) = VAW_FMA(VAW_ATAN_PIO2_A, VAW_ATAN_PIO2_B, -(p)); \
    } while (0)

#endif
