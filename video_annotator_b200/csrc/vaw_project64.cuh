// vaw_project64.cuh -- the projection of /root/reference/opencv/createMap.cl:15-49 in double precision, for a
// (possibly fractional) output position: the anchors and certificates of the piece-table builder
// (vaw_pieces.cu) and the per-pixel evaluation of uncertified pieces under the projection pairs createMap.cl
// does not have.
//
// Projection pairs (vaw_params::projection; CameraModel, opencv/FrameSourceWarp.hpp:23-26; the `in_p` /
// `out_p` = rect | fish options the wider toolchain passes to its external filter, src/render.ts:611-618):
//   0  fisheye input, rectilinear output  -- createMap.cl, the reference's only pair
//   bit 0 set: rectilinear (pinhole) input:  m = c + f * (q.xy / q.z)          (no atan step)
//   bit 1 set: fisheye (equidistant) output: the pixel offset from the centre, in focal lengths, is the
//              ANGLE of the ray from the optical axis: d = (p sin|p| / |p|, cos|p|) instead of (p, 1)
#pragma once
#include <cuda_runtime.h>
#include "vaw_pieces.cuh"

namespace vaw {

struct RotD { double r[9]; };

struct Ray { double mx, my, q0, q1, q2; };

// atan(t) / t on t in [0, 1] as a polynomial in s = t^2 (tools/fit_atan64.py, degree 13: relative error
// 1.8e-12, i.e. < 1e-8 px at 4K; the anchors need ~1e-10).  The library atan() costs a double-precision
// division subroutine per call; together with 1 / q2 those calls were 16 % of the builder's instructions.
__device__ __forceinline__ double atan_over_t(double s)
{
    const double c[14] = {
        9.99999999998195555e-01, -3.33333332624654421e-01, 1.99999953399910668e-01, -1.42855923902450388e-01,
        1.11094283515041928e-01, -9.07679894442711965e-02, 7.61425097816378765e-02, -6.36609554917869219e-02,
        5.04556591859826667e-02, -3.52678487665738852e-02, 1.99375722877763104e-02, -8.24198211827211098e-03,
        2.16282427498169566e-03, -2.66606699012429877e-04};
    double p = c[13];
#pragma unroll
    for (int i = 12; i >= 0; --i) p = fma(p, s, c[i]);
    return p;
}

// 1 / x for x in the certified range [2^-6, 2^6]: single-precision seed, two Newton steps (relative error
// 2^-23 -> 2^-46 -> double rounding).  Outside that range the piece is not certified anyway.
__device__ __forceinline__ double rcp_newton64(double x)
{
    double y = (double)__frcp_rn((float)x);
    y = fma(y, fma(-x, y, 1.0), y);
    y = fma(y, fma(-x, y, 1.0), y);
    return y;
}

// the rotated ray of output position (u, v): createMap.cl:15-30, or the fisheye-output variant
__device__ __forceinline__ Ray ray_only(const GeomD& g, const RotD& R, double u, double v)
{
    double x = (u - g.mcx) * g.inv_mfx, y = (v - g.mcy) * g.inv_mfy, z = 1.0;
    if (g.projection & 2) {  // equidistant output: |p| is the angle from the axis
        const double th2 = x * x + y * y, th = sqrt(th2);
        double sn, cs;
        sincos(th, &sn, &cs);
        const double sinc = th > 1e-4 ? sn / th : 1.0 - th2 * (1.0 / 6.0 - th2 * (1.0 / 120.0));
        x *= sinc; y *= sinc; z = cs;
    }
    Ray o;
    o.q0 = R.r[0] * x + R.r[1] * y + R.r[2] * z;
    o.q1 = R.r[3] * x + R.r[4] * y + R.r[5] * z;
    o.q2 = R.r[6] * x + R.r[7] * y + R.r[8] * z;
    o.mx = o.my = 0.0;
    return o;
}

// createMap.cl:15-49 in double precision (and its siblings for the other projection pairs)
__device__ __forceinline__ Ray project(const GeomD& g, const RotD& R, double u, double v)
{
    Ray o = ray_only(g, R, u, v);
    // one reciprocal per divisor (1e-16 relative error; the anchors need ~1e-10)
    const double iq = rcp_newton64(o.q2);
    const double c0 = o.q0 * iq, c1 = o.q1 * iq;
    double k = 1.0;  // rectilinear input: no radial step
    if (!(g.projection & 1)) {
        const double r2 = c0 * c0 + c1 * c1;
        const double ir = rsqrt(r2);
        const double r = r2 * ir;
        // theta / r: atan(r) / r for r <= 1, (pi/2 - atan(1/r)) / r beyond; the series near the axis avoids 0 * inf
        // (the reference's NaN at r == 0 is reproduced by sending the piece that contains the axis to the per-pixel path)
        if (r > 1.0) {
            const double t2 = ir * ir;
            k = (1.5707963267948966 - ir * atan_over_t(t2)) * ir;
        } else if (r > 1e-4) {
            k = atan_over_t(r2);
        } else {
            k = 1.0 - r2 * (1.0 / 3.0 - r2 * 0.2);
        }
        if (g.has_dist) {  // extension: cv::fisheye distortion, theta_d / r = (theta / r) (1 + k1 theta^2 + ... + k4 theta^8)
            const double th2 = k * k * r2;
            k *= 1.0 + th2 * (g.kd[0] + th2 * (g.kd[1] + th2 * (g.kd[2] + th2 * g.kd[3])));
        }
    }
    o.mx = g.scx + c0 * k * g.sfx;
    o.my = g.scy + c1 * k * g.sfy;
    return o;
}

}  // namespace vaw
