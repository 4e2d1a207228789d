// vaw_poly.cuh -- device pieces shared by the polynomial-coordinate kernels (vaw_poly.cu):
// collapsing a piece polynomial onto a lane's columns, per-row coordinates, the unchecked
// integer samplers (global memory and shared memory), packing and stores.
//
// Everything here implements cv::remap's INTER_LINEAR fixed-point filter as called at
// /root/reference/opencv/FrameSourceWarp.cpp:306-312 (see vaw_sample.cuh) on the map of
// vaw_pieces.cuh; the NV12 chroma rule is oracle/nv12_warp_ref.c's.
#pragma once
#include <stdint.h>
#include "vaw_coords.cuh"
#include "vaw_pieces.cuh"
#include "vaw_sample.cuh"
#include "vaw_cubic.cuh"
#include "vaw_internal.h"
#include "vaw_project64.cuh"

namespace vaw {

constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23: float add -> round-half-even integer in the mantissa
constexpr int kMagicBits = 0x4B400000;
constexpr unsigned kMagicShift = (unsigned)(kMagicBits >> 5);  // bias left in (bits >> 5)

// (x, y) pairs throughout: FFMA2 / FADD2 (sm_100) do both coordinates in one issue slot and are
// two ordinary IEEE fp32 operations, so nothing changes numerically.
struct ColPoly {
    float2 a[4][kNv];  // [column][power of t]
    float2 base;
};

__device__ __forceinline__ float2 pair(float v) { return make_float2(v, v); }

// Horner in s for one column: coefficients (x, y) of t^k, k = 0..3
__device__ __forceinline__ void collapse_column(const float2 (&c)[kNu][kNv], float s, float2 (&out)[kNv])
{
    const float2 ss = pair(s);
#pragma unroll
    for (int k = 0; k < kNv; ++k) {
        float2 acc = c[kDegU][k];
#pragma unroll
        for (int i = kDegU - 1; i >= 0; --i) acc = __ffma2_rn(acc, ss, c[i][k]);
        out[k] = acc;
    }
}

__device__ __forceinline__ void load_coeffs(const PieceRec* __restrict__ rec, float2 (&c)[kNu][kNv])
{
    const float4* r4 = reinterpret_cast<const float4*>(rec);
#pragma unroll
    for (int q = 0; q < 12; ++q) {
        const float4 v = __ldg(r4 + q);
        float2* dst = &c[0][0] + 2 * q;
        dst[0] = make_float2(v.x, v.y);
        dst[1] = make_float2(v.z, v.w);
    }
}

__device__ __forceinline__ float2 load_base(const PieceRec* __restrict__ rec)
{
    const float4 tail = __ldg(reinterpret_cast<const float4*>(rec) + 12);
    return make_float2(tail.x, tail.y);
}

// Collapse the piece polynomial onto this lane's four columns (80 FFMA2 per piece).
__device__ __forceinline__ void derive(const PieceRec* __restrict__ rec, int lane, ColPoly& cp)
{
    float2 c[kNu][kNv];
    load_coeffs(rec, c);
    cp.base = load_base(rec);
#pragma unroll
    for (int j = 0; j < 4; ++j)
        collapse_column(c, ((float)(4 * lane + j) - 63.5f) * 0.015625f, cp.a[j]);  // s is exact
}

// fp32 coordinates (x, y) of the lane's 4 columns on one row; t = (dv - t_off) * t_scale (exact).
#ifndef VAW_SCALAR_COLS
#define VAW_SCALAR_COLS 0  // analysis: evaluate the last N column slots with scalar FFMA instead of FFMA2
#endif
__device__ __forceinline__ void row_coords(const ColPoly& cp, float t, float2 (&m)[4])
{
    const float2 tt = pair(t);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j >= 4 - VAW_SCALAR_COLS) {
            float px = __fmaf_rn(cp.a[j][3].x, t, cp.a[j][2].x), py = __fmaf_rn(cp.a[j][3].y, t, cp.a[j][2].y);
            px = __fmaf_rn(px, t, cp.a[j][1].x); py = __fmaf_rn(py, t, cp.a[j][1].y);
            px = __fmaf_rn(px, t, cp.a[j][0].x); py = __fmaf_rn(py, t, cp.a[j][0].y);
            m[j] = make_float2(__fadd_rn(cp.base.x, px), __fadd_rn(cp.base.y, py));
        } else {
            float2 p = __ffma2_rn(cp.a[j][3], tt, cp.a[j][2]);
            p = __ffma2_rn(p, tt, cp.a[j][1]);
            p = __ffma2_rn(p, tt, cp.a[j][0]);
            m[j] = __fadd2_rn(cp.base, p);  // the map value: rounded once to fp32
        }
    }
}

__device__ __forceinline__ float row_t(const Geom& g, int dv) { return __fmul_rn(__fsub_rn((float)dv, g.t_off), g.t_scale); }

// Twice the NV12 chroma coordinate of a luma quad: chroma_coord() = ((sum * 0.25) - 0.5) * 0.5 where
// the two multiplications are exact, so z = fma(sum, 0.25, -0.5) carries the single rounding and
// rint(32 * coordinate) = rint(16 * z).
__device__ __forceinline__ float2 chroma_z(float2 m00, float2 m01, float2 m10, float2 m11)
{
    return __ffma2_rn(__fadd2_rn(__fadd2_rn(m00, m01), __fadd2_rn(m10, m11)), pair(0.25f), pair(-0.5f));
}
// rint(32 m) in the mantissa: fma(m, 32, 1.5 * 2^23) for both coordinates (16 for chroma z)
__device__ __forceinline__ int2 fix_bits(float2 m, float scale)
{
    const float2 s = __ffma2_rn(m, pair(scale), pair(kMagic));
    return make_int2(__float_as_int(s.x), __float_as_int(s.y));
}

// ---- the integer blend --------------------------------------------------------------------
// returns 1024 * result + fraction (+512 already added); the caller shifts by 10
__device__ __forceinline__ int blend_y(int t00, int t01, int t10, int t11, int ax, int ay)
{
    const int wx = 32 - ax, wy = 32 - ay;
    const int top = t00 * wx + t01 * ax, bot = t10 * wx + t11 * ax;
    return top * wy + bot * ay + 512;
}

// taps are (U | V << 8) pairs.  The four-weight sum is exact in integers, so the order of the two
// blends is free: vertical first on U (bits 0..15) and V (bits 16..31) at once, then the horizontal
// blend of each channel as one two-way dot product (IDP.2A) -- same value as blend_y's order.
__device__ __forceinline__ unsigned blend_uv(unsigned t00, unsigned t01, unsigned t10, unsigned t11,
                                             unsigned ax, unsigned ay)
{
    const unsigned wy = 32u - ay;
    const unsigned s00 = __byte_perm(t00, 0, 0x4140), s01 = __byte_perm(t01, 0, 0x4140);
    const unsigned s10 = __byte_perm(t10, 0, 0x4140), s11 = __byte_perm(t11, 0, 0x4140);
    // the rounding constant rides along as +16 in every 16-bit half: the horizontal weights sum to 32, so
    // the dot products below come out as sum + 512 without a register for the accumulator (<= 8176 per half)
    const unsigned left = s00 * wy + (s10 * ay + 0x00100010u), right = s01 * wy + (s11 * ay + 0x00100010u);
    const unsigned wpair = 32u + 255u * ax;                                   // (32 - ax) | ax << 8
    const unsigned u = __dp2a_lo(__byte_perm(left, right, 0x5410), wpair, 0u) >> 10;
    const unsigned v = __dp2a_lo(__byte_perm(left, right, 0x7632), wpair, 0u) >> 2;
    return u | (v & 0xff00u);
}

// ---- samplers without border tests, taps from global memory ---------------------------------
// rint(32 m) sits in the mantissa of fma(m, 32, 1.5*2^23); >> 5 keeps a constant bias that is
// folded into `bias` (all offset arithmetic is modulo 2^32 and the true offset fits).
__device__ __forceinline__ int luma_gmem(const uint8_t* __restrict__ plane, unsigned pitch, unsigned bias, float2 m)
{
    const int2 bb = fix_bits(m, 32.0f);
    const int bx = bb.x, by = bb.y;
    const unsigned off = (unsigned)(by >> 5) * pitch + ((unsigned)(bx >> 5) + bias);
    const uint8_t* p = plane + off;
    const uint8_t* q = p + pitch;
    return blend_y(__ldg(p), __ldg(p + 1), __ldg(q), __ldg(q + 1), bx & 31, by & 31);
}

__device__ __forceinline__ unsigned chroma_gmem(const uint8_t* __restrict__ plane, unsigned pitch, unsigned bias,
                                                float2 z)
{
    const int2 bb = fix_bits(z, 16.0f);
    const int bx = bb.x, by = bb.y;
    const unsigned off = (unsigned)(by >> 5) * pitch + (((unsigned)(bx >> 5) + bias) << 1);
    const uint8_t* p = plane + off;
    const uint8_t* q = p + pitch;
    return blend_uv(__ldg(reinterpret_cast<const uint16_t*>(p)), __ldg(reinterpret_cast<const uint16_t*>(p + 2)),
                    __ldg(reinterpret_cast<const uint16_t*>(q)), __ldg(reinterpret_cast<const uint16_t*>(q + 2)),
                    bx & 31, by & 31);
}

// ---- the same, taps from a staged tile in shared memory (row pitch PL bytes) ---------------
template <int IMM>
__device__ __forceinline__ unsigned lds_u8(unsigned addr)
{
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(IMM));
    return v;
}
template <int IMM>
__device__ __forceinline__ unsigned lds_u16(unsigned addr)
{
    unsigned v;
    asm volatile("ld.shared.u16 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(IMM));
    return v;
}

// ---- packing and stores ------------------------------------------------------------------------
__device__ __forceinline__ unsigned pack4(int a0, int a1, int a2, int a3)
{
    // each a_i = 1024 * result + fraction, result <= 255
    const unsigned lo = __byte_perm((unsigned)a0 >> 10, (unsigned)a1 >> 10, 0x0040);
    const unsigned hi = __byte_perm((unsigned)a2 >> 10, (unsigned)a3 >> 10, 0x0040);
    return __byte_perm(lo, hi, 0x5410);
}

template <bool kRagged>
__device__ __forceinline__ void store_word(uint8_t* p, unsigned word, int valid)
{
    if (!kRagged || (valid >= 4 && (reinterpret_cast<uintptr_t>(p) & 3) == 0)) {
        *reinterpret_cast<unsigned*>(p) = word;
    } else {
        for (int i = 0; i < valid && i < 4; ++i) p[i] = (uint8_t)(word >> (8 * i));
    }
}

struct PlaneRefs {
    const uint8_t* y;   // luma plane of this frame
    const uint8_t* uv;  // chroma plane of this frame
    uint8_t* dst;       // output frame
};

// Checked sampling of one row pair from given coordinates (mixed pieces, per-pixel fallback).  kMode 1: cv::INTER_NEAREST
// = the same integer filter on coordinates rounded to whole pixels (luma: the map values; chroma: the chroma
// coordinate derived from the UNROUNDED luma map); kMode 2: the context's INTER_CUBIC / INTER_LANCZOS4 table filter.
enum { kModeLinear = 0, kModeNearest = 1, kModeTable = 2 };
template <int kMode = kModeLinear>
__device__ __forceinline__ void sample_rows_checked(const Geom& g, const PlaneRefs& f, int u0, int v0,
                                                    const float2 (&m)[2][4])
{
    constexpr bool kNearest = kMode == kModeNearest;
    const int border_y = g.border & 255;
    const unsigned border_uv = (g.border >> 8) & 0xffffu;
    const int valid = g.out_w - u0;
    unsigned yw[2] = {0u, 0u};
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (kMode == kModeTable)
                yw[r] |= sample_hi<1>(f.y, g.src_pitch, g.src_w, g.src_h, m[r][i].x, m[r][i].y, (unsigned)border_y, g.cubic_tab, g.tab_ks) << (8 * i);
            else
                yw[r] |= (unsigned)sample_c1(f.y, g.src_pitch, g.src_w, g.src_h, kNearest ? nearest_coord(m[r][i].x) : m[r][i].x,
                                             kNearest ? nearest_coord(m[r][i].y) : m[r][i].y, border_y) << (8 * i);
        }
    unsigned cw = 0u;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const float cx = chroma_coord(m[0][2 * q].x, m[0][2 * q + 1].x, m[1][2 * q].x, m[1][2 * q + 1].x);
        const float cy = chroma_coord(m[0][2 * q].y, m[0][2 * q + 1].y, m[1][2 * q].y, m[1][2 * q + 1].y);
        if (kMode == kModeTable)
            cw |= sample_hi<2>(f.uv, g.src_pitch, g.src_w >> 1, g.src_h >> 1, cx, cy, border_uv, g.cubic_tab, g.tab_ks) << (16 * q);
        else
            cw |= sample_c2(f.uv, g.src_pitch, g.src_w >> 1, g.src_h >> 1, kNearest ? nearest_coord(cx) : cx,
                            kNearest ? nearest_coord(cy) : cy, border_uv) << (16 * q);
    }
    if (valid > 0) {
        store_word<true>(f.dst + (size_t)v0 * g.dst_pitch + u0, yw[0], valid);
        store_word<true>(f.dst + (size_t)(v0 + 1) * g.dst_pitch + u0, yw[1], valid);
        store_word<true>(f.dst + (size_t)(g.out_h + (v0 >> 1)) * g.dst_pitch + u0, cw, valid);
    }
}

// Per-pixel op-for-op coordinates of a row pair (pieces without a polynomial certificate).
__device__ __forceinline__ void exact_rows(const Geom& g, const Rot& R, int u_lo, int u0, int v0, float2 (&m)[2][4])
{
    if (g.projection != 0) {
        // the projection pairs createMap.cl does not have: no fp32 operation order to follow, so the few
        // uncertified pieces get the double-precision projection, rounded once
        GeomD d;
        d.scx = g.scx; d.scy = g.scy; d.sfx = g.sfx; d.sfy = g.sfy;
        d.mcx = g.mcx; d.mcy = g.mcy; d.mfx = g.mfx; d.mfy = g.mfy;
        d.inv_mfx = 1.0 / d.mfx; d.inv_mfy = 1.0 / d.mfy;
#pragma unroll
        for (int i = 0; i < 4; ++i) d.kd[i] = g.kd[i];
        d.has_dist = g.has_dist;
        d.projection = g.projection;
        RotD Rd;
#pragma unroll
        for (int i = 0; i < 9; ++i) Rd.r[i] = (double)R.r[i];
        for (int r = 0; r < 2; ++r)
            for (int i = 0; i < 4; ++i) {
                const Ray p = project(d, Rd, (double)(u0 + i), (double)(v0 + r));
                m[r][i] = make_float2((float)p.mx, (float)p.my);
            }
        return;
    }
    const float4 xs = __ldg(reinterpret_cast<const float4*>(g.xtab + u0));
    const float2 ys = __ldg(reinterpret_cast<const float2*>(g.ytab + v0));
    const ColTerms c[4] = {col_terms(xs.x, R), col_terms(xs.y, R), col_terms(xs.z, R), col_terms(xs.w, R)};
    const RowTerms w[2] = {row_terms(ys.x, R), row_terms(ys.y, R)};
    const int u_hi = min(u_lo + kPieceW, g.out_w) - 1;
    const int v_hi = min(v0 + 1, g.out_h - 1);
    if (fast_path_ok(u_lo, u_hi, v0, v_hi, R, g)) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) map_eval<true>(c[i], w[r], R, g, m[r][i].x, m[r][i].y);
    } else {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) map_eval<false>(c[i], w[r], R, g, m[r][i].x, m[r][i].y);
    }
}

// ---- per-warp row walk with taps from global memory ---------------------------------------------
__device__ __forceinline__ Rot load_rot(const FrameBatch& b, int frame)
{
    if (b.rots == nullptr) return b.rot0;
    Rot R;
    const float* p = b.rots + (size_t)frame * 9;
#pragma unroll
    for (int i = 0; i < 9; ++i) R.r[i] = __ldg(p + i);
    return R;
}

struct RowPtrs {  // output pointers of the lane's 4-pixel group for the current row pair
    uint8_t *y0, *y1, *c;
    size_t step_y, step_c;  // advance per row pair
};

// nrows (even) rows starting at piece row dv0, taps from global memory, no border tests.
template <bool kRagged>
__device__ __forceinline__ void band_gmem(const Geom& g, const ColPoly& cp, const PlaneRefs& f, int dv0, int nrows,
                                          RowPtrs& o, int valid)
{
    const unsigned pitch = (unsigned)g.src_pitch;
    const unsigned bias_y = 0u - kMagicShift * pitch - kMagicShift;         // offset = iy*pitch + ix
    const unsigned bias_c = 0u - kMagicShift * (pitch >> 1) - kMagicShift;  // offset = iy*pitch + 2*(ix + bias)
    // t = (dv - t_off) * t_scale is a small dyadic rational: stepping it by t_scale is exact
    float t = row_t(g, dv0);
    const float dt = g.t_scale, dt2 = __fadd_rn(g.t_scale, g.t_scale);
#pragma unroll 1
    for (int dv = dv0; dv < dv0 + nrows; dv += 2) {
        float2 m[2][4];
        row_coords(cp, t, m[0]);
        row_coords(cp, __fadd_rn(t, dt), m[1]);
        t = __fadd_rn(t, dt2);
        int acc[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[r][i] = luma_gmem(f.y, pitch, bias_y, m[r][i]);
        unsigned cw = 0u;
#pragma unroll
        for (int q = 0; q < 2; ++q)
            cw |= chroma_gmem(f.uv, pitch, bias_c, chroma_z(m[0][2 * q], m[0][2 * q + 1], m[1][2 * q], m[1][2 * q + 1]))
                  << (16 * q);
        if (!kRagged || valid > 0) {
            store_word<kRagged>(o.y0, pack4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]), valid);
            store_word<kRagged>(o.y1, pack4(acc[1][0], acc[1][1], acc[1][2], acc[1][3]), valid);
            store_word<kRagged>(o.c, cw, valid);
        }
        o.y0 += o.step_y; o.y1 += o.step_y; o.c += o.step_c;
    }
}


// ---- mbarrier / TMA primitives -----------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
#ifndef VAW_ABL_NO_INIT_FENCE  // analysis only (timing of the fence; the init must be fenced before the async proxy uses the barrier)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    } while (!ok);
}
// The same with a suspend-time hint (nanoseconds): the waiting warp is parked by the hardware for up to that
// long per try instead of spinning through the issue slots of the warps that work.
__device__ __forceinline__ void mbar_wait_parked(unsigned mbar, unsigned parity, unsigned hint_ns)
{
    unsigned ok;
    do {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok) : "r"(mbar), "r"(parity), "r"(hint_ns) : "memory");
    } while (!ok);
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion on mbar
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

}  // namespace vaw
