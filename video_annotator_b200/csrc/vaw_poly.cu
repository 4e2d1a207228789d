// vaw_poly.cu -- fused map + remap for NV12, coordinates from the per-piece polynomials
// (variant POLY: taps gathered through L1/L2).
//
// Replaces FrameSourceWarp::warp_frame's two passes
// (/root/reference/opencv/FrameSourceWarp.cpp:272-314) like vaw_kernels.cu does, but is
// built around the instruction budget of a B200 (vaw_pieces.cuh): a warp owns one
// 128x32-pixel piece; lane l owns the four columns 4l..4l+3 and walks down the 32 rows two
// at a time (two luma rows + one chroma row per step, 4-byte stores -> every store
// instruction of the warp writes one full 128-byte line).  Before the walk each lane
// collapses the piece's tensor polynomial onto its four columns (160 FMAs, amortised over
// 128 pixels), so a coordinate costs 3 FMAs + 1 add per pixel; the 1/32-px fixed-point
// conversion is a single FMA with the 1.5*2^23 constant.  Pieces the builder classified
// as interior take a sampler without border tests; pure-border pieces are just filled;
// the rest use the checked sampler; pieces without a polynomial certificate fall back to
// the per-pixel op-for-op evaluation (vaw_coords.cuh).
#include <stdint.h>
#include "vaw_internal.h"
#include "vaw_pieces.cuh"
#include "vaw_sample.cuh"

namespace vaw {

namespace {

constexpr int kWarps = 4;            // warps per CTA, stacked vertically (4 pieces)
constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23: float add -> round-half-even integer in the mantissa
constexpr int kMagicBits = 0x4B400000;

struct ColPoly {
    float a[2][4][kNv];  // [coordinate][column][power of t]
    float bx, by;
};

__device__ __forceinline__ Rot load_rot(const FrameBatch& b, int frame)
{
    if (b.rots == nullptr) return b.rot0;
    Rot R;
    const float* p = b.rots + (size_t)frame * 9;
#pragma unroll
    for (int i = 0; i < 9; ++i) R.r[i] = __ldg(p + i);
    return R;
}

// Collapse the piece polynomial onto this lane's four columns.
__device__ __forceinline__ void derive(const PieceRec* __restrict__ rec, int lane, ColPoly& cp)
{
    const float4* r4 = reinterpret_cast<const float4*>(rec);
    float c[2][kNu][kNv];
#pragma unroll
    for (int q = 0; q < 12; ++q) {
        const float4 v = __ldg(r4 + q);
        float* dst = &c[0][0][0] + 4 * q;
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
    const float4 tail = __ldg(r4 + 12);
    cp.bx = tail.x;
    cp.by = tail.y;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float s = ((float)(4 * lane + j) - 63.5f) * 0.015625f;  // exact
#pragma unroll
        for (int co = 0; co < 2; ++co)
#pragma unroll
            for (int k = 0; k < kNv; ++k) {
                float acc = c[co][kDegU][k];
#pragma unroll
                for (int i = kDegU - 1; i >= 0; --i) acc = __fmaf_rn(acc, s, c[co][i][k]);
                cp.a[co][j][k] = acc;
            }
    }
}

// fp32 coordinates of the lane's 4 columns on one row (dv = row inside the piece).
__device__ __forceinline__ void row_coords(const ColPoly& cp, int dv, float (&mx)[4], float (&my)[4])
{
    const float t = ((float)dv - 15.5f) * 0.0625f;  // exact
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float px = __fmaf_rn(cp.a[0][j][3], t, cp.a[0][j][2]);
        px = __fmaf_rn(px, t, cp.a[0][j][1]);
        px = __fmaf_rn(px, t, cp.a[0][j][0]);
        mx[j] = __fadd_rn(cp.bx, px);  // the map value: rounded once to fp32
        float py = __fmaf_rn(cp.a[1][j][3], t, cp.a[1][j][2]);
        py = __fmaf_rn(py, t, cp.a[1][j][1]);
        py = __fmaf_rn(py, t, cp.a[1][j][0]);
        my[j] = __fadd_rn(cp.by, py);
    }
}

// ---- samplers without border tests (piece certified interior) --------------------------
// rint(32 m) sits in the mantissa of fma(m, 32, 1.5*2^23); >> 5 keeps a constant bias that
// is folded into `bias` (all offset arithmetic is modulo 2^32 and the true offset fits).
__device__ __forceinline__ int luma_fast(const uint8_t* __restrict__ plane, unsigned pitch, unsigned bias,
                                         float mx, float my)
{
    const int bx = __float_as_int(__fmaf_rn(mx, 32.0f, kMagic));
    const int by = __float_as_int(__fmaf_rn(my, 32.0f, kMagic));
    const int ax = bx & 31, ay = by & 31;
    const unsigned off = (unsigned)(by >> 5) * pitch + ((unsigned)(bx >> 5) + bias);
    const uint8_t* p = plane + off;
    const uint8_t* q = p + pitch;
    const int t00 = __ldg(p), t01 = __ldg(p + 1), t10 = __ldg(q), t11 = __ldg(q + 1);
    const int wx = 32 - ax, wy = 32 - ay;
    const int top = t00 * wx + t01 * ax, bot = t10 * wx + t11 * ax;
    return top * wy + bot * ay + 512;  // caller shifts by 10
}

// chroma: plane of (U,V) byte pairs; returns U | V << 8
// zx, zy = 2 * chroma coordinate (see chroma_z): rint(32 c) = rint(16 z), exact scalings.
__device__ __forceinline__ unsigned chroma_fast(const uint8_t* __restrict__ plane, unsigned pitch, unsigned bias,
                                                float zx, float zy)
{
    const int bx = __float_as_int(__fmaf_rn(zx, 16.0f, kMagic));
    const int by = __float_as_int(__fmaf_rn(zy, 16.0f, kMagic));
    const unsigned ax = bx & 31, ay = by & 31;
    const unsigned off = (unsigned)(by >> 5) * pitch + (((unsigned)(bx >> 5) + bias) << 1);
    const uint8_t* p = plane + off;
    const uint8_t* q = p + pitch;
    const unsigned t00 = __ldg(reinterpret_cast<const uint16_t*>(p));
    const unsigned t01 = __ldg(reinterpret_cast<const uint16_t*>(p + 2));
    const unsigned t10 = __ldg(reinterpret_cast<const uint16_t*>(q));
    const unsigned t11 = __ldg(reinterpret_cast<const uint16_t*>(q + 2));
    // U in bits 0..15, V in bits 16..31: the horizontal blend runs on both at once
    const unsigned wx = 32u - ax, wy = 32u - ay;
    const unsigned s00 = __byte_perm(t00, 0, 0x4140), s01 = __byte_perm(t01, 0, 0x4140);
    const unsigned s10 = __byte_perm(t10, 0, 0x4140), s11 = __byte_perm(t11, 0, 0x4140);
    const unsigned top = s00 * wx + s01 * ax, bot = s10 * wx + s11 * ax;  // <= 8160 per half
    const unsigned u = ((top & 0xffffu) * wy + (bot & 0xffffu) * ay + 512u) >> 10;
    const unsigned v = ((top >> 16) * wy + (bot >> 16) * ay + 512u) >> 10;
    return u | (v << 8);
}

// Twice the NV12 chroma coordinate of a luma quad: chroma_coord() = ((sum * 0.25) - 0.5) * 0.5 where
// the two multiplications are exact, so z = fma(sum, 0.25, -0.5) carries the single rounding.
__device__ __forceinline__ float chroma_z(float m00, float m01, float m10, float m11)
{
    return __fmaf_rn(__fadd_rn(__fadd_rn(m00, m01), __fadd_rn(m10, m11)), 0.25f, -0.5f);
}

__device__ __forceinline__ void store4(uint8_t* p, unsigned word, int valid)
{
    if (valid >= 4) {
        *reinterpret_cast<unsigned*>(p) = word;
    } else {
        for (int i = 0; i < valid; ++i) p[i] = (uint8_t)(word >> (8 * i));
    }
}

__device__ __forceinline__ unsigned pack4(int a0, int a1, int a2, int a3)
{
    // each a_i = 1024 * result + fraction, result <= 255
    const unsigned lo = __byte_perm((unsigned)a0 >> 10, (unsigned)a1 >> 10, 0x0040);
    const unsigned hi = __byte_perm((unsigned)a2 >> 10, (unsigned)a3 >> 10, 0x0040);
    return __byte_perm(lo, hi, 0x5410);
}

struct PlaneRefs {
    const uint8_t* y;    // luma plane of this frame
    const uint8_t* uv;   // chroma plane of this frame
    uint8_t* dst;        // output frame
};

// Checked sampling of one row pair from given coordinates (mixed pieces, exact fallback).
__device__ __forceinline__ void sample_rows_checked(const Geom& g, const PlaneRefs& f, int u0, int v0,
                                                    const float (&mx)[2][4], const float (&my)[2][4])
{
    const int border_y = g.border & 255;
    const unsigned border_uv = (g.border >> 8) & 0xffffu;
    const int valid = g.out_w - u0;
    unsigned yw[2] = {0u, 0u};
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i)
            yw[r] |= (unsigned)sample_c1(f.y, g.src_pitch, g.src_w, g.src_h, mx[r][i], my[r][i], border_y) << (8 * i);
    unsigned cw = 0u;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const float cx = chroma_coord(mx[0][2 * q], mx[0][2 * q + 1], mx[1][2 * q], mx[1][2 * q + 1]);
        const float cy = chroma_coord(my[0][2 * q], my[0][2 * q + 1], my[1][2 * q], my[1][2 * q + 1]);
        cw |= sample_c2(f.uv, g.src_pitch, g.src_w >> 1, g.src_h >> 1, cx, cy, border_uv) << (16 * q);
    }
    if (valid > 0) {
        store4(f.dst + (size_t)v0 * g.dst_pitch + u0, yw[0], valid);
        store4(f.dst + (size_t)(v0 + 1) * g.dst_pitch + u0, yw[1], valid);
        store4(f.dst + (size_t)(g.out_h + (v0 >> 1)) * g.dst_pitch + u0, cw, valid);
    }
}

// Per-pixel op-for-op coordinates of a row pair (pieces without a polynomial certificate).
__device__ __forceinline__ void exact_rows(const Geom& g, const Rot& R, int u_lo, int u0, int v0,
                                           float (&mx)[2][4], float (&my)[2][4])
{
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
            for (int i = 0; i < 4; ++i) map_eval<true>(c[i], w[r], R, g, mx[r][i], my[r][i]);
    } else {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) map_eval<false>(c[i], w[r], R, g, mx[r][i], my[r][i]);
    }
}

__global__ void __launch_bounds__(32 * kWarps)
warp_nv12_poly_kernel(const Geom g, const FrameBatch b, const PieceRec* __restrict__ table)
{
    const int lane = threadIdx.x;
    const int px = blockIdx.x, py = blockIdx.y * kWarps + threadIdx.y, frame = blockIdx.z;
    const int npx = pieces_x(g.out_w), npy = pieces_y(g.out_h);
    if (py >= npy) return;  // warp-uniform
    const PieceRec* rec = table + ((size_t)frame * npy + py) * npx + px;
    const unsigned flags = __ldg(&rec->flags);
    const int u_lo = px * kPieceW, u0 = u_lo + 4 * lane, v_base = py * kPieceH;
    const int rows = min(kPieceH, g.out_h - v_base);  // even for NV12
    const int valid = g.out_w - u0;

    PlaneRefs f;
    f.y = b.src + (size_t)frame * b.src_frame_stride;
    f.uv = f.y + (size_t)g.src_pitch * g.src_h;
    f.dst = b.dst + (size_t)frame * b.dst_frame_stride;

    if (flags & kPieceOutside) {  // pure border: nothing to compute
        const unsigned yw = (g.border & 255u) * 0x01010101u;
        const unsigned cw = ((g.border >> 8) & 0xffffu) * 0x00010001u;
        if (valid > 0)
            for (int dv = 0; dv < rows; dv += 2) {
                const int v0 = v_base + dv;
                store4(f.dst + (size_t)v0 * g.dst_pitch + u0, yw, valid);
                store4(f.dst + (size_t)(v0 + 1) * g.dst_pitch + u0, yw, valid);
                store4(f.dst + (size_t)(g.out_h + (v0 >> 1)) * g.dst_pitch + u0, cw, valid);
            }
        return;
    }

    if (!(flags & kPiecePoly)) {  // op-for-op per pixel
        const Rot R = load_rot(b, frame);
        for (int dv = 0; dv < rows; dv += 2) {
            float mx[2][4], my[2][4];
            exact_rows(g, R, u_lo, u0, v_base + dv, mx, my);
            sample_rows_checked(g, f, u0, v_base + dv, mx, my);
        }
        return;
    }

    ColPoly cp;
    derive(rec, lane, cp);

    if (!(flags & kPieceInterior)) {  // polynomial coordinates, checked sampler
        for (int dv = 0; dv < rows; dv += 2) {
            float mx[2][4], my[2][4];
            row_coords(cp, dv, mx[0], my[0]);
            row_coords(cp, dv + 1, mx[1], my[1]);
            sample_rows_checked(g, f, u0, v_base + dv, mx, my);
        }
        return;
    }

    // interior: no border tests
    const unsigned pitch = (unsigned)g.src_pitch;
    const unsigned kb = (unsigned)(kMagicBits >> 5);
    const unsigned bias_y = 0u - kb * pitch - kb;        // luma: offset = iy*pitch + ix
    const unsigned bias_c = 0u - kb * (pitch >> 1) - kb;  // chroma: offset = iy*pitch + 2*(ix + bias), pitch even
#pragma unroll 1
    for (int dv = 0; dv < rows; dv += 2) {
        float mx[2][4], my[2][4];
        row_coords(cp, dv, mx[0], my[0]);
        row_coords(cp, dv + 1, mx[1], my[1]);
        int acc[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[r][i] = luma_fast(f.y, pitch, bias_y, mx[r][i], my[r][i]);
        unsigned cw = 0u;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float zx = chroma_z(mx[0][2 * q], mx[0][2 * q + 1], mx[1][2 * q], mx[1][2 * q + 1]);
            const float zy = chroma_z(my[0][2 * q], my[0][2 * q + 1], my[1][2 * q], my[1][2 * q + 1]);
            cw |= chroma_fast(f.uv, pitch, bias_c, zx, zy) << (16 * q);
        }
        const int v0 = v_base + dv;
        if (valid > 0) {
            store4(f.dst + (size_t)v0 * g.dst_pitch + u0, pack4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]), valid);
            store4(f.dst + (size_t)(v0 + 1) * g.dst_pitch + u0, pack4(acc[1][0], acc[1][1], acc[1][2], acc[1][3]), valid);
            store4(f.dst + (size_t)(g.out_h + (v0 >> 1)) * g.dst_pitch + u0, cw, valid);
        }
    }
}

// The map the kernel above samples with: plane 0 luma, plane 1 chroma.
__global__ void __launch_bounds__(32 * kWarps)
dump_coords_poly_kernel(const Geom g, const Rot R, const PieceRec* __restrict__ table, int plane,
                        float* __restrict__ map_x, float* __restrict__ map_y, int map_pitch)
{
    const int lane = threadIdx.x;
    const int px = blockIdx.x, py = blockIdx.y * kWarps + threadIdx.y;
    const int npx = pieces_x(g.out_w), npy = pieces_y(g.out_h);
    if (py >= npy) return;
    const PieceRec* rec = table + (size_t)py * npx + px;
    const unsigned flags = __ldg(&rec->flags);
    const int u_lo = px * kPieceW, u0 = u_lo + 4 * lane, v_base = py * kPieceH;
    const int rows = min(kPieceH, g.out_h - v_base);
    ColPoly cp;
    if (flags & kPiecePoly) derive(rec, lane, cp);
    for (int dv = 0; dv < rows; dv += 2) {
        float mx[2][4], my[2][4];
        const int v0 = v_base + dv;
        if (flags & kPiecePoly) {
            row_coords(cp, dv, mx[0], my[0]);
            row_coords(cp, dv + 1, mx[1], my[1]);
        } else {
            exact_rows(g, R, u_lo, u0, v0, mx, my);
        }
        if (plane == 0) {
            for (int r = 0; r < 2; ++r)
                for (int i = 0; i < 4; ++i)
                    if (u0 + i < g.out_w && v0 + r < g.out_h) {
                        map_x[(size_t)(v0 + r) * map_pitch + u0 + i] = mx[r][i];
                        map_y[(size_t)(v0 + r) * map_pitch + u0 + i] = my[r][i];
                    }
        } else {
            for (int q = 0; q < 2; ++q)
                if (u0 + 2 * q + 1 < g.out_w && v0 + 1 < g.out_h) {
                    const size_t o = (size_t)(v0 >> 1) * map_pitch + (u0 >> 1) + q;
                    map_x[o] = chroma_coord(mx[0][2 * q], mx[0][2 * q + 1], mx[1][2 * q], mx[1][2 * q + 1]);
                    map_y[o] = chroma_coord(my[0][2 * q], my[0][2 * q + 1], my[1][2 * q], my[1][2 * q + 1]);
                }
        }
    }
}

}  // namespace

cudaError_t launch_warp_nv12_poly(const Geom& g, const FrameBatch& b, const PieceRec* table, cudaStream_t st)
{
    dim3 block(32, kWarps);
    dim3 grid(pieces_x(g.out_w), (pieces_y(g.out_h) + kWarps - 1) / kWarps, b.n_frames);
    warp_nv12_poly_kernel<<<grid, block, 0, st>>>(g, b, table);
    return cudaGetLastError();
}

cudaError_t launch_dump_coords_poly(const Geom& g, const Rot& rot, const PieceRec* table, int plane,
                                    float* map_x, float* map_y, int map_pitch, cudaStream_t st)
{
    dim3 block(32, kWarps);
    dim3 grid(pieces_x(g.out_w), (pieces_y(g.out_h) + kWarps - 1) / kWarps, 1);
    dump_coords_poly_kernel<<<grid, block, 0, st>>>(g, rot, table, plane, map_x, map_y, map_pitch);
    return cudaGetLastError();
}

}  // namespace vaw
