// vaw_tex.cu -- fused map + remap for NV12 with the bilinear blend done by the texture units
// (variant TEX).
//
// Replaces FrameSourceWarp::warp_frame's two passes
// (/root/reference/opencv/FrameSourceWarp.cpp:272-314).  Coordinates are the per-piece
// polynomials of vaw_pieces.cuh, exactly as in vaw_tile.cu.  What changes is who filters: the
// integer sampler of vaw_tile.cuh spends ~24 issue slots per luma pixel on addresses, four
// shared-memory loads and the 10-bit blend, and that -- not HBM -- bounds variant TILED.  The
// texture unit does address, taps and blend for one instruction per pixel.
//
// Why the result still follows cv::remap's fixed-point filter (INTER_LINEAR, 1/32-px coordinates,
// 10-bit weights, (sum + 512) >> 10): the coordinate handed to the unit is already rounded to
// 1/32 px the way cv::remap rounds it (round-half-even of 32 m, in the mantissa of an fp32 add).
// The unit quantises the fractional position to 1/256, which represents k/32 exactly, so its
// weights are cv::remap's (32-a)/32, a/32 and the filtered value is S/1024/255 with
// S = the exact 10-bit weighted sum.  Scaling by 255 and rounding half up gives (S + 512) >> 10.
// Only certified interior pieces come here (every tap of every pixel inside the source), so no
// border handling is involved; every other piece stays with vaw_tile.cu (FrameBatch::skip_interior).
#include <cuda_runtime.h>
#include <stdint.h>
#include "vaw_internal.h"
#include "vaw_poly.cuh"
#include "vaw_tile.cuh"

namespace vaw {

namespace {

constexpr int kWarps = 4;
#ifndef VAW_TEX_CTAS
#define VAW_TEX_CTAS 8  // resident CTAs per SM the kernel is sized for (registers)
#endif

// cv::remap's rounding of the coordinate to 1/32 px, then the texture coordinate of that position:
// rint(scale * m) / 32 + off (texel centres sit at integer + 0.5; off.y also carries the frame's row
// offset).  s = magic + rint(scale * m) exactly, so s / 32 + (off - magic / 32) is the wanted value
// with a single (exact) rounding: it has 5 fractional bits below 2^17.  `offm` = off - magic / 32.
__device__ __forceinline__ float2 tex_coord(float2 m, float scale, float2 offm)
{
    const float2 s = __ffma2_rn(m, pair(scale), pair(kMagic));   // round-half-even integer in the mantissa
    return __ffma2_rn(s, pair(0.03125f), offm);
}

// filtered values v = S / 1024 / 255 -> (S + 512) >> 10 in the low byte of each result:
// 255 v + 2^-11 is never a tie (S / 1024 lies on a 2^-10 grid), so round-to-nearest of it is round-half-up of S / 1024
__device__ __forceinline__ uint2 to_u8x2(float a, float b)
{
    const float2 y = __ffma2_rn(make_float2(a, b), pair(255.0f), pair(0.00048828125f));
    const float2 z = __fadd2_rn(y, pair(kMagic));
    return make_uint2(__float_as_uint(z.x), __float_as_uint(z.y));
}


// nrows (even) rows starting at piece row dv0; o.y0 / o.y1 / o.c point at column 2*lane of the piece.
template <bool kRagged>
__device__ __forceinline__ void rows_tex(const Geom& g, const ColPoly& cp, cudaTextureObject_t tex_y,
                                         cudaTextureObject_t tex_c, float2 off_y, float2 off_c, int dv0, int nrows,
                                         RowPtrs& o, bool in_a, bool in_b)
{
    // t = (dv - t_off) * t_scale is a small dyadic rational: stepping it by t_scale is exact
    float t = row_t(g, dv0);
    const float dt = g.t_scale, dt2 = __fadd_rn(g.t_scale, g.t_scale);
#pragma unroll 1
    for (int dv = dv0; dv < dv0 + nrows; dv += 2) {
        float2 m[2][4];
        row_coords(cp, t, m[0]);
        row_coords(cp, __fadd_rn(t, dt), m[1]);
        t = __fadd_rn(t, dt2);
        float v[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 c = tex_coord(m[r][i], 32.0f, off_y);
                v[r][i] = tex2D<float>(tex_y, c.x, c.y);
            }
        float2 cv[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float2 z = chroma_z(m[0][2 * q], m[0][2 * q + 1], m[1][2 * q], m[1][2 * q + 1]);
            const float2 c = tex_coord(z, 16.0f, off_c);
            cv[q] = tex2D<float2>(tex_c, c.x, c.y);
        }
        uint2 y[2][2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            y[r][0] = to_u8x2(v[r][0], v[r][1]);
            y[r][1] = to_u8x2(v[r][2], v[r][3]);
        }
        const uint2 c0 = to_u8x2(cv[0].x, cv[0].y), c1 = to_u8x2(cv[1].x, cv[1].y);
        store_pair<kRagged>(o.y0, y[0][0].x & 255u, y[0][0].y, in_a);
        store_pair<kRagged>(o.y0 + 64, y[0][1].x & 255u, y[0][1].y, in_b);
        store_pair<kRagged>(o.y1, y[1][0].x & 255u, y[1][0].y, in_a);
        store_pair<kRagged>(o.y1 + 64, y[1][1].x & 255u, y[1][1].y, in_b);
        store_pair<kRagged>(o.c, c0.x & 255u, c0.y, in_a);
        store_pair<kRagged>(o.c + 64, c1.x & 255u, c1.y, in_b);
        o.y0 += o.step_y; o.y1 += o.step_y; o.c += o.step_c;
    }
}

}  // namespace

__global__ void __launch_bounds__(32 * kWarps, VAW_TEX_CTAS)
warp_nv12_tex_kernel(const Geom g, const FrameBatch b, const PieceRec* __restrict__ table,
                     const __grid_constant__ TexSet ts)
{
    __shared__ float4 coefs[8 * 32];
    const int lane = threadIdx.x, w = threadIdx.y;
    const int px = blockIdx.x, py = blockIdx.y, frame = blockIdx.z;
    const int ph = g.piece_h, rows_per_warp = ph / kWarps;
    const int npx = (int)gridDim.x, npy = (int)gridDim.y;  // = pieces_x(out_w), pieces_y(out_h, ph): the launch grid, no division
    const PieceRec* rec = table + ((size_t)frame * npy + py) * npx + px;
    const unsigned flags = __ldg(&rec->flags);
    if ((flags & (kPiecePoly | kPieceInterior)) != (kPiecePoly | kPieceInterior)) return;  // vaw_tile.cu's

    // ---- collapse the polynomial: warp w does column slot j = w for every lane ------------------
    {
        float2 c[kNu][kNv], a[kNv];
        load_coeffs(rec, c);
        collapse_column(c, ((float)pair_column(lane, w) - 63.5f) * 0.015625f, a);  // s is exact
        coefs[(2 * w) * 32 + lane] = make_float4(a[0].x, a[0].y, a[1].x, a[1].y);
        coefs[(2 * w + 1) * 32 + lane] = make_float4(a[2].x, a[2].y, a[3].x, a[3].y);
    }
    __syncthreads();
    ColPoly cp;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 lo = coefs[(2 * j) * 32 + lane], hi = coefs[(2 * j + 1) * 32 + lane];
        cp.a[j][0] = make_float2(lo.x, lo.y); cp.a[j][1] = make_float2(lo.z, lo.w);
        cp.a[j][2] = make_float2(hi.x, hi.y); cp.a[j][3] = make_float2(hi.z, hi.w);
    }
    cp.base = load_base(rec);

    const int u_lo = px * kPieceW, v_base = py * ph;
    const int rows = min(ph, g.out_h - v_base);
    const int dv0 = w * rows_per_warp;
    const int my_rows = max(0, min(rows_per_warp, rows - dv0));
    if (my_rows <= 0) return;

    const int group = frame / ts.group_frames, fin = frame - group * ts.group_frames;
    const cudaTextureObject_t tex_y = ts.y[group], tex_c = ts.uv[group];
    const float row0 = (float)(fin * ts.frame_rows);                    // exact: < 65000
    constexpr float kM32 = kMagic * 0.03125f;                           // 393216
    const float2 off_y = make_float2(0.5f - kM32, row0 + 0.5f - kM32);  // exact: |.| < 2^19, one fractional bit
    const float2 off_c = make_float2(0.5f - kM32, row0 + (float)g.src_h + 0.5f - kM32);

    uint8_t* dst = b.dst + (size_t)frame * b.dst_frame_stride;
    RowPtrs o;
    o.y0 = dst + (size_t)(v_base + dv0) * g.dst_pitch + u_lo + 2 * lane;
    o.y1 = o.y0 + g.dst_pitch;
    o.c = dst + (size_t)(g.out_h + ((v_base + dv0) >> 1)) * g.dst_pitch + u_lo + 2 * lane;
    o.step_y = 2 * (size_t)g.dst_pitch;
    o.step_c = (size_t)g.dst_pitch;
    const bool in_a = u_lo + 2 * lane < g.out_w, in_b = u_lo + 64 + 2 * lane < g.out_w;  // widths are even
    const bool pair_ok = ((reinterpret_cast<uintptr_t>(dst) | (uintptr_t)g.dst_pitch) & 1) == 0 &&
                         u_lo + kPieceW <= g.out_w;

    if (pair_ok) rows_tex<false>(g, cp, tex_y, tex_c, off_y, off_c, dv0, my_rows, o, in_a, in_b);
    else rows_tex<true>(g, cp, tex_y, tex_c, off_y, off_c, dv0, my_rows, o, in_a, in_b);
}

cudaError_t launch_warp_nv12_tex(const Geom& g, const FrameBatch& b, const PieceRec* table, const TexSet& ts,
                                 cudaStream_t st)
{
    dim3 block(32, kWarps);
    dim3 grid(pieces_x(g.out_w), pieces_y(g.out_h, g.piece_h), b.n_frames);
    warp_nv12_tex_kernel<<<grid, block, 0, st>>>(g, b, table, ts);
    return cudaGetLastError();
}

}  // namespace vaw
