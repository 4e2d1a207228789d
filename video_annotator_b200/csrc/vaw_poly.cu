// vaw_poly.cu -- fused map + remap for NV12 with coordinates from the per-piece polynomials,
// taps gathered from global memory through L1/L2 (variant POLY), and the coordinate dump.
//
// Replaces FrameSourceWarp::warp_frame's two passes
// (/root/reference/opencv/FrameSourceWarp.cpp:272-314) like vaw_kernels.cu does, but is
// built around the instruction budget of a B200 (vaw_pieces.cuh): a warp owns one piece of
// 128 x PH pixels; lane l owns the four columns 4l..4l+3 and walks down the rows two at a
// time (two luma rows + one chroma row per step, 4-byte stores -> every store instruction
// of the warp writes one full 128-byte line).  Before the walk each lane collapses the
// piece's tensor polynomial onto its four columns (160 FMAs, amortised over 4*PH pixels), so
// a coordinate costs 3 FMAs + 1 add per pixel; the 1/32-px fixed-point conversion is a
// single FMA with the 1.5*2^23 constant.  Pieces the builder classified as interior take a
// sampler without border tests; pure-border pieces are just filled; the rest use the
// checked sampler; pieces without a polynomial certificate fall back to the per-pixel
// op-for-op evaluation (vaw_coords.cuh).  The shared-memory variant is vaw_tile.cu.
#include <stdint.h>
#include "vaw_internal.h"
#include "vaw_poly.cuh"

namespace vaw {

namespace {

constexpr int kWarps = 4;  // warps per CTA, stacked vertically (4 pieces)

__global__ void __launch_bounds__(32 * kWarps)
warp_nv12_poly_kernel(const Geom g, const FrameBatch b, const PieceRec* __restrict__ table)
{
    const int lane = threadIdx.x;
    const int px = blockIdx.x, py = blockIdx.y * kWarps + threadIdx.y, frame = blockIdx.z;
    const int ph = g.piece_h;
    const int npx = pieces_x(g.out_w), npy = pieces_y(g.out_h, ph);
    if (py >= npy) return;  // warp-uniform; no CTA-wide barrier is used anywhere below
    const PieceRec* rec = table + ((size_t)frame * npy + py) * npx + px;
    const unsigned flags = __ldg(&rec->flags);
    const int u_lo = px * kPieceW, u0 = u_lo + 4 * lane, v_base = py * ph;
    const int rows = min(ph, g.out_h - v_base);  // even for NV12
    const int valid = g.out_w - u0;

    PlaneRefs f;
    f.y = b.src + (size_t)frame * b.src_frame_stride;
    f.uv = f.y + (size_t)g.src_pitch * g.src_h;
    f.dst = b.dst + (size_t)frame * b.dst_frame_stride;

    RowPtrs o;
    o.y0 = f.dst + (size_t)v_base * g.dst_pitch + u0;
    o.y1 = o.y0 + g.dst_pitch;
    o.c = f.dst + (size_t)(g.out_h + (v_base >> 1)) * g.dst_pitch + u0;
    o.step_y = 2 * (size_t)g.dst_pitch;
    o.step_c = (size_t)g.dst_pitch;

    if (flags & kPieceOutside) {  // pure border: nothing to compute
        const unsigned yw = (g.border & 255u) * 0x01010101u;
        const unsigned cw = ((g.border >> 8) & 0xffffu) * 0x00010001u;
        if (valid > 0)
            for (int dv = 0; dv < rows; dv += 2) {
                store_word<true>(o.y0, yw, valid);
                store_word<true>(o.y1, yw, valid);
                store_word<true>(o.c, cw, valid);
                o.y0 += o.step_y; o.y1 += o.step_y; o.c += o.step_c;
            }
        return;
    }

    if (!(flags & kPiecePoly)) {  // op-for-op per pixel
        const Rot R = load_rot(b, frame);
        for (int dv = 0; dv < rows; dv += 2) {
            float2 m[2][4];
            exact_rows(g, R, u_lo, u0, v_base + dv, m);
            sample_rows_checked(g, f, u0, v_base + dv, m);
        }
        return;
    }

    ColPoly cp;
    derive(rec, lane, cp);

    if (!(flags & kPieceInterior)) {  // polynomial coordinates, checked sampler
        for (int dv = 0; dv < rows; dv += 2) {
            float2 m[2][4];
            row_coords(cp, row_t(g, dv), m[0]);
            row_coords(cp, row_t(g, dv + 1), m[1]);
            sample_rows_checked(g, f, u0, v_base + dv, m);
        }
        return;
    }

    // 4-byte stores need a 4-byte aligned frame base and row pitch and a full piece
    const bool word_ok = ((reinterpret_cast<uintptr_t>(f.dst) | (uintptr_t)g.dst_pitch) & 3) == 0;
    if (word_ok && u_lo + kPieceW <= g.out_w)
        band_gmem<false>(g, cp, f, 0, rows, o, valid);
    else
        band_gmem<true>(g, cp, f, 0, rows, o, valid);
}

// The map the kernel above samples with: plane 0 luma, plane 1 chroma.
__global__ void __launch_bounds__(32 * kWarps)
dump_coords_poly_kernel(const Geom g, const Rot R, const PieceRec* __restrict__ table, int plane,
                        float* __restrict__ map_x, float* __restrict__ map_y, int map_pitch)
{
    const int lane = threadIdx.x;
    const int px = blockIdx.x, py = blockIdx.y * kWarps + threadIdx.y;
    const int ph = g.piece_h;
    const int npx = pieces_x(g.out_w), npy = pieces_y(g.out_h, ph);
    if (py >= npy) return;
    const PieceRec* rec = table + (size_t)py * npx + px;
    const unsigned flags = __ldg(&rec->flags);
    const int u_lo = px * kPieceW, u0 = u_lo + 4 * lane, v_base = py * ph;
    const int rows = min(ph, g.out_h - v_base);
    ColPoly cp;
    if (flags & kPiecePoly) derive(rec, lane, cp);
    for (int dv = 0; dv < rows; dv += 2) {
        float2 m[2][4];
        const int v0 = v_base + dv;
        if (flags & kPiecePoly) {
            row_coords(cp, row_t(g, dv), m[0]);
            row_coords(cp, row_t(g, dv + 1), m[1]);
        } else {
            exact_rows(g, R, u_lo, u0, v0, m);
        }
        if (plane == 0) {
            for (int r = 0; r < 2; ++r)
                for (int i = 0; i < 4; ++i)
                    if (u0 + i < g.out_w && v0 + r < g.out_h) {
                        map_x[(size_t)(v0 + r) * map_pitch + u0 + i] = m[r][i].x;
                        map_y[(size_t)(v0 + r) * map_pitch + u0 + i] = m[r][i].y;
                    }
        } else {
            for (int q = 0; q < 2; ++q)
                if (u0 + 2 * q + 1 < g.out_w && v0 + 1 < g.out_h) {
                    const size_t o = (size_t)(v0 >> 1) * map_pitch + (u0 >> 1) + q;
                    map_x[o] = chroma_coord(m[0][2 * q].x, m[0][2 * q + 1].x, m[1][2 * q].x, m[1][2 * q + 1].x);
                    map_y[o] = chroma_coord(m[0][2 * q].y, m[0][2 * q + 1].y, m[1][2 * q].y, m[1][2 * q + 1].y);
                }
        }
    }
}

}  // namespace

cudaError_t launch_warp_nv12_poly(const Geom& g, const FrameBatch& b, const PieceRec* table, cudaStream_t st)
{
    dim3 block(32, kWarps);
    dim3 grid(pieces_x(g.out_w), (pieces_y(g.out_h, g.piece_h) + kWarps - 1) / kWarps, b.n_frames);
    warp_nv12_poly_kernel<<<grid, block, 0, st>>>(g, b, table);
    return cudaGetLastError();
}

cudaError_t launch_dump_coords_poly(const Geom& g, const Rot& rot, const PieceRec* table, int plane,
                                    float* map_x, float* map_y, int map_pitch, cudaStream_t st)
{
    dim3 block(32, kWarps);
    dim3 grid(pieces_x(g.out_w), (pieces_y(g.out_h, g.piece_h) + kWarps - 1) / kWarps, 1);
    dump_coords_poly_kernel<<<grid, block, 0, st>>>(g, rot, table, plane, map_x, map_y, map_pitch);
    return cudaGetLastError();
}

}  // namespace vaw
