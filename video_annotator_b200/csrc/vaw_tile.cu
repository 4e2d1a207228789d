// vaw_tile.cu -- fused map + remap for NV12, source tiles staged in shared memory by TMA
// (variant TILED).
//
// Replaces FrameSourceWarp::warp_frame's two passes
// (/root/reference/opencv/FrameSourceWarp.cpp:272-314).  Same arithmetic as vaw_poly.cu
// (coordinates from the per-piece polynomials of vaw_pieces.cuh, cv::remap's integer
// filter); what changes is where the taps come from.
//
// Why shared memory when L1 already hits 94 %: instructions and latency, not bytes.  A tap
// address in global memory is 64-bit (5 extra integer instructions per pixel), the LSU takes
// one global load per ~1.8 cycles per SM, and an L1 miss stalls a warp for an L2 round trip.
// From shared memory the four taps are LDS [a], [a+1], [a+PL], [a+PL+1] off one 32-bit IMAD,
// and the staging costs no issue slots because the TMA engine does it.
//
// One CTA (4 warps) owns one 128x32-pixel piece:
//   - thread 0 arms an mbarrier and issues cp.async.bulk.tensor loads (3-D tensor map over
//     the NV12 clip: bytes/4 x rows x frames, boxes of PL/4 x 8 rows) for the source
//     rectangle the piece's taps touch -- luma rows, then the chroma rows of the same map;
//   - meanwhile warp w collapses the piece polynomial onto column j = w of every lane and
//     the four warps exchange the results through shared memory (the collapse is amortised
//     over the whole piece, not repeated per warp);
//   - pieces that straddle the frame border get the out-of-frame cells of the tile
//     overwritten with the border value, so the sampler needs no border tests at all
//     (cv::remap's BORDER_CONSTANT replaces each out-of-image tap by the border value, which
//     is exactly what reading a border-filled cell does);
//   - warp w then walks rows 8w..8w+7 two at a time.
// Pieces whose rectangle does not fit the tile budget, pieces without a polynomial
// certificate and pure-border pieces take the paths of vaw_poly.cu.
#include <cuda.h>
#include <stdint.h>
#include <algorithm>
#include <atomic>
#include "vaw_internal.h"
#include "vaw_poly.cuh"
#include "vaw_tile.cuh"

namespace vaw {

namespace {

constexpr int kWarps = 4;
#ifndef VAW_TILE_HOIST
#define VAW_TILE_HOIST 0  // 1: also request the coefficients at CTA entry (measured slower: 0.763 ms against 0.738 ms)
#endif
#ifndef VAW_TILE_PREFETCH
#define VAW_TILE_PREFETCH 1
#endif
#ifndef VAW_TILE_CTAS
#define VAW_TILE_CTAS 6  // resident CTAs per SM the kernel is sized for (registers and shared memory)
#endif
// shared memory: [column polynomials exchanged between the warps | mbarrier | tile (TMA: 128-byte aligned)]
constexpr int kCoefBytes = 8 * 32 * 16;
constexpr int kTileOffset = kCoefBytes + 128;

}  // namespace

__global__ void __launch_bounds__(32 * kWarps, VAW_TILE_CTAS)
warp_nv12_tile_kernel(const Geom g, const FrameBatch b, const PieceRec* __restrict__ table,
                      const __grid_constant__ TileMaps maps)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x, w = threadIdx.y, tid = w * 32 + lane;
    const int px = blockIdx.x, py = blockIdx.y, frame = blockIdx.z;
    const int ph = g.piece_h, rows_per_warp = ph / kWarps;  // 32 / 16 / 8 rows per piece -> 8 / 4 / 2 per warp
    const int npx = (int)gridDim.x, npy = (int)gridDim.y;  // = pieces_x(out_w), pieces_y(out_h, ph): the launch grid, no division
    const PieceRec* rec = table + ((size_t)frame * npy + py) * npx + px;
#if VAW_TILE_PREFETCH
    // pull the coefficient lines of the record into L1 while the flags / box round trip is in flight: the
    // coefficient loads that follow the TMA issue then hit L1 instead of paying a second L2 round trip
    asm volatile("prefetch.global.L1 [%0];" ::"l"(rec));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(rec) + 128));
#endif
    // flags and box are requested together (one L2 round trip instead of two dependent ones: 0.745 -> 0.738 ms)
    const float4 rec_tail = __ldg(reinterpret_cast<const float4*>(rec) + 12);  // base.x, base.y, flags, pad
    const int4 raw = __ldg(reinterpret_cast<const int4*>(rec) + 14);          // how to stage the source box (PieceStage)
    float2 c[kNu][kNv];
#if VAW_TILE_HOIST
    load_coeffs(rec, c);
#endif
    const unsigned flags = __float_as_uint(rec_tail.z);
    const int u_lo = px * kPieceW, u0 = u_lo + 4 * lane, v_base = py * ph;
    const int rows = min(ph, g.out_h - v_base);  // even for NV12
    const int dv0 = w * rows_per_warp;
    const int my_rows = max(0, min(rows_per_warp, rows - dv0));
    const int valid = g.out_w - u0;

    if (b.skip_interior && (flags & (kPiecePoly | kPieceInterior)) == (kPiecePoly | kPieceInterior))
        return;  // variant TEX: this piece belongs to the texture kernel

    if (flags & kPieceOutside) {  // pure border: nothing to compute (before any of the sampling paths' pointer set-up)
        uint8_t* const dst = b.dst + (size_t)frame * b.dst_frame_stride;
        const unsigned yw = (g.border & 255u) * 0x01010101u;
        const unsigned cw = ((g.border >> 8) & 0xffffu) * 0x00010001u;
        if (((reinterpret_cast<uintptr_t>(dst) | (uintptr_t)g.dst_pitch) & 15) == 0 && u_lo + kPieceW <= g.out_w) {
            // 16 bytes per lane: 8 lanes per row, 4 rows per store instruction
            const int sub = lane >> 3, col = (lane & 7) * 16;
            uint8_t* yrow = dst + (size_t)(v_base + dv0 + sub) * g.dst_pitch + u_lo + col;
            uint8_t* crow = dst + (size_t)(g.out_h + ((v_base + dv0) >> 1) + sub) * g.dst_pitch + u_lo + col;
            const uint4 y4 = make_uint4(yw, yw, yw, yw), c4 = make_uint4(cw, cw, cw, cw);
            for (int r = sub; r < my_rows; r += 4, yrow += 4 * (size_t)g.dst_pitch) *reinterpret_cast<uint4*>(yrow) = y4;
            for (int r = sub; r < my_rows / 2; r += 4, crow += 4 * (size_t)g.dst_pitch) *reinterpret_cast<uint4*>(crow) = c4;
            return;
        }
        uint8_t* y0 = dst + (size_t)(v_base + dv0) * g.dst_pitch + u0;
        uint8_t* cc = dst + (size_t)(g.out_h + ((v_base + dv0) >> 1)) * g.dst_pitch + u0;
        if (valid > 0)
            for (int dv = 0; dv < my_rows; dv += 2) {
                store_word<true>(y0, yw, valid);
                store_word<true>(y0 + g.dst_pitch, yw, valid);
                store_word<true>(cc, cw, valid);
                y0 += 2 * (size_t)g.dst_pitch; cc += g.dst_pitch;
            }
        return;
    }

    PlaneRefs f;
    f.y = b.src + (size_t)frame * b.src_frame_stride;
    f.uv = f.y + (size_t)g.src_pitch * g.src_h;
    f.dst = b.dst + (size_t)frame * b.dst_frame_stride;

    RowPtrs o;
    o.y0 = f.dst + (size_t)(v_base + dv0) * g.dst_pitch + u0;
    o.y1 = o.y0 + g.dst_pitch;
    o.c = f.dst + (size_t)(g.out_h + ((v_base + dv0) >> 1)) * g.dst_pitch + u0;
    o.step_y = 2 * (size_t)g.dst_pitch;
    o.step_c = (size_t)g.dst_pitch;
    const bool word_ok = ((reinterpret_cast<uintptr_t>(f.dst) | (uintptr_t)g.dst_pitch) & 3) == 0 &&
                         u_lo + kPieceW <= g.out_w;

    if (!(flags & kPiecePoly)) {  // op-for-op per pixel
        const Rot R = load_rot(b, frame);
        for (int dv = dv0; dv < dv0 + my_rows; dv += 2) {
            float2 m[2][4];
            exact_rows(g, R, u_lo, u0, v_base + dv, m);
            sample_rows_checked(g, f, u0, v_base + dv, m);
        }
        return;
    }

    // ---- the tile of this piece's source rectangle, as the builder laid it out (block-uniform) ------
    const int lx0 = (int16_t)(raw.x & 0xffff), by0 = raw.x >> 16;
    const int cbx0 = (int16_t)(raw.y & 0xffff), cy0 = raw.y >> 16;
    const int pl = raw.z & 0xffff, nr8 = (raw.z >> 16) & 0xffff, cnr8 = raw.w & 0xffff;
    const bool fits = maps.enabled && pl != 0 && pl * (nr8 + cnr8) <= maps.tile_cap;

    if (!fits) {  // gather from global memory like variant POLY (each warp collapses for itself)
        if (my_rows <= 0) return;
        ColPoly cp;
        derive(rec, lane, cp);
        if (flags & kPieceInterior) {
            if (word_ok) band_gmem<false>(g, cp, f, dv0, my_rows, o, valid);
            else band_gmem<true>(g, cp, f, dv0, my_rows, o, valid);
        } else {
            for (int dv = dv0; dv < dv0 + my_rows; dv += 2) {
                float2 m[2][4];
                row_coords(cp, row_t(g, dv), m[0]);
                row_coords(cp, row_t(g, dv + 1), m[1]);
                sample_rows_checked(g, f, u0, v_base + dv, m);
            }
        }
        return;
    }

    uint8_t* ltile = smem + kTileOffset;
    uint8_t* ctile = ltile + nr8 * pl;
    float4* coefs = reinterpret_cast<float4*>(smem);
    const unsigned mbar = smem_u32(smem + kCoefBytes);

    // ---- thread 0: arm the barrier, launch the tile loads ---------------------------------------
    if (tid == 0) {
        mbar_init(mbar, 1);
        mbar_expect_tx(mbar, (unsigned)(pl * (nr8 + cnr8)));
        const int mi = (pl - kTileMinPitch) / kTilePitchStep;
        const CUtensorMap *map = &maps.m[mi], *map32 = &maps.m32[mi];
        const unsigned l0 = smem_u32(ltile), c0 = smem_u32(ctile);
        int k = 0;
        for (; k + 32 <= nr8; k += 32) tma_load_3d(l0 + (unsigned)(k * pl), map32, lx0 >> 2, by0 + k, frame + b.tma_frame0, mbar);
        for (; k < nr8; k += 8) tma_load_3d(l0 + (unsigned)(k * pl), map, lx0 >> 2, by0 + k, frame + b.tma_frame0, mbar);
        for (k = 0; k + 32 <= cnr8; k += 32)
            tma_load_3d(c0 + (unsigned)(k * pl), map32, cbx0 >> 2, g.src_h + cy0 + k, frame + b.tma_frame0, mbar);
        for (; k < cnr8; k += 8)
            tma_load_3d(c0 + (unsigned)(k * pl), map, cbx0 >> 2, g.src_h + cy0 + k, frame + b.tma_frame0, mbar);
    }

    // ---- collapse the polynomial: warp w does column j = w for every lane -------------------------
    {
        float2 a[kNv];
#if !VAW_TILE_HOIST
        load_coeffs(rec, c);
#endif
        collapse_column(c, ((float)pair_column(lane, w) - 63.5f) * 0.015625f, a);  // s is exact
        coefs[(2 * w) * 32 + lane] = make_float4(a[0].x, a[0].y, a[1].x, a[1].y);
        coefs[(2 * w + 1) * 32 + lane] = make_float4(a[2].x, a[2].y, a[3].x, a[3].y);
    }
    __syncthreads();  // column polynomials exchanged; the barrier init is visible to every thread
    ColPoly cp;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 lo = coefs[(2 * j) * 32 + lane], hi = coefs[(2 * j + 1) * 32 + lane];
        cp.a[j][0] = make_float2(lo.x, lo.y); cp.a[j][1] = make_float2(lo.z, lo.w);
        cp.a[j][2] = make_float2(hi.x, hi.y); cp.a[j][3] = make_float2(hi.z, hi.w);
    }
    cp.base = make_float2(rec_tail.x, rec_tail.y);

#ifndef VAW_ABL_NO_TMA_WAIT
    mbar_wait(mbar, 0);  // the tile has landed
#endif

    if (!(flags & kPieceInterior)) {  // straddles the frame border: paint the outside cells
        const unsigned by_ = g.border & 255u, bu = (g.border >> 8) & 255u, bv = (g.border >> 16) & 255u;
        fill_border(ltile, pl, nr8, by0, g.src_h, lx0, g.src_w, by_ * 0x01010101u, tid, 32 * kWarps);
        fill_border(ctile, pl, cnr8, cy0, g.src_h >> 1, cbx0, g.src_w, (bu | (bv << 8)) * 0x00010001u, tid, 32 * kWarps);
        __syncthreads();
    }

    if (my_rows <= 0) return;
    // tap address = (iy - y0) * pl + (ix - x0) + tile, with the >>5 bias of the magic constant folded in
    const unsigned upl = (unsigned)pl;
    const unsigned lconst = smem_u32(ltile) - (unsigned)by0 * upl - (unsigned)lx0 - kMagicShift * upl - kMagicShift;
    const unsigned cconst = ((smem_u32(ctile) - (unsigned)cy0 * upl - (unsigned)cbx0 - kMagicShift * upl) >> 1) - kMagicShift;
    // the staged path uses the pair mapping: re-base the output pointers on column 2 * lane
    const int shift = 2 * lane - 4 * lane;
    o.y0 += shift; o.y1 += shift; o.c += shift;
    const bool in_a = u_lo + 2 * lane < g.out_w, in_b = u_lo + 64 + 2 * lane < g.out_w;  // widths are even
    const bool pair_ok = ((reinterpret_cast<uintptr_t>(f.dst) | (uintptr_t)g.dst_pitch) & 1) == 0 &&
                         u_lo + kPieceW <= g.out_w;
    const TileBounds tb = {smem_u32(ltile), smem_u32(ltile) + (unsigned)(nr8 * pl), smem_u32(ctile),
                           smem_u32(ctile) + (unsigned)(cnr8 * pl)};
    if (pair_ok) rows_tile<false>(g, cp, lconst, cconst, upl, dv0, my_rows, o, in_a, in_b, tb);
    else rows_tile<true>(g, cp, lconst, cconst, upl, dv0, my_rows, o, in_a, in_b, tb);
}

long long tile_oob_count()
{
#ifdef VAW_BOUNDS_CHECK
    unsigned long long v = 0;
    if (cudaMemcpyFromSymbol(&v, g_oob_taps, sizeof v) != cudaSuccess) return -2;
    return (long long)v;
#else
    return -1;
#endif
}

int tile_smem_bytes(int tile_cap) { return kTileOffset + tile_cap; }

// Host mirror of the kernel's tile sizing (same integer arithmetic).
int tile_need_bytes(const PieceRec& rec)
{
    if (!(rec.flags & kPiecePoly) || (rec.flags & kPieceOutside)) return 0;
    const PieceBox& b = rec.box;
    const int lx0 = b.x0 & ~15, wb = (b.x1 - lx0 + 16) & ~15;
    const int cbx0 = (2 * b.cx0) & ~15, cwb = (2 * b.cx1 + 2 - cbx0 + 15) & ~15;
    const int nr8 = (b.y1 - b.y0 + 8) & ~7, cnr8 = (b.cy1 - b.cy0 + 8) & ~7;
    const int pl = std::max(kTileMinPitch, (std::max(wb, cwb) + 31) & ~31);
    if (pl > kTileMaxPitch || nr8 <= 0 || cnr8 <= 0) return 0x7fffffff;
    return pl * (nr8 + cnr8);
}

cudaError_t launch_warp_nv12_tile(const Geom& g, const FrameBatch& b, const PieceRec* table, const TileMaps& maps,
                                  cudaStream_t st)
{
    // once per device (the clip scheduler launches from one host thread per device); ordinals
    // beyond the table simply set the attribute on every launch
    static std::atomic<bool> configured[64];
    int dev = 0;
    cudaGetDevice(&dev);
    const bool tracked = dev >= 0 && dev < 64;
    if (!tracked || !configured[dev].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(warp_nv12_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             tile_smem_bytes(kTileCapMax));
        if (e != cudaSuccess) return e;
        if (tracked) configured[dev].store(true, std::memory_order_release);
    }
    dim3 block(32, kWarps);
    dim3 grid(pieces_x(g.out_w), pieces_y(g.out_h, g.piece_h), b.n_frames);
    warp_nv12_tile_kernel<<<grid, block, tile_smem_bytes(maps.tile_cap), st>>>(g, b, table, maps);
    return cudaGetLastError();
}

}  // namespace vaw
