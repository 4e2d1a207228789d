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
#include <cstdlib>
#include <atomic>
#include "vaw_internal.h"
#include "vaw_poly.cuh"
#include "vaw_tile.cuh"

namespace vaw {

namespace {

#ifndef VAW_QUAD_WARPS
#define VAW_QUAD_WARPS 4  // warps per CTA of the quadrant kernel: 4 (64 columns x PH/2 rows each) or 2 (64 columns x PH rows each)
#endif
constexpr int kWarps = VAW_QUAD_WARPS;
static_assert(kWarps == 2 || kWarps == 4, "two warps side by side, one or two deep");

}  // namespace

// Quadrant kernel: the round-1 tile kernel (one CTA per piece, four columns per lane) re-cut for occupancy.
//
// What round 1's profile said (profiles/r01_ncu_summary.txt): the sampler is issue- and latency-bound,
// issue slots 71 % busy with 24 resident warps per SM, and both limits on the warp count -- 80
// registers and 37 KB of shared memory per CTA -- came from the lane mapping: four columns per lane
// = 32 registers of column polynomials, and a 4 KB exchange buffer + barrier to share the collapse
// between the warps.  Here warp w owns a QUADRANT of the piece (64 columns x PH/2 rows), lane l the
// two adjacent columns 2l, 2l+1 of it:
//   - 16 registers of column polynomials -> the kernel fits 64 registers, 8 CTAs (32 warps) per SM;
//   - each warp collapses its own two columns (40 FFMA2, coefficients straight from L1): no exchange
//     buffer, no barrier between the collapse and the sampling, shared memory = 128 bytes + tile;
//   - tile rows come in multiples of 4 (32/8/4-row TMA boxes), so the C3 tiles fit 8 per SM;
//   - per row pair a lane makes 2 x 2 luma samples and the one chroma sample of that quad; a warp
//     store writes 64 contiguous bytes, the LDS of a warp stay within one 128-byte window per row
//     exactly as with the round-1 pair mapping (2 columns per lane).
constexpr int kQuadRecOffset = 16, kQuadTileOffset = 256;  // [tile mbarrier | record mbarrier | record (240 B) | tile (TMA: 128-byte aligned)]

// nrows (even) rows starting at piece row dv0 for the lane's two columns; taps from the staged tile.
// py / pc: the lane's column pair in the first luma row / chroma row of the band.
#ifndef VAW_PREFETCH_AFTER_BARRIER
#define VAW_PREFETCH_AFTER_BARRIER 0
#endif
#ifndef VAW_ONE_WAITER
#define VAW_ONE_WAITER 0  // 1: one thread polls the record / tile mbarriers and the others park at a CTA barrier behind it
                          // (removes the try_wait loops, 3.3 % of the executed instructions -- and measures 0.6 % SLOWER:
                          // the extra barrier costs more than the polling)
#endif
#ifndef VAW_VECTOR_CONSTS
#define VAW_VECTOR_CONSTS 1  // the 32.0 / 16.0 scale factors of the samplers in vector registers (0: re-made from uniforms every row pair)
#endif
#ifndef VAW_RUNNING_PTRS
#define VAW_RUNNING_PTRS 0
#endif
#ifndef VAW_SAMPLER
#define VAW_SAMPLER 3  // 1: luma_tile / chroma_tile (shift + mask + six multiply-adds per sample); 3: luma_tile3 / chroma_tile3
#endif
#if VAW_SAMPLER == 3
using TapConst = FloorConst;
#else
using TapConst = unsigned;
#endif

// kMode: 0 cv::INTER_LINEAR, 1 cv::INTER_NEAREST, 2 cv::INTER_CUBIC, 3 cv::INTER_LANCZOS4 (2, 3: vaw_tile.cuh's table samplers
// on tiles staged with a halo)
template <bool kRagged, int kMode>
__device__ __forceinline__ void rows_quad(const Geom& g, const ColPoly2& cp, const TapConst& lconst, const TapConst& cconst, unsigned pl,
                                          int dv0, int nrows, uint8_t* __restrict__ py, uint8_t* __restrict__ pc,
                                          bool inside, const TileBounds& tb)
{
    float t = row_t(g, dv0);
    const float dt = g.t_scale, dt2 = __fadd_rn(g.t_scale, g.t_scale);  // stepping the dyadic t is exact
    // the pitch in a VECTOR register (threadIdx.x >> 5 is 0, which ptxas does not know): with a uniform-register
    // multiplicand IMAD.WIDE takes no 64-bit addend and every store address costs three instructions instead of one
    const unsigned dpitch = (unsigned)g.dst_pitch + (threadIdx.x >> 5);
    const unsigned hpitch = dpitch >> 1;
    const unsigned long long gy0 = (unsigned long long)__cvta_generic_to_global(py);
    const unsigned long long gc = (unsigned long long)__cvta_generic_to_global(pc);
#if VAW_RUNNING_PTRS
    unsigned long long q0 = gy0, q1 = gy0 + (unsigned)g.dst_pitch, qc = gc;
    const unsigned long long step1 = (unsigned)g.dst_pitch, step2 = 2ull * (unsigned)g.dst_pitch;
#endif
#ifndef VAW_UNROLL2
#define VAW_UNROLL2 1  // two row pairs per trip: 135.5 instead of 138 instructions per row pair, same registers
#endif
#if VAW_UNROLL2
#pragma unroll 2
#else
#pragma unroll 1
#endif
    for (unsigned j2 = opaque_zero(); j2 < (unsigned)nrows; j2 += 2) {  // rows j2, j2 + 1 of the band
        const float2 t0 = pair(t), t1 = pair(__fadd_rn(t, dt));
        t = __fadd_rn(t, dt2);
        const float2 m00 = col_coord(cp.a[0], cp.base, t0), m01 = col_coord(cp.a[1], cp.base, t0);
        const float2 m10 = col_coord(cp.a[0], cp.base, t1), m11 = col_coord(cp.a[1], cp.base, t1);
#if VAW_SAMPLER == 3
        unsigned y00, y01, y10, y11, c;
        if (kMode >= 2) {  // cv::INTER_CUBIC / cv::INTER_LANCZOS4: lconst / cconst carry the block's top-left tap
            constexpr int kKs = kMode == 2 ? 4 : 8;
            y00 = luma_tile_hi<kKs>(lconst, pl, m00, g.cubic_tab, tb); y01 = luma_tile_hi<kKs>(lconst, pl, m01, g.cubic_tab, tb);
            y10 = luma_tile_hi<kKs>(lconst, pl, m10, g.cubic_tab, tb); y11 = luma_tile_hi<kKs>(lconst, pl, m11, g.cubic_tab, tb);
            c = chroma_tile_hi<kKs>(cconst, pl, chroma_z(m00, m01, m10, m11), g.cubic_tab, tb);
        } else if (kMode == 1) {  // cv::INTER_NEAREST: one tap per sample (lconst / cconst carry the nearest row constants)
            y00 = luma_tile_nearest(lconst.row0, pl, m00); y01 = luma_tile_nearest(lconst.row0, pl, m01);
            y10 = luma_tile_nearest(lconst.row0, pl, m10); y11 = luma_tile_nearest(lconst.row0, pl, m11);
            c = chroma_tile_nearest(cconst.row0, pl, chroma_z(m00, m01, m10, m11));
        } else {
            y00 = luma_tile3(lconst, pl, m00, tb) >> 10; y01 = luma_tile3(lconst, pl, m01, tb) >> 10;
            y10 = luma_tile3(lconst, pl, m10, tb) >> 10; y11 = luma_tile3(lconst, pl, m11, tb) >> 10;
            c = chroma_tile3(cconst, pl, chroma_z(m00, m01, m10, m11), tb);  // U | V << 8
        }
#else
        const unsigned y00 = (unsigned)luma_tile(lconst, pl, m00, tb) >> 10, y01 = (unsigned)luma_tile(lconst, pl, m01, tb) >> 10;
        const unsigned y10 = (unsigned)luma_tile(lconst, pl, m10, tb) >> 10, y11 = (unsigned)luma_tile(lconst, pl, m11, tb) >> 10;
        const unsigned c = chroma_tile(cconst, pl, chroma_z(m00, m01, m10, m11), tb);  // U | V << 8
#endif
        // one IMAD.WIDE per store address: distinct multiplicands keep ptxas from sharing the product and adding the
        // three 64-bit bases to it afterwards (two more instructions per address)
#if VAW_RUNNING_PTRS
        const unsigned long long r0 = q0, r1 = q1, rc = qc;
        q0 += step2; q1 += step2; qc += step1;
#else
        const unsigned long long r0 = row_ptr(gy0, j2, dpitch), r1 = row_ptr(gy0, j2 + 1u, dpitch);
        const unsigned long long rc = kRagged ? row_ptr(gc, j2 >> 1, dpitch) : row_ptr(gc, j2, hpitch);  // even pitch
#endif
        if (!kRagged) {
            stg_u16(r0, y00 | (y01 << 8));
            stg_u16(r1, y10 | (y11 << 8));
            stg_u16(rc, c);
        } else if (inside) {
            stg_u8(r0, y00); stg_u8(r0 + 1, y01);
            stg_u8(r1, y10); stg_u8(r1 + 1, y11);
            stg_u8(rc, c); stg_u8(rc + 1, c >> 8);
        }
    }
}

// One piece, from its record in shared memory (`rs`) to the stores.  (A separate function since round 2's persistent-
// kernel experiment: resident CTAs pulling pieces from an atomic queue with the next record prefetched measured
// 0.736 ms against 0.660 ms -- more instructions, not fewer, and longer tile waits)  `rec` = the same record in the table (the gather fallbacks read it from there);
// `tile_parity` = phase of the tile mbarrier (smem + 0) this piece's loads complete.  Returns whether the piece was staged
// (i.e. whether that phase was consumed).
template <int kMode>
__device__ __forceinline__ bool quad_piece(const Geom& g, const FrameBatch& b, const PieceRec* __restrict__ rec,
                                           const PieceRec* rs, const TileMaps& maps, uint8_t* smem, int px, int py,
                                           int frame, unsigned tile_parity, int lane, int w, int tid)
{
    const int ph = g.piece_h;
    const float4 rec_tail = *(reinterpret_cast<const float4*>(rs) + 12);  // base.x, base.y, flags, pad
    const int4 raw = *(reinterpret_cast<const int4*>(rs) + 14);          // PieceStage

    const unsigned flags = __float_as_uint(rec_tail.z);
    const int u_lo = px * kPieceW, v_base = py * ph;
    const int rows = min(ph, g.out_h - v_base);  // even for NV12
    uint8_t* const dst = b.dst + (size_t)frame * b.dst_frame_stride;

    if (b.skip_interior && (flags & (kPiecePoly | kPieceInterior)) == (kPiecePoly | kPieceInterior))
        return false;  // variant TEX: this piece belongs to the texture kernel

    if (flags & kPieceOutside) {  // pure border: 128 threads fill the piece
        const unsigned yw = (g.border & 255u) * 0x01010101u;
        const unsigned cw = ((g.border >> 8) & 0xffffu) * 0x00010001u;
        if (((reinterpret_cast<uintptr_t>(dst) | (uintptr_t)g.dst_pitch) & 15) == 0 && u_lo + kPieceW <= g.out_w) {
            // 16 bytes per lane: 8 lanes per row, 16 rows per pass of the CTA
            constexpr int kPass = 4 * kWarps;
            const int sub = tid >> 3, col = (tid & 7) * 16;
            const uint4 y4 = make_uint4(yw, yw, yw, yw), c4 = make_uint4(cw, cw, cw, cw);
            uint8_t* yrow = dst + (size_t)(v_base + sub) * g.dst_pitch + u_lo + col;
            for (int r = sub; r < rows; r += kPass, yrow += kPass * (size_t)g.dst_pitch) *reinterpret_cast<uint4*>(yrow) = y4;
            uint8_t* crow = dst + (size_t)(g.out_h + (v_base >> 1) + sub) * g.dst_pitch + u_lo + col;
            for (int r = sub; r < rows / 2; r += kPass, crow += kPass * (size_t)g.dst_pitch) *reinterpret_cast<uint4*>(crow) = c4;
            return false;
        }
        // ragged / unaligned: 4 columns per lane, rows split over the warps like the round-1 kernel
        const int u0 = u_lo + 4 * lane, valid = g.out_w - u0, rpw = ph / kWarps;
        const int dv0 = w * rpw, my = max(0, min(rpw, rows - dv0));
        uint8_t* y0 = dst + (size_t)(v_base + dv0) * g.dst_pitch + u0;
        uint8_t* cc = dst + (size_t)(g.out_h + ((v_base + dv0) >> 1)) * g.dst_pitch + u0;
        if (valid > 0)
            for (int dv = 0; dv < my; dv += 2) {
                store_word<true>(y0, yw, valid);
                store_word<true>(y0 + g.dst_pitch, yw, valid);
                store_word<true>(cc, cw, valid);
                y0 += 2 * (size_t)g.dst_pitch; cc += g.dst_pitch;
            }
        return false;
    }

    // ---- the tile of this piece's source rectangle, as the builder laid it out (block-uniform) ------
    const int lx0 = (int16_t)(raw.x & 0xffff), by0 = raw.x >> 16;
    const int cbx0 = (int16_t)(raw.y & 0xffff), cy0 = raw.y >> 16;
    const int pl = raw.z & 0xffff, nrows = (raw.z >> 16) & 0xffff, cnrows = raw.w & 0xffff;
    constexpr bool kNearest = kMode == 1;
    constexpr int kChecked = kMode >= 2 ? (int)kModeTable : kMode;  // the per-pixel fallbacks' filter
    constexpr int kHalo = kMode == 2 ? 1 : (kMode == 3 ? 3 : 0);    // = GeomD::halo of the builder
    const bool staged = (flags & kPiecePoly) && maps.enabled && pl != 0 &&
                        pl * (nrows + cnrows) + (kMode >= 2 ? kTileSlack : 0) <= maps.tile_cap;

    if (!staged) {
        // pieces without a polynomial certificate or whose box does not fit the tile: the gather paths of
        // variant POLY with the 4-columns-per-lane mapping (warp w walks rows [w PH/4, (w+1) PH/4))
        const int rpw = ph / kWarps, dv0 = w * rpw, my_rows = max(0, min(rpw, rows - dv0));
        const int u0 = u_lo + 4 * lane, valid = g.out_w - u0;
        if (my_rows <= 0) return false;
        PlaneRefs f;
        f.y = b.src + (size_t)frame * b.src_frame_stride;
        f.uv = f.y + (size_t)g.src_pitch * g.src_h;
        f.dst = dst;
        if (!(flags & kPiecePoly)) {  // op-for-op per pixel
            const Rot R = load_rot(b, frame);
            for (int dv = dv0; dv < dv0 + my_rows; dv += 2) {
                float2 m[2][4];
                exact_rows(g, R, u_lo, u0, v_base + dv, m);
                sample_rows_checked<kChecked>(g, f, u0, v_base + dv, m);
            }
            return false;
        }
        ColPoly cp;
        derive(rec, lane, cp);
        if (kMode == 0 && (flags & kPieceInterior)) {
            RowPtrs o;
            o.y0 = dst + (size_t)(v_base + dv0) * g.dst_pitch + u0;
            o.y1 = o.y0 + g.dst_pitch;
            o.c = dst + (size_t)(g.out_h + ((v_base + dv0) >> 1)) * g.dst_pitch + u0;
            o.step_y = 2 * (size_t)g.dst_pitch;
            o.step_c = (size_t)g.dst_pitch;
            const bool word_ok = ((reinterpret_cast<uintptr_t>(dst) | (uintptr_t)g.dst_pitch) & 3) == 0 &&
                                 u_lo + kPieceW <= g.out_w;
            if (word_ok) band_gmem<false>(g, cp, f, dv0, my_rows, o, valid);
            else band_gmem<true>(g, cp, f, dv0, my_rows, o, valid);
        } else {
            for (int dv = dv0; dv < dv0 + my_rows; dv += 2) {
                float2 m[2][4];
                row_coords(cp, row_t(g, dv), m[0]);
                row_coords(cp, row_t(g, dv + 1), m[1]);
                sample_rows_checked<kChecked>(g, f, u0, v_base + dv, m);
            }
        }
        return false;
    }

    uint8_t* ltile = smem + kQuadTileOffset;
    const unsigned mbar = smem_u32(smem);
    uint8_t* ctile = ltile + nrows * pl;

    // ---- thread 0: launch the tile loads ------------------------------------------------------------
#ifdef VAW_ABL_NO_TMA  // analysis only: no tile loads (the loop samples whatever the shared memory holds)
    if (tid == 0 && g.out_w < 0) {
#else
    if (tid == 0) {
#endif
        mbar_expect_tx(mbar, (unsigned)(pl * (nrows + cnrows)));
        const int mi = (pl - kTileMinPitch) / kTilePitchStep;
        const CUtensorMap *map = &maps.m[mi], *map32 = &maps.m32[mi], *map4 = &maps.m4[mi];
        const unsigned l0 = smem_u32(ltile), c0 = smem_u32(ctile);
        const int z = frame + b.tma_frame0;
        int k = 0;
        for (; k + 32 <= nrows; k += 32) tma_load_3d(l0 + (unsigned)(k * pl), map32, lx0 >> 2, by0 + k, z, mbar);
        for (; k + 8 <= nrows; k += 8) tma_load_3d(l0 + (unsigned)(k * pl), map, lx0 >> 2, by0 + k, z, mbar);
        if (k < nrows) tma_load_3d(l0 + (unsigned)(k * pl), map4, lx0 >> 2, by0 + k, z, mbar);
        for (k = 0; k + 32 <= cnrows; k += 32) tma_load_3d(c0 + (unsigned)(k * pl), map32, cbx0 >> 2, g.src_h + cy0 + k, z, mbar);
        for (; k + 8 <= cnrows; k += 8) tma_load_3d(c0 + (unsigned)(k * pl), map, cbx0 >> 2, g.src_h + cy0 + k, z, mbar);
        if (k < cnrows) tma_load_3d(c0 + (unsigned)(k * pl), map4, cbx0 >> 2, g.src_h + cy0 + k, z, mbar);
#if VAW_PREFETCH_NEXT_FRAME
        // the same boxes of the NEXT frame into L2: consecutive frames' rotations differ by a fraction of a degree, so
        // the next frame's piece at this position reads (nearly) this rectangle about one frame's worth of CTAs from now
        if (frame + 1 < (int)gridDim.z) {
            for (k = 0; k + 32 <= nrows; k += 32) tma_prefetch_3d(map32, lx0 >> 2, by0 + k, z + 1);
            for (; k + 8 <= nrows; k += 8) tma_prefetch_3d(map, lx0 >> 2, by0 + k, z + 1);
            if (k < nrows) tma_prefetch_3d(map4, lx0 >> 2, by0 + k, z + 1);
            for (k = 0; k + 32 <= cnrows; k += 32) tma_prefetch_3d(map32, cbx0 >> 2, g.src_h + cy0 + k, z + 1);
            for (; k + 8 <= cnrows; k += 8) tma_prefetch_3d(map, cbx0 >> 2, g.src_h + cy0 + k, z + 1);
            if (k < cnrows) tma_prefetch_3d(map4, cbx0 >> 2, g.src_h + cy0 + k, z + 1);
        }
#endif
    }

    // ---- this warp's quadrant; every lane collapses the polynomial onto its two columns ---------------
    const int wx = w & 1, wy = w >> 1, hrows = ph / (kWarps / 2);
    const int col0 = 64 * wx + 2 * lane;  // within the piece
    ColPoly2 cp;
    derive2(rs, col0, cp);
    cp.base = make_float2(rec_tail.x, rec_tail.y);

#ifndef VAW_ABL_NO_TMA
#if VAW_ONE_WAITER
    if (tid == 0) mbar_wait_parked(mbar, tile_parity, 4000);  // the tile has landed
    __syncthreads();
#else
    mbar_wait_parked(mbar, tile_parity, 4000);  // the tile has landed (the warp is parked, not spinning, until then)
#endif
#endif
#ifdef VAW_ABL_NO_LOOP  // analysis only: the per-piece set-up without the row loop
    if (cp.a[0][0].x + cp.a[1][3].y == 12345.f) dst[0] = 1;
    return false;
#endif

    if (!(flags & kPieceInterior)) {  // straddles the frame border: paint the outside cells
        const unsigned by_ = g.border & 255u, bu = (g.border >> 8) & 255u, bv = (g.border >> 16) & 255u;
        fill_border(ltile, pl, nrows, by0, g.src_h, lx0, g.src_w, by_ * 0x01010101u, tid, 32 * kWarps);
        fill_border(ctile, pl, cnrows, cy0, g.src_h >> 1, cbx0, g.src_w, (bu | (bv << 8)) * 0x00010001u, tid, 32 * kWarps);
        __syncthreads();
    }

    const int dv0 = wy * hrows;
    const int my_rows = max(0, min(hrows, rows - dv0));  // 0 for a warp below the frame's last row: its row loop does not run
    // (no early return here: a branch on a per-warp value in front of the tap-row constants below can cost them their
    // uniform registers -- everything behind it is divergent as far as ptxas can tell)
    // tap address = (iy - y0) * pl + (ix - x0) + tile, with the >>5 bias of the magic constant folded in
#if VAW_SHFL_UNIFORM
    const unsigned upl = __shfl_sync(0xffffffffu, (unsigned)pl, 0);  // warp-uniform as far as ptxas is concerned: a uniform register
#else
    const unsigned upl = (unsigned)pl;
#endif
#if VAW_SAMPLER == 3
    // floor constants: tile origin (and, for luma, the tile's shared-memory address) folded into the round-down FMA
    const unsigned never = (unsigned)g.out_w >> 31;  // 0 at run time: constants built from it stay loop-invariant uniforms
#if VAW_VECTOR_CONSTS
    const unsigned vnever = threadIdx.x >> 5;        // 0 as well, but per thread as far as ptxas knows: a vector register
#else
    const unsigned vnever = never;
#endif
    const FloorConst lconst = floor_const(-lx0 - kHalo, -by0 - kHalo, smem_u32(ltile) - 0x40000000u, upl, __uint_as_float(0x42000000u | vnever), (unsigned)raw.w >> 31);  // (the stage's zero pad: with g.out_w >> 31 ptxas keeps the luma pair in vector registers)
    const FloorConst cconst = floor_const(-(cbx0 >> 1) - kHalo, -cy0 - kHalo, smem_u32(ctile) - 0x80000000u, upl, __uint_as_float(0x41800000u | vnever), never);
    // INTER_NEAREST: the row constants of luma_tile_nearest / chroma_tile_nearest instead (same uniform-register treatment)
    const FloorConst lnear = floor_const(0, 0, smem_u32(ltile) - (unsigned)by0 * upl - (unsigned)lx0 - (unsigned)kMagicBits * (upl + 1u), upl, 0.f, (unsigned)raw.w >> 31);
    const FloorConst cnear = floor_const(0, 0, smem_u32(ctile) - (unsigned)cy0 * upl - (unsigned)cbx0 - (unsigned)kMagicBits * (upl + 2u), upl, 0.f, never);
#else
    const unsigned lconst = smem_u32(ltile) - (unsigned)by0 * upl - (unsigned)lx0 - kMagicShift * upl - kMagicShift;
    const unsigned cconst = ((smem_u32(ctile) - (unsigned)cy0 * upl - (unsigned)cbx0 - kMagicShift * upl) >> 1) - kMagicShift;
#endif
    const int u0 = u_lo + col0;
    const bool inside = u0 < g.out_w;  // widths are even: the pair is inside or outside together
    const bool pair_ok = ((reinterpret_cast<uintptr_t>(dst) | (uintptr_t)g.dst_pitch) & 1) == 0 &&
                         u_lo + kPieceW <= g.out_w;
    const TileBounds tb = {smem_u32(ltile), smem_u32(ltile) + (unsigned)(nrows * pl), smem_u32(ctile),
                           smem_u32(ctile) + (unsigned)(cnrows * pl)};
    uint8_t* const oy = dst + (size_t)(v_base + dv0) * g.dst_pitch + u0;
    uint8_t* const oc = dst + (size_t)(g.out_h + ((v_base + dv0) >> 1)) * g.dst_pitch + u0;
#if VAW_SAMPLER == 3
    if (kNearest) {
        if (pair_ok) rows_quad<false, 1>(g, cp, lnear, cnear, upl, dv0, my_rows, oy, oc, inside, tb);
        else rows_quad<true, 1>(g, cp, lnear, cnear, upl, dv0, my_rows, oy, oc, inside, tb);
        return true;
    }
#endif
    if (pair_ok) rows_quad<false, kMode>(g, cp, lconst, cconst, upl, dv0, my_rows, oy, oc, inside, tb);
    else rows_quad<true, kMode>(g, cp, lconst, cconst, upl, dv0, my_rows, oy, oc, inside, tb);
    return true;  // this piece's loads completed a phase of the tile mbarrier
}

// kCtas = resident CTAs per SM the instantiation is compiled for, i.e. its register budget: 8 (64 registers), 7 (72:
// the C3 tiles let seven CTAs share an SM, and the eight registers the 64-register build gives away buy nothing
// there -- measured 0.651 against 0.661 ms) or 6 (80 registers, no spills in the fallback paths) where shared
// memory allows six or fewer anyway (C5 +2.6 %, C2 +3 %).
template <int kCtas, int kMode = 0>
__global__ void __launch_bounds__(32 * kWarps, kCtas)
warp_nv12_quad_kernel(const Geom g, const FrameBatch b, const PieceRec* __restrict__ table,
                      const __grid_constant__ TileMaps maps)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x, w = threadIdx.y, tid = w * 32 + lane;
    const int px = blockIdx.x, py = blockIdx.y, frame = blockIdx.z;
    const int ph = g.piece_h;
    const int npx = (int)gridDim.x, npy = (int)gridDim.y;
    const PieceRec* rec = table + ((size_t)frame * npy + py) * npx + px;
    // The record comes in by ONE bulk copy into shared memory (240 bytes, its own mbarrier); everything below reads
    // it from there.  Round 2's profile: warps spent 60 % of their lifetime before the row loop, most of it on the
    // ~40 dependent global loads of the record (flags, stage, 36 coefficient loads that missed L1).  The record of
    // the same piece of the NEXT frame is pulled into L2 now: by the time its CTA starts (about one frame's worth
    // of CTAs later) the table entry the builder wrote has long been evicted by the frame data streaming through.
    const unsigned mbar = smem_u32(smem), mbar_rec = mbar + 8, rec_s = mbar + kQuadRecOffset;
    if (tid == 0) {
        mbar_init(mbar, 1);
        mbar_init(mbar_rec, 1);
        mbar_expect_tx(mbar_rec, (unsigned)sizeof(PieceRec));
        bulk_g2s(rec_s, rec, (unsigned)sizeof(PieceRec), mbar_rec);
#if !VAW_PREFETCH_AFTER_BARRIER
        if (frame + 1 < (int)gridDim.z) {
            const char* nxt = reinterpret_cast<const char*>(rec + (size_t)npy * npx);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + 128));
        }
#endif
    }
#if VAW_ONE_WAITER
    // ONE thread polls the mbarrier, the others park at the CTA barrier behind it (round 2's profile: the try_wait
    // loops of all four warps were 3.3 % of the kernel's executed instructions)
    if (tid == 0) mbar_wait_parked(mbar_rec, 0, 4000);
    __syncthreads();
#else
    __syncthreads();  // the barriers are initialised for every warp
#if VAW_PREFETCH_AFTER_BARRIER  // (A/B: the next frame's record prefetch behind the CTA barrier instead of in front of it)
    if (tid == 0 && frame + 1 < (int)gridDim.z) {
        const char* nxt = reinterpret_cast<const char*>(rec + (size_t)npy * npx);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + 128));
    }
#endif
    mbar_wait_parked(mbar_rec, 0, 4000);
#endif
    const PieceRec* rs = reinterpret_cast<const PieceRec*>(smem + kQuadRecOffset);
    quad_piece<kMode>(g, b, rec, rs, maps, smem, px, py, frame, 0u, lane, w, tid);
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

int tile_smem_bytes(int tile_cap) { return kQuadTileOffset + tile_cap; }

// Host mirror of the kernel's tile sizing (same integer arithmetic).
int tile_need_bytes(const PieceRec& rec, int halo)
{
    if (!(rec.flags & kPiecePoly) || (rec.flags & kPieceOutside)) return 0;
    const PieceBox& b = rec.box;
    const int lx0 = b.x0 & ~15, wb = (b.x1 - lx0 + 16) & ~15;
    const int cbx0 = (2 * b.cx0) & ~15, cwb = (2 * b.cx1 + 2 - cbx0 + 15) & ~15;
    const int nrows = (b.y1 - b.y0 + 4) & ~3, cnrows = (b.cy1 - b.cy0 + 4) & ~3;
    const int pl = std::max(kTileMinPitch, (std::max(wb, cwb) + 31) & ~31);
    if (pl > kTileMaxPitch || nrows <= 0 || cnrows <= 0) return 0x7fffffff;
    return pl * (nrows + cnrows) + (halo ? kTileSlack : 0);  // (the box already includes the halo: GeomD::halo)
}

constexpr int kCubicCtas = 4, kLanczosCtas = 4;  // register budget of the table-filter instantiations (128: vaw_api.cu's choose_tile_cap sizes for four CTAs)
template <int kCtas, int kMode = 0>
static cudaError_t configure_quad()
{
    cudaError_t e = cudaFuncSetAttribute(warp_nv12_quad_kernel<kCtas, kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         tile_smem_bytes(kTileCapMax));
    // all of the SM's shared memory for tiles: the taps never go through L1
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(warp_nv12_quad_kernel<kCtas, kMode>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    return e;
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
        cudaError_t e = configure_quad<8>();
        if (e == cudaSuccess) e = configure_quad<7>();
        if (e == cudaSuccess) e = configure_quad<6>();
        if (e == cudaSuccess) e = configure_quad<8, 1>();
        if (e == cudaSuccess) e = configure_quad<kCubicCtas, 2>();
        if (e == cudaSuccess) e = configure_quad<kLanczosCtas, 3>();
        if (e != cudaSuccess) return e;
        if (tracked) configured[dev].store(true, std::memory_order_release);
    }
    dim3 block(32, kWarps);
    dim3 grid(pieces_x(g.out_w), pieces_y(g.out_h, g.piece_h), b.n_frames);
    // the instantiation whose register budget matches the CTAs the tile capacity lets share an SM
    const int smem = tile_smem_bytes(maps.tile_cap);
    // cv::INTER_NEAREST: one light instantiation (64 registers are plenty for one tap per sample)
    // cv::INTER_CUBIC / cv::INTER_LANCZOS4: one instantiation each
    if (g.cubic_tab) {
        // shared memory for the resident CTAs' tiles, the rest of the SM's 256 KB to L1 (the weight table lives there)
        const int smem_h = smem;  // (the slack behind the last tile row is part of the tile capacity: quad_piece, tile_need_bytes)
        int pct = (int)(((long long)std::max(1, maps.table_ctas) * (smem_h + 1024) * 100 + (kSmemPerSM - 1)) / kSmemPerSM);
        if (getenv("VAW_EXPERIMENT_FULL_SMEM")) pct = 100;
        pct = std::min(100, std::max(1, pct));
        cudaError_t e = g.tab_ks == 4
            ? cudaFuncSetAttribute(warp_nv12_quad_kernel<kCubicCtas, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, pct)
            : cudaFuncSetAttribute(warp_nv12_quad_kernel<kLanczosCtas, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        if (e != cudaSuccess) return e;
        if (g.tab_ks == 4) warp_nv12_quad_kernel<kCubicCtas, 2><<<grid, block, smem_h, st>>>(g, b, table, maps);
        else warp_nv12_quad_kernel<kLanczosCtas, 3><<<grid, block, smem_h, st>>>(g, b, table, maps);
    }
    else if (g.nearest) warp_nv12_quad_kernel<8, 1><<<grid, block, smem, st>>>(g, b, table, maps);
    else if (maps.tile_cap <= tile_cap_for_ctas(8, kQuadTileOffset)) warp_nv12_quad_kernel<8><<<grid, block, smem, st>>>(g, b, table, maps);
    else if (maps.tile_cap <= tile_cap_for_ctas(7, kQuadTileOffset)) warp_nv12_quad_kernel<7><<<grid, block, smem, st>>>(g, b, table, maps);
    else warp_nv12_quad_kernel<6><<<grid, block, smem, st>>>(g, b, table, maps);
    return cudaGetLastError();
}

}  // namespace vaw
