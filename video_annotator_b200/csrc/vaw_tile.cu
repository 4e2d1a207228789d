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
#include <type_traits>
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
constexpr int kQuadRecOffset = 16, kQuadTileOffset = 640;  // [tile mbarrier | record mbarrier | record (240 B) | ... | tile (TMA: 128-byte aligned)]
// the persistent kernel's bookkeeping in the same 640 bytes: [tile mbarrier | 2 record mbarriers | pad | 2 piece slots (int4) | 2 records]
constexpr int kPersistSlotOffset = 32, kPersistRecOffset = 64, kPersistRecStride = 240;
static_assert(kPersistRecOffset + 2 * kPersistRecStride <= kQuadTileOffset, "bookkeeping fits in front of the tile");

struct ColPoly2 {
    float2 a[2][kNv];  // [column][power of t]
    float2 base;
};

// Collapse the piece polynomial onto the lane's two columns: per power of t one Horner chain in s per
// column -- the operation order of collapse_column(), so the coefficients equal derive()'s bit for bit
// (vaw_dump_coords runs derive()).  Coefficients are read per power of t (6 x 8-byte broadcast loads
// from L1) so that at most 12 of the record's 48 coefficient registers are live at a time.
__device__ __forceinline__ void derive2(const PieceRec* __restrict__ rec, int col0, ColPoly2& cp)
{
    const float2* c2 = reinterpret_cast<const float2*>(rec);  // c[i][k] at index i * kNv + k
    const float2 s0 = pair(((float)col0 - 63.5f) * 0.015625f), s1 = pair(((float)(col0 + 1) - 63.5f) * 0.015625f);
#pragma unroll
    for (int k = 0; k < kNv; ++k) {
        float2 ci[kNu];
#pragma unroll
        for (int i = 0; i < kNu; ++i) ci[i] = c2[i * kNv + k];
        float2 a0 = ci[kDegU], a1 = ci[kDegU];
#pragma unroll
        for (int i = kDegU - 1; i >= 0; --i) {
            a0 = __ffma2_rn(a0, s0, ci[i]);
            a1 = __ffma2_rn(a1, s1, ci[i]);
        }
        cp.a[0][k] = a0;
        cp.a[1][k] = a1;
    }
}

__device__ __forceinline__ float2 col_coord(const float2 (&a)[kNv], float2 base, float2 tt)
{
    float2 p = __ffma2_rn(a[3], tt, a[2]);
    p = __ffma2_rn(p, tt, a[1]);
    p = __ffma2_rn(p, tt, a[0]);
    return __fadd2_rn(base, p);  // the map value: rounded once to fp32
}

// base + index * pitch as ONE 64-bit multiply-add (IMAD.WIDE): the compiler's strength-reduced running
// pointers cost four instructions per store here (add, add-with-carry and two moves to re-pair registers).
__device__ __forceinline__ unsigned long long row_ptr(unsigned long long base, unsigned index, unsigned pitch)
{
    unsigned long long a;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(a) : "r"(index), "r"(pitch), "l"(base));
    return a;
}
__device__ __forceinline__ void stg_u16(unsigned long long gaddr, unsigned v)
{
    asm volatile("st.global.u16 [%0], %1;" ::"l"(gaddr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void stg_u8(unsigned long long gaddr, unsigned v)
{
    asm volatile("st.global.u8 [%0], %1;" ::"l"(gaddr), "h"((unsigned short)(v & 255u)) : "memory");
}
// A zero the compiler cannot see through: a loop counter started from it stays in a vector register, so
// that counter * pitch + pointer is one IMAD.WIDE per store instead of uniform-datapath arithmetic plus
// a two-instruction 64-bit vector add.
__device__ __forceinline__ unsigned opaque_zero()
{
    unsigned z;
    asm volatile("mov.u32 %0, 0;" : "=r"(z));
    return z;
}

// nrows (even) rows starting at piece row dv0 for the lane's two columns; taps from the staged tile.
// py / pc: the lane's column pair in the first luma row / chroma row of the band.
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

template <bool kRagged>
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
        const unsigned y00 = luma_tile3(lconst, pl, m00, tb) >> 10, y01 = luma_tile3(lconst, pl, m01, tb) >> 10;
        const unsigned y10 = luma_tile3(lconst, pl, m10, tb) >> 10, y11 = luma_tile3(lconst, pl, m11, tb) >> 10;
        const unsigned c = chroma_tile3(cconst, pl, chroma_z(m00, m01, m10, m11), tb);  // U | V << 8
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

// One piece, from its record in shared memory (`rs`) to the stores: the body shared by the one-piece-per-CTA kernel
// and the persistent kernel.  `rec` = the same record in the table (the gather fallbacks read it from there);
// `tile_parity` = phase of the tile mbarrier (smem + 0) this piece's loads complete.  Returns whether the piece was staged
// (i.e. whether that phase was consumed).
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
    const bool staged = (flags & kPiecePoly) && maps.enabled && pl != 0 && pl * (nrows + cnrows) <= maps.tile_cap;

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
                sample_rows_checked(g, f, u0, v_base + dv, m);
            }
            return false;
        }
        ColPoly cp;
        derive(rec, lane, cp);
        if (flags & kPieceInterior) {
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
                sample_rows_checked(g, f, u0, v_base + dv, m);
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
    // (no early return here: inside the persistent kernel's piece loop a branch on a per-warp value puts everything
    // behind it into a region ptxas treats as divergent, and the tap-row constants below then lose their uniform registers)
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
    const FloorConst lconst = floor_const(-lx0, -by0, smem_u32(ltile) - 0x40000000u, upl, __uint_as_float(0x42000000u | vnever), (unsigned)raw.w >> 31);  // (the stage's zero pad: with g.out_w >> 31 ptxas keeps the luma pair in vector registers)
    const FloorConst cconst = floor_const(-(cbx0 >> 1), -cy0, smem_u32(ctile) - 0x80000000u, upl, __uint_as_float(0x41800000u | vnever), never);
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
    if (pair_ok) rows_quad<false>(g, cp, lconst, cconst, upl, dv0, my_rows, oy, oc, inside, tb);
    else rows_quad<true>(g, cp, lconst, cconst, upl, dv0, my_rows, oy, oc, inside, tb);
    return true;  // this piece's loads completed a phase of the tile mbarrier
}

// kCtas = resident CTAs per SM the instantiation is compiled for, i.e. its register budget: 8 (64 registers), 7 (72:
// the C3 tiles let seven CTAs share an SM, and the eight registers the 64-register build gives away buy nothing
// there -- measured 0.651 against 0.661 ms) or 6 (80 registers, no spills in the fallback paths) where shared
// memory allows six or fewer anyway (C5 +2.6 %, C2 +3 %).
template <int kCtas>
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
        if (frame + 1 < (int)gridDim.z) {
            const char* nxt = reinterpret_cast<const char*>(rec + (size_t)npy * npx);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + 128));
        }
    }
#if VAW_ONE_WAITER
    // ONE thread polls the mbarrier, the others park at the CTA barrier behind it (round 2's profile: the try_wait
    // loops of all four warps were 3.3 % of the kernel's executed instructions)
    if (tid == 0) mbar_wait_parked(mbar_rec, 0, 4000);
    __syncthreads();
#else
    __syncthreads();  // the barriers are initialised for every warp
    mbar_wait_parked(mbar_rec, 0, 4000);
#endif
    const PieceRec* rs = reinterpret_cast<const PieceRec*>(smem + kQuadRecOffset);
    quad_piece(g, b, rec, rs, maps, smem, px, py, frame, 0u, lane, w, tid);
}

// Persistent form of the quadrant kernel: kCtas CTAs per SM stay resident and pull pieces from a queue in table order
// (an atomic counter in the pad word of the table's first record, which the builder zeroes), kPersistChunk pieces per ticket.
// What it removes from the one-piece-per-CTA kernel: the CTA launch between two pieces of a shared-memory slot, the
// entry code of four warps per piece, and the wait for the record -- the record of the NEXT piece is copied into the
// second record buffer while the current piece is sampled.  The tile itself cannot be fetched ahead (one tile per CTA
// is what shared memory holds at 7 CTAs per SM).
#ifndef VAW_PERSIST_ONE_COPY
#define VAW_PERSIST_ONE_COPY 1
#endif
#ifndef VAW_PERSIST_CHUNK
#define VAW_PERSIST_CHUNK 4
#endif
constexpr int kPersistChunk = VAW_PERSIST_CHUNK;

template <int kCtas>
__global__ void __launch_bounds__(32 * kWarps, kCtas)
warp_nv12_quad_persist_kernel(const Geom g, const FrameBatch b, const PieceRec* __restrict__ table,
                              const __grid_constant__ TileMaps maps, const int npx, const int npy)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x, w = threadIdx.y, tid = w * 32 + lane;
    const unsigned sbase = smem_u32(smem);
    const unsigned mbar = sbase, mbar_rec0 = sbase + 8;
    const unsigned per_frame = (unsigned)(npx * npy), total = per_frame * (unsigned)b.n_frames;
    unsigned* const queue = const_cast<unsigned*>(&table[0].pad);
    int4* const slots = reinterpret_cast<int4*>(smem + kPersistSlotOffset);  // written by thread 0, read after the next CTA barrier

    // thread 0's private queue state: pieces [cur, end) of the ticket in hand and the first piece of the next ticket,
    // whose atomic was issued when the ticket in hand was taken up -- kPersistChunk pieces before its result is needed, so
    // the round trip never sits in front of a tile load.  (One piece per ticket serialises the whole launch on the
    // counter: 130 k same-address atomics took 1.04 ms, 8 ns each.)
    unsigned cur = 0, end = 0, ahead = 0xffffffffu;
    auto take_ticket = [&]() -> unsigned { return atomicAdd(queue, (unsigned)kPersistChunk); };
    auto next_piece = [&]() -> unsigned {  // >= total: the queue is drained
        if (cur == end) {
            cur = ahead;
            end = cur < total ? min(cur + (unsigned)kPersistChunk, total) : cur;
            if (cur >= total) return 0xffffffffu;
            ahead = take_ticket();
        }
        return cur++;
    };
    auto post = [&](unsigned q, int buf) {  // thread 0 only: announce piece q in slot `buf` and start its record copy
        if (q >= total) {
            slots[buf] = make_int4(-1, 0, 0, 0);
            return;
        }
        const unsigned frame = q / per_frame, r = q - frame * per_frame;
        const unsigned py = r / (unsigned)npx, px = r - py * (unsigned)npx;
        slots[buf] = make_int4((int)q, (int)px, (int)py, (int)frame);
        const unsigned mb = mbar_rec0 + 8u * (unsigned)buf;
        mbar_expect_tx(mb, (unsigned)sizeof(PieceRec));
        bulk_g2s(sbase + kPersistRecOffset + (unsigned)(buf * kPersistRecStride), table + q, (unsigned)sizeof(PieceRec), mb);
        if (frame + 1 < (unsigned)b.n_frames) {  // the same piece of the next frame: into L2 (see the one-piece kernel)
            const char* nxt = reinterpret_cast<const char*>(table + q + per_frame);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + 128));
        }
    };

    if (tid == 0) {
        mbar_init(mbar, 1);
        mbar_init(mbar_rec0, 1);
        mbar_init(mbar_rec0 + 8, 1);
        ahead = take_ticket();
        post(next_piece(), 0);
    }
    __syncthreads();

    unsigned tile_uses = 0;  // staged pieces so far: phase of the tile mbarrier
    // One piece from record buffer kBuf.  The buffer index is a compile-time constant (the loop below is unrolled by
    // hand over the two buffers): with `it & 1` ptxas no longer proves the record's address -- and with it the tile
    // pitch and the tap-row constants loaded from it -- warp-uniform, and the row loop grows from 271 to 296
    // instructions per two row pairs.
#if VAW_PERSIST_ONE_COPY
    auto step = [&](int kBuf, unsigned it) -> bool {
#else
    auto step = [&](auto buf_c, unsigned it) -> bool {
        constexpr int kBuf = decltype(buf_c)::value;
#endif
        const int4 slot = slots[kBuf];
        if (__any_sync(0xffffffffu, slot.x < 0)) return false;  // queue drained (block-uniform; the vote says so to ptxas)
        if (tid == 0) post(next_piece(), kBuf ^ 1);
        mbar_wait_parked(mbar_rec0 + 8u * (unsigned)kBuf, (it >> 1) & 1u, 4000);
        const PieceRec* rs = reinterpret_cast<const PieceRec*>(smem + kPersistRecOffset + kBuf * kPersistRecStride);
        // block-uniform mirror of quad_piece()'s `staged` decision (its own return value is per-thread as far as ptxas
        // can tell, and a loop-carried per-thread phase costs the row loop its uniform registers)
        const unsigned fl = rs->flags;
        const bool uses_tile = (fl & kPiecePoly) && !(fl & kPieceOutside) && maps.enabled && rs->stage.pl != 0 &&
                               (int)rs->stage.pl * ((int)rs->stage.nrows + (int)rs->stage.cnrows) <= maps.tile_cap &&
                               !(b.skip_interior && (fl & kPieceInterior));
        quad_piece(g, b, table + slot.x, rs, maps, smem, slot.y, slot.z, slot.w, tile_uses & 1u, lane, w, tid);
        tile_uses += uses_tile ? 1u : 0u;
        // every warp is done with the tile, the record and the slot; generic-proxy accesses to the tile (taps, border
        // fill) are ordered before the next piece's asynchronous tile writes
#ifndef VAW_PERSIST_NO_FENCE  // (analysis only)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
        __syncthreads();
        return true;
    };
#if VAW_PERSIST_ONE_COPY
#pragma unroll 1
    for (unsigned it = 0;; ++it)
        if (!step((int)(it & 1u), it)) break;
#else
    for (unsigned it = 0;; it += 2) {
        if (!step(std::integral_constant<int, 0>{}, it)) break;
        if (!step(std::integral_constant<int, 1>{}, it + 1)) break;
    }
#endif
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
int tile_need_bytes(const PieceRec& rec)
{
    if (!(rec.flags & kPiecePoly) || (rec.flags & kPieceOutside)) return 0;
    const PieceBox& b = rec.box;
    const int lx0 = b.x0 & ~15, wb = (b.x1 - lx0 + 16) & ~15;
    const int cbx0 = (2 * b.cx0) & ~15, cwb = (2 * b.cx1 + 2 - cbx0 + 15) & ~15;
    const int nrows = (b.y1 - b.y0 + 4) & ~3, cnrows = (b.cy1 - b.cy0 + 4) & ~3;
    const int pl = std::max(kTileMinPitch, (std::max(wb, cwb) + 31) & ~31);
    if (pl > kTileMaxPitch || nrows <= 0 || cnrows <= 0) return 0x7fffffff;
    return pl * (nrows + cnrows);
}

template <int kCtas>
static cudaError_t configure_quad()
{
    cudaError_t e = cudaFuncSetAttribute(warp_nv12_quad_kernel<kCtas>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         tile_smem_bytes(kTileCapMax));
    // all of the SM's shared memory for tiles: the taps never go through L1
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(warp_nv12_quad_kernel<kCtas>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    return e;
}
template <int kCtas>
static cudaError_t configure_persist()
{
    cudaError_t e = cudaFuncSetAttribute(warp_nv12_quad_persist_kernel<kCtas>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         tile_smem_bytes(kTileCapMax));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(warp_nv12_quad_persist_kernel<kCtas>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
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
        if (e != cudaSuccess) return e;
        if (tracked) configured[dev].store(true, std::memory_order_release);
    }
    dim3 block(32, kWarps);
    dim3 grid(pieces_x(g.out_w), pieces_y(g.out_h, g.piece_h), b.n_frames);
    static const int persist_mode = [] { const char* e = getenv("VAW_PERSIST"); return e ? atoi(e) : 0; }();
    if (persist_mode) {
        static std::atomic<int> sms[64];
        int n_sm = tracked ? sms[dev].load(std::memory_order_acquire) : 0;
        if (n_sm == 0) {
            cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
            cudaError_t e = configure_persist<8>();
            if (e == cudaSuccess) e = configure_persist<7>();
            if (e == cudaSuccess) e = configure_persist<6>();
            if (e != cudaSuccess) return e;
            if (tracked) sms[dev].store(n_sm, std::memory_order_release);
        }
        const int smem = tile_smem_bytes(maps.tile_cap);
        const int npx = (int)grid.x, npy = (int)grid.y;
        const long long pieces = (long long)npx * npy * b.n_frames;
        auto slots = [&](int ctas) { return (unsigned)std::min<long long>((long long)n_sm * ctas, (pieces + kPersistChunk - 1) / kPersistChunk); };
        if (maps.tile_cap <= tile_cap_for_ctas(8, kQuadTileOffset)) warp_nv12_quad_persist_kernel<8><<<slots(8), block, smem, st>>>(g, b, table, maps, npx, npy);
        else if (maps.tile_cap <= tile_cap_for_ctas(7, kQuadTileOffset)) warp_nv12_quad_persist_kernel<7><<<slots(7), block, smem, st>>>(g, b, table, maps, npx, npy);
        else {
            int ctas = 6;
            while (ctas > 1 && maps.tile_cap > tile_cap_for_ctas(ctas, kQuadTileOffset)) --ctas;
            warp_nv12_quad_persist_kernel<6><<<slots(ctas), block, smem, st>>>(g, b, table, maps, npx, npy);
        }
        return cudaGetLastError();
    }
    // the instantiation whose register budget matches the CTAs the tile capacity lets share an SM
    const int smem = tile_smem_bytes(maps.tile_cap);
    if (maps.tile_cap <= tile_cap_for_ctas(8, kQuadTileOffset)) warp_nv12_quad_kernel<8><<<grid, block, smem, st>>>(g, b, table, maps);
    else if (maps.tile_cap <= tile_cap_for_ctas(7, kQuadTileOffset)) warp_nv12_quad_kernel<7><<<grid, block, smem, st>>>(g, b, table, maps);
    else warp_nv12_quad_kernel<6><<<grid, block, smem, st>>>(g, b, table, maps);
    return cudaGetLastError();
}

}  // namespace vaw
