// vaw_packed_tile.cu -- fused map + remap for interleaved GRAY8 / BGR24 frames, source tiles staged in shared
// memory by TMA: the quadrant kernel of vaw_tile.cu for the reference's LITERAL frame format.
//
// The reference warps BGR frames: cvtColor(COLOR_YUV2BGR_NV12), then remap on 8UC3
// (/root/reference/opencv/FrameSourceWarp.cpp:399-401 and :306-312, map from createMap.cl).  Round 1 and most
// of round 2 ran that format on the per-pixel-tap kernel (vaw_kernels.cu: coordinates op for op, twelve
// global loads per pixel, 12.7 k frames/s at 4K = 10 % of the roofline for its bytes).  Here it takes the NV12
// path's machinery instead:
//   - coordinates from the per-piece polynomial table (vaw_pieces.cuh; the map is format-independent);
//   - the piece's source box (pixels x0..x1, rows y0..y1 of the record) staged as bytes
//     [kCn x0 & ~15, ...) x rows by cp.async.bulk.tensor over the clip viewed as (pitch / 8, H, frames)
//     8-byte elements (a BGR box is up to 3 x wider than a luma box: 4-byte elements cap a TMA box at 1024 bytes);
//   - 128 x 16-pixel pieces for BGR (a 32-row piece needs ~60 KB of tile: three CTAs per SM);
//   - per sample the round-down-FMA floor and the 16-bit-half blend of vaw_tile.cuh, once per channel off
//     ONE tap address (LDS.U8 [a + c], [a + kCn + c], [a + pitch + c], [a + pitch + kCn + c]).
// Pieces without a polynomial certificate, pieces whose box does not fit the tile and layouts the TMA engine
// cannot address are sampled per pixel from global memory (checked taps), pure-border pieces are a fill.
// Output sizes may be odd here (1759 x 998 is the reference's own C1 case): ragged stores are per byte.
#include <cuda.h>
#include <stdint.h>
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include "vaw_internal.h"
#include "vaw_poly.cuh"
#include "vaw_tile.cuh"

namespace vaw {

namespace {

constexpr int kWarps = 4;
constexpr int kRecOffset = 16, kTileOffset = 256;  // [tile mbarrier | record mbarrier | record (240 B) | tile (128-byte aligned)]

// Tile layout of a piece's source box for kCn interleaved channels (same integer arithmetic on the host:
// packed_tile_need_bytes): rows are `pl` bytes apart (a multiple of 64), start at source byte bx0 (a multiple of
// 16) of row box.y0 and come in whole 4-row boxes.
template <int kCn>
__host__ __device__ inline bool packed_stage(const PieceBox& b, int tile_cap, int& bx0, int& pl, int& nrows)
{
    bx0 = (kCn * (int)b.x0) & ~15;
    const int wb = (kCn * ((int)b.x1 + 1) - bx0 + 15) & ~15;
    pl = (wb + 63) & ~63;
    if (pl < kPackedMinPitch) pl = kPackedMinPitch;
    nrows = ((int)b.y1 - (int)b.y0 + 4) & ~3;
    // (+16: the BGR sampler reads whole words, up to 6 bytes past the last tap of the last row)
    return pl <= kPackedMaxPitch && nrows > 0 && nrows < 4096 && pl * nrows + 16 <= tile_cap;
}

template <int kCn>
__device__ __forceinline__ void store_px_checked(uint8_t* row, int u, int out_w, unsigned v)
{
    if (u >= out_w) return;
#pragma unroll
    for (int c = 0; c < kCn; ++c) row[(size_t)u * kCn + c] = (uint8_t)(v >> (8 * c));
}

template <int kCn, int kMode>
__device__ __forceinline__ unsigned sample_checked(const Geom& g, const uint8_t* __restrict__ src, float mx, float my)
{
    if (kMode >= 2)  // INTER_CUBIC / INTER_LANCZOS4: the context's table filter, per tap from global memory
        return sample_hi<kCn>(src, g.src_pitch, g.src_w, g.src_h, mx, my, kCn == 1 ? (g.border & 255u) : (g.border & 0xffffffu),
                              g.cubic_tab, g.tab_ks);
    if (g.nearest) { mx = nearest_coord(mx); my = nearest_coord(my); }  // INTER_NEAREST = the filter on whole-pixel coordinates
    if (kCn == 1) return (unsigned)sample_c1(src, g.src_pitch, g.src_w, g.src_h, mx, my, g.border & 255);
    return sample_c3(src, g.src_pitch, g.src_w, g.src_h, mx, my, g.border & 0xffffffu);
}

// One sample from the staged tile: B | G << 8 | R << 16 (or Y).  fc.c folds -box.x0 / -box.y0 into the round-down
// FMA, fc.row0 / row1 = tile address + (kCn x0 - bx0) - kCn 2^30 (+ pitch): see luma_tile3 (vaw_tile.cuh).
template <int kCn>
__device__ __forceinline__ unsigned packed_tile_sample(const FloorConst& fc, unsigned pl, float2 m, const TileBounds& tb)
{
    const float2 s = __ffma2_rn(m, pair(fc.scale), pair(kMagic));
    const float2 fl = __ffma2_rd(s, pair(kFloorScale), fc.c);
    const unsigned a = kCn == 1 ? imad_u32(__float_as_uint(fl.y), pl, __float_as_uint(fl.x))
                                : imad_u32(__float_as_uint(fl.x), (unsigned)kCn, __float_as_uint(fl.y) * pl);
    const unsigned a0 = a + fc.row0, a1 = a + fc.row1;
#ifdef VAW_BOUNDS_CHECK
    check_taps(a0, a1, 2u * kCn, tb.l_lo, tb.l_hi);
#endif
    const unsigned ax = __float_as_uint(s.x) & 31u, ay = __float_as_uint(s.y) & 31u;
    const unsigned wx = 32u - ax, wy = imad_u32(ay, 255u, 32u);  // (32 - ay) | ay << 8
    unsigned out = 0;
#pragma unroll
    for (int c = 0; c < kCn; ++c) {
        unsigned t00, t01, t10, t11;
        if (c == 0) { t00 = lds_u8<0>(a0); t01 = lds_u8<kCn>(a0); t10 = lds_u8<0>(a1); t11 = lds_u8<kCn>(a1); }
        else if (c == 1) { t00 = lds_u8<1>(a0); t01 = lds_u8<kCn + 1>(a0); t10 = lds_u8<1>(a1); t11 = lds_u8<kCn + 1>(a1); }
        else { t00 = lds_u8<2>(a0); t01 = lds_u8<kCn + 2>(a0); t10 = lds_u8<2>(a1); t11 = lds_u8<kCn + 2>(a1); }
        const unsigned left = __byte_perm(t00, t10, 0x5410), right = __byte_perm(t01, t11, 0x5410);  // top | bottom << 16
        const unsigned h = imad_u32(right, ax, left * wx);
        const unsigned v = __dp2a_lo(h, wy, 512u) >> 10;
        out |= v << (8 * c);
    }
    return out;
}

// cv::INTER_NEAREST: the sample at (cvRound(x), cvRound(y)) (see luma_tile_nearest, vaw_tile.cuh);
// row = tile address - y0 * pitch - bx0 - 0x4B400000 * (pitch + kCn)
template <int kCn>
__device__ __forceinline__ unsigned packed_tile_nearest(unsigned row, unsigned pl, float2 m)
{
    const float2 s = __fadd2_rn(m, pair(kMagic));
    const unsigned a = (kCn == 1 ? imad_u32(__float_as_uint(s.y), pl, __float_as_uint(s.x))
                                 : imad_u32(__float_as_uint(s.x), (unsigned)kCn, __float_as_uint(s.y) * pl)) + row;
    unsigned out = lds_u8<0>(a);
    if (kCn == 3) out |= (lds_u8<1>(a) << 8) | (lds_u8<2>(a) << 16);
    return out;
}

// The same for BGR with HALF the shared-memory instructions (ncu on the byte-load form: l1tex 92 % busy, mio_throttle the
// top stall, issue slots 45 % busy -- twelve LDS.U8 per sample, ~3 wavefronts each at 6 x magnification bytes between
// lanes).  The six bytes B0 G0 R0 B1 G1 R1 of a tap row come in as three aligned words and two funnel shifts; the two
// rows are widened to 16-bit halves (three PRMT each), blended vertically as halves (six IMAD), regrouped per channel
// (three PRMT) and blended horizontally with the rounding constant by one IDP.2A per channel.  Same integers:
// (t00 wy + t10 ay) wx + (t01 wy + t11 ay) ax + 512 is the exact four-weight sum in either order.
__device__ __forceinline__ unsigned lds_u32(unsigned addr, unsigned row)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr + row));
    return v;
}
__device__ __forceinline__ unsigned bgr_tile_sample(const FloorConst& fc, unsigned pl, float2 m, const TileBounds& tb)
{
    const float2 s = __ffma2_rn(m, pair(fc.scale), pair(kMagic));
    const float2 fl = __ffma2_rd(s, pair(kFloorScale), fc.c);
    const unsigned a0 = imad_u32(__float_as_uint(fl.x), 3u, __float_as_uint(fl.y) * pl) + fc.row0;  // byte address of B0, top row
#ifdef VAW_BOUNDS_CHECK
    check_taps(a0, a0 + pl, 6u, tb.l_lo, tb.l_hi);
#endif
    const unsigned w0 = a0 & ~3u, sh = (a0 & 3u) * 8u;
    const unsigned t0 = lds_u32(w0, 0u), t1 = lds_u32(w0, 4u), t2 = lds_u32(w0, 8u);
    const unsigned b0 = lds_u32(w0, pl), b1 = lds_u32(w0, pl + 4u), b2 = lds_u32(w0, pl + 8u);
    const unsigned lt = __funnelshift_r(t0, t1, sh), ht = __funnelshift_r(t1, t2, sh);  // B0 G0 R0 B1 | G1 R1 . .
    const unsigned lb = __funnelshift_r(b0, b1, sh), hb = __funnelshift_r(b1, b2, sh);
    const unsigned ax = __float_as_uint(s.x) & 31u, ay = __float_as_uint(s.y) & 31u, wy = 32u - ay;
    // vertical blend in 16-bit halves: B0 | G0, R0 | B1, G1 | R1  (each <= 8160)
    const unsigned v0 = imad_u32(__byte_perm(lb, 0u, 0x4140), ay, __byte_perm(lt, 0u, 0x4140) * wy);
    const unsigned v1 = imad_u32(__byte_perm(lb, 0u, 0x4342), ay, __byte_perm(lt, 0u, 0x4342) * wy);
    const unsigned v2 = imad_u32(__byte_perm(hb, 0u, 0x4140), ay, __byte_perm(ht, 0u, 0x4140) * wy);
    const unsigned wxp = imad_u32(ax, 255u, 32u);  // (32 - ax) | ax << 8
    const unsigned b = __dp2a_lo(__byte_perm(v0, v1, 0x7610), wxp, 512u) >> 10;  // left B | right B << 16
    const unsigned g = __dp2a_lo(__byte_perm(v0, v2, 0x5432), wxp, 512u) >> 10;
    const unsigned r = __dp2a_lo(__byte_perm(v1, v2, 0x7610), wxp, 512u) >> 10;
    return b | (g << 8) | (r << 16);
}

// cv::INTER_CUBIC / cv::INTER_LANCZOS4 for BGR from the staged tile (the scheme of luma_tile_hi, vaw_tile.cuh): the
// twelve bytes B0 G0 R0 ... B3 G3 R3 of four taps come in as four aligned words and three funnel shifts, two PRMT per
// channel gather its four taps into one register, two IDP.2A (signed 16-bit weight pairs x unsigned tap bytes)
// accumulate them; a Lanczos4 row is two such groups.  fc carries the block's top-left tap (halo folded into the
// floor constants).
template <int kKs>
__device__ __forceinline__ unsigned bgr_tile_hi(const FloorConst& fc, unsigned pl, float2 m, const int16_t* __restrict__ tab,
                                                const TileBounds& tb)
{
    const float2 s = __ffma2_rn(m, pair(fc.scale), pair(kMagic));
    const float2 fl = __ffma2_rd(s, pair(kFloorScale), fc.c);
    const unsigned a0 = imad_u32(__float_as_uint(fl.x), 3u, __float_as_uint(fl.y) * pl) + fc.row0;  // B of the top-left tap
#ifdef VAW_BOUNDS_CHECK
    check_taps(a0, a0 + (unsigned)(kKs - 1) * pl, 3u * (unsigned)kKs, tb.l_lo, tb.l_hi);
#endif
    const unsigned aw = a0 & ~3u, sh = a0 << 3;
    const unsigned idx = ((__float_as_uint(s.y) & 31u) << 5) | (__float_as_uint(s.x) & 31u);
    const uint4* __restrict__ wt = reinterpret_cast<const uint4*>(tab) + idx * (unsigned)(kKs * kKs / 8);
    int sb = 1 << 14, sg = 1 << 14, sr = 1 << 14;
    unsigned row = aw;
    constexpr int kRowsPerLoad = 16 / kKs;  // 16 weights per 256-bit load: four rows of a 4 x 4 block, two of an 8 x 8 block
#pragma unroll
    for (int r0 = 0; r0 < kKs; r0 += kRowsPerLoad) {
        const Weights8 q = ldg_weights8(wt + 2 * (r0 / kRowsPerLoad));
#pragma unroll
        for (int k = 0; k < kRowsPerLoad; ++k, row += pl) {
#pragma unroll
            for (int grp = 0; grp < kKs / 4; ++grp) {  // four taps = twelve bytes
                const unsigned w0 = lds_w32(row + 12u * grp), w1 = lds_w32(row + 12u * grp + 4u), w2 = lds_w32(row + 12u * grp + 8u),
                               w3 = lds_w32(row + 12u * grp + 12u);
                const unsigned t0 = __funnelshift_r(w0, w1, sh), t1 = __funnelshift_r(w1, w2, sh), t2 = __funnelshift_r(w2, w3, sh);
                // t0 = B0 G0 R0 B1, t1 = G1 R1 B2 G2, t2 = R2 B3 G3 R3
                const unsigned qb = __byte_perm(__byte_perm(t0, t1, 0x0630), t2, 0x5210);
                const unsigned qg = __byte_perm(__byte_perm(t0, t1, 0x0741), t2, 0x6210);
                const unsigned qr = __byte_perm(__byte_perm(t0, t1, 0x0052), t2, 0x7410);
                const unsigned wa = q.w[(kKs / 2) * k + 2 * grp], wb = q.w[(kKs / 2) * k + 2 * grp + 1];
                sb = dp2a_hi_su(wb, qb, dp2a_lo_su(wa, qb, sb));
                sg = dp2a_hi_su(wb, qg, dp2a_lo_su(wa, qg, sg));
                sr = dp2a_hi_su(wb, qr, dp2a_lo_su(wa, qr, sr));
            }
        }
    }
    return sat_u8_q15(sb) | (sat_u8_q15(sg) << 8) | (sat_u8_q15(sr) << 16);
}

// Overwrite the cells of the staged tile that lie outside the source with the border colour (cv::remap's
// BORDER_CONSTANT replaces each out-of-image tap; the colour has period kCn along a row, so the 4-byte word at
// tile word index wd starts with channel (bx0 + 4 wd) mod kCn).  Tile row r <-> source row y0 + r, tile byte c <->
// source byte bx0 + c, valid in [0, n_bytes).  Warps run along the rows, lanes along the words.
template <int kCn>
__device__ __forceinline__ void fill_border_packed(uint8_t* tile, int pl, int tile_rows, int y0, int n_rows, int bx0,
                                                   int n_bytes, unsigned border, int lane, int w)
{
    unsigned pw[3];  // the word that starts with channel 0, 1, 2
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        unsigned v = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) v |= ((border >> (8 * ((r + i) % kCn))) & 255u) << (8 * i);
        pw[r] = v;
    }
    const int wpr = pl >> 2;
    const int r_lo = min(max(-y0, 0), tile_rows), r_hi = min(max(n_rows - y0, r_lo), tile_rows);
    const int c_lo = min(max(-bx0, 0), pl), c_hi = min(max(n_bytes - bx0, c_lo), pl);
    const int w_lo = c_lo >> 2, w_hi = (c_hi + 3) >> 2;  // words [w_lo, w_hi) hold at least one inside byte
    int bm = bx0 % kCn;
    if (bm < 0) bm += kCn;
    const int rot0 = kCn == 1 ? 0 : (bm + lane) % kCn;  // channel of the first byte of word `lane`; 4 = 1 mod 3: +1 per word
    unsigned* words = reinterpret_cast<unsigned*>(tile);
    for (int r = w; r < tile_rows; r += kWarps) {
        const bool row_out = r < r_lo || r >= r_hi;
        unsigned* row = words + r * wpr;
        int rot = rot0;
        for (int wd = lane; wd < wpr; wd += 32) {
            if (row_out || wd < w_lo || wd >= w_hi) row[wd] = pw[rot];
            if (kCn != 1) rot = (rot + 32 % kCn) % kCn;
        }
        if (!row_out && (c_hi & 3) && lane == 0) {  // the inside / outside seam falls into word w_hi - 1
            uint8_t* bytes = tile + r * pl;
            for (int c = c_hi; c < min(4 * w_hi, pl); ++c) bytes[c] = (uint8_t)(border >> (8 * (((bm + c) % kCn + kCn) % kCn)));
        }
    }
}

// nrows rows starting at piece row dv0 for the lane's two columns (u0, u0 + 1); taps from the staged tile.
// kMode: 0 cv::INTER_LINEAR, 1 cv::INTER_NEAREST, 2 cv::INTER_CUBIC, 3 cv::INTER_LANCZOS4
template <int kCn, bool kRagged, int kMode>
__device__ __forceinline__ void rows_packed(const Geom& g, const ColPoly2& cp, const FloorConst& fc, unsigned pl, int dv0,
                                            int nrows, uint8_t* __restrict__ out0, int u0, const TileBounds& tb)
{
    float t = row_t(g, dv0);
    const float dt = g.t_scale, dt2 = __fadd_rn(g.t_scale, g.t_scale);
    const unsigned dpitch = (unsigned)g.dst_pitch + (threadIdx.x >> 5);  // a vector register (see rows_quad)
    const unsigned long long g0 = (unsigned long long)__cvta_generic_to_global(out0);
#if VAW_PACKED_UNROLL2
#pragma unroll 2
#else
#pragma unroll 1
#endif
    for (unsigned j2 = opaque_zero(); j2 < (unsigned)nrows; j2 += 2) {
        const float2 t0 = pair(t), t1 = pair(__fadd_rn(t, dt));
        t = __fadd_rn(t, dt2);
#if VAW_BGR_BYTE_TAPS  // (analysis: the twelve-byte-load form)
        constexpr bool kWords = false;
#else
        constexpr bool kWords = kCn == 3;
#endif
        unsigned v00, v01, v10, v11;
        if (kMode >= 2) {
            const float2 m00 = col_coord(cp.a[0], cp.base, t0), m01 = col_coord(cp.a[1], cp.base, t0);
            const float2 m10 = col_coord(cp.a[0], cp.base, t1), m11 = col_coord(cp.a[1], cp.base, t1);
            if (kCn == 1) {
                constexpr int kKs = kMode == 2 ? 4 : 8;
                v00 = luma_tile_hi<kKs>(fc, pl, m00, g.cubic_tab, tb); v01 = luma_tile_hi<kKs>(fc, pl, m01, g.cubic_tab, tb);
                v10 = luma_tile_hi<kKs>(fc, pl, m10, g.cubic_tab, tb); v11 = luma_tile_hi<kKs>(fc, pl, m11, g.cubic_tab, tb);
            } else {
                constexpr int kKs = kMode == 2 ? 4 : 8;
                v00 = bgr_tile_hi<kKs>(fc, pl, m00, g.cubic_tab, tb); v01 = bgr_tile_hi<kKs>(fc, pl, m01, g.cubic_tab, tb);
                v10 = bgr_tile_hi<kKs>(fc, pl, m10, g.cubic_tab, tb); v11 = bgr_tile_hi<kKs>(fc, pl, m11, g.cubic_tab, tb);
            }
        } else if (kMode == 1) {
            v00 = packed_tile_nearest<kCn>(fc.row0, pl, col_coord(cp.a[0], cp.base, t0));
            v01 = packed_tile_nearest<kCn>(fc.row0, pl, col_coord(cp.a[1], cp.base, t0));
            v10 = packed_tile_nearest<kCn>(fc.row0, pl, col_coord(cp.a[0], cp.base, t1));
            v11 = packed_tile_nearest<kCn>(fc.row0, pl, col_coord(cp.a[1], cp.base, t1));
        } else {
        v00 = kWords ? bgr_tile_sample(fc, pl, col_coord(cp.a[0], cp.base, t0), tb) : packed_tile_sample<kCn>(fc, pl, col_coord(cp.a[0], cp.base, t0), tb);
        v01 = kWords ? bgr_tile_sample(fc, pl, col_coord(cp.a[1], cp.base, t0), tb) : packed_tile_sample<kCn>(fc, pl, col_coord(cp.a[1], cp.base, t0), tb);
        v10 = kWords ? bgr_tile_sample(fc, pl, col_coord(cp.a[0], cp.base, t1), tb) : packed_tile_sample<kCn>(fc, pl, col_coord(cp.a[0], cp.base, t1), tb);
        v11 = kWords ? bgr_tile_sample(fc, pl, col_coord(cp.a[1], cp.base, t1), tb) : packed_tile_sample<kCn>(fc, pl, col_coord(cp.a[1], cp.base, t1), tb);
        }
        const unsigned long long r0 = row_ptr(g0, j2, dpitch), r1 = row_ptr(g0, j2 + 1u, dpitch);
        if (!kRagged) {
            if (kCn == 1) {
                stg_u16(r0, v00 | (v01 << 8));
                stg_u16(r1, v10 | (v11 << 8));
            } else {  // B0 G0 | R0 B1 | G1 R1
                stg_u16(r0, v00); stg_u16(r0 + 2, (v00 >> 16) | (v01 << 8)); stg_u16(r0 + 4, v01 >> 8);
                stg_u16(r1, v10); stg_u16(r1 + 2, (v10 >> 16) | (v11 << 8)); stg_u16(r1 + 4, v11 >> 8);
            }
        } else {
            const bool in0 = u0 < g.out_w, in1 = u0 + 1 < g.out_w, second = j2 + 1u < (unsigned)nrows;
#pragma unroll
            for (int c = 0; c < kCn; ++c) {
                if (in0) stg_u8(r0 + c, v00 >> (8 * c));
                if (in1) stg_u8(r0 + kCn + c, v01 >> (8 * c));
                if (in0 && second) stg_u8(r1 + c, v10 >> (8 * c));
                if (in1 && second) stg_u8(r1 + kCn + c, v11 >> (8 * c));
            }
        }
    }
}

template <int kCn, int kCtas, int kMode = 0>
__global__ void __launch_bounds__(32 * kWarps, kCtas)
warp_packed_tile_kernel(const Geom g, const FrameBatch b, const PieceRec* __restrict__ table,
                        const __grid_constant__ PackedMaps maps)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x, w = threadIdx.y, tid = w * 32 + lane;
    const int px = blockIdx.x, py = blockIdx.y, frame = blockIdx.z;
    const int ph = g.piece_h;
    const int npx = (int)gridDim.x, npy = (int)gridDim.y;
    const PieceRec* rec = table + ((size_t)frame * npy + py) * npx + px;
    const unsigned mbar = smem_u32(smem), mbar_rec = mbar + 8, rec_s = mbar + kRecOffset;
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
    __syncthreads();
    mbar_wait_parked(mbar_rec, 0, 4000);
    const PieceRec* rs = reinterpret_cast<const PieceRec*>(smem + kRecOffset);
    const float4 rec_tail = *(reinterpret_cast<const float4*>(rs) + 12);  // base.x, base.y, flags, pad
    const unsigned flags = __float_as_uint(rec_tail.z);
    const int u_lo = px * kPieceW, v_base = py * ph;
    const int rows = min(ph, g.out_h - v_base);
    const uint8_t* const src = b.src + (size_t)frame * b.src_frame_stride;
    uint8_t* const dst = b.dst + (size_t)frame * b.dst_frame_stride;

    if (flags & kPieceOutside) {  // pure border
        const unsigned border = kCn == 1 ? (g.border & 255u) : (g.border & 0xffffffu);
        if (((reinterpret_cast<uintptr_t>(dst) | (uintptr_t)g.dst_pitch) & 3) == 0 && u_lo + kPieceW <= g.out_w) {
            // 4-byte words: 32 kCn per row, the word at index j starts with channel (4 j) mod kCn = j mod 3
            unsigned pw[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                unsigned v = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) v |= ((border >> (8 * ((r + i) % kCn))) & 255u) << (8 * i);
                pw[r] = v;
            }
            for (int r = w; r < rows; r += kWarps) {
                unsigned* row = reinterpret_cast<unsigned*>(dst + (size_t)(v_base + r) * g.dst_pitch + (size_t)u_lo * kCn);
#pragma unroll
                for (int k = 0; k < kCn; ++k) row[lane + 32 * k] = pw[kCn == 1 ? 0 : (lane + 32 * k) % 3];
            }
            return;
        }
        for (int r = w; r < rows; r += kWarps) {
            uint8_t* row = dst + (size_t)(v_base + r) * g.dst_pitch;
#pragma unroll
            for (int k = 0; k < 4; ++k) store_px_checked<kCn>(row, u_lo + lane + 32 * k, g.out_w, border);
        }
        return;
    }

    // ---- the tile of this piece's source box ---------------------------------------------------------
    const int4 rawbox = *(reinterpret_cast<const int4*>(rs) + 13);  // PieceBox: x0 x1 | y0 y1 | cx0 cx1 | cy0 cy1
    PieceBox box;
    box.x0 = (int16_t)(rawbox.x & 0xffff); box.x1 = (int16_t)(rawbox.x >> 16);
    box.y0 = (int16_t)(rawbox.y & 0xffff); box.y1 = (int16_t)(rawbox.y >> 16);
    int bx0 = 0, pl = 0, nrows = 0;
    const bool staged = (flags & kPiecePoly) && maps.enabled && packed_stage<kCn>(box, maps.tile_cap, bx0, pl, nrows);

    if (!staged) {
        // per pixel from global memory, four columns per lane, warp w walks rows [w PH/4, (w + 1) PH/4)
        const int rpw = ph / kWarps, dv0 = w * rpw, my_rows = max(0, min(rpw, rows - dv0));
        const int u0 = u_lo + 4 * lane;
        if (my_rows <= 0) return;
        ColPoly cp;
        Rot R;
        if (flags & kPiecePoly) derive(rec, lane, cp);
        else R = load_rot(b, frame);
        for (int dv = dv0; dv < dv0 + my_rows; dv += 2) {
            float2 m[2][4];
            if (flags & kPiecePoly) {
                row_coords(cp, row_t(g, dv), m[0]);
                row_coords(cp, row_t(g, dv + 1), m[1]);
            } else {
                exact_rows(g, R, u_lo, u0, v_base + dv, m);
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                if (dv + r >= dv0 + my_rows) break;
                uint8_t* row = dst + (size_t)(v_base + dv + r) * g.dst_pitch;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    store_px_checked<kCn>(row, u0 + i, g.out_w, sample_checked<kCn, kMode>(g, src, m[r][i].x, m[r][i].y));
            }
        }
        return;
    }

    uint8_t* tile = smem + kTileOffset;
    if (tid == 0) {
        mbar_expect_tx(mbar, (unsigned)(pl * nrows));
        const int mi = (pl - kPackedMinPitch) / kPackedPitchStep;
        const CUtensorMap *map16 = &maps.m16[mi], *map4 = &maps.m4[mi];
        const unsigned t0 = smem_u32(tile);
        const int z = frame + b.tma_frame0;
        int k = 0;
        for (; k + 16 <= nrows; k += 16) tma_load_3d(t0 + (unsigned)(k * pl), map16, bx0 >> 3, box.y0 + k, z, mbar);
        for (; k < nrows; k += 4) tma_load_3d(t0 + (unsigned)(k * pl), map4, bx0 >> 3, box.y0 + k, z, mbar);
    }

    // ---- this warp's quadrant; every lane collapses the polynomial onto its two columns ---------------
    const int wx = w & 1, wy = w >> 1, hrows = ph >> 1;
    const int col0 = 64 * wx + 2 * lane;
    ColPoly2 cp;
    derive2(rs, col0, cp);
    cp.base = make_float2(rec_tail.x, rec_tail.y);

    mbar_wait_parked(mbar, 0, 4000);  // the tile has landed

    if (!(flags & kPieceInterior)) {  // straddles the frame border: paint the outside cells
        fill_border_packed<kCn>(tile, pl, nrows, box.y0, g.src_h, bx0, g.src_w * kCn,
                                kCn == 1 ? (g.border & 255u) : (g.border & 0xffffffu), lane, w);
        __syncthreads();
    }

    const int dv0 = wy * hrows;
    const int my_rows = max(0, min(hrows, rows - dv0));
    const unsigned upl = (unsigned)pl;
    // tap address = (iy - y0) pl + kCn (ix - x0) + (kCn x0 - bx0) + tile; x carries kCn * 2^30 out of the floor trick
    // 0 at run time, unknown to ptxas (the zero pad of the record's stage): keeps the tap-row constants in uniform registers
    const unsigned never = (unsigned)(*(reinterpret_cast<const int4*>(rs) + 14)).w >> 31;
    const unsigned vnever = threadIdx.x >> 5;
    constexpr int kHalo = kMode == 2 ? 1 : (kMode == 3 ? 3 : 0);  // = GeomD::halo: the box carries it, the block starts there
    const FloorConst fc = floor_const(-(int)box.x0 - kHalo, -(int)box.y0 - kHalo,
                                      smem_u32(tile) + (unsigned)(kCn * (int)box.x0 - bx0) - (unsigned)kCn * 0x40000000u, upl,
                                      __uint_as_float(0x42000000u | vnever), never);
    const int u0 = u_lo + col0;
    const bool pair_ok = ((reinterpret_cast<uintptr_t>(dst) | (uintptr_t)g.dst_pitch) & 1) == 0 &&
                         u_lo + kPieceW <= g.out_w && (rows & 1) == 0 && (hrows & 1) == 0;
    const TileBounds tb = {smem_u32(tile), smem_u32(tile) + (unsigned)(nrows * pl), 0u, 0u};
    uint8_t* const out0 = dst + (size_t)(v_base + dv0) * g.dst_pitch + (size_t)u0 * kCn;
    if (kMode == 1) {
        const FloorConst fn = floor_const(0, 0, smem_u32(tile) - (unsigned)box.y0 * upl - (unsigned)bx0 - (unsigned)kMagicBits * (upl + (unsigned)kCn),
                                          upl, 0.f, never);
        if (pair_ok) rows_packed<kCn, false, 1>(g, cp, fn, upl, dv0, my_rows, out0, u0, tb);
        else rows_packed<kCn, true, 1>(g, cp, fn, upl, dv0, my_rows, out0, u0, tb);
        return;
    }
    if (pair_ok) rows_packed<kCn, false, kMode>(g, cp, fc, upl, dv0, my_rows, out0, u0, tb);
    else rows_packed<kCn, true, kMode>(g, cp, fc, upl, dv0, my_rows, out0, u0, tb);
}

template <int kCn, int kCtas, int kMode = 0>
cudaError_t configure_packed()
{
    cudaError_t e = cudaFuncSetAttribute(warp_packed_tile_kernel<kCn, kCtas, kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kTileOffset + kTileCapMax);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(warp_packed_tile_kernel<kCn, kCtas, kMode>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    return e;
}
// INTER_CUBIC / INTER_LANCZOS4: shared memory for the resident CTAs' tiles, the rest of the SM to L1 (the weight table)
template <int kCn, int kMode>
cudaError_t launch_packed_table(const Geom& g, const FrameBatch& b, const PieceRec* table, const PackedMaps& maps, dim3 grid,
                                dim3 block, int smem, cudaStream_t st)
{
    int pct = (int)(((long long)std::max(1, maps.table_ctas) * (smem + 1024) * 100 + (kSmemPerSM - 1)) / kSmemPerSM);
    if (getenv("VAW_EXPERIMENT_FULL_SMEM")) pct = 100;
    pct = std::min(100, std::max(1, pct));
    cudaError_t e = cudaFuncSetAttribute(warp_packed_tile_kernel<kCn, 4, kMode>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    if (e != cudaSuccess) return e;
    warp_packed_tile_kernel<kCn, 4, kMode><<<grid, block, smem, st>>>(g, b, table, maps);
    return cudaGetLastError();
}

}  // namespace

int packed_tile_need_bytes(const PieceRec& rec, int channels)
{
    if (!(rec.flags & kPiecePoly) || (rec.flags & kPieceOutside)) return 0;
    int bx0, pl, nrows;
    const bool ok = channels == 1 ? packed_stage<1>(rec.box, 0x7fffffff, bx0, pl, nrows)
                                  : packed_stage<3>(rec.box, 0x7fffffff, bx0, pl, nrows);
    return ok ? pl * nrows : 0x7fffffff;
}

int packed_tile_smem_bytes(int tile_cap) { return kTileOffset + tile_cap; }

long long packed_tile_oob_count()  // this translation unit's own counter of the instrumented build (vaw_tile.cuh)
{
#ifdef VAW_BOUNDS_CHECK
    unsigned long long v = 0;
    if (cudaMemcpyFromSymbol(&v, g_oob_taps, sizeof v) != cudaSuccess) return -2;
    return (long long)v;
#else
    return -1;
#endif
}

cudaError_t launch_warp_packed_tile(const Geom& g, const FrameBatch& b, const PieceRec* table, const PackedMaps& maps,
                                    int channels, cudaStream_t st)
{
    static std::atomic<bool> configured[64];
    int dev = 0;
    cudaGetDevice(&dev);
    const bool tracked = dev >= 0 && dev < 64;
    if (!tracked || !configured[dev].load(std::memory_order_acquire)) {
        cudaError_t e = configure_packed<1, 7>();
        if (e == cudaSuccess) e = configure_packed<1, 6>();
        if (e == cudaSuccess) e = configure_packed<1, 4>();
        if (e == cudaSuccess) e = configure_packed<3, 7>();
        if (e == cudaSuccess) e = configure_packed<3, 6>();
        if (e == cudaSuccess) e = configure_packed<3, 4>();
        if (e == cudaSuccess) e = configure_packed<1, 7, 1>();
        if (e == cudaSuccess) e = configure_packed<3, 7, 1>();
        if (e == cudaSuccess) e = configure_packed<1, 4, 2>();
        if (e == cudaSuccess) e = configure_packed<1, 4, 3>();
        if (e == cudaSuccess) e = configure_packed<3, 4, 2>();
        if (e == cudaSuccess) e = configure_packed<3, 4, 3>();
        if (e != cudaSuccess) return e;
        if (tracked) configured[dev].store(true, std::memory_order_release);
    }
    dim3 block(32, kWarps);
    dim3 grid(pieces_x(g.out_w), pieces_y(g.out_h, g.piece_h), b.n_frames);
    const int smem = kTileOffset + maps.tile_cap;
    // the instantiation whose register budget matches the CTAs the tile capacity lets share an SM: 7 (72 registers),
    // 6 (80) or 4 and fewer (128)
    const int ctas = maps.tile_cap <= tile_cap_for_ctas(7, kTileOffset) ? 7 : (maps.tile_cap <= tile_cap_for_ctas(6, kTileOffset) ? 6 : 4);
    if (g.cubic_tab) {  // cv::INTER_CUBIC / cv::INTER_LANCZOS4
        if (channels == 1) return g.tab_ks == 4 ? launch_packed_table<1, 2>(g, b, table, maps, grid, block, smem, st)
                                                : launch_packed_table<1, 3>(g, b, table, maps, grid, block, smem, st);
        return g.tab_ks == 4 ? launch_packed_table<3, 2>(g, b, table, maps, grid, block, smem, st)
                             : launch_packed_table<3, 3>(g, b, table, maps, grid, block, smem, st);
    }
    if (g.nearest) {  // cv::INTER_NEAREST: one light instantiation per format
        if (channels == 1) warp_packed_tile_kernel<1, 7, 1><<<grid, block, smem, st>>>(g, b, table, maps);
        else warp_packed_tile_kernel<3, 7, 1><<<grid, block, smem, st>>>(g, b, table, maps);
        return cudaGetLastError();
    }
    if (channels == 1) {
        if (ctas == 7) warp_packed_tile_kernel<1, 7><<<grid, block, smem, st>>>(g, b, table, maps);
        else if (ctas == 6) warp_packed_tile_kernel<1, 6><<<grid, block, smem, st>>>(g, b, table, maps);
        else warp_packed_tile_kernel<1, 4><<<grid, block, smem, st>>>(g, b, table, maps);
    } else {
        if (ctas == 7) warp_packed_tile_kernel<3, 7><<<grid, block, smem, st>>>(g, b, table, maps);
        else if (ctas == 6) warp_packed_tile_kernel<3, 6><<<grid, block, smem, st>>>(g, b, table, maps);
        else warp_packed_tile_kernel<3, 4><<<grid, block, smem, st>>>(g, b, table, maps);
    }
    return cudaGetLastError();
}

}  // namespace vaw
