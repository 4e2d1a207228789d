// vaw_tile.cuh -- device pieces shared by the shared-memory variants (vaw_tile.cu: one CTA per
// piece; vaw_pipe.cu: persistent producer/consumer pipeline): the TMA tensor load, the samplers
// that read taps from a staged tile, the pair lane mapping, the border fill.
// Same arithmetic as vaw_poly.cuh: cv::remap's integer filter
// (/root/reference/opencv/FrameSourceWarp.cpp:306-312) on the map of vaw_pieces.cuh.
#pragma once
#include <cuda.h>
#include <stdint.h>
#include "vaw_poly.cuh"

namespace vaw {

__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, int x, int y, int z, unsigned mbar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(mbar)
        : "memory");
}

// Instrumented build (-DVAW_BOUNDS_CHECK, tests only: compute-sanitizer is not available on the GPU
// pool): every tap address of the staged samplers is checked against the plane's tile and counted
// when it falls outside; vaw_debug_oob_count() returns the count (-1 when not instrumented).
struct TileBounds { unsigned l_lo, l_hi, c_lo, c_hi; };  // shared-memory byte ranges [lo, hi)
#ifdef VAW_BOUNDS_CHECK
static __device__ unsigned long long g_oob_taps;
__device__ __forceinline__ void check_taps(unsigned a0, unsigned a1, unsigned width, unsigned lo, unsigned hi)
{
    if (a0 < lo || a0 + width > hi || a1 < lo || a1 + width > hi) atomicAdd(&g_oob_taps, 1ull);
}
#endif

__device__ __forceinline__ int luma_tile(unsigned lconst, unsigned pl, float2 m, const TileBounds& tb)
{
    const int2 bb = fix_bits(m, 32.0f);
    const int bx = bb.x, by = bb.y;
    const unsigned a0 = (unsigned)(by >> 5) * pl + ((unsigned)(bx >> 5) + lconst);
    const unsigned a1 = a0 + pl;
#ifdef VAW_BOUNDS_CHECK
    check_taps(a0, a1, 2u, tb.l_lo, tb.l_hi);
#endif
#ifdef VAW_ABL_NO_LDS  // analysis only: taps derived from the address instead of loaded
    return blend_y(a0 & 255, (a0 >> 1) & 255, a1 & 255, (a1 >> 3) & 255, bx & 31, by & 31);
#else
    return blend_y(lds_u8<0>(a0), lds_u8<1>(a0), lds_u8<0>(a1), lds_u8<1>(a1), bx & 31, by & 31);
#endif
}

__device__ __forceinline__ unsigned chroma_tile(unsigned cconst, unsigned pl, float2 z, const TileBounds& tb)
{
    const int2 bb = fix_bits(z, 16.0f);
    const int bx = bb.x, by = bb.y;
    const unsigned a0 = (unsigned)(by >> 5) * pl + (((unsigned)(bx >> 5) + cconst) << 1);
    const unsigned a1 = a0 + pl;
#ifdef VAW_BOUNDS_CHECK
    check_taps(a0, a1, 4u, tb.c_lo, tb.c_hi);
#endif
#ifdef VAW_ABL_NO_LDS
    return blend_uv(a0 & 0xffff, (a0 >> 1) & 0xffff, a1 & 0xffff, (a1 >> 3) & 0xffff, bx & 31, by & 31);
#else
    return blend_uv(lds_u16<0>(a0), lds_u16<2>(a0), lds_u16<0>(a1), lds_u16<2>(a1), bx & 31, by & 31);
#endif
}

// Lane -> column mapping of the staged path: lane l owns luma columns 2l, 2l+1 and 64+2l, 64+2l+1
// of the piece (slot j -> column 2l + (j & 1) + 64 (j >> 1)).  One LDS instruction then serves 32
// pixels that are 2 columns apart: at the C3 centre (1.84 source px per output px) its addresses
// span 118 bytes = 30 banks, i.e. one shared-memory wavefront.  With four consecutive columns per
// lane the same instruction spans 236 bytes and needs two or more (measured 2.5 wavefronts per
// LDS, shared-memory pipe 72 % busy).  Quads (j = 0,1 and j = 2,3) stay inside a lane, so the NV12
// chroma rule needs no shuffles; stores become 2 bytes per lane (64 contiguous bytes per warp).
__device__ __forceinline__ int pair_column(int lane, int j) { return 2 * lane + (j & 1) + 64 * (j >> 1); }

template <bool kRagged>
__device__ __forceinline__ void store_pair(uint8_t* p, unsigned lo, unsigned hi, bool inside)
{
    if (!kRagged) {
        *reinterpret_cast<uint16_t*>(p) = (uint16_t)(lo | (hi << 8));
    } else if (inside) {
        p[0] = (uint8_t)lo;
        p[1] = (uint8_t)hi;
    }
}

// nrows (even) rows starting at piece row dv0, taps from the staged tile.  o.y0 / o.y1 / o.c point
// at column 2*lane of the piece.
template <bool kRagged>
__device__ __forceinline__ void rows_tile(const Geom& g, const ColPoly& cp, unsigned lconst, unsigned cconst,
                                          unsigned pl, int dv0, int nrows, RowPtrs& o, bool in_a, bool in_b,
                                          const TileBounds& tb)
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
        unsigned y[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) y[r][i] = (unsigned)luma_tile(lconst, pl, m[r][i], tb) >> 10;
        unsigned c[2];
#pragma unroll
        for (int q = 0; q < 2; ++q)
#ifdef VAW_ABL_NO_CHROMA
            c[q] = y[0][q] | (y[1][q] << 8);
#else
            c[q] = chroma_tile(cconst, pl, chroma_z(m[0][2 * q], m[0][2 * q + 1], m[1][2 * q], m[1][2 * q + 1]), tb);
#endif
#ifdef VAW_ABL_NO_STORE  // analysis only: results stay live, (almost) nothing is written
        if ((y[0][0] ^ y[0][1] ^ y[0][2] ^ y[0][3] ^ y[1][0] ^ y[1][1] ^ y[1][2] ^ y[1][3] ^ c[0] ^ c[1]) == 0x12345u)
#endif
        {
        store_pair<kRagged>(o.y0, y[0][0], y[0][1], in_a);
        store_pair<kRagged>(o.y0 + 64, y[0][2], y[0][3], in_b);
        store_pair<kRagged>(o.y1, y[1][0], y[1][1], in_a);
        store_pair<kRagged>(o.y1 + 64, y[1][2], y[1][3], in_b);
        store_pair<kRagged>(o.c, c[0] & 255u, c[0] >> 8, in_a);
        store_pair<kRagged>(o.c + 64, c[1] & 255u, c[1] >> 8, in_b);
        }
        o.y0 += o.step_y; o.y1 += o.step_y; o.c += o.step_c;
    }
}

// Overwrite the cells of a staged plane that lie outside the source with the border value.
// Tile row r <-> source row y0 + r (valid in [0, n_rows)); tile byte c <-> source byte x0 + c
// (valid in [0, n_bytes)); `pattern` = the border replicated over 4 bytes (x0 is a multiple of 4).
// Only the outside rows and the outside column strips are visited.
__device__ __forceinline__ void fill_border(uint8_t* tile, int pl, int tile_rows, int y0, int n_rows, int x0,
                                            int n_bytes, unsigned pattern, int tid, int nthr)
{
    const int wpr = pl >> 2;                                  // words per tile row
    const int r_lo = min(max(-y0, 0), tile_rows);             // rows [0, r_lo) are above the plane
    const int r_hi = min(max(n_rows - y0, r_lo), tile_rows);  // rows [r_hi, tile_rows) are below it
    unsigned* w = reinterpret_cast<unsigned*>(tile);
    for (int idx = tid; idx < r_lo * wpr; idx += nthr) w[idx] = pattern;
    for (int idx = r_hi * wpr + tid; idx < tile_rows * wpr; idx += nthr) w[idx] = pattern;
    const int c_lo = min(max(-x0, 0), pl);                    // bytes [0, c_lo) are left of the plane (a multiple of 4)
    const int c_hi = min(max(n_bytes - x0, c_lo), pl);        // bytes [c_hi, pl) are right of it
    if (c_lo == 0 && c_hi == pl) return;
    // whole outside words of an inside row: [0, w_lo) and [w_hi, wpr); then the bytes [c_hi, 4 w_hi) of a
    // plane whose width is not a multiple of 4.  No division: lanes run along the words, warps along the
    // rows -- or one row per thread when the strips are only a few words wide.
    const int w_lo = c_lo >> 2, w_hi = min((c_hi + 3) >> 2, wpr);
    const int nw = w_lo + (wpr - w_hi);
    if (nw > 8) {
        const int sub = tid & 31, nwarp = nthr >> 5;
        for (int r = r_lo + (tid >> 5); r < r_hi; r += nwarp) {
            unsigned* row = w + r * wpr;
            for (int j = sub; j < nw; j += 32) row[j < w_lo ? j : w_hi + (j - w_lo)] = pattern;
        }
    } else if (nw > 0) {
        for (int r = r_lo + tid; r < r_hi; r += nthr) {
            unsigned* row = w + r * wpr;
            for (int j = 0; j < nw; ++j) row[j < w_lo ? j : w_hi + (j - w_lo)] = pattern;
        }
    }
    if (c_hi & 3) {
        const int end = min(4 * w_hi, pl);
        for (int r = r_lo + tid; r < r_hi; r += nthr)
            for (int c = c_hi; c < end; ++c) tile[r * pl + c] = (uint8_t)(pattern >> (8 * (c & 3)));
    }
}


}  // namespace vaw
