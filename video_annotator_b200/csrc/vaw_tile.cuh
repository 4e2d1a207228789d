// vaw_tile.cuh -- device pieces of the shared-memory variant (vaw_tile.cu: one CTA per piece): the TMA
// tensor load, the samplers that read taps from a staged tile, the border fill.
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

// The same box pulled into L2 only (no shared-memory destination, no completion to wait for).
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int x, int y, int z)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z) : "memory");
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

// ---- chroma sampler, second formulation (round 2) -----------------------------------------------------------------
// The U,V pairs of one row are one register [U0 V0 U1 V1]; the horizontal blend of either channel is one IDP.4A
// against [wx 0 ax 0] / [0 wx 0 ax] -- no byte widening, no two-way dot products and their re-packing: 9 fewer
// ALU-pipe instructions per chroma sample than blend_uv (the 16-lane ALU pipe is the busiest one in the row loop).
// Same integers: horizontal sums (<= 8160), then top * wy + bot * ay + 512, >> 10.
__device__ __forceinline__ unsigned imad_u32(unsigned a, unsigned b, unsigned c)
{
    unsigned d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__device__ __forceinline__ unsigned chroma_tile4a(unsigned cconst, unsigned pl, float2 z, const TileBounds& tb)
{
    const int2 bb = fix_bits(z, 16.0f);
    const unsigned bx = (unsigned)bb.x, by = (unsigned)bb.y;
    const unsigned a0 = (unsigned)(bb.y >> 5) * pl + (((unsigned)(bb.x >> 5) + cconst) << 1);
    const unsigned a1 = a0 + pl;
#ifdef VAW_BOUNDS_CHECK
    check_taps(a0, a1, 4u, tb.c_lo, tb.c_hi);
#endif
    const unsigned t00 = lds_u16<0>(a0), t01 = lds_u16<2>(a0), t10 = lds_u16<0>(a1), t11 = lds_u16<2>(a1);
    const unsigned r0 = imad_u32(t01, 65536u, t00), r1 = imad_u32(t11, 65536u, t10);  // [U0 V0 U1 V1] of either row
    const unsigned ax = bx & 31u, ay = by & 31u;
    const unsigned wu = imad_u32(ax, 65535u, 32u);  // bytes [32 - ax, 0, ax, 0]
    const unsigned wv = wu << 8;                    // bytes [0, 32 - ax, 0, ax]
    const unsigned utop = __dp4a(r0, wu, 0u), vtop = __dp4a(r0, wv, 0u), ubot = __dp4a(r1, wu, 0u), vbot = __dp4a(r1, wv, 0u);
    const unsigned wy = 32u - ay;
    const unsigned u = imad_u32(utop, wy, imad_u32(ubot, ay, 512u)) >> 10;
    const unsigned v = imad_u32(vtop, wy, imad_u32(vbot, ay, 512u)) >> 2;
    return u | (v & 0xff00u);  // U | V << 8
}

// ---- samplers, third formulation (round 2, the row loop of the quadrant kernel) ----------------------------------
// Same integers as luma_tile / chroma_tile, eight fewer instructions per row pair:
//   * floor(n / 32) of BOTH coordinates in one FFMA2 with round-down: s = 32 m + 1.5 * 2^23 holds n = rint(32 m) in
//     its mantissa; s * 2^-27 + c is exact inside the FMA and, rounded toward -inf to the 2^-22 ulp of [2, 4), leaves
//     bits 0x40000000 + floor(n / 32) + D, D = the integer folded into c (minus the tile origin, plus the tile's
//     shared-memory address for x).  0x40000000 * pitch vanishes modulo 2^32 (pitches are multiples of 32), so the
//     tap address is ONE multiply-add of the two bit patterns; the 0x40000000 that x carries is cancelled by the
//     block-uniform `ubase` inside the load's address (register + uniform register).
//   * horizontal blend of both tap rows at once in 16-bit halves, vertical blend + rounding constant as one IDP.2A.
#ifndef VAW_SHFL_UNIFORM
#define VAW_SHFL_UNIFORM 1
#endif
struct FloorConst { float2 c; unsigned row0, row1; float scale; };  // c = 1.90625 + D * 2^-22 per coordinate; row0 / row1 = tile
                                                      // address of the upper / lower tap row minus what x carries
constexpr float kFloorScale = 7.450580596923828125e-09f;  // 2^-27
constexpr float kFloorBias = 1.90625f;                    // 2 - 1.5 * 2^23 * 2^-27
constexpr float kFloorUnit = 2.384185791015625e-07f;      // 2^-22

// A value the compiler front end cannot split or fold (ptxas still sees a plain move): keeps `tile address - 2^30`
// one loop-invariant uniform instead of an immediate added to every tap address.
__device__ __forceinline__ unsigned opaque_u32(unsigned v)
{
    unsigned r;
    asm("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ FloorConst floor_const(int dx, int dy, unsigned row0, unsigned pl, float scale, unsigned never)
{
    FloorConst f;
    f.scale = scale;
    f.c = make_float2(__fmaf_rn((float)dx, kFloorUnit, kFloorBias), __fmaf_rn((float)dy, kFloorUnit, kFloorBias));  // exact
    // `never` is 0 at run time but unknown to ptxas: it cannot take the sums apart again, so each stays ONE uniform
    // register that the loads add for free ([register + uniform register + immediate])
    f.row0 = opaque_u32(row0) | never;
    f.row1 = opaque_u32(row0 + pl) | never;
#if VAW_SHFL_UNIFORM
    // a shuffle from a fixed lane tells ptxas the value is warp-uniform (all lanes hold the same one anyway)
    f.row0 = __shfl_sync(0xffffffffu, f.row0, 0);
    f.row1 = __shfl_sync(0xffffffffu, f.row1, 0);
#endif
    return f;
}

__device__ __forceinline__ unsigned luma_tile3(const FloorConst& fc, unsigned pl, float2 m, const TileBounds& tb)
{
    const float2 s = __ffma2_rn(m, pair(fc.scale), pair(kMagic));
    const float2 fl = __ffma2_rd(s, pair(kFloorScale), fc.c);
    const unsigned a = imad_u32(__float_as_uint(fl.y), pl, __float_as_uint(fl.x));
    const unsigned a0 = a + fc.row0, a1 = a + fc.row1;  // register + uniform register inside the loads' addresses
#ifdef VAW_BOUNDS_CHECK
    check_taps(a0, a1, 2u, tb.l_lo, tb.l_hi);
#endif
    const unsigned t00 = lds_u8<0>(a0), t01 = lds_u8<1>(a0), t10 = lds_u8<0>(a1), t11 = lds_u8<1>(a1);
    const unsigned ax = __float_as_uint(s.x) & 31u, ay = __float_as_uint(s.y) & 31u;
    const unsigned left = __byte_perm(t00, t10, 0x5410), right = __byte_perm(t01, t11, 0x5410);  // top | bottom << 16
    const unsigned h = imad_u32(right, ax, left * (32u - ax));  // both rows, <= 8160 per half
    return __dp2a_lo(h, imad_u32(ay, 255u, 32u), 512u);       // top * (32 - ay) + bottom * ay + 512
}

__device__ __forceinline__ unsigned chroma_tile3(const FloorConst& fc, unsigned pl, float2 z, const TileBounds& tb)
{
    const float2 s = __ffma2_rn(z, pair(fc.scale), pair(kMagic));
    const float2 fl = __ffma2_rd(s, pair(kFloorScale), fc.c);
    const unsigned a = imad_u32(__float_as_uint(fl.x), 2u, __float_as_uint(fl.y) * pl);
    const unsigned a0 = a + fc.row0, a1 = a + fc.row1;
#ifdef VAW_BOUNDS_CHECK
    check_taps(a0, a1, 4u, tb.c_lo, tb.c_hi);
#endif
    return blend_uv(lds_u16<0>(a0), lds_u16<2>(a0), lds_u16<0>(a1), lds_u16<2>(a1), __float_as_uint(s.x) & 31u,
                    __float_as_uint(s.y) & 31u);
}

// ---- cv::INTER_NEAREST from the staged tile (FrameSourceWarp.hpp:90's `interpolation`): cv::remap takes the sample at
// (cvRound(x), cvRound(y)), i.e. round-half-even of the map value itself -- one magic-constant add per coordinate pair,
// one multiply-add for the address, one load.  `row` = tile address - y0 * pitch - x0 - 0x4B400000 * (pitch + step)
// (block-uniform; the bit patterns' exponent part cancels modulo 2^32).  The tap always lies inside the box the tile
// was staged for (floor(m) <= rint(m) <= floor(m) + 1).
__device__ __forceinline__ unsigned luma_tile_nearest(unsigned row, unsigned pl, float2 m)
{
    const float2 s = __fadd2_rn(m, pair(kMagic));
    return lds_u8<0>(imad_u32(__float_as_uint(s.y), pl, __float_as_uint(s.x)) + row);
}
// z = twice the chroma coordinate (chroma_z): the coordinate itself is z / 2 exactly; returns U | V << 8
__device__ __forceinline__ unsigned chroma_tile_nearest(unsigned row, unsigned pl, float2 z)
{
    const float2 s = __fadd2_rn(__fmul2_rn(z, pair(0.5f)), pair(kMagic));
    return lds_u16<0>(imad_u32(__float_as_uint(s.x), 2u, __float_as_uint(s.y) * pl) + row);
}

// ---- cv::INTER_CUBIC / cv::INTER_LANCZOS4 from the staged tile (FrameSourceWarp.hpp:90's `interpolation`) ---------
// cv::remap's kKs x kKs fixed-point filter (vaw_cubic.cuh has the scheme and the weight tables): the coordinate is
// rounded to 1/32 px exactly as for the bilinear filter, the block of taps starts kKs / 2 - 1 samples up and left of
// floor(), the kKs^2 weights (shorts scaled by 2^15, one table entry per pair of 5-bit fractions) come from global
// memory through L1 (16-byte loads), the result is saturate_cast<uchar>((sum + 2^14) >> 15).  The piece's tile was
// staged with that halo (GeomD::halo), border cells painted, so there are no range tests here either.
//   * taps: a row of kKs bytes starts at an arbitrary byte address -> the aligned words around it (kKs / 4 + 1 LDS.32)
//     and one funnel shift per word put the row into registers four taps at a time;
//   * multiply-accumulate: IDP.2A with signed 16-bit weight pairs against unsigned tap bytes (dp2a.lo / .hi): two taps per
//     instruction, the accumulator threaded through.
// The aligned loads read up to 4 bytes past the last tap of a row (into the next tile row, or into the kTileSlack
// bytes behind the tile).
constexpr int kTileSlack = 16;

__device__ __forceinline__ unsigned lds_w32(unsigned addr)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// c + w.lo16 * t.byte0 + w.hi16 * t.byte1 (weights signed, taps unsigned); _hi: bytes 2 and 3
__device__ __forceinline__ int dp2a_lo_su(unsigned w, unsigned t, int c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(t), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_su(unsigned w, unsigned t, int c)
{
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(t), "r"(c));
    return d;
}
// The weights of one table entry: sm_100's 256-bit global load (SASS LDG.E.ENL2.256) fetches 16 shorts in one
// instruction.  The threads of a warp read 32 different entries, so every load instruction costs the L1 a tag
// look-up per thread: halving the instruction count is what counts (VAW_TABLE_LDG256=0: 128-bit loads, for the A/B).
#ifndef VAW_TABLE_LDG256
#define VAW_TABLE_LDG256 1
#endif
struct Weights8 { unsigned w[8]; };  // 16 shorts: two rows of a 4 x 4 block, or one row of an 8 x 8 block
__device__ __forceinline__ Weights8 ldg_weights8(const uint4* __restrict__ p)
{
    Weights8 r;
#if VAW_TABLE_LDG256
    asm("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "l"(p));
#else
    const uint4 a = __ldg(p), b = __ldg(p + 1);
    r.w[0] = a.x; r.w[1] = a.y; r.w[2] = a.z; r.w[3] = a.w; r.w[4] = b.x; r.w[5] = b.y; r.w[6] = b.z; r.w[7] = b.w;
#endif
    return r;
}

__device__ __forceinline__ unsigned sat_u8_q15(int sum) { return (unsigned)min(255, max(0, sum >> 15)); }

// fc: floor constants with the halo folded into the origin (the address is that of the block's top-left tap)
template <int kKs>
__device__ __forceinline__ unsigned luma_tile_hi(const FloorConst& fc, unsigned pl, float2 m, const int16_t* __restrict__ tab,
                                                 const TileBounds& tb)
{
    const float2 s = __ffma2_rn(m, pair(fc.scale), pair(kMagic));
    const float2 fl = __ffma2_rd(s, pair(kFloorScale), fc.c);
    const unsigned a0 = imad_u32(__float_as_uint(fl.y), pl, __float_as_uint(fl.x)) + fc.row0;
#ifdef VAW_BOUNDS_CHECK
    check_taps(a0, a0 + (unsigned)(kKs - 1) * pl, (unsigned)kKs, tb.l_lo, tb.l_hi);
#endif
    const unsigned aw = a0 & ~3u, sh = a0 << 3;  // the funnel shift takes its count modulo 32: 8 (a0 & 3)
    const unsigned idx = ((__float_as_uint(s.y) & 31u) << 5) | (__float_as_uint(s.x) & 31u);
    const uint4* __restrict__ wt = reinterpret_cast<const uint4*>(tab) + idx * (unsigned)(kKs * kKs / 8);
    int sum = 1 << 14;
    unsigned row = aw;
    if (kKs == 4) {
        const Weights8 q0 = ldg_weights8(wt);  // the whole 4 x 4 block: 16 shorts
        const unsigned wp[8] = {q0.w[0], q0.w[1], q0.w[2], q0.w[3], q0.w[4], q0.w[5], q0.w[6], q0.w[7]};
#pragma unroll
        for (int r = 0; r < 4; ++r, row += pl) {
            const unsigned t = __funnelshift_r(lds_w32(row), lds_w32(row + 4u), sh);
            sum = dp2a_hi_su(wp[2 * r + 1], t, dp2a_lo_su(wp[2 * r], t, sum));
        }
    } else {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const Weights8 q = ldg_weights8(wt + 2 * rr);  // rows 2 rr, 2 rr + 1
#pragma unroll
            for (int k = 0; k < 2; ++k, row += pl) {
                const unsigned w0 = lds_w32(row), w1 = lds_w32(row + 4u), w2 = lds_w32(row + 8u);
                const unsigned t0 = __funnelshift_r(w0, w1, sh), t1 = __funnelshift_r(w1, w2, sh);
                sum = dp2a_hi_su(q.w[4 * k + 1], t0, dp2a_lo_su(q.w[4 * k], t0, sum));
                sum = dp2a_hi_su(q.w[4 * k + 3], t1, dp2a_lo_su(q.w[4 * k + 2], t1, sum));
            }
        }
    }
    return sat_u8_q15(sum);
}

// z = twice the chroma coordinate (chroma_z); a row of the block is kKs (U, V) pairs; returns U | V << 8
template <int kKs>
__device__ __forceinline__ unsigned chroma_tile_hi(const FloorConst& fc, unsigned pl, float2 z, const int16_t* __restrict__ tab,
                                                   const TileBounds& tb)
{
    const float2 s = __ffma2_rn(z, pair(fc.scale), pair(kMagic));
    const float2 fl = __ffma2_rd(s, pair(kFloorScale), fc.c);
    const unsigned a0 = imad_u32(__float_as_uint(fl.x), 2u, __float_as_uint(fl.y) * pl) + fc.row0;  // even
#ifdef VAW_BOUNDS_CHECK
    check_taps(a0, a0 + (unsigned)(kKs - 1) * pl, 2u * (unsigned)kKs, tb.c_lo, tb.c_hi);
#endif
    const unsigned aw = a0 & ~3u, sh = a0 << 3;  // 0 or 16
    const unsigned idx = ((__float_as_uint(s.y) & 31u) << 5) | (__float_as_uint(s.x) & 31u);
    const uint4* __restrict__ wt = reinterpret_cast<const uint4*>(tab) + idx * (unsigned)(kKs * kKs / 8);
    int su = 1 << 14, sv = 1 << 14;
    unsigned row = aw;
    if (kKs == 4) {
        const Weights8 q0 = ldg_weights8(wt);
        const unsigned wp[8] = {q0.w[0], q0.w[1], q0.w[2], q0.w[3], q0.w[4], q0.w[5], q0.w[6], q0.w[7]};
#pragma unroll
        for (int r = 0; r < 4; ++r, row += pl) {
            const unsigned w0 = lds_w32(row), w1 = lds_w32(row + 4u), w2 = lds_w32(row + 8u);
            const unsigned t0 = __funnelshift_r(w0, w1, sh), t1 = __funnelshift_r(w1, w2, sh);  // U0 V0 U1 V1 | U2 V2 U3 V3
            const unsigned u = __byte_perm(t0, t1, 0x6420), v = __byte_perm(t0, t1, 0x7531);
            su = dp2a_hi_su(wp[2 * r + 1], u, dp2a_lo_su(wp[2 * r], u, su));
            sv = dp2a_hi_su(wp[2 * r + 1], v, dp2a_lo_su(wp[2 * r], v, sv));
        }
    } else {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const Weights8 q = ldg_weights8(wt + 2 * rr);  // rows 2 rr, 2 rr + 1
#pragma unroll
            for (int k = 0; k < 2; ++k, row += pl) {
                const unsigned w0 = lds_w32(row), w1 = lds_w32(row + 4u), w2 = lds_w32(row + 8u), w3 = lds_w32(row + 12u),
                               w4 = lds_w32(row + 16u);
                const unsigned t0 = __funnelshift_r(w0, w1, sh), t1 = __funnelshift_r(w1, w2, sh);
                const unsigned t2 = __funnelshift_r(w2, w3, sh), t3 = __funnelshift_r(w3, w4, sh);
                const unsigned u0 = __byte_perm(t0, t1, 0x6420), v0 = __byte_perm(t0, t1, 0x7531);
                const unsigned u1 = __byte_perm(t2, t3, 0x6420), v1 = __byte_perm(t2, t3, 0x7531);
                su = dp2a_hi_su(q.w[4 * k + 1], u0, dp2a_lo_su(q.w[4 * k], u0, su));
                su = dp2a_hi_su(q.w[4 * k + 3], u1, dp2a_lo_su(q.w[4 * k + 2], u1, su));
                sv = dp2a_hi_su(q.w[4 * k + 1], v0, dp2a_lo_su(q.w[4 * k], v0, sv));
                sv = dp2a_hi_su(q.w[4 * k + 3], v1, dp2a_lo_su(q.w[4 * k + 2], v1, sv));
            }
        }
    }
    return sat_u8_q15(su) | (sat_u8_q15(sv) << 8);
}

// ---- pieces shared by the quadrant kernels (vaw_tile.cu: NV12; vaw_packed_tile.cu: GRAY8 / BGR24) ----------------------
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

// Pair lane mapping of the texture variant (vaw_tex.cu): lane l owns luma columns 2l, 2l+1 and 64+2l, 64+2l+1
// of the piece (slot j -> column 2l + (j & 1) + 64 (j >> 1)); stores are 2 bytes per lane.
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
