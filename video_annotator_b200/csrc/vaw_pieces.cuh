// vaw_pieces.cuh -- piecewise-polynomial form of the createMap projection.
//
// Why it exists: on a B200 the fused warp is not bound by HBM but by instruction issue.
// At 70 % of the measured copy bandwidth there are ~24 issue slots per output pixel, and
// the op-for-op evaluation of /root/reference/opencv/createMap.cl:15-49 (3 divides, sqrt,
// atan) costs ~50 of them.  The map is smooth, so it is evaluated ONCE per piece of
// 128 x PH output pixels (PH = 32, 16 or 8 rows, chosen from the focal lengths so that the
// truncation error stays below ~1e-5 px) on a sparse anchor grid in double precision and
// interpolated inside the piece by a tensor polynomial (degree 5 along u, 3 along v).
// Per pixel that leaves 3 FMAs per coordinate.
//
// Accuracy contract (tests/test_gpu_parity.py): a piece takes the polynomial path only
// if (a) the projection is regular over it (q.z bounded away from 0, optical axis not
// inside -- the reference yields NaN at r == 0, createMap.cl:38-39), and (b) the
// polynomial agrees with the exact projection at interior check points to 5e-5 px
// (3e-6 .. 4e-5 px in practice: it is the interpolation error of the end intervals).
// Everything else is evaluated per pixel with the op-for-op sequence (vaw_coords.cuh).
// The coordinate the sampler uses is always the fp32 value base + offset, rounded once;
// vaw_dump_coords returns exactly those values, so "cv::remap on the same map" is
// well defined for either path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vaw {

constexpr int kPieceW = 128;   // output pixels per piece along u (= one warp row: 32 lanes x 4)
constexpr int kPieceHMax = 32; // rows per piece: 32, 16 or 8
constexpr int kDegU = 5;       // polynomial degree along u  (6 anchors, spacing 25.6 px)
constexpr int kDegV = 3;       // polynomial degree along v  (4 anchors, spacing PH/3 rows)
constexpr int kNu = kDegU + 1, kNv = kDegV + 1;
constexpr int kStageMinPitch = 128, kStageMaxPitch = 448;  // tile row pitches (= kTileMinPitch / kTileMaxPitch, vaw_internal.h)

enum : uint32_t {
    kPiecePoly = 1u,      // polynomial coordinates certified for this piece
    kPieceInterior = 2u,  // every tap of every pixel (luma and chroma) lies inside the source
    kPieceOutside = 4u    // every tap of every pixel lies outside: the piece is pure border
};

// Source rectangle (inclusive, in samples of the plane) that the taps of one piece touch; it may
// reach outside the frame for pieces that straddle the border.
struct PieceBox {
    int16_t x0, x1, y0, y1;      // luma
    int16_t cx0, cx1, cy0, cy1;  // chroma (U,V pairs)
};

// How the samplers stage the piece's source box in shared memory (same integer arithmetic for every
// variant, done once here instead of by every thread of the sampler): tile rows are `pl` bytes apart,
// start at source byte lx0 / cbx0 (multiples of 16) and row by0 / cy0, and come in whole 4-row boxes
// (loaded as 32-row, then 8-row, then 4-row TMA boxes).
// The pitch is a multiple of 128 bytes (32 banks) when that fits the tile capacity -- the lanes of one
// LDS then keep distinct banks however many source rows they straddle -- else the tightest multiple of 32.
struct PieceStage {
    int16_t lx0, by0, cbx0, cy0;
    uint16_t pl;          // 0: the box cannot be staged (wider than the largest tile pitch)
    uint16_t nrows, cnrows;   // luma / chroma tile rows
    uint16_t pad;
};

// One record per (frame, piece): 240 bytes, 16-byte aligned.
// coordinate = base + sum_{i<=5, j<=3} c[i][j] * s^i * t^j,
//   s = (du - 63.5) / 64, t = (dv - (PH-1)/2) * 2/PH, (du, dv) = pixel offset inside the piece.
// x and y coefficients are interleaved so that both coordinates run through the packed
// FFMA2 / FADD2 instructions of sm_100 (one issue slot for two fp32 operations).
struct PieceRec {
    float2 c[kNu][kNv];    // (x, y) coefficient of s^i t^j
    float2 base;           // integers: base + offset rounds once to the fp32 coordinate
    uint32_t flags;
    uint32_t pad;
    PieceBox box;      // valid for certified pieces that are not pure border
    PieceStage stage;  // derived from `box`
};
static_assert(sizeof(PieceRec) == 240, "piece record layout");

// Lagrange -> monomial conversion matrices for the anchor nodes (computed on the host in
// double precision, passed by value to the builder kernel).
struct PieceBasis {
    double mu[kNu][kNu];  // monomial coefficient i = sum_a mu[i][a] * f(node a)
    double mv[kNv][kNv];
};

struct GeomD {  // the 8 fp32 scalars of FrameSourceWarp.cpp:283-290, widened exactly
    double scx, scy, sfx, sfy, mcx, mcy, mfx, mfy;
    double inv_mfx, inv_mfy;
    double kd[4];  // cv::fisheye distortion k1..k4 of the input camera (extension; zeros = createMap.cl)
    int src_w, src_h, out_w, out_h;
    int piece_h;
    int has_dist;  // any kd != 0
    int tile_cap;  // bytes of shared memory a tile may take with a 128-byte-multiple pitch (0: always the tightest pitch)
    int projection;  // 0 = createMap.cl (fisheye in, rectilinear out); bit 0: rectilinear input; bit 1: fisheye output
    int pitch64;  // tile pitch policy: 0 = a multiple of 128 bytes when the tile capacity allows, else the tightest multiple of 32;
                  // 1 = the smallest pitch that is 64 modulo 128 (adjacent tile rows sit 16 banks apart: lanes that share a
                  // source column in neighbouring rows no longer collide)
    int halo;  // extra taps on every side of the bilinear pair that the staged sampler reads: 0 (INTER_LINEAR / INTER_NEAREST),
               // 1 (INTER_CUBIC: 4 x 4 taps from floor - 1) or 3 (INTER_LANCZOS4: 8 x 8 taps from floor - 3); the source boxes
               // and the interior / outside classification of the pieces account for it
};

inline __host__ __device__ int pieces_x(int out_w) { return (out_w + kPieceW - 1) / kPieceW; }
inline __host__ __device__ int pieces_y(int out_h, int ph) { return (out_h + ph - 1) / ph; }

cudaError_t launch_build_pieces(const GeomD& g, const PieceBasis& basis, const float* rots,
                                const float* rot0, int n_frames, PieceRec* table, cudaStream_t st);

}  // namespace vaw
