// vaw_bgr.cu -- NV12 in, BGR out in ONE launch: the reference's literal per-frame pipeline
//   cvtColor(frame, bgr, COLOR_YUV2BGR_NV12)                 /root/reference/opencv/FrameSourceWarp.cpp:399-401
//   remap(bgr, out, map_x, map_y, INTER_LINEAR) on 8UC3      :306-312 (the map from createMap.cl)
// fused: no BGR frame and no map ever exist in memory (the reference writes and re-reads both:
// 3 + 3 bytes per source pixel and 8 bytes per output pixel).
//
// Order of operations = the reference's, so the result is bit-exact with converting the whole frame
// first: every bilinear tap is one SOURCE pixel converted with OpenCV's 20-bit BT.601 (vaw_cvt.cuh; its
// chroma is the NV12 pair of the tap's own 2x2 block), taps outside the image take the BGR border value
// (cv::remap's BORDER_CONSTANT acts on the converted image), and the three channels go through
// cv::remap's integer filter (vaw_sample.cuh).  Coordinates come from the per-piece polynomials of
// vaw_pieces.cuh (variant POLY; pieces without a certificate are evaluated op for op), one warp per
// 128 x PH piece, four columns per lane.  Algorithmic bytes: 1.5 per source pixel + 3 per output pixel.
#include <stdint.h>
#include "vaw_internal.h"
#include "vaw_poly.cuh"
#include "vaw_cvt.cuh"

namespace vaw {

namespace {

constexpr int kWarps = 4;

struct Nv12Src {
    const uint8_t* y;
    const uint8_t* uv;
    int pitch, w, h;
};

// the source pixel (x, y) as B | G << 8 | R << 16, or the border when it lies outside
__device__ __forceinline__ unsigned tap_bgr(const Nv12Src& s, int x, int y, unsigned border)
{
    if ((unsigned)x >= (unsigned)s.w || (unsigned)y >= (unsigned)s.h) return border;
    const int yy = __ldg(s.y + (size_t)y * s.pitch + x);
    const unsigned c = __ldg(reinterpret_cast<const uint16_t*>(s.uv + (size_t)(y >> 1) * s.pitch) + (x >> 1));
    unsigned b, g, r;
    yuv_pixel(yy, (int)(c & 255u) - 128, (int)(c >> 8) - 128, b, g, r);
    return b | (g << 8) | (r << 16);
}

// cv::remap INTER_LINEAR / BORDER_CONSTANT of the (virtual) BGR image at (mx, my)
__device__ __forceinline__ unsigned sample_bgr(const Nv12Src& s, float mx, float my, unsigned border)
{
    const int sx = fix5(mx), sy = fix5(my);
    const int ix = sx >> 5, iy = sy >> 5, ax = sx & 31, ay = sy & 31;
    const unsigned t00 = tap_bgr(s, ix, iy, border), t01 = tap_bgr(s, ix + 1, iy, border);
    const unsigned t10 = tap_bgr(s, ix, iy + 1, border), t11 = tap_bgr(s, ix + 1, iy + 1, border);
    // B and R blended together in 16-bit halves (each partial sum <= 255 * 32 * 32 needs 18 bits: two steps
    // with the first one (<= 8160) in halves, the second per channel), G on its own
    const unsigned wx = 32u - (unsigned)ax, wy = 32u - (unsigned)ay;
    const unsigned br0 = (t00 & 0xff00ffu) * wx + (t01 & 0xff00ffu) * (unsigned)ax;  // top row: B | R << 16
    const unsigned br1 = (t10 & 0xff00ffu) * wx + (t11 & 0xff00ffu) * (unsigned)ax;
    const unsigned g0 = ((t00 >> 8) & 255u) * wx + ((t01 >> 8) & 255u) * (unsigned)ax;
    const unsigned g1 = ((t10 >> 8) & 255u) * wx + ((t11 >> 8) & 255u) * (unsigned)ax;
    const unsigned b = ((br0 & 0xffffu) * wy + (br1 & 0xffffu) * (unsigned)ay + 512u) >> 10;
    const unsigned r = ((br0 >> 16) * wy + (br1 >> 16) * (unsigned)ay + 512u) >> 10;
    const unsigned g = (g0 * wy + g1 * (unsigned)ay + 512u) >> 10;
    return b | (g << 8) | (r << 16);
}

__device__ __forceinline__ void store_bgr4(uint8_t* row, const unsigned (&px)[4], int valid, bool aligned)
{
    if (aligned && valid >= 4) {  // 12 bytes = three 4-byte words
        unsigned* o = reinterpret_cast<unsigned*>(row);
        o[0] = px[0] | (px[1] << 24);
        o[1] = (px[1] >> 8) | (px[2] << 16);
        o[2] = (px[2] >> 16) | (px[3] << 8);
    } else {
        for (int i = 0; i < valid && i < 4; ++i) {
            row[3 * i] = (uint8_t)px[i];
            row[3 * i + 1] = (uint8_t)(px[i] >> 8);
            row[3 * i + 2] = (uint8_t)(px[i] >> 16);
        }
    }
}

__global__ void __launch_bounds__(32 * kWarps)
warp_nv12_to_bgr_kernel(const Geom g, const FrameBatch b, const PieceRec* __restrict__ table)
{
    const int lane = threadIdx.x;
    const int px = blockIdx.x, py = blockIdx.y * kWarps + threadIdx.y, frame = blockIdx.z;
    const int ph = g.piece_h;
    const int npx = pieces_x(g.out_w), npy = pieces_y(g.out_h, ph);
    if (py >= npy) return;  // warp-uniform
    const PieceRec* rec = table + ((size_t)frame * npy + py) * npx + px;
    const unsigned flags = __ldg(&rec->flags);
    const int u_lo = px * kPieceW, u0 = u_lo + 4 * lane, v_base = py * ph;
    const int rows = min(ph, g.out_h - v_base);
    const int valid = g.out_w - u0;
    const unsigned border = g.border & 0xffffffu;
    uint8_t* dst = b.dst + (size_t)frame * b.dst_frame_stride;
    const bool aligned = ((reinterpret_cast<uintptr_t>(dst) | (uintptr_t)g.dst_pitch) & 3) == 0;  // 12 u0 is a multiple of 4

    if (flags & kPieceOutside) {  // every tap outside: the border colour
        const unsigned fill[4] = {border, border, border, border};
        if (valid > 0)
            for (int dv = 0; dv < rows; ++dv) store_bgr4(dst + (size_t)(v_base + dv) * g.dst_pitch + (size_t)u0 * 3, fill, valid, aligned);
        return;
    }
    Nv12Src s;
    s.y = b.src + (size_t)frame * b.src_frame_stride;
    s.uv = s.y + (size_t)g.src_pitch * g.src_h;
    s.pitch = g.src_pitch; s.w = g.src_w; s.h = g.src_h;

    ColPoly cp;
    const bool poly = (flags & kPiecePoly) != 0;
    Rot R;
    if (poly) derive(rec, lane, cp);
    else R = load_rot(b, frame);
    for (int dv = 0; dv < rows; dv += 2) {  // row pairs: the per-pixel path evaluates two rows at a time
        float2 m[2][4];
        if (poly) {
            row_coords(cp, row_t(g, dv), m[0]);
            row_coords(cp, row_t(g, dv + 1), m[1]);
        } else {
            exact_rows(g, R, u_lo, u0, v_base + dv, m);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (dv + r >= rows) break;
            unsigned out[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) out[i] = sample_bgr(s, m[r][i].x, m[r][i].y, border);
            if (valid > 0) store_bgr4(dst + (size_t)(v_base + dv + r) * g.dst_pitch + (size_t)u0 * 3, out, valid, aligned);
        }
    }
}

}  // namespace

cudaError_t launch_warp_nv12_to_bgr(const Geom& g, const FrameBatch& b, const PieceRec* table, cudaStream_t st)
{
    dim3 block(32, kWarps);
    dim3 grid(pieces_x(g.out_w), (pieces_y(g.out_h, g.piece_h) + kWarps - 1) / kWarps, b.n_frames);
    warp_nv12_to_bgr_kernel<<<grid, block, 0, st>>>(g, b, table);
    return cudaGetLastError();
}

}  // namespace vaw
