// vaw_poly.cu -- fused map + remap for NV12 with coordinates from the per-piece polynomials.
//   variant POLY : taps gathered from global memory through L1/L2;
//   variant TILED: the source rectangle of every 8-row band is first copied into shared
//                  memory by the TMA engine (cp.async.bulk, one row per lane, completion on a
//                  per-warp mbarrier) and the taps are read from there.
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
// op-for-op evaluation (vaw_coords.cuh).
//
// Why shared memory when L1 already hits 94 %: instructions, not bytes.  A tap address in
// global memory is 64-bit (5 extra integer instructions per pixel) and the LSU accepts one
// global load per 1.8 cycles per SM; from shared memory the four taps are LDS [a], [a+1],
// [a+PL], [a+PL+1] off one 32-bit IMAD, and the staging itself costs no issue slots
// because the TMA engine does it.  Each warp is its own pipeline (no CTA barrier): it
// issues the row copies of a band, waits on its mbarrier, samples 8 rows, repeats; the
// other warps of the SM cover the wait.
#include <stdint.h>
#include "vaw_internal.h"
#include "vaw_poly.cuh"

namespace vaw {

namespace {

constexpr int kWarps = 4;  // warps per CTA, stacked vertically (4 pieces)
// per-warp staging space (variant TILED)
constexpr int kLumaCap = 7680, kChromaCap = 3840;
constexpr int kWarpSmem = 16 + kLumaCap + kChromaCap;  // mbarrier + luma + chroma

__device__ __forceinline__ Rot load_rot(const FrameBatch& b, int frame)
{
    if (b.rots == nullptr) return b.rot0;
    Rot R;
    const float* p = b.rots + (size_t)frame * 9;
#pragma unroll
    for (int i = 0; i < 9; ++i) R.r[i] = __ldg(p + i);
    return R;
}

struct RowPtrs {  // output pointers of the lane's 4-pixel group for the current row pair
    uint8_t *y0, *y1, *c;
    size_t step_y, step_c;  // advance per row pair
};

// nrows (even) rows starting at piece row dv0, taps from global memory, no border tests.
template <bool kRagged>
__device__ __forceinline__ void band_gmem(const Geom& g, const ColPoly& cp, const PlaneRefs& f, int dv0, int nrows,
                                          RowPtrs& o, int valid)
{
    const unsigned pitch = (unsigned)g.src_pitch;
    const unsigned bias_y = 0u - kMagicShift * pitch - kMagicShift;         // offset = iy*pitch + ix
    const unsigned bias_c = 0u - kMagicShift * (pitch >> 1) - kMagicShift;  // offset = iy*pitch + 2*(ix + bias)
#pragma unroll 1
    for (int dv = dv0; dv < dv0 + nrows; dv += 2) {
        float mx[2][4], my[2][4];
        row_coords(cp, row_t(g, dv), mx[0], my[0]);
        row_coords(cp, row_t(g, dv + 1), mx[1], my[1]);
        int acc[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[r][i] = luma_gmem(f.y, pitch, bias_y, mx[r][i], my[r][i]);
        unsigned cw = 0u;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float zx = chroma_z(mx[0][2 * q], mx[0][2 * q + 1], mx[1][2 * q], mx[1][2 * q + 1]);
            const float zy = chroma_z(my[0][2 * q], my[0][2 * q + 1], my[1][2 * q], my[1][2 * q + 1]);
            cw |= chroma_gmem(f.uv, pitch, bias_c, zx, zy) << (16 * q);
        }
        if (!kRagged || valid > 0) {
            store_word<kRagged>(o.y0, pack4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]), valid);
            store_word<kRagged>(o.y1, pack4(acc[1][0], acc[1][1], acc[1][2], acc[1][3]), valid);
            store_word<kRagged>(o.c, cw, valid);
        }
        o.y0 += o.step_y; o.y1 += o.step_y; o.c += o.step_c;
    }
}

// The same with taps from the staged tile (row pitch PL bytes for both planes).
template <int PL, bool kRagged>
__device__ __forceinline__ void band_smem(const Geom& g, const ColPoly& cp, unsigned lconst, unsigned cconst, int dv0,
                                          int nrows, RowPtrs& o, int valid)
{
#pragma unroll 1
    for (int dv = dv0; dv < dv0 + nrows; dv += 2) {
        float mx[2][4], my[2][4];
        row_coords(cp, row_t(g, dv), mx[0], my[0]);
        row_coords(cp, row_t(g, dv + 1), mx[1], my[1]);
        int acc[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[r][i] = luma_smem<PL>(lconst, mx[r][i], my[r][i]);
        unsigned cw = 0u;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float zx = chroma_z(mx[0][2 * q], mx[0][2 * q + 1], mx[1][2 * q], mx[1][2 * q + 1]);
            const float zy = chroma_z(my[0][2 * q], my[0][2 * q + 1], my[1][2 * q], my[1][2 * q + 1]);
            cw |= chroma_smem<PL>(cconst, zx, zy) << (16 * q);
        }
        if (!kRagged || valid > 0) {
            store_word<kRagged>(o.y0, pack4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]), valid);
            store_word<kRagged>(o.y1, pack4(acc[1][0], acc[1][1], acc[1][2], acc[1][3]), valid);
            store_word<kRagged>(o.c, cw, valid);
        }
        o.y0 += o.step_y; o.y1 += o.step_y; o.c += o.step_c;
    }
}

struct Stage {  // per-warp staging state (variant TILED)
    unsigned mbar, lbuf, cbuf, parity;
};

// Copy the band's source rectangle into shared memory and sample from there; returns false
// (nothing done) when the rectangle does not fit, and the caller gathers from global memory.
template <int PL, bool kRagged>
__device__ __forceinline__ void band_staged(const Geom& g, const ColPoly& cp, const PlaneRefs& f, const BandBox& bb,
                                            int lx0, int wb, int nr, int cbx0, int cwb, int cnr, Stage& st, int lane,
                                            int dv0, int nrows, RowPtrs& o, int valid)
{
    __syncwarp();  // every lane is done reading the previous band
    const size_t pitch = (size_t)g.src_pitch;
    for (int r = lane; r < nr; r += 32)
        bulk_g2s(st.lbuf + (unsigned)(r * PL), f.y + (size_t)(bb.y0 + r) * pitch + lx0, (unsigned)wb, st.mbar);
    for (int r = lane; r < cnr; r += 32)
        bulk_g2s(st.cbuf + (unsigned)(r * PL), f.uv + (size_t)(bb.cy0 + r) * pitch + cbx0, (unsigned)cwb, st.mbar);
    if (lane == 0) mbar_expect_tx(st.mbar, (unsigned)(nr * wb + cnr * cwb));
    const unsigned lconst = st.lbuf - (unsigned)bb.y0 * PL - (unsigned)lx0 - kMagicShift * PL - kMagicShift;
    const unsigned cconst = ((st.cbuf - (unsigned)bb.cy0 * PL - (unsigned)cbx0 - kMagicShift * PL) >> 1) - kMagicShift;
    mbar_wait(st.mbar, st.parity);
    st.parity ^= 1u;
    band_smem<PL, kRagged>(g, cp, lconst, cconst, dv0, nrows, o, valid);
}

template <bool kStaged, bool kRagged>
__device__ __forceinline__ void interior_piece(const Geom& g, const ColPoly& cp, const PlaneRefs& f,
                                               const PieceRec* __restrict__ rec, Stage& st, int lane, int rows,
                                               RowPtrs& o, int valid)
{
    if (!kStaged) {
        band_gmem<kRagged>(g, cp, f, 0, rows, o, valid);
        return;
    }
    for (int dv0 = 0; dv0 < rows; dv0 += kBandH) {
        const int nrows = min(kBandH, rows - dv0);
        const int4 raw = __ldg(reinterpret_cast<const int4*>(&rec->band[dv0 / kBandH]));
        BandBox bb;
        bb.x0 = (int16_t)(raw.x & 0xffff); bb.x1 = (int16_t)(raw.x >> 16);
        bb.y0 = (int16_t)(raw.y & 0xffff); bb.y1 = (int16_t)(raw.y >> 16);
        bb.cx0 = (int16_t)(raw.z & 0xffff); bb.cx1 = (int16_t)(raw.z >> 16);
        bb.cy0 = (int16_t)(raw.w & 0xffff); bb.cy1 = (int16_t)(raw.w >> 16);
        const int lx0 = bb.x0 & ~15, wb = (bb.x1 - lx0 + 16) & ~15, nr = bb.y1 - bb.y0 + 1;
        const int cbx0 = (2 * bb.cx0) & ~15, cwb = (2 * bb.cx1 + 2 - cbx0 + 15) & ~15, cnr = bb.cy1 - bb.cy0 + 1;
        const int need = max(wb, cwb);
        const bool sane = nr > 0 && cnr > 0 && bb.x0 >= 0 && bb.y0 >= 0 && bb.cx0 >= 0 && bb.cy0 >= 0;
        if (sane && need <= 160 && nr * 160 <= kLumaCap && cnr * 160 <= kChromaCap)
            band_staged<160, kRagged>(g, cp, f, bb, lx0, wb, nr, cbx0, cwb, cnr, st, lane, dv0, nrows, o, valid);
        else if (sane && need <= 288 && nr * 288 <= kLumaCap && cnr * 288 <= kChromaCap)
            band_staged<288, kRagged>(g, cp, f, bb, lx0, wb, nr, cbx0, cwb, cnr, st, lane, dv0, nrows, o, valid);
        else if (sane && need <= 416 && nr * 416 <= kLumaCap && cnr * 416 <= kChromaCap)
            band_staged<416, kRagged>(g, cp, f, bb, lx0, wb, nr, cbx0, cwb, cnr, st, lane, dv0, nrows, o, valid);
        else
            band_gmem<kRagged>(g, cp, f, dv0, nrows, o, valid);
    }
}

template <bool kStaged>
__global__ void __launch_bounds__(32 * kWarps)
warp_nv12_poly_kernel(const Geom g, const FrameBatch b, const PieceRec* __restrict__ table)
{
    extern __shared__ __align__(16) uint8_t smem[];
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
            float mx[2][4], my[2][4];
            exact_rows(g, R, u_lo, u0, v_base + dv, mx, my);
            sample_rows_checked(g, f, u0, v_base + dv, mx, my);
        }
        return;
    }

    ColPoly cp;
    derive(rec, lane, cp);

    if (!(flags & kPieceInterior)) {  // polynomial coordinates, checked sampler
        for (int dv = 0; dv < rows; dv += 2) {
            float mx[2][4], my[2][4];
            row_coords(cp, row_t(g, dv), mx[0], my[0]);
            row_coords(cp, row_t(g, dv + 1), mx[1], my[1]);
            sample_rows_checked(g, f, u0, v_base + dv, mx, my);
        }
        return;
    }

    Stage st{};
    if (kStaged) {
        uint8_t* mine = smem + (size_t)threadIdx.y * kWarpSmem;
        st.mbar = smem_u32(mine);
        st.lbuf = st.mbar + 16;
        st.cbuf = st.lbuf + kLumaCap;
        st.parity = 0;
        if (lane == 0) mbar_init(st.mbar, 1);
        __syncwarp();
    }
    // 4-byte stores need a 4-byte aligned frame base and row pitch and a full piece
    const bool word_ok = ((reinterpret_cast<uintptr_t>(f.dst) | (uintptr_t)g.dst_pitch) & 3) == 0;
    if (word_ok && u_lo + kPieceW <= g.out_w)
        interior_piece<kStaged, false>(g, cp, f, rec, st, lane, rows, o, valid);
    else
        interior_piece<kStaged, true>(g, cp, f, rec, st, lane, rows, o, valid);
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
        float mx[2][4], my[2][4];
        const int v0 = v_base + dv;
        if (flags & kPiecePoly) {
            row_coords(cp, row_t(g, dv), mx[0], my[0]);
            row_coords(cp, row_t(g, dv + 1), mx[1], my[1]);
        } else {
            exact_rows(g, R, u_lo, u0, v0, mx, my);
        }
        if (plane == 0) {
            for (int r = 0; r < 2; ++r)
                for (int i = 0; i < 4; ++i)
                    if (u0 + i < g.out_w && v0 + r < g.out_h) {
                        map_x[(size_t)(v0 + r) * map_pitch + u0 + i] = mx[r][i];
                        map_y[(size_t)(v0 + r) * map_pitch + u0 + i] = my[r][i];
                    }
        } else {
            for (int q = 0; q < 2; ++q)
                if (u0 + 2 * q + 1 < g.out_w && v0 + 1 < g.out_h) {
                    const size_t o = (size_t)(v0 >> 1) * map_pitch + (u0 >> 1) + q;
                    map_x[o] = chroma_coord(mx[0][2 * q], mx[0][2 * q + 1], mx[1][2 * q], mx[1][2 * q + 1]);
                    map_y[o] = chroma_coord(my[0][2 * q], my[0][2 * q + 1], my[1][2 * q], my[1][2 * q + 1]);
                }
        }
    }
}

}  // namespace

cudaError_t launch_warp_nv12_poly(const Geom& g, const FrameBatch& b, const PieceRec* table, bool staged,
                                  cudaStream_t st)
{
    dim3 block(32, kWarps);
    dim3 grid(pieces_x(g.out_w), (pieces_y(g.out_h, g.piece_h) + kWarps - 1) / kWarps, b.n_frames);
    if (staged)
        warp_nv12_poly_kernel<true><<<grid, block, kWarps * kWarpSmem, st>>>(g, b, table);
    else
        warp_nv12_poly_kernel<false><<<grid, block, 0, st>>>(g, b, table);
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
