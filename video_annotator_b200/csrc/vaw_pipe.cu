// vaw_pipe.cu -- fused map + remap for NV12 as a persistent producer/consumer pipeline
// (variant PIPE).  Same arithmetic and same bytes as vaw_tile.cu (variant TILED); what changes
// is the schedule.
//
// Replaces FrameSourceWarp::warp_frame's two passes
// (/root/reference/opencv/FrameSourceWarp.cpp:272-314).
//
// Why: with one CTA per piece (vaw_tile.cu) every CTA starts with a serial chain -- load the
// piece record from L2, issue the TMA loads, wait for the tile, exchange the column polynomials --
// during which its four warps issue nothing; ncu attributes a quarter of all warp time to it
// (long_scoreboard + barrier) and the issue ports stay at ~71 %.  Here each CTA is persistent:
//   - one PRODUCER warp pulls the next piece index from a global counter (so all CTAs advance
//     along one frontier and the L2 working set stays about one source frame), copies the
//     piece record into the stage, and issues the cp.async.bulk.tensor loads of the source box
//     into the stage's tile buffer (completion on the stage's `full` mbarrier);
//   - four CONSUMER warps wait on `full`, collapse the polynomial onto their columns, sample
//     the 32 rows of the piece from the staged tile and release the stage (`empty` mbarrier).
// With two stages the loads of piece k+1 fly while piece k is being sampled, and the per-CTA set-up
// is paid once per launch.
//
// Measured (B200, C3, 64 frames): 0.86 ms against 0.74 ms for TILED, so TILED stays the default and
// this variant is kept for A/B.  The reason is shared memory: a tile is 28-32 KB, so an SM holds
// six of them either way; TILED spends them on six CTAs = 24 sampling warps whose start-up bubbles
// overlap each other, PIPE on 3 CTAs x 2 stages = 12 sampling warps, and one piece of look-ahead
// does not cover a tile load that mostly comes from HBM (the consumers still wait on `full` for
// 23 % of their time, ncu).  A deeper ring needs smaller tiles (half-height pieces): round 2.
#include <cuda.h>
#include <stdint.h>
#include "vaw_internal.h"
#include "vaw_poly.cuh"
#include "vaw_tile.cuh"

namespace vaw {

namespace {

constexpr int kConsumers = 4;                       // consumer warps (8 rows of the piece each)
constexpr int kRowsPerWarp = kPieceHMax / kConsumers;
constexpr int kThreads = 32 * (kConsumers + 1);     // + 1 producer warp
constexpr int kStages = 2;
constexpr int kCoefBytes = 8 * 32 * 16;             // column polynomials exchanged between the consumer warps

// Stage header written by the producer next to the copied piece record.
struct StageHead {
    int idx;          // flattened piece index, -1 = no more work
    int mode;         // kModeStaged, kModeDirect
    int pl, nr8, cnr8, lx0, cbx0;
    int pad;
};
enum { kModeDirect = 0, kModeStaged = 1 };

// shared memory layout (bytes):
//   [0, 64)                     mbarriers: full[kStages], empty[kStages]
//   [64, 64 + 2*kCoefBytes)     two exchange buffers for the column polynomials (alternating pieces)
//   then per stage: StageHead (32) + PieceRec (224) = 256, then the tile (tile_cap, 128-byte aligned)
constexpr int kBarOffset = 0;
constexpr int kCoefOffset = 128;
constexpr int kStageOffset = kCoefOffset + 2 * kCoefBytes;
__host__ __device__ inline int stage_bytes(int tile_cap) { return 256 + tile_cap; }

__device__ __forceinline__ void mbar_arrive(unsigned mbar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void consumer_sync()  // the four consumer warps only
{
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kConsumers) : "memory");
}

}  // namespace

__global__ void __launch_bounds__(kThreads)
warp_nv12_pipe_kernel(const Geom g, const FrameBatch b, const PieceRec* __restrict__ table,
                      unsigned* __restrict__ counter, const __grid_constant__ TileMaps maps)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int npx = pieces_x(g.out_w), npy = pieces_y(g.out_h, kPieceHMax);
    const int total = b.n_frames * npy * npx;
    const int sbytes = stage_bytes(maps.tile_cap);
    const unsigned bar0 = smem_u32(smem + kBarOffset);
    auto full_bar = [&](int s) { return bar0 + 8u * (unsigned)s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (unsigned)(kStages + s); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), 1);            // the producer's arrive (+ the TMA bytes)
            mbar_init(empty_bar(s), kConsumers);  // one arrive per consumer warp
        }
    }
    __syncthreads();

    if (w == kConsumers) {
        // ================================ producer warp ==========================================
        // Pieces are pulled from the global queue four at a time (one atomic per batch; lane i < 4
        // prefetches the flags and box of piece i), so the queue and record latencies are paid once
        // per batch.  Pure-border pieces never reach the consumers: the producer fills them itself.
        constexpr int kBatch = 4;
        const bool word_base_ok =
            ((reinterpret_cast<uintptr_t>(b.dst) | (uintptr_t)g.dst_pitch | (uintptr_t)b.dst_frame_stride) & 3) == 0;
        int k = 0;  // stages handed to the consumers so far
        bool more = true;
        while (more) {
            int first = 0;
            if (lane == 0) first = (int)atomicAdd(counter, (unsigned)kBatch);
            first = __shfl_sync(0xffffffffu, first, 0);
            unsigned my_flags = 0;
            if (lane < kBatch && first + lane < total) my_flags = __ldg(&table[first + lane].flags);
            for (int i = 0; i < kBatch; ++i) {
                const int idx = first + i;
                if (idx >= total) { more = false; break; }
                const unsigned flags = __shfl_sync(0xffffffffu, my_flags, i);
                const int frame = idx / (npy * npx);
                if (flags & kPieceOutside) {  // pure border: fill it here (32 rows x 128 px + chroma)
                    const int rem = idx - frame * npy * npx;
                    const int py = rem / npx, px = rem - py * npx;
                    const int u0 = px * kPieceW + 4 * lane, v_base = py * kPieceHMax;
                    const int rows = min(kPieceHMax, g.out_h - v_base), valid = g.out_w - u0;
                    uint8_t* dst = b.dst + (size_t)frame * b.dst_frame_stride;
                    uint8_t* yrow = dst + (size_t)v_base * g.dst_pitch + u0;
                    uint8_t* crow = dst + (size_t)(g.out_h + (v_base >> 1)) * g.dst_pitch + u0;
                    const unsigned yw = (g.border & 255u) * 0x01010101u;
                    const unsigned cw = ((g.border >> 8) & 0xffffu) * 0x00010001u;
                    const bool fast = word_base_ok && valid >= 4;
                    if (valid > 0) {
                        for (int r = 0; r < rows; ++r, yrow += g.dst_pitch) {
                            if (fast) *reinterpret_cast<unsigned*>(yrow) = yw; else store_word<true>(yrow, yw, valid);
                        }
                        for (int r = 0; r < rows / 2; ++r, crow += g.dst_pitch) {
                            if (fast) *reinterpret_cast<unsigned*>(crow) = cw; else store_word<true>(crow, cw, valid);
                        }
                    }
                    continue;
                }
                const int s = k % kStages;
                const unsigned round = (unsigned)(k / kStages);
                uint8_t* stage = smem + kStageOffset + s * sbytes;
                // start the record load before waiting for the stage
                float4 recpart = make_float4(0.f, 0.f, 0.f, 0.f);
                if (lane < 14) recpart = __ldg(reinterpret_cast<const float4*>(table + idx) + lane);
                if (k >= kStages) mbar_wait(empty_bar(s), (round - 1u) & 1u);  // consumers released the stage
                StageHead* head = reinterpret_cast<StageHead*>(stage);
                if (lane < 14) reinterpret_cast<float4*>(stage + 32)[lane] = recpart;
                __syncwarp();
                const PieceRec* lrec = reinterpret_cast<const PieceRec*>(stage + 32);
                const PieceBox box = lrec->box;
                int mode = kModeDirect, pl = 0, nr8 = 0, cnr8 = 0, lx0 = 0, cbx0 = 0;
                if (flags & kPiecePoly) {
                    lx0 = box.x0 & ~15;
                    const int wb = (box.x1 - lx0 + 16) & ~15;
                    cbx0 = (2 * box.cx0) & ~15;
                    const int cwb = (2 * box.cx1 + 2 - cbx0 + 15) & ~15;
                    nr8 = (box.y1 - box.y0 + 8) & ~7;
                    cnr8 = (box.cy1 - box.cy0 + 8) & ~7;
                    const int need = max(wb, cwb);
                    const int pl128 = (need + 127) & ~127, pl32 = max(kTileMinPitch, (need + 31) & ~31);
                    pl = (pl128 <= kTileMaxPitch && pl128 * (nr8 + cnr8) <= maps.tile_cap) ? pl128 : pl32;
                    if (maps.enabled && pl <= kTileMaxPitch && nr8 > 0 && cnr8 > 0 && pl * (nr8 + cnr8) <= maps.tile_cap)
                        mode = kModeStaged;
                }
                if (lane == 0) {
                    head->idx = idx; head->mode = mode; head->pl = pl; head->nr8 = nr8; head->cnr8 = cnr8;
                    head->lx0 = lx0; head->cbx0 = cbx0;
                }
                __syncwarp();
                if (mode == kModeStaged) {
                    // lanes issue the 8-row boxes in parallel: lane r -> luma box r, then chroma boxes
                    const unsigned l0 = smem_u32(stage + 256), c0 = l0 + (unsigned)(nr8 * pl);
                    const CUtensorMap* map = &maps.m[(pl - kTileMinPitch) / kTilePitchStep];
                    if (lane == 0) mbar_expect_tx(full_bar(s), (unsigned)(pl * (nr8 + cnr8)));  // arrive + expected bytes
                    __syncwarp();
                    const int nl = nr8 >> 3, nc = cnr8 >> 3;
                    for (int q = lane; q < nl + nc; q += 32) {
                        if (q < nl) tma_load_3d(l0 + (unsigned)(q * 8 * pl), map, lx0 >> 2, box.y0 + 8 * q, frame, full_bar(s));
                        else tma_load_3d(c0 + (unsigned)((q - nl) * 8 * pl), map, cbx0 >> 2, g.src_h + box.cy0 + 8 * (q - nl), frame, full_bar(s));
                    }
                } else if (lane == 0) {
                    mbar_arrive(full_bar(s));
                }
                ++k;
            }
        }
        // tell the consumers there is no more work
        {
            const int s = k % kStages;
            const unsigned round = (unsigned)(k / kStages);
            if (k >= kStages) mbar_wait(empty_bar(s), (round - 1u) & 1u);
            if (lane == 0) {
                reinterpret_cast<StageHead*>(smem + kStageOffset + s * sbytes)->idx = -1;
                mbar_arrive(full_bar(s));
            }
        }
        return;
    }

    // ==================================== consumer warps =========================================
    const int dv0 = w * kRowsPerWarp;
    const bool even_ok = ((reinterpret_cast<uintptr_t>(b.dst) | (uintptr_t)g.dst_pitch | (uintptr_t)b.dst_frame_stride) & 1) == 0;
    const bool word_base_ok = ((reinterpret_cast<uintptr_t>(b.dst) | (uintptr_t)g.dst_pitch | (uintptr_t)b.dst_frame_stride) & 3) == 0;
    for (int k = 0;; ++k) {
        const int s = k % kStages;
        const unsigned round = (unsigned)(k / kStages);
        mbar_wait(full_bar(s), round & 1u);  // record copied, tile landed
        uint8_t* stage = smem + kStageOffset + s * sbytes;
        const StageHead head = *reinterpret_cast<const StageHead*>(stage);
        if (head.idx < 0) break;
        const PieceRec* rec = reinterpret_cast<const PieceRec*>(stage + 32);  // in shared memory
        const unsigned flags = rec->flags;
        const int frame = head.idx / (npy * npx);
        const int rem = head.idx - frame * npy * npx;
        const int py = rem / npx, px = rem - py * npx;
        const int u_lo = px * kPieceW, u0 = u_lo + 4 * lane, v_base = py * kPieceHMax;
        const int rows = min(kPieceHMax, g.out_h - v_base);  // even for NV12
        const int my_rows = max(0, min(kRowsPerWarp, rows - dv0));
        const int valid = g.out_w - u0;

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
        const bool word_ok = word_base_ok && u_lo + kPieceW <= g.out_w;

        if (flags & kPieceOutside) {  // pure border: nothing to compute
            const unsigned yw = (g.border & 255u) * 0x01010101u;
            const unsigned cw = ((g.border >> 8) & 0xffffu) * 0x00010001u;
            if (valid > 0)
                for (int dv = 0; dv < my_rows; dv += 2) {
                    store_word<true>(o.y0, yw, valid);
                    store_word<true>(o.y1, yw, valid);
                    store_word<true>(o.c, cw, valid);
                    o.y0 += o.step_y; o.y1 += o.step_y; o.c += o.step_c;
                }
        } else if (!(flags & kPiecePoly)) {  // op-for-op per pixel
            const Rot R = load_rot(b, frame);
            for (int dv = dv0; dv < dv0 + my_rows; dv += 2) {
                float2 m[2][4];
                exact_rows(g, R, u_lo, u0, v_base + dv, m);
                sample_rows_checked(g, f, u0, v_base + dv, m);
            }
        } else if (head.mode != kModeStaged) {  // does not fit a tile: gather from global memory
            if (my_rows > 0) {
                ColPoly cp;
                derive(table + head.idx, lane, cp);
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
            }
        } else {
            // ---- staged: collapse (warp w does column slot j = w), exchange, sample ---------------
            float4* coefs = reinterpret_cast<float4*>(smem + kCoefOffset + (k & 1) * kCoefBytes);
            {
                float2 a[kNv];
                collapse_column(rec->c, ((float)pair_column(lane, w) - 63.5f) * 0.015625f, a);  // s is exact
                coefs[(2 * w) * 32 + lane] = make_float4(a[0].x, a[0].y, a[1].x, a[1].y);
                coefs[(2 * w + 1) * 32 + lane] = make_float4(a[2].x, a[2].y, a[3].x, a[3].y);
            }
            consumer_sync();
            ColPoly cp;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 lo = coefs[(2 * j) * 32 + lane], hi = coefs[(2 * j + 1) * 32 + lane];
                cp.a[j][0] = make_float2(lo.x, lo.y); cp.a[j][1] = make_float2(lo.z, lo.w);
                cp.a[j][2] = make_float2(hi.x, hi.y); cp.a[j][3] = make_float2(hi.z, hi.w);
            }
            cp.base = rec->base;
            const PieceBox box = rec->box;
            const int pl = head.pl;
            uint8_t* ltile = stage + 256;
            uint8_t* ctile = ltile + head.nr8 * pl;
            if (!(flags & kPieceInterior)) {  // straddles the frame border: paint the outside cells
                const unsigned by_ = g.border & 255u, bu = (g.border >> 8) & 255u, bv = (g.border >> 16) & 255u;
                fill_border(ltile, pl, head.nr8, box.y0, g.src_h, head.lx0, g.src_w, by_ * 0x01010101u, threadIdx.x, 32 * kConsumers);
                fill_border(ctile, pl, head.cnr8, box.cy0, g.src_h >> 1, head.cbx0, g.src_w, (bu | (bv << 8)) * 0x00010001u,
                            threadIdx.x, 32 * kConsumers);
                consumer_sync();
            }
            if (my_rows > 0) {
                const unsigned upl = (unsigned)pl;
                const unsigned lconst = smem_u32(ltile) - (unsigned)box.y0 * upl - (unsigned)head.lx0 - kMagicShift * upl - kMagicShift;
                const unsigned cconst = ((smem_u32(ctile) - (unsigned)box.cy0 * upl - (unsigned)head.cbx0 - kMagicShift * upl) >> 1) - kMagicShift;
                const int shift = 2 * lane - 4 * lane;  // the staged path uses the pair mapping
                o.y0 += shift; o.y1 += shift; o.c += shift;
                const bool in_a = u_lo + 2 * lane < g.out_w, in_b = u_lo + 64 + 2 * lane < g.out_w;
                const TileBounds tb = {smem_u32(ltile), smem_u32(ltile) + (unsigned)(head.nr8 * pl), smem_u32(ctile),
                                       smem_u32(ctile) + (unsigned)(head.cnr8 * pl)};
                if (even_ok && u_lo + kPieceW <= g.out_w) rows_tile<false>(g, cp, lconst, cconst, upl, dv0, my_rows, o, in_a, in_b, tb);
                else rows_tile<true>(g, cp, lconst, cconst, upl, dv0, my_rows, o, in_a, in_b, tb);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(s));  // this warp is done with the stage
    }
}

int pipe_smem_bytes(int tile_cap) { return kStageOffset + kStages * stage_bytes(tile_cap); }

cudaError_t launch_warp_nv12_pipe(const Geom& g, const FrameBatch& b, const PieceRec* table, unsigned* counter,
                                  const TileMaps& maps, cudaStream_t st)
{
    static bool configured[64] = {};  // per device; a benign race: the attribute call is idempotent
    static int sm_count[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(warp_nv12_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 << 10);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const int smem = pipe_smem_bytes(maps.tile_cap);
    if (smem > (227 << 10)) return cudaErrorInvalidValue;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, warp_nv12_pipe_kernel, kThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const long long total = (long long)pieces_x(g.out_w) * pieces_y(g.out_h, kPieceHMax) * b.n_frames;
    long long ctas = (long long)sm_count[dev] * per_sm;
    if (ctas > total) ctas = total;
    if (ctas < 1) ctas = 1;
    e = cudaMemsetAsync(counter, 0, sizeof(unsigned), st);
    if (e != cudaSuccess) return e;
    warp_nv12_pipe_kernel<<<(unsigned)ctas, kThreads, smem, st>>>(g, b, table, counter, maps);
    return cudaGetLastError();
}

}  // namespace vaw
