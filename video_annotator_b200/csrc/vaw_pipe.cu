// vaw_pipe.cu -- fused map + remap for NV12 as one persistent producer/consumer pipeline per SM
// (variant PIPE).  Same arithmetic and same bytes as vaw_tile.cu (variant TILED); what changes
// is the schedule.
//
// Replaces FrameSourceWarp::warp_frame's two passes
// (/root/reference/opencv/FrameSourceWarp.cpp:272-314).
//
// Why: with one CTA per piece (vaw_tile.cu) every CTA starts with a serial chain -- load the
// piece record from L2, issue the TMA loads, wait for the tile, exchange the column polynomials --
// during which its four warps issue nothing; ncu attributes a quarter of all warp time to it and
// the issue ports stay at ~71 %.  Pure-border pieces still cost a CTA launch each.  Here:
//   - ONE CTA per SM: a PRODUCER warp, two ISSUER warps, a FILLER warp and kGroups CONSUMER GROUPS
//     of four warps;
//   - the FILLER warp writes the pure-border pieces (they need no source);
//   - the PRODUCER pulls pieces from a global queue (eight per atomic, so all SMs advance along one
//     frontier and the L2 working set stays about one source frame); lane i prepares piece i of the
//     ticket (position, box, tile size) and the warp then hands out, in queue order, a stage
//     descriptor and exactly the bytes the piece's source box needs from a RING in shared memory
//     (~198 KB; space is reclaimed in stage order as the consumers release stages);
//   - ISSUER j takes the described stages j, j+2, ...: expected bytes, a bulk copy of the piece
//     record, the cp.async.bulk.tensor loads of the box (32-row boxes, then 8-row boxes), all
//     completing on the stage's `full` mbarrier;
//   - consumer group g takes stages g, g+kGroups, ...: waits on `full`, collapses the polynomial,
//     samples the 32 rows from the staged tile, releases the stage (`empty` mbarrier).
//
// What the measurements say (B200, C3, 64 frames, against TILED measured in the same runs; DESIGN.md has the
// table): shared memory is the currency.  A C3 tile is 28 KB and the ring holds seven.  Four groups (16
// sampling warps, 96 registers, three tiles of look-ahead) run at TILED's speed with two thirds of its
// sampling warps (0.726 ms against 0.699 ms); five groups (two tiles ahead) 0.759 ms and six groups (one
// tile ahead) 0.827 ms even with their registers raised by setmaxnreg -- the consumers then wait for their
// loads up to 30 % of the time.  Staging half pieces instead (13-16 KB tiles) needs per-half source boxes
// from the builder (+50 % builder time) and more live state than 72 registers hold: 1.14 ms.  A single
// producer warp doing everything (a ~230-instruction dependent chain per piece) was the limit before the
// issuer warps existed: 0.95 ms.  Where tiles are too large for TILED's six per SM (C5, 50 KB: 0.464 ms
// against 0.505 ms) or small enough for the ring to run four pieces ahead (C2, 16 KB: 0.204 against
// 0.212 ms) the pipeline is the faster variant, and AUTO selects it there (vaw_create).
#include <cuda.h>
#include <atomic>
#include <stdint.h>
#include "vaw_internal.h"
#include "vaw_poly.cuh"
#include "vaw_tile.cuh"

namespace vaw {

namespace {

#ifndef VAW_PIPE_GROUPS
#define VAW_PIPE_GROUPS 4
#endif
constexpr int kGroups = VAW_PIPE_GROUPS;            // consumer groups per CTA
constexpr int kGroupWarps = 4;                      // warps per group
constexpr int kRowsPerWarp = kPieceHMax / kGroupWarps;      // 8 rows of the piece per warp
constexpr int kIssuers = 2;                                 // warps that issue the loads of the stages the producer described
constexpr int kIssuerWarp0 = 1;                             // warps 0..3 (one warpgroup): producer, two issuers, filler
constexpr int kFillerWarp = 3;                              // fills the pure-border pieces
constexpr int kFirstConsumer = 4;
#ifndef VAW_PIPE_PARK_NS
#define VAW_PIPE_PARK_NS 2000      // suspend-time hint of the pipeline's mbarrier waits (0: plain try_wait spin)
#endif
#ifndef VAW_PIPE_AUX_REGS
#define VAW_PIPE_AUX_REGS 40       // registers the four auxiliary warps keep (setmaxnreg.dec)
#endif
#ifndef VAW_PIPE_CONSUMER_REGS
#define VAW_PIPE_CONSUMER_REGS 0   // registers the consumer warps grow to (setmaxnreg.inc); 0 = leave the launch allocation
#endif
constexpr int kThreads = 32 * (2 + kIssuers + kGroups * kGroupWarps);  // registers are granted per 4 warps: 20 warps -> 96 each, 24 -> 80, 28 -> 72
constexpr int kSlots = 16;                          // stage descriptors (<= 32: one lane per slot in the producer)
constexpr int kBatch = 8;                           // pieces per queue ticket
constexpr int kCoefBytes = 8 * 32 * 16;             // column polynomials exchanged inside a group
constexpr int kSlotBytes = 32 + (int)sizeof(PieceRec);

// Stage descriptor written by the producer; the piece record is copied next to it.
struct SlotHead {
    int idx;            // flattened piece index, -1 = no more work
    int frame;
    int pxy;            // piece column | piece row << 16
    int plmode;         // tile row pitch | mode << 16
    int rows;           // luma tile rows | chroma tile rows << 16
    int lorg, corg;     // first source byte of the tile rows (int16) | first source row (int16) << 16, luma / chroma
    unsigned tile_off;  // byte offset of the tile in the ring
};
static_assert(sizeof(SlotHead) == 32, "slot head layout");
enum { kModeDirect = 0, kModeStaged = 1 };  // direct: no tile, the taps are gathered from global memory

// shared memory layout (bytes)
constexpr int kBarOffset = 0;                                  // full[kSlots], empty[kSlots], ready[kSlots]
constexpr int kSlotOffset = 1024;                              // per slot: SlotHead (32) + PieceRec (256)
constexpr int kCoefOffset = kSlotOffset + kSlots * kSlotBytes; // one exchange buffer per group
constexpr int kRingOffset = (kCoefOffset + kGroups * kCoefBytes + 1023) & ~1023;
constexpr int kSmemBytes = 227 << 10;
constexpr int kRingBytes = kSmemBytes - kRingOffset;
static_assert(24 * kSlots <= kSlotOffset, "mbarrier area");
static_assert(kSlots <= 32 && kBatch <= 32, "the producer keeps one slot / one piece of a ticket per lane");
static_assert(kGroups >= 1 && kGroups <= 15, "one named barrier per consumer group (ids 1..15)");
static_assert(kSlotBytes % 16 == 0 && kCoefOffset % 16 == 0, "alignment");

__device__ __forceinline__ void pipe_wait(unsigned mbar, unsigned parity)
{
#if VAW_PIPE_PARK_NS
    mbar_wait_parked(mbar, parity, VAW_PIPE_PARK_NS);
#else
    mbar_wait(mbar, parity);
#endif
}
__device__ __forceinline__ void mbar_arrive(unsigned mbar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void group_sync(int g)  // the four warps of consumer group g
{
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(32 * kGroupWarps) : "memory");
}
__device__ __forceinline__ void fence_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// The tile of one piece as the builder laid it out (PieceStage): returns the bytes needed, 0 if it cannot be staged.
__device__ __forceinline__ unsigned plan_tile(int4 raw, int enabled, int& pl, int& rows, int& lorg, int& corg)
{
    lorg = raw.x;                      // lx0 | by0 << 16
    corg = raw.y;                      // cbx0 | cy0 << 16
    pl = raw.z & 0xffff;
    const int nrows = (raw.z >> 16) & 0xffff, cnrows = raw.w & 0xffff;
    rows = nrows | (cnrows << 16);
    if (!enabled || pl == 0 || pl * (nrows + cnrows) > kRingBytes / 2) return 0u;
    return (unsigned)(pl * (nrows + cnrows) + 127) & ~127u;
}

// A piece without a tile (no polynomial certificate, or a box the ring cannot hold): per-pixel
// coordinates and/or taps gathered from global memory, as in vaw_poly.cu.
// (Out of line it was slower: 0.771 ms against 0.732 ms at four groups.)
__device__ __forceinline__ void direct_piece(const Geom& g, const FrameBatch& b, const PieceRec* __restrict__ table, int idx,
                                          unsigned flags, int frame, int u_lo, int v_base, int dv0, int my_rows, int lane,
                                          uint8_t* dst, bool word_base_ok)
{
    const int u0 = u_lo + 4 * lane, valid = g.out_w - u0;
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
    } else if (my_rows > 0) {  // no tile (layout or size): gather from global memory
        ColPoly cp;
        derive(table + idx, lane, cp);
        if (flags & kPieceInterior) {
            RowPtrs o;
            o.y0 = dst + (size_t)(v_base + dv0) * g.dst_pitch + u0;
            o.y1 = o.y0 + g.dst_pitch;
            o.c = dst + (size_t)(g.out_h + ((v_base + dv0) >> 1)) * g.dst_pitch + u0;
            o.step_y = 2 * (size_t)g.dst_pitch;
            o.step_c = (size_t)g.dst_pitch;
            if (word_base_ok && u_lo + kPieceW <= g.out_w) band_gmem<false>(g, cp, f, dv0, my_rows, o, valid);
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
}

}  // namespace

__global__ void __launch_bounds__(kThreads, 1)
warp_nv12_pipe_kernel(const Geom g, const FrameBatch b, const PieceRec* __restrict__ table,
                      unsigned* __restrict__ counter, const __grid_constant__ TileMaps maps)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int npx = pieces_x(g.out_w), npy = pieces_y(g.out_h, kPieceHMax);
    const int per_frame = npy * npx;
    const int total = b.n_frames * per_frame;
    const unsigned bar0 = smem_u32(smem + kBarOffset);
    auto full_bar = [&](int s) { return bar0 + 8u * (unsigned)s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (unsigned)(kSlots + s); };
    auto ready_bar = [&](int s) { return bar0 + 8u * (unsigned)(2 * kSlots + s); };
    uint8_t* const ring = smem + kRingOffset;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kSlots; ++s) {
            mbar_init(full_bar(s), 1);             // the issuer's arrive (+ the TMA bytes)
            mbar_init(empty_bar(s), kGroupWarps);  // one arrive per consumer warp of the group
            mbar_init(ready_bar(s), 1);            // the producer's arrive: the stage is described
        }
    }
    __syncthreads();

    if (w == 0) {
        // ================================ producer warp ==========================================
#if VAW_PIPE_CONSUMER_REGS
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(VAW_PIPE_AUX_REGS));  // hand registers to the consumers
#endif
        int k = 0;        // stages handed out so far
        int k_head = 0;   // oldest stage whose ring space is still accounted as in use
        unsigned tail = 0, used = 0;
        unsigned my_size = 0;  // lane s: bytes (tile + wrap padding) held by slot s
        auto release_oldest = [&]() {
            const int s = k_head % kSlots;
            pipe_wait(empty_bar(s), (unsigned)(k_head / kSlots) & 1u);
            used -= __shfl_sync(0xffffffffu, my_size, s);
            ++k_head;
        };
        // metadata of a ticket's pieces: lane i < kBatch holds the flags and the stage descriptor of piece first + i
        auto fetch = [&](int& first, unsigned& flags, int4& box) {
            int f = 0;
            if (lane == 0) f = (int)atomicAdd(counter, (unsigned)kBatch);
            first = __shfl_sync(0xffffffffu, f, 0);
            flags = kPieceOutside;  // lanes without a piece: nothing to stage
            box = make_int4(0, 0, 0, 0);
            if (lane < kBatch && first + lane < total) {
                const int4* r4 = reinterpret_cast<const int4*>(table + first + lane);
                flags = (unsigned)__ldg(r4 + 12).z;
                box = __ldg(r4 + 14);  // PieceStage
            }
        };
        int first_n; unsigned flags_n; int4 box_n;
        fetch(first_n, flags_n, box_n);
        for (;;) {
            const int first = first_n;
            const unsigned flags_l = flags_n;
            const int4 box = box_n;
            if (first >= total) break;
            fetch(first_n, flags_n, box_n);  // the next ticket's latency hides behind this one's work
            // ---- lane i < kBatch prepares piece first + i: position and the tile of its source box ----
            const int idx_l = first + lane;
            const int frame_l = idx_l / per_frame, rem_l = idx_l - frame_l * per_frame;
            const int py_l = rem_l / npx, px_l = rem_l - py_l * npx;
            int pl_l = 0, rows_l = 0, lorg_l = 0, corg_l = 0, mode_l = kModeDirect;
            unsigned need_l = 0u;
            if ((flags_l & (kPiecePoly | kPieceOutside)) == kPiecePoly) {
                need_l = plan_tile(box, maps.enabled, pl_l, rows_l, lorg_l, corg_l);
                if (need_l) mode_l = kModeStaged;
            }
            // pure-border pieces are the filler warp's; every other piece gets a stage, in queue order
            unsigned todo = __ballot_sync(0xffffffffu, !(flags_l & kPieceOutside));
            while (todo) {
                const int i = __ffs((int)todo) - 1;
                todo &= todo - 1;
                const unsigned need = __shfl_sync(0xffffffffu, need_l, i);
                // ---- a descriptor and ring space: reclaim released stages in order until both exist ----
                if (used == 0) tail = 0;
                const bool wrap = need && tail + need > (unsigned)kRingBytes;
                const unsigned padding = wrap ? (unsigned)kRingBytes - tail : 0u;
                while (k - k_head >= kSlots || used + padding + need > (unsigned)kRingBytes) release_oldest();
                const int s = k % kSlots;
                const unsigned off = wrap ? 0u : tail;
                tail = off + need;
                used += padding + need;
                if (lane == s) my_size = padding + need;
                if (lane == i) {  // the piece's own lane describes the stage; an issuer warp starts its loads
                    int4* head = reinterpret_cast<int4*>(smem + kSlotOffset + s * kSlotBytes);
                    head[0] = make_int4(idx_l, frame_l, px_l | (py_l << 16), pl_l | (mode_l << 16));
                    head[1] = make_int4(rows_l, lorg_l, corg_l, (int)off);
                    mbar_arrive(ready_bar(s));
                }
                __syncwarp();
                ++k;
            }
        }
        // one terminal stage per consumer group (idx -1), then one per issuer warp (idx -2)
        for (int t = 0; t < kGroups + kIssuers; ++t, ++k) {
            while (k - k_head >= kSlots) release_oldest();
            const int s = k % kSlots;
            if (lane == s) my_size = 0;
            if (lane == 0) {
                reinterpret_cast<SlotHead*>(smem + kSlotOffset + s * kSlotBytes)->idx = t < kGroups ? -1 : -2;
                mbar_arrive(ready_bar(s));
            }
            __syncwarp();
        }
        return;
    }

    if (w >= kIssuerWarp0 && w < kIssuerWarp0 + kIssuers) {
        // ================================ issuer warps ===========================================
#if VAW_PIPE_CONSUMER_REGS
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(VAW_PIPE_AUX_REGS));  // hand registers to the consumers
#endif
        // Issuer j starts the loads of stages j, j + kIssuers, ...: the record copy and the
        // TMA boxes of the tile (whole 32-row boxes, then 8-row boxes for the rest; luma first, then chroma).
        for (int k = w - kIssuerWarp0;; k += kIssuers) {
            const int s = k % kSlots;
            pipe_wait(ready_bar(s), (unsigned)(k / kSlots) & 1u);
            const uint8_t* slot = smem + kSlotOffset + s * kSlotBytes;
            const SlotHead head = *reinterpret_cast<const SlotHead*>(slot);
            const unsigned fb = full_bar(s);
            if (head.idx == -2) break;  // consecutive stages alternate between the issuers: one of these each
            if (head.idx < 0) {         // a consumer group's terminal stage: pass it on
                if (lane == 0) mbar_arrive(fb);
                continue;
            }
            if (lane == 0) {
                const int pl = head.plmode & 0xffff, mode = head.plmode >> 16;
                const int nrows = head.rows & 0xffff, cnrows = head.rows >> 16;
                {
                    // arrive + expected bytes: the record copy and, when staged, the boxes
                    mbar_expect_tx(fb, (unsigned)sizeof(PieceRec) + (mode == kModeStaged ? (unsigned)(pl * (nrows + cnrows)) : 0u));
                    bulk_g2s(smem_u32(slot + 32), table + head.idx, (unsigned)sizeof(PieceRec), fb);
                    if (mode == kModeStaged) {
                        const int lx0 = (int16_t)(head.lorg & 0xffff), by0 = head.lorg >> 16;
                        const int cbx0 = (int16_t)(head.corg & 0xffff), cy0 = head.corg >> 16;
                        const unsigned l0 = smem_u32(ring + head.tile_off), c0 = l0 + (unsigned)(nrows * pl);
                        const int mi = (pl - kTileMinPitch) / kTilePitchStep;
                        const CUtensorMap *map = &maps.m[mi], *map32 = &maps.m32[mi], *map4 = &maps.m4[mi];
                        int r = 0;
                        for (; r + 32 <= nrows; r += 32) tma_load_3d(l0 + (unsigned)(r * pl), map32, lx0 >> 2, by0 + r, head.frame + b.tma_frame0, fb);
                        for (; r + 8 <= nrows; r += 8) tma_load_3d(l0 + (unsigned)(r * pl), map, lx0 >> 2, by0 + r, head.frame + b.tma_frame0, fb);
                        if (r < nrows) tma_load_3d(l0 + (unsigned)(r * pl), map4, lx0 >> 2, by0 + r, head.frame + b.tma_frame0, fb);
                        for (r = 0; r + 32 <= cnrows; r += 32)
                            tma_load_3d(c0 + (unsigned)(r * pl), map32, cbx0 >> 2, g.src_h + cy0 + r, head.frame + b.tma_frame0, fb);
                        for (; r + 8 <= cnrows; r += 8)
                            tma_load_3d(c0 + (unsigned)(r * pl), map, cbx0 >> 2, g.src_h + cy0 + r, head.frame + b.tma_frame0, fb);
                        if (r < cnrows) tma_load_3d(c0 + (unsigned)(r * pl), map4, cbx0 >> 2, g.src_h + cy0 + r, head.frame + b.tma_frame0, fb);
                    }
                }
            }
            __syncwarp();
        }
        return;
    }

    if (w == kFillerWarp) {
        // ================================ filler warp ============================================
#if VAW_PIPE_CONSUMER_REGS
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(VAW_PIPE_AUX_REGS));  // hand registers to the consumers
#endif
        // Pure-border pieces need no source: this warp writes them, 32 flags per load, independent of
        // the queue (static stride over all pieces of the launch).
        const bool vec_ok =
            ((reinterpret_cast<uintptr_t>(b.dst) | (uintptr_t)g.dst_pitch | (uintptr_t)b.dst_frame_stride) & 15) == 0;
        const bool pair_ok =
            ((reinterpret_cast<uintptr_t>(b.dst) | (uintptr_t)g.dst_pitch | (uintptr_t)b.dst_frame_stride) & 1) == 0;
        const unsigned yw = (g.border & 255u) * 0x01010101u;
        const unsigned cw = ((g.border >> 8) & 0xffffu) * 0x00010001u;
        for (int base = (int)blockIdx.x * 32; base < total; base += (int)gridDim.x * 32) {
            unsigned fl = 0;
            if (base + lane < total) fl = __ldg(&table[base + lane].flags);
            unsigned todo = __ballot_sync(0xffffffffu, (fl & kPieceOutside) != 0);
            while (todo) {
                const int idx = base + __ffs((int)todo) - 1;
                todo &= todo - 1;
                const int frame = idx / per_frame, rem = idx - frame * per_frame;
                const int py = rem / npx, px = rem - py * npx;
                const int v_base = py * kPieceHMax;
                const int rows = min(kPieceHMax, g.out_h - v_base);
                uint8_t* dst = b.dst + (size_t)frame * b.dst_frame_stride;
                if (vec_ok && px * kPieceW + kPieceW <= g.out_w) {
                    // 16 bytes per lane: 8 lanes per row, 4 rows per store instruction
                    const int sub = lane >> 3, col = (lane & 7) * 16;
                    uint8_t* yrow = dst + (size_t)(v_base + sub) * g.dst_pitch + px * kPieceW + col;
                    uint8_t* crow = dst + (size_t)(g.out_h + (v_base >> 1) + sub) * g.dst_pitch + px * kPieceW + col;
                    const uint4 y4 = make_uint4(yw, yw, yw, yw), c4 = make_uint4(cw, cw, cw, cw);
                    for (int r = sub; r < rows; r += 4, yrow += 4 * (size_t)g.dst_pitch) *reinterpret_cast<uint4*>(yrow) = y4;
                    for (int r = sub; r < rows / 2; r += 4, crow += 4 * (size_t)g.dst_pitch) *reinterpret_cast<uint4*>(crow) = c4;
                } else if (pair_ok && px * kPieceW + kPieceW <= g.out_w) {
                    // even base / pitch only (e.g. a 2482-byte pitch): 2 bytes per lane, two stores per row
                    uint8_t* yrow = dst + (size_t)v_base * g.dst_pitch + px * kPieceW + 2 * lane;
                    uint8_t* crow = dst + (size_t)(g.out_h + (v_base >> 1)) * g.dst_pitch + px * kPieceW + 2 * lane;
                    for (int r = 0; r < rows; ++r, yrow += g.dst_pitch) {
                        *reinterpret_cast<uint16_t*>(yrow) = (uint16_t)yw;
                        *reinterpret_cast<uint16_t*>(yrow + 64) = (uint16_t)yw;
                    }
                    for (int r = 0; r < rows / 2; ++r, crow += g.dst_pitch) {
                        *reinterpret_cast<uint16_t*>(crow) = (uint16_t)cw;
                        *reinterpret_cast<uint16_t*>(crow + 64) = (uint16_t)cw;
                    }
                } else {
                    const int u0 = px * kPieceW + 4 * lane, valid = g.out_w - u0;
                    uint8_t* yrow = dst + (size_t)v_base * g.dst_pitch + u0;
                    uint8_t* crow = dst + (size_t)(g.out_h + (v_base >> 1)) * g.dst_pitch + u0;
                    if (valid > 0) {
                        for (int r = 0; r < rows; ++r, yrow += g.dst_pitch) store_word<true>(yrow, yw, valid);
                        for (int r = 0; r < rows / 2; ++r, crow += g.dst_pitch) store_word<true>(crow, cw, valid);
                    }
                }
            }
        }
        return;
    }

    // ==================================== consumer groups ========================================
#if VAW_PIPE_CONSUMER_REGS
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(VAW_PIPE_CONSUMER_REGS));  // from the auxiliary warpgroup
#endif
    const int grp = (w - kFirstConsumer) / kGroupWarps, wg = (w - kFirstConsumer) % kGroupWarps, gtid = wg * 32 + lane;
    const bool even_ok = ((reinterpret_cast<uintptr_t>(b.dst) | (uintptr_t)g.dst_pitch | (uintptr_t)b.dst_frame_stride) & 1) == 0;
    const bool word_base_ok = ((reinterpret_cast<uintptr_t>(b.dst) | (uintptr_t)g.dst_pitch | (uintptr_t)b.dst_frame_stride) & 3) == 0;
    float4* const coefs = reinterpret_cast<float4*>(smem + kCoefOffset + grp * kCoefBytes);
    for (int k = grp;; k += kGroups) {
        const int s = k % kSlots;
        pipe_wait(full_bar(s), (unsigned)(k / kSlots) & 1u);  // record copied, tile landed
        const uint8_t* slot = smem + kSlotOffset + s * kSlotBytes;
        const SlotHead head = *reinterpret_cast<const SlotHead*>(slot);
        if (head.idx < 0) break;
        const PieceRec* rec = reinterpret_cast<const PieceRec*>(slot + 32);  // in shared memory
        const unsigned flags = rec->flags;
        const int frame = head.frame;
        const int py = head.pxy >> 16, px = head.pxy & 0xffff;
        const int u_lo = px * kPieceW, v_base = py * kPieceHMax;
        const int rows = min(kPieceHMax, g.out_h - v_base);  // even for NV12
        const int dv0 = wg * kRowsPerWarp;
        const int my_rows = max(0, min(kRowsPerWarp, rows - dv0));
        uint8_t* const dst = b.dst + (size_t)frame * b.dst_frame_stride;
        bool wrote_tile = false;

        if ((head.plmode >> 16) != kModeStaged) {
            direct_piece(g, b, table, head.idx, flags, frame, u_lo, v_base, dv0, my_rows, lane, dst, word_base_ok);
        } else {
            // ---- staged: collapse (warp wg does column slot j = wg), exchange, sample -------------
            {
                float2 a[kNv];
                collapse_column(rec->c, ((float)pair_column(lane, wg) - 63.5f) * 0.015625f, a);  // s is exact
                coefs[(2 * wg) * 32 + lane] = make_float4(a[0].x, a[0].y, a[1].x, a[1].y);
                coefs[(2 * wg + 1) * 32 + lane] = make_float4(a[2].x, a[2].y, a[3].x, a[3].y);
            }
            group_sync(grp);
            ColPoly cp;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 lo = coefs[(2 * j) * 32 + lane], hi = coefs[(2 * j + 1) * 32 + lane];
                cp.a[j][0] = make_float2(lo.x, lo.y); cp.a[j][1] = make_float2(lo.z, lo.w);
                cp.a[j][2] = make_float2(hi.x, hi.y); cp.a[j][3] = make_float2(hi.z, hi.w);
            }
            cp.base = rec->base;
            const int pl = head.plmode & 0xffff;
            const int nrows = head.rows & 0xffff, cnrows = head.rows >> 16;
            const int lx0 = (int16_t)(head.lorg & 0xffff), by0 = head.lorg >> 16;
            const int cbx0 = (int16_t)(head.corg & 0xffff), cy0 = head.corg >> 16;
            uint8_t* ltile = ring + head.tile_off;
            uint8_t* ctile = ltile + nrows * pl;
            if (!(flags & kPieceInterior)) {  // straddles the frame border: paint the outside cells
                const unsigned by_ = g.border & 255u, bu = (g.border >> 8) & 255u, bv = (g.border >> 16) & 255u;
                fill_border(ltile, pl, nrows, by0, g.src_h, lx0, g.src_w, by_ * 0x01010101u, gtid, 32 * kGroupWarps);
                fill_border(ctile, pl, cnrows, cy0, g.src_h >> 1, cbx0, g.src_w, (bu | (bv << 8)) * 0x00010001u,
                            gtid, 32 * kGroupWarps);
                wrote_tile = true;
            }
            group_sync(grp);  // the exchange buffer is free again; the painted cells are visible
            if (my_rows > 0) {
                RowPtrs o;  // the staged path uses the pair mapping: column 2 * lane of the piece
                o.y0 = dst + (size_t)(v_base + dv0) * g.dst_pitch + u_lo + 2 * lane;
                o.y1 = o.y0 + g.dst_pitch;
                o.c = dst + (size_t)(g.out_h + ((v_base + dv0) >> 1)) * g.dst_pitch + u_lo + 2 * lane;
                o.step_y = 2 * (size_t)g.dst_pitch;
                o.step_c = (size_t)g.dst_pitch;
                const unsigned upl = (unsigned)pl;
                const unsigned lconst = smem_u32(ltile) - (unsigned)by0 * upl - (unsigned)lx0 - kMagicShift * upl - kMagicShift;
                const unsigned cconst = ((smem_u32(ctile) - (unsigned)cy0 * upl - (unsigned)cbx0 - kMagicShift * upl) >> 1) - kMagicShift;
                const bool in_a = u_lo + 2 * lane < g.out_w, in_b = u_lo + 64 + 2 * lane < g.out_w;
                const TileBounds tb = {smem_u32(ltile), smem_u32(ltile) + (unsigned)(nrows * pl), smem_u32(ctile),
                                       smem_u32(ctile) + (unsigned)(cnrows * pl)};
                if (even_ok && u_lo + kPieceW <= g.out_w) rows_tile<false>(g, cp, lconst, cconst, upl, dv0, my_rows, o, in_a, in_b, tb);
                else rows_tile<true>(g, cp, lconst, cconst, upl, dv0, my_rows, o, in_a, in_b, tb);
            }
        }
        if (wrote_tile) fence_async_smem();  // generic-proxy writes before the TMA reuses the bytes
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(s));  // this warp is done with the stage
    }
}

int pipe_smem_bytes(int tile_cap) { (void)tile_cap; return kSmemBytes; }

cudaError_t launch_warp_nv12_pipe(const Geom& g, const FrameBatch& b, const PieceRec* table, unsigned* counter,
                                  const TileMaps& maps, cudaStream_t st)
{
    // once per device: the SM count is published before the flag (one host thread per device launches
    // concurrently in the clip scheduler); ordinals beyond the table query and set on every launch
    static std::atomic<int> sm_count[64];
    int dev = 0;
    cudaGetDevice(&dev);
    const bool tracked = dev >= 0 && dev < 64;
    int sms = tracked ? sm_count[dev].load(std::memory_order_acquire) : 0;
    if (sms == 0) {
        cudaError_t e = cudaFuncSetAttribute(warp_nv12_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        if (sms < 1) sms = 1;
        if (tracked) sm_count[dev].store(sms, std::memory_order_release);
    }
    const long long total = (long long)pieces_x(g.out_w) * pieces_y(g.out_h, kPieceHMax) * b.n_frames;
    long long ctas = sms;
    if (ctas > (total + kBatch - 1) / kBatch) ctas = (total + kBatch - 1) / kBatch;
    if (ctas < 1) ctas = 1;
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned), st);
    if (e != cudaSuccess) return e;
    warp_nv12_pipe_kernel<<<(unsigned)ctas, kThreads, kSmemBytes, st>>>(g, b, table, counter, maps);
    return cudaGetLastError();
}

}  // namespace vaw
