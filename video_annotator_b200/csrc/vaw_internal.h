// vaw_internal.h -- launchers shared between the kernels and the C-ABI layer.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "vaw_coords.cuh"
#include "vaw_pieces.cuh"

namespace vaw {

struct FrameBatch {
    const uint8_t* src;
    uint8_t* dst;
    size_t src_frame_stride, dst_frame_stride;
    const float* rots;  // device, 9 floats per frame, or nullptr -> rot0 for every frame
    Rot rot0;
    int n_frames;
    int skip_interior;  // variant TEX: certified interior pieces are sampled by the texture kernel
    int tma_frame0;     // frame index of src within the clip the tensor maps were encoded for (split batches)
};

// xtab[u] = (u - mcx)/mfx for u < n_x, ytab[v] = (v - mcy)/mfy for v < n_y (createMap.cl:16-17)
cudaError_t launch_ray_tables(float* xtab, int n_x, float* ytab, int n_y, float mcx, float mfx,
                              float mcy, float mfy, cudaStream_t st);

// Fused map + remap, NV12 luma and chroma in one launch (variant GATHER).
cudaError_t launch_warp_nv12_gather(const Geom& g, const FrameBatch& b, cudaStream_t st);
// Fused map + remap for interleaved 1- or 3-channel frames (GRAY8 / BGR24).
cudaError_t launch_warp_packed_gather(const Geom& g, const FrameBatch& b, int channels,
                                      cudaStream_t st);
// Fused map + remap, NV12, coordinates from the per-piece polynomial table (variant POLY).
cudaError_t launch_warp_nv12_poly(const Geom& g, const FrameBatch& b, const PieceRec* table, cudaStream_t st);

// NV12 in, BGR24 out in one launch (vaw_bgr.cu): cvtColor(COLOR_YUV2BGR_NV12) + 3-channel remap, coordinates
// from the piece table.
cudaError_t launch_warp_nv12_to_bgr(const Geom& g, const FrameBatch& b, const PieceRec* table, cudaStream_t st);

// Variant TILED (vaw_tile.cu): tensor maps over the NV12 clip, viewed as a 3-D tensor of 4-byte
// elements (pitch/4 x 3H/2 rows x frames), one per tile row pitch (box = pitch/4 x 8 rows).
constexpr int kTileMinPitch = kStageMinPitch, kTileMaxPitch = kStageMaxPitch, kTilePitchStep = 32;
constexpr int kTileCapMin = 16 << 10, kTileCapMax = 160 << 10;  // per-CTA tile bytes (chosen per geometry)
constexpr int kTileWidths = (kTileMaxPitch - kTileMinPitch) / kTilePitchStep + 1;
struct alignas(64) TileMaps {
    CUtensorMap m[kTileWidths];    // boxes of 8 rows
    CUtensorMap m32[kTileWidths];  // boxes of 32 rows: a tile is loaded as 32-row boxes, then 8-row, then 4-row boxes
    CUtensorMap m4[kTileWidths];   // boxes of 4 rows (tile rows come in multiples of 4)
    int enabled;   // 0: no maps (layout not TMA-compatible) -> every piece gathers from global memory
    int tile_cap;  // bytes of shared memory for the luma + chroma tile of one CTA
    int table_ctas;  // INTER_CUBIC / INTER_LANCZOS4: CTAs per SM the capacity was sized for (the launcher's L1 / shared split)
    int pad[13];
};
// Tile capacity that lets `ctas` CTAs of the tile kernel share one SM: 227 KB of shared memory, 1 KB
// reserved per CTA by the system, 128 bytes of the kernel's own bookkeeping; a multiple of 128.
constexpr int kSmemPerSM = 227 << 10;
inline int tile_cap_for_ctas(int ctas, int bookkeeping) { return ((kSmemPerSM / ctas - 1024 - bookkeeping) / 128) * 128; }
// bytes of tile a piece needs for its source box (what the kernel computes), 0 if it has none
int tile_need_bytes(const PieceRec& rec, int halo = 0);  // halo = GeomD::halo the table was built with
// dynamic shared memory of the tile kernel for a given tile capacity
int tile_smem_bytes(int tile_cap);
// out-of-tile tap count of the instrumented build (-DVAW_BOUNDS_CHECK), -1 when not instrumented
long long tile_oob_count();
cudaError_t launch_warp_nv12_tile(const Geom& g, const FrameBatch& b, const PieceRec* table, const TileMaps& maps,
                                  cudaStream_t st);
// Quadrant kernel for interleaved GRAY8 / BGR24 frames (vaw_packed_tile.cu): tensor maps over the clip viewed as
// (pitch / 8, H, frames) 8-byte elements, one per tile row pitch (128 ... 1536 bytes in steps of 64), boxes of 16 and of 4 rows.
constexpr int kPackedMinPitch = 128, kPackedMaxPitch = 1536, kPackedPitchStep = 64;
constexpr int kPackedWidths = (kPackedMaxPitch - kPackedMinPitch) / kPackedPitchStep + 1;
struct alignas(64) PackedMaps {
    CUtensorMap m16[kPackedWidths];
    CUtensorMap m4[kPackedWidths];
    int enabled;   // 0: layout not TMA-compatible -> every piece is sampled per pixel from global memory
    int tile_cap;  // bytes of shared memory for the tile of one CTA
    int table_ctas;  // INTER_CUBIC / INTER_LANCZOS4: CTAs per SM the capacity was sized for (the launcher's L1 / shared split)
    int pad[13];
};
int packed_tile_need_bytes(const PieceRec& rec, int channels);
int packed_tile_smem_bytes(int tile_cap);
long long packed_tile_oob_count();  // -1 when not instrumented
cudaError_t launch_warp_packed_tile(const Geom& g, const FrameBatch& b, const PieceRec* table, const PackedMaps& maps,
                                    int channels, cudaStream_t st);
// Variant TEX (vaw_tex.cu): certified interior pieces are filtered by the texture units.  The clip
// is viewed as pitch-linear 2-D textures over groups of `group_frames` whole frames (a texture is at
// most 65000 rows high): y[k] = 8-bit luma view, uv[k] = 2 x 8-bit view of the same rows (chroma
// texel x of global row r = bytes 2x, 2x+1 of that row).  Frame f sits at texture rows
// (f % group_frames) * frame_rows ...; its chroma plane starts src_h rows further down.
constexpr int kTexGroups = 32;
struct alignas(64) TexSet {
    unsigned long long y[kTexGroups], uv[kTexGroups];
    int enabled;       // 0: layout not texturable -> every piece takes the TILED path
    int group_frames;  // frames per texture
    int frame_rows;    // texture rows from one frame to the next (frame stride / pitch)
    int pad[13];
};
cudaError_t launch_warp_nv12_tex(const Geom& g, const FrameBatch& b, const PieceRec* table, const TexSet& ts,
                                 cudaStream_t st);

// The map the POLY kernel samples with (table built for `rot`, one frame).
cudaError_t launch_dump_coords_poly(const Geom& g, const Rot& rot, const PieceRec* table, int plane,
                                    float* map_x, float* map_y, int map_pitch, cudaStream_t st);
// The map createMap.cl would write; plane 0 luma, plane 1 NV12 chroma.
cudaError_t launch_dump_coords(const Geom& g, const Rot& rot, int plane, float* map_x, float* map_y,
                               int map_pitch, cudaStream_t st);
// cv::remap(INTER_LINEAR, BORDER_CONSTANT) with an explicit map, cn = 1..3.
cudaError_t launch_remap(const uint8_t* src, int src_w, int src_h, int src_pitch, int cn,
                         const float* map_x, const float* map_y, int rows, int cols, int map_pitch,
                         uint8_t* dst, int dst_pitch, unsigned border, cudaStream_t st);
// cv::cvtColor(COLOR_YUV2BGR_NV12) for n_frames frames (FrameSourceWarp.cpp:399-401).
cudaError_t launch_nv12_to_bgr(const uint8_t* src, int w, int h, int src_pitch, size_t src_stride, uint8_t* dst,
                               int dst_pitch, size_t dst_stride, int n_frames, cudaStream_t st);
// Integer synthetic content (mirror of oracle/synth_ref.c).
cudaError_t launch_synth_nv12(uint8_t* dst, int w, int h, int pitch, size_t frame_stride,
                              int first_index, int n_frames, uint32_t seed, int white,
                              cudaStream_t st);
// Device self-test of the Fast-mode primitives against the IEEE intrinsics.
cudaError_t launch_selftest_math(uint32_t seed, unsigned long long n_per_thread,
                                 unsigned long long* mismatches /* device, 4 counters */,
                                 cudaStream_t st);

}  // namespace vaw
