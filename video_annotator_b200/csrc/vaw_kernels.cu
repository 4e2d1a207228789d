// vaw_kernels.cu -- fused map + remap kernels for sm_100a (variant GATHER), plus the
// coordinate dump, ray tables, synthetic frames and the math self-test.
//
// Replaces FrameSourceWarp::warp_frame's two device passes
// (/root/reference/opencv/FrameSourceWarp.cpp:272-314): the createMap OpenCL kernel
// (opencv/createMap.cl) and cv::remap.  One launch covers a whole batch of frames,
// luma and chroma planes together; no map buffer exists.
//
// Work decomposition (NV12): a thread owns a 4x2 block of luma pixels -- two 2x2
// quads -- and the two chroma samples under them, because the chroma coordinate is
// the mean of its quad's four luma coordinates (nv12 semantics).  A warp
// is 32 lanes side by side: 128 luma pixels x 2 rows, so every global store
// instruction of the warp writes one full, contiguous 128-byte line (4 bytes per
// lane), and neighbouring lanes gather from neighbouring source bytes.  A CTA is
// 8 warps stacked vertically (128 x 16 luma pixels).  blockIdx.z is the frame, the
// slowest grid dimension, so CTAs of one frame are adjacent in launch order and the
// L2 working set stays about one source frame.
#include <stdint.h>
#include "vaw_internal.h"
#include "vaw_sample.cuh"
#include "vaw_cubic.cuh"
#include "vaw_synth.cuh"
#include "vaw_cvt.cuh"

namespace vaw {

constexpr int kTileW = 128;  // luma pixels per warp row
constexpr int kWarpsPerCta = 8;

__device__ __forceinline__ Rot load_rot(const FrameBatch& b, int frame)
{
    if (b.rots == nullptr) return b.rot0;
    Rot R;
    const float* p = b.rots + (size_t)frame * 9;
#pragma unroll
    for (int i = 0; i < 9; ++i) R.r[i] = __ldg(p + i);
    return R;
}

// Coordinates of the thread's 4x2 luma block: mx[row][col], my[row][col].
template <bool kFast>
__device__ __forceinline__ void block_coords(const Geom& g, const Rot& R, int u0, int v0,
                                             float (&mx)[2][4], float (&my)[2][4])
{
    const float4 xs = __ldg(reinterpret_cast<const float4*>(g.xtab + u0));
    const float2 ys = __ldg(reinterpret_cast<const float2*>(g.ytab + v0));
    const ColTerms c[4] = {col_terms(xs.x, R), col_terms(xs.y, R), col_terms(xs.z, R),
                           col_terms(xs.w, R)};
    const RowTerms w[2] = {row_terms(ys.x, R), row_terms(ys.y, R)};
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) map_eval<kFast>(c[i], w[r], R, g, mx[r][i], my[r][i]);
}

// 4 bytes to dst + off, of which `valid` (0..4) lie inside the image.
__device__ __forceinline__ void store4(uint8_t* p, unsigned word, int valid)
{
    if (valid >= 4 && ((reinterpret_cast<uintptr_t>(p) & 3) == 0)) {
        *reinterpret_cast<unsigned*>(p) = word;
    } else {
        for (int i = 0; i < valid && i < 4; ++i) p[i] = (uint8_t)(word >> (8 * i));
    }
}

template <bool kFast>
__device__ __forceinline__ void warp_nv12_body(const Geom& g, const Rot& R, const uint8_t* src,
                                               uint8_t* dst, int u0, int v0)
{
    float mx[2][4], my[2][4];
    block_coords<kFast>(g, R, u0, v0, mx, my);

    const int border_y = g.border & 255;
    const unsigned border_uv = (g.border >> 8) & 0xffffu;
    const uint8_t* src_uv = src + (size_t)g.src_pitch * g.src_h;
    const int valid = g.out_w - u0;  // >= 1 for active threads; out_w is even

    unsigned yw[2] = {0u, 0u};
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i)
            yw[r] |= (g.cubic_tab ? sample_hi<1>(src, g.src_pitch, g.src_w, g.src_h, mx[r][i], my[r][i], (unsigned)border_y, g.cubic_tab, g.tab_ks)
                                  : (unsigned)sample_c1(src, g.src_pitch, g.src_w, g.src_h,
                                                        g.nearest ? nearest_coord(mx[r][i]) : mx[r][i],
                                                        g.nearest ? nearest_coord(my[r][i]) : my[r][i], border_y))
                     << (8 * i);
    unsigned cw = 0u;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        float cx = chroma_coord(mx[0][2 * q], mx[0][2 * q + 1], mx[1][2 * q], mx[1][2 * q + 1]);
        float cy = chroma_coord(my[0][2 * q], my[0][2 * q + 1], my[1][2 * q], my[1][2 * q + 1]);
        if (g.nearest) { cx = nearest_coord(cx); cy = nearest_coord(cy); }
        cw |= (g.cubic_tab ? sample_hi<2>(src_uv, g.src_pitch, g.src_w >> 1, g.src_h >> 1, cx, cy, border_uv, g.cubic_tab, g.tab_ks)
                           : sample_c2(src_uv, g.src_pitch, g.src_w >> 1, g.src_h >> 1, cx, cy, border_uv))
              << (16 * q);
    }
    if (valid > 0) {
        store4(dst + (size_t)v0 * g.dst_pitch + u0, yw[0], valid);
        store4(dst + (size_t)(v0 + 1) * g.dst_pitch + u0, yw[1], valid);
        store4(dst + (size_t)(g.out_h + (v0 >> 1)) * g.dst_pitch + u0, cw, valid);
    }
}

__global__ void __launch_bounds__(32 * kWarpsPerCta)
warp_nv12_gather_kernel(const Geom g, const FrameBatch b)
{
    const int frame = blockIdx.z;
    const int u_lo = blockIdx.x * kTileW;
    const int u0 = u_lo + threadIdx.x * 4;
    const int v0 = (blockIdx.y * kWarpsPerCta + threadIdx.y) * 2;
    if (v0 >= g.out_h) return;  // warp-uniform
    const Rot R = load_rot(b, frame);
    const uint8_t* src = b.src + (size_t)frame * b.src_frame_stride;
    uint8_t* dst = b.dst + (size_t)frame * b.dst_frame_stride;
    const int u_hi = min(u_lo + kTileW, g.out_w) - 1;
    if (fast_path_ok(u_lo, u_hi, v0, v0 + 1, R, g))
        warp_nv12_body<true>(g, R, src, dst, u0, v0);
    else
        warp_nv12_body<false>(g, R, src, dst, u0, v0);
}

cudaError_t launch_warp_nv12_gather(const Geom& g, const FrameBatch& b, cudaStream_t st)
{
    dim3 block(32, kWarpsPerCta);
    dim3 grid((g.out_w + kTileW - 1) / kTileW, (g.out_h / 2 + kWarpsPerCta - 1) / kWarpsPerCta,
              b.n_frames);
    warp_nv12_gather_kernel<<<grid, block, 0, st>>>(g, b);
    return cudaGetLastError();
}

// ---- interleaved 1- / 3-channel frames (GRAY8, and the reference's literal BGR24) -------
template <bool kFast, int kCn>
__device__ __forceinline__ void warp_packed_body(const Geom& g, const Rot& R, const uint8_t* src,
                                                 uint8_t* dst, int u0, int v0)
{
    float mx[2][4], my[2][4];
    block_coords<kFast>(g, R, u0, v0, mx, my);
    const int valid = min(g.out_w - u0, 4);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (v0 + r >= g.out_h) break;
        uint8_t* row = dst + (size_t)(v0 + r) * g.dst_pitch + (size_t)u0 * kCn;
        if (kCn == 1) {
            unsigned wv = 0u;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                wv |= (g.cubic_tab ? sample_hi<1>(src, g.src_pitch, g.src_w, g.src_h, mx[r][i], my[r][i], g.border & 255u, g.cubic_tab, g.tab_ks)
                                   : (unsigned)sample_c1(src, g.src_pitch, g.src_w, g.src_h,
                                                         g.nearest ? nearest_coord(mx[r][i]) : mx[r][i],
                                                         g.nearest ? nearest_coord(my[r][i]) : my[r][i], g.border & 255))
                      << (8 * i);
            if (valid > 0) store4(row, wv, valid);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                unsigned px = g.cubic_tab
                                  ? sample_hi<3>(src, g.src_pitch, g.src_w, g.src_h, mx[r][i], my[r][i], g.border & 0xffffffu, g.cubic_tab, g.tab_ks)
                                  : sample_c3(src, g.src_pitch, g.src_w, g.src_h, g.nearest ? nearest_coord(mx[r][i]) : mx[r][i],
                                              g.nearest ? nearest_coord(my[r][i]) : my[r][i], g.border & 0xffffffu);
                if (i < valid) {
                    row[3 * i] = (uint8_t)px;
                    row[3 * i + 1] = (uint8_t)(px >> 8);
                    row[3 * i + 2] = (uint8_t)(px >> 16);
                }
            }
        }
    }
}

template <int kCn>
__global__ void __launch_bounds__(32 * kWarpsPerCta)
warp_packed_gather_kernel(const Geom g, const FrameBatch b)
{
    const int frame = blockIdx.z;
    const int u_lo = blockIdx.x * kTileW;
    const int u0 = u_lo + threadIdx.x * 4;
    const int v0 = (blockIdx.y * kWarpsPerCta + threadIdx.y) * 2;
    if (v0 >= g.out_h) return;
    const Rot R = load_rot(b, frame);
    const uint8_t* src = b.src + (size_t)frame * b.src_frame_stride;
    uint8_t* dst = b.dst + (size_t)frame * b.dst_frame_stride;
    const int u_hi = min(u_lo + kTileW, g.out_w) - 1;
    const int v_hi = min(v0 + 1, g.out_h - 1);
    if (fast_path_ok(u_lo, u_hi, v0, v_hi, R, g))
        warp_packed_body<true, kCn>(g, R, src, dst, u0, v0);
    else
        warp_packed_body<false, kCn>(g, R, src, dst, u0, v0);
}

cudaError_t launch_warp_packed_gather(const Geom& g, const FrameBatch& b, int channels,
                                      cudaStream_t st)
{
    dim3 block(32, kWarpsPerCta);
    dim3 grid((g.out_w + kTileW - 1) / kTileW, ((g.out_h + 1) / 2 + kWarpsPerCta - 1) / kWarpsPerCta,
              b.n_frames);
    if (channels == 1) warp_packed_gather_kernel<1><<<grid, block, 0, st>>>(g, b);
    else warp_packed_gather_kernel<3><<<grid, block, 0, st>>>(g, b);
    return cudaGetLastError();
}

// ---- coordinate dump: what createMap.cl:42-49 would have stored ----------------------
__global__ void __launch_bounds__(32 * kWarpsPerCta)
dump_coords_kernel(const Geom g, const Rot R, int plane, float* map_x, float* map_y, int map_pitch)
{
    const int u_lo = blockIdx.x * kTileW;
    const int u0 = u_lo + threadIdx.x * 4;
    const int v0 = (blockIdx.y * kWarpsPerCta + threadIdx.y) * 2;
    if (v0 >= g.out_h) return;
    const int u_hi = min(u_lo + kTileW, g.out_w) - 1;
    const int v_hi = min(v0 + 1, g.out_h - 1);
    float mx[2][4], my[2][4];
    if (fast_path_ok(u_lo, u_hi, v0, v_hi, R, g)) block_coords<true>(g, R, u0, v0, mx, my);
    else block_coords<false>(g, R, u0, v0, mx, my);
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
                size_t o = (size_t)(v0 >> 1) * map_pitch + (u0 >> 1) + q;
                map_x[o] = chroma_coord(mx[0][2 * q], mx[0][2 * q + 1], mx[1][2 * q], mx[1][2 * q + 1]);
                map_y[o] = chroma_coord(my[0][2 * q], my[0][2 * q + 1], my[1][2 * q], my[1][2 * q + 1]);
            }
    }
}

cudaError_t launch_dump_coords(const Geom& g, const Rot& rot, int plane, float* map_x, float* map_y,
                               int map_pitch, cudaStream_t st)
{
    dim3 block(32, kWarpsPerCta);
    dim3 grid((g.out_w + kTileW - 1) / kTileW, ((g.out_h + 1) / 2 + kWarpsPerCta - 1) / kWarpsPerCta, 1);
    dump_coords_kernel<<<grid, block, 0, st>>>(g, rot, plane, map_x, map_y, map_pitch);
    return cudaGetLastError();
}

// ---- unfused sampler: cv::remap with an explicit map -------------------------------------
// The reference's second pass (cv::remap at FrameSourceWarp.cpp:306-312) on its own, for
// callers that already hold a map and for pinning the integer filter against cv::remap's
// golden vectors (NaN / inf / edge coordinates that a camera geometry rarely produces).
template <int kCn>
__global__ void __launch_bounds__(256)
remap_kernel(const uint8_t* __restrict__ src, int src_w, int src_h, int src_pitch,
             const float* __restrict__ map_x, const float* __restrict__ map_y, int rows, int cols,
             int map_pitch, uint8_t* __restrict__ dst, int dst_pitch, unsigned border)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= cols || y >= rows) return;
    const float mx = __ldg(map_x + (size_t)y * map_pitch + x);
    const float my = __ldg(map_y + (size_t)y * map_pitch + x);
    uint8_t* o = dst + (size_t)y * dst_pitch + (size_t)x * kCn;
    if (kCn == 1) {
        o[0] = (uint8_t)sample_c1(src, src_pitch, src_w, src_h, mx, my, border & 255);
    } else if (kCn == 2) {
        unsigned v = sample_c2(src, src_pitch, src_w, src_h, mx, my, border & 0xffffu);
        o[0] = (uint8_t)v;
        o[1] = (uint8_t)(v >> 8);
    } else {
        unsigned v = sample_c3(src, src_pitch, src_w, src_h, mx, my, border & 0xffffffu);
        o[0] = (uint8_t)v;
        o[1] = (uint8_t)(v >> 8);
        o[2] = (uint8_t)(v >> 16);
    }
}

cudaError_t launch_remap(const uint8_t* src, int src_w, int src_h, int src_pitch, int cn,
                         const float* map_x, const float* map_y, int rows, int cols, int map_pitch,
                         uint8_t* dst, int dst_pitch, unsigned border, cudaStream_t st)
{
    dim3 block(256), grid((cols + 255) / 256, rows);
    if (cn == 1)
        remap_kernel<1><<<grid, block, 0, st>>>(src, src_w, src_h, src_pitch, map_x, map_y, rows, cols, map_pitch, dst, dst_pitch, border);
    else if (cn == 2)
        remap_kernel<2><<<grid, block, 0, st>>>(src, src_w, src_h, src_pitch, map_x, map_y, rows, cols, map_pitch, dst, dst_pitch, border);
    else
        remap_kernel<3><<<grid, block, 0, st>>>(src, src_w, src_h, src_pitch, map_x, map_y, rows, cols, map_pitch, dst, dst_pitch, border);
    return cudaGetLastError();
}

// ---- NV12 -> BGR: cv::cvtColor(COLOR_YUV2BGR_NV12) ------------------------------------------
// The conversion the reference runs on every frame before buffering and warping it
// (FrameSourceWarp.cpp:399-401).  OpenCV's 20-bit fixed-point BT.601 (oracle/cvt_ref.c).
// A thread converts 4 x 2 luma samples (two UV pairs): two 4-byte luma loads, one 4-byte chroma
// load, six 4-byte stores; a warp row is 128 pixels = 384 contiguous output bytes.  HBM-bound:
// 1.5 bytes in, 3 bytes out per pixel.
__global__ void __launch_bounds__(256)
nv12_to_bgr_kernel(const uint8_t* __restrict__ src, int w, int h, int src_pitch, size_t src_stride,
                   uint8_t* __restrict__ dst, int dst_pitch, size_t dst_stride, int aligned)
{
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int y0 = (blockIdx.y * 8 + threadIdx.y) * 2;
    if (x0 >= w || y0 >= h) return;
    const uint8_t* s = src + (size_t)blockIdx.z * src_stride;
    uint8_t* d = dst + (size_t)blockIdx.z * dst_stride;
    const uint8_t* uvp = s + (size_t)(h + (y0 >> 1)) * src_pitch + x0;
    unsigned yw[2], uvw;
    const int n = min(4, w - x0);  // w is even: n is 2 or 4
    if (aligned && n == 4) {
        yw[0] = __ldg(reinterpret_cast<const unsigned*>(s + (size_t)y0 * src_pitch + x0));
        yw[1] = __ldg(reinterpret_cast<const unsigned*>(s + (size_t)(y0 + 1) * src_pitch + x0));
        uvw = __ldg(reinterpret_cast<const unsigned*>(uvp));
    } else {
        yw[0] = yw[1] = uvw = 0;
        for (int i = 0; i < n; ++i) {
            yw[0] |= (unsigned)__ldg(s + (size_t)y0 * src_pitch + x0 + i) << (8 * i);
            yw[1] |= (unsigned)__ldg(s + (size_t)(y0 + 1) * src_pitch + x0 + i) << (8 * i);
            uvw |= (unsigned)__ldg(uvp + i) << (8 * i);
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        unsigned px[4][3];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int u = (int)((uvw >> (16 * (i >> 1))) & 255) - 128, v = (int)((uvw >> (16 * (i >> 1) + 8)) & 255) - 128;
            yuv_pixel((int)((yw[r] >> (8 * i)) & 255), u, v, px[i][0], px[i][1], px[i][2]);
        }
        uint8_t* o = d + (size_t)(y0 + r) * dst_pitch + (size_t)x0 * 3;
        if (aligned && n == 4) {
            unsigned* o32 = reinterpret_cast<unsigned*>(o);
            o32[0] = px[0][0] | (px[0][1] << 8) | (px[0][2] << 16) | (px[1][0] << 24);
            o32[1] = px[1][1] | (px[1][2] << 8) | (px[2][0] << 16) | (px[2][1] << 24);
            o32[2] = px[2][2] | (px[3][0] << 8) | (px[3][1] << 16) | (px[3][2] << 24);
        } else {
            for (int i = 0; i < n; ++i)
                for (int c = 0; c < 3; ++c) o[3 * i + c] = (uint8_t)px[i][c];
        }
    }
}

cudaError_t launch_nv12_to_bgr(const uint8_t* src, int w, int h, int src_pitch, size_t src_stride, uint8_t* dst,
                               int dst_pitch, size_t dst_stride, int n_frames, cudaStream_t st)
{
    const int aligned = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)src_pitch | (uintptr_t)src_stride |
                          reinterpret_cast<uintptr_t>(dst) | (uintptr_t)dst_pitch | (uintptr_t)dst_stride) & 3) == 0;
    dim3 block(32, 8), grid((w + 127) / 128, (h / 2 + 7) / 8, n_frames);
    nv12_to_bgr_kernel<<<grid, block, 0, st>>>(src, w, h, src_pitch, src_stride, dst, dst_pitch, dst_stride, aligned);
    return cudaGetLastError();
}

// ---- ray tables ----------------------------------------------------------------------
__global__ void ray_tables_kernel(float* xtab, int n_x, float* ytab, int n_y, float mcx, float mfx,
                                  float mcy, float mfy)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    // the reference's index is `short` (createMap.cl:10-11); sizes are capped at 32766
    if (i < n_x) xtab[i] = ray_component((int)(short)i, mcx, mfx);
    if (i < n_y) ytab[i] = ray_component((int)(short)i, mcy, mfy);
}

cudaError_t launch_ray_tables(float* xtab, int n_x, float* ytab, int n_y, float mcx, float mfx,
                              float mcy, float mfy, cudaStream_t st)
{
    int n = n_x > n_y ? n_x : n_y;
    ray_tables_kernel<<<(n + 255) / 256, 256, 0, st>>>(xtab, n_x, ytab, n_y, mcx, mfx, mcy, mfy);
    return cudaGetLastError();
}

// ---- synthetic frames ------------------------------------------------------------------
__global__ void synth_nv12_kernel(uint8_t* dst, int w, int h, int pitch, size_t frame_stride,
                                  int first_index, uint32_t seed, int white)
{
    const int rows = h + h / 2;
    const int xb = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int row = blockIdx.y;
    const int n = first_index + blockIdx.z;
    if (xb >= w || row >= rows) return;
    const int plane = row >= h;
    const int y = plane ? row - h : row;
    uint8_t* p = dst + (size_t)blockIdx.z * frame_stride + (size_t)row * pitch + xb;
    for (int i = 0; i < 4 && xb + i < w; ++i) p[i] = synth_byte(plane, y, xb + i, n, seed, white);
}

cudaError_t launch_synth_nv12(uint8_t* dst, int w, int h, int pitch, size_t frame_stride,
                              int first_index, int n_frames, uint32_t seed, int white,
                              cudaStream_t st)
{
    dim3 block(128);
    dim3 grid((w / 4 + 127) / 128 + 1, h + h / 2, n_frames);
    synth_nv12_kernel<<<grid, block, 0, st>>>(dst, w, h, pitch, frame_stride, first_index, seed, white);
    return cudaGetLastError();
}

// ---- self-test of the Fast-mode primitives ------------------------------------------------
__device__ __forceinline__ uint32_t rng_next(uint32_t& s)
{
    s ^= s << 13; s ^= s >> 17; s ^= s << 5;
    return s;
}
__device__ __forceinline__ float rng_float(uint32_t& s, int e_lo, int e_hi)
{
    // random mantissa, exponent uniform in [e_lo, e_hi], random sign
    uint32_t m = rng_next(s);
    uint32_t e = (uint32_t)(e_lo + (int)(rng_next(s) % (uint32_t)(e_hi - e_lo + 1)) + 127);
    return __uint_as_float((m & 0x807fffffu) | (e << 23));
}

__global__ void selftest_math_kernel(uint32_t seed, unsigned long long n_per_thread,
                                     unsigned long long* mism)
{
    uint32_t s = seed ^ (0x9E3779B9u * (blockIdx.x * blockDim.x + threadIdx.x + 1));
    if (s == 0) s = 1;
    unsigned long long bad_div = 0, bad_sqrt = 0, bad_rcp = 0, bad_k = 0;
    for (unsigned long long i = 0; i < n_per_thread; ++i) {
        // operand ranges the certificate fast_path_ok() admits
        float b = fabsf(rng_float(s, -6, 5));
        float a = rng_float(s, -40, 5);
        float y = rcp_newton(b);
        bad_rcp += __float_as_uint(y) != __float_as_uint(__frcp_rn(b));
        bad_div += __float_as_uint(div_with_rcp(a, b, y)) != __float_as_uint(__fdiv_rn(a, b));
        float q = fabsf(rng_float(s, -52, 25));
        bad_sqrt += __float_as_uint(sqrt_newton(q)) != __float_as_uint(__fsqrt_rn(q));
        // k = atan(r)/r with the shared reciprocal
        float r = fabsf(rng_float(s, -26, 13));
        float yr = rcp_newton(r);
        bool big = r > 1.0f;
        float kf = div_with_rcp(atan_reduced(big ? yr : r, big), r, yr);
        float ke = __fdiv_rn(vaw_atanf_pos(r), r);
        bad_k += __float_as_uint(kf) != __float_as_uint(ke);
    }
    atomicAdd(&mism[0], bad_rcp);
    atomicAdd(&mism[1], bad_div);
    atomicAdd(&mism[2], bad_sqrt);
    atomicAdd(&mism[3], bad_k);
}

cudaError_t launch_selftest_math(uint32_t seed, unsigned long long n_per_thread,
                                 unsigned long long* mismatches, cudaStream_t st)
{
    selftest_math_kernel<<<148 * 4, 256, 0, st>>>(seed, n_per_thread, mismatches);
    return cudaGetLastError();
}

}  // namespace vaw
