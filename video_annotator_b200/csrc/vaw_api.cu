// vaw_api.cu -- the C-ABI of libvaw.so (declared in include/vaw.h).
//
// Replaces the state and the call FrameSourceWarp builds/makes for its warp
// (/root/reference/opencv/FrameSourceWarp.cpp:199-226 constructor, :272-314 warp_frame):
// a context holds the camera scalars cast to float exactly where the reference casts to
// cl_float (:283-299), the per-column/per-row ray tables, and -- for the host-buffer
// entry point -- a small ring of device staging buffers and streams.  No map buffers.
// There is no CPU fallback anywhere in this file.
#include <cuda.h>
#include <cuda_runtime.h>
#include <sched.h>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <utility>
#include "../../include/vaw.h"
#include "vaw_internal.h"
#include "vaw_cubic.cuh"

static_assert(sizeof(vaw_params) == 128 && sizeof(vaw_camera) == 120, "C-ABI struct layout");

namespace {

constexpr int kStages = 8;                      // host-path pipeline depth: upper bound (ctx->host_stages are used)
constexpr int kHeadFrames = 8;                  // frames whose table is built in line; the rest overlaps their sampling
constexpr size_t kChunkBytes = 32u << 20;       // default target bytes of source frames per chunk (small: short pipeline fill and drain)
thread_local std::string g_create_error;

struct Stage {
    cudaStream_t stream = nullptr;
    uint8_t *dev_in = nullptr, *dev_out = nullptr, *pin_in = nullptr, *pin_out = nullptr;
    float *dev_rot = nullptr, *pin_rot = nullptr;
    vaw::PieceRec* pieces = nullptr;  // this stage's polynomial table (chunk_frames frames)
    // pending output of the chunk in flight on this stage
    uint8_t* host_dst = nullptr;
    size_t out_bytes = 0;
    bool busy = false, out_staged = false;
};

}  // namespace

struct vaw_ctx {
    vaw_params p{};
    int device = 0;
    vaw::Geom g{};
    float *xtab = nullptr, *ytab = nullptr;
    int16_t* cubic_tab = nullptr;  // INTER_CUBIC weights (device)
    int channels = 1;      // bytes per pixel of a SOURCE row (NV12 and GRAY8: 1, BGR24: 3)
    int dst_channels = 1;  // ... of an OUTPUT row (differs for VAW_FORMAT_NV12_TO_BGR24: NV12 in, BGR24 out)
    int src_format = 0, dst_format = 0;  // VAW_FORMAT_NV12 / BGR24 / GRAY8 of the two sides
    size_t src_frame_bytes = 0, dst_frame_bytes = 0;  // tightly packed
    uint64_t launches = 0;
    std::string err;
    // variant POLY: per-piece polynomial tables (vaw_pieces.cuh)
    int variant = VAW_VARIANT_GATHER;  // resolved (never AUTO)
    vaw::GeomD gd{};
    vaw::PieceBasis basis{};
    size_t pieces_per_frame = 0;
    vaw::PieceRec* table = nullptr;    // for vaw_warp / vaw_warp_batch
    size_t table_frames = 0;
    cudaEvent_t table_free = nullptr;  // recorded after the last kernel that read `table`
    cudaStream_t table_stream = nullptr;
    bool table_used = false;
    vaw::PieceRec* dump_table = nullptr;
    // NV12 -> BGR24, variant TILED: converted frames of the chunk in flight (kept small enough to stay in L2)
    uint8_t* bgr_scratch = nullptr;
    int bgr_chunk = 0, bgr_pitch = 0;
    size_t bgr_stride = 0;
    // variant TILED: tensor maps, cached per source layout (encoding 11 maps costs ~10 us)
    struct MapEntry { const void* src = nullptr; int pitch = 0; size_t stride = 0; int frames = 0; vaw::TileMaps maps{}; };
    MapEntry map_cache[4];
    struct PackedEntry { const void* src = nullptr; int pitch = 0; size_t stride = 0; int frames = 0; vaw::PackedMaps maps{}; };
    PackedEntry packed_cache[2];  // the same for GRAY8 / BGR24 clips (vaw_packed_tile.cu)
    int packed_next = 0;
    int map_next = 0;
    int tile_cap = 32 << 10;  // chosen at creation from the pieces' source boxes
    int table_ctas = 0;       // INTER_CUBIC / INTER_LANCZOS4 on staged tiles: CTAs per SM the tile capacity was sized for
    // vaw_bind_clip: a slab of equally spaced source frames; launches inside it share its tensor maps
    const uint8_t* clip_base = nullptr;
    int clip_pitch = 0, clip_slots = 0;
    size_t clip_stride = 0;
    long long tile_need = 0;  // largest tile a piece of the unrotated geometry needs (bytes)
    // variant TEX: texture objects over the clip, cached per source layout
    struct TexEntry { const void* src = nullptr; int pitch = 0; size_t stride = 0; int frames = 0; vaw::TexSet set{}; int n_groups = 0; };
    static constexpr int kTexCache = 8;
    TexEntry tex_cache[kTexCache];
    int tex_next = 0;
    int tex_align = 512, tex_pitch_align = 32;
    // builder off the critical path: the table of all but the first kHeadFrames frames of a batch is
    // built on a high-priority side stream while the sampler already works on the head frames
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool split_builder = false;  // option "split_builder": measured neutral on B200 (the two kernels share the issue slots)
    // option "time_kernels": CUDA-event stamps around the kernels of every launch (bench.py's roofline)
    static constexpr int kTimeRing = 512;
    bool time_kernels = false;
    cudaEvent_t* tev = nullptr;  // 4 per launch: start, after the table builder, after the texture kernel, after the warp kernel
    static constexpr int kTev = 4;
    uint64_t timed_launches = 0;
    // host path
    Stage stage[kStages];
    int chunk_frames = 0;
    int host_stages = 4;                  // chunks in flight (option "host_stages", 2..8)
    size_t host_chunk_bytes = kChunkBytes;  // option "host_chunk_mb"
    bool host_ready = false;
};

namespace {

int fail(vaw_ctx* ctx, int code, const std::string& msg)
{
    if (ctx) ctx->err = msg;
    else g_create_error = msg;
    return code;
}

int cuda_fail(vaw_ctx* ctx, cudaError_t e, const char* what)
{
    return fail(ctx, VAW_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define VAW_CUDA(ctx, call)                                        \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call);   \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

vaw::Rot rot_from_double(const double r[9])
{
    vaw::Rot R;
    for (int i = 0; i < 9; ++i) R.r[i] = (float)r[i];  // the (cl_float) cast, FrameSourceWarp.cpp:291-299
    return R;
}

bool centre_ok(float c) { return c == 0.0f || std::fabs(c) >= 8.67361738e-19f /* 2^-60 */; }

int check_buffers(vaw_ctx* ctx, const void* src, int src_pitch, void* dst, int dst_pitch)
{
    if (!ctx) return VAW_ERR_INVALID;
    if (!src || !dst) return fail(ctx, VAW_ERR_INVALID, "null frame pointer");
    if (src_pitch < ctx->p.src_width * ctx->channels || dst_pitch < ctx->p.out_width * ctx->dst_channels)
        return fail(ctx, VAW_ERR_INVALID, "pitch smaller than a row");
    if (ctx->src_format == VAW_FORMAT_NV12 &&
        ((src_pitch & 1) || (reinterpret_cast<uintptr_t>(src) & 1)))
        return fail(ctx, VAW_ERR_INVALID, "NV12 source base and pitch must be even");
    return VAW_OK;
}

// Inverse of the Vandermonde matrix of `n` nodes: monomial coefficient i = sum_a inv[i][a] f(node a).
template <int N>
void invert_vandermonde(const long double (&node)[N], double (&inv)[N][N])
{
    long double a[N][2 * N];
    for (int r = 0; r < N; ++r) {
        long double p = 1.0L;
        for (int c = 0; c < N; ++c) { a[r][c] = p; p *= node[r]; }
        for (int c = 0; c < N; ++c) a[r][N + c] = (r == c) ? 1.0L : 0.0L;
    }
    for (int col = 0; col < N; ++col) {
        int piv = col;
        for (int r = col + 1; r < N; ++r)
            if (fabsl(a[r][col]) > fabsl(a[piv][col])) piv = r;
        for (int c = 0; c < 2 * N; ++c) std::swap(a[col][c], a[piv][c]);
        const long double d = a[col][col];
        for (int c = 0; c < 2 * N; ++c) a[col][c] /= d;
        for (int r = 0; r < N; ++r)
            if (r != col) {
                const long double f = a[r][col];
                for (int c = 0; c < 2 * N; ++c) a[r][c] -= f * a[col][c];
            }
    }
    // a[:, N:] = V^-1 with V[r][c] = node_r^c, i.e. coef = V^-1 f
    for (int i = 0; i < N; ++i)
        for (int r = 0; r < N; ++r) inv[i][r] = (double)a[i][N + r];
}

void make_basis(vaw::PieceBasis& b, int ph)
{
    long double su[vaw::kNu], sv[vaw::kNv];
    for (int a = 0; a < vaw::kNu; ++a) su[a] = ((128.0L * a) / vaw::kDegU - 63.5L) / 64.0L;
    for (int a = 0; a < vaw::kNv; ++a) sv[a] = (((long double)ph * a) / vaw::kDegV - 0.5L * (ph - 1)) * (2.0L / ph);
    invert_vandermonde<vaw::kNu>(su, b.mu);
    invert_vandermonde<vaw::kNv>(sv, b.mv);
}

// Make ctx->table hold `n_frames` frames and safe to overwrite from stream `st`.
int acquire_table(vaw_ctx* ctx, int n_frames, cudaStream_t st)
{
    if ((size_t)n_frames > ctx->table_frames) {
        if (ctx->table) {
            VAW_CUDA(ctx, cudaDeviceSynchronize());
            cudaFree(ctx->table);
            ctx->table = nullptr;
            ctx->table_frames = 0;
        }
        size_t want = (size_t)n_frames < 8 ? 8 : (size_t)n_frames;
        VAW_CUDA(ctx, cudaMalloc(&ctx->table, want * ctx->pieces_per_frame * sizeof(vaw::PieceRec)));
        ctx->table_frames = want;
        ctx->table_used = false;
    }
    // the table is rebuilt by every call: order it after the previous reader when streams differ
    if (ctx->table_used && ctx->table_stream != st) VAW_CUDA(ctx, cudaStreamWaitEvent(st, ctx->table_free, 0));
    return VAW_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled()
{
    // function-local static: initialised once, thread-safe (several host threads create contexts
    // concurrently in the clip scheduler)
    static const EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            return reinterpret_cast<EncodeTiledFn>(p);
        cudaGetLastError();
        return nullptr;
    }();
    return fn;
}

// Tensor maps for a clip of `frames` NV12 frames at `src` (vaw_tile.cu).  The TMA engine needs a
// 16-byte aligned base, pitch and frame stride; other layouts get maps.enabled = 0.
const vaw::TileMaps& tile_maps(vaw_ctx* ctx, const uint8_t* src, int pitch, size_t stride, int frames)
{
    for (vaw_ctx::MapEntry& e : ctx->map_cache)
        if (e.src == src && e.pitch == pitch && e.stride == stride && e.frames == frames) return e.maps;
    vaw_ctx::MapEntry& e = ctx->map_cache[ctx->map_next];
    ctx->map_next = (ctx->map_next + 1) % 4;
    e.src = src; e.pitch = pitch; e.stride = stride; e.frames = frames;
    e.maps.enabled = 0;
    e.maps.tile_cap = ctx->tile_cap;
    e.maps.table_ctas = ctx->table_ctas;
    const int rows_total = ctx->p.src_height + ctx->p.src_height / 2;
    EncodeTiledFn enc = encode_tiled();
    const bool ok = enc && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (pitch & 15) == 0 && (stride & 15) == 0 &&
                    pitch / 4 >= vaw::kTileMaxPitch / 4 && rows_total >= 32 && (frames == 1 || stride >= (size_t)pitch);
    if (!ok) return e.maps;
    const cuuint64_t dims[3] = {(cuuint64_t)(pitch / 4), (cuuint64_t)rows_total, (cuuint64_t)frames};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch,
                                   (cuuint64_t)(stride ? stride : (size_t)pitch * (size_t)rows_total)};
    const cuuint32_t estr[3] = {1, 1, 1};
    for (int i = 0; i < vaw::kTileWidths; ++i) {
        const cuuint32_t box[3] = {(cuuint32_t)((vaw::kTileMinPitch + i * vaw::kTilePitchStep) / 4), 8, 1};
        CUresult r = enc(&e.maps.m[i], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t*>(src), dims, strides,
                         box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return e.maps;
        const cuuint32_t box32[3] = {box[0], 32, 1};
        r = enc(&e.maps.m32[i], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t*>(src), dims, strides,
                box32, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return e.maps;
        const cuuint32_t box4[3] = {box[0], 4, 1};
        r = enc(&e.maps.m4[i], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t*>(src), dims, strides,
                box4, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return e.maps;
    }
    e.maps.enabled = 1;
    return e.maps;
}

// Tensor maps for a clip of `frames` interleaved GRAY8 / BGR24 frames at `src` (vaw_packed_tile.cu).
const vaw::PackedMaps& packed_maps(vaw_ctx* ctx, const uint8_t* src, int pitch, size_t stride, int frames)
{
    for (vaw_ctx::PackedEntry& e : ctx->packed_cache)
        if (e.src == src && e.pitch == pitch && e.stride == stride && e.frames == frames && e.maps.tile_cap == ctx->tile_cap) return e.maps;
    vaw_ctx::PackedEntry& e = ctx->packed_cache[ctx->packed_next];
    ctx->packed_next = (ctx->packed_next + 1) % 2;
    e.src = src; e.pitch = pitch; e.stride = stride; e.frames = frames;
    e.maps.enabled = 0;
    e.maps.tile_cap = ctx->tile_cap;
    e.maps.table_ctas = ctx->table_ctas;
    const int rows_total = ctx->p.src_height;
    EncodeTiledFn enc = encode_tiled();
    const bool ok = enc && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (pitch & 15) == 0 && (stride & 15) == 0 &&
                    pitch >= 16 && rows_total >= 1 && (frames == 1 || stride >= (size_t)pitch);
    if (!ok) return e.maps;
    const cuuint64_t dims[3] = {(cuuint64_t)(pitch / 8), (cuuint64_t)rows_total, (cuuint64_t)frames};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch,
                                   (cuuint64_t)(stride ? stride : (size_t)pitch * (size_t)rows_total)};
    const cuuint32_t estr[3] = {1, 1, 1};
    for (int i = 0; i < vaw::kPackedWidths; ++i) {
        const cuuint32_t box16[3] = {(cuuint32_t)((vaw::kPackedMinPitch + i * vaw::kPackedPitchStep) / 8), 16, 1};
        CUresult r = enc(&e.maps.m16[i], CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<uint8_t*>(src), dims, strides,
                         box16, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return e.maps;
        const cuuint32_t box4[3] = {box16[0], 4, 1};
        r = enc(&e.maps.m4[i], CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<uint8_t*>(src), dims, strides,
                box4, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return e.maps;
    }
    e.maps.enabled = 1;
    return e.maps;
}

void destroy_tex_entry(vaw_ctx::TexEntry& e)
{
    for (int k = 0; k < e.n_groups; ++k) {
        if (e.set.y[k]) cudaDestroyTextureObject((cudaTextureObject_t)e.set.y[k]);
        if (e.set.uv[k]) cudaDestroyTextureObject((cudaTextureObject_t)e.set.uv[k]);
    }
    e = vaw_ctx::TexEntry{};
}

// Texture objects for a clip of `frames` NV12 frames at `src` (vaw_tex.cu): pitch-linear 2-D
// textures over groups of whole frames.  Needs the texture base / pitch alignment of the device and
// a frame stride that is a whole number of rows; other layouts get set.enabled = 0.  At most
// kTexGroups * group_frames frames per launch (the caller splits longer clips).
const vaw::TexSet& tex_set(vaw_ctx* ctx, const uint8_t* src, int pitch, size_t stride, int frames)
{
    for (vaw_ctx::TexEntry& e : ctx->tex_cache)
        if (e.src == src && e.pitch == pitch && e.stride == stride && e.frames == frames) return e.set;
    vaw_ctx::TexEntry& e = ctx->tex_cache[ctx->tex_next];
    ctx->tex_next = (ctx->tex_next + 1) % vaw_ctx::kTexCache;
    if (e.n_groups) {
        cudaDeviceSynchronize();  // a kernel in flight may still read the evicted objects
        destroy_tex_entry(e);
    }
    e.src = src; e.pitch = pitch; e.stride = stride; e.frames = frames;
    e.set.enabled = 0;
    const int rows_total = ctx->p.src_height + ctx->p.src_height / 2;
    const size_t fstride = frames > 1 ? stride : (size_t)pitch * rows_total;
    if (pitch <= 0 || (pitch % ctx->tex_pitch_align) || (reinterpret_cast<uintptr_t>(src) % ctx->tex_align) ||
        fstride % (size_t)pitch || fstride / (size_t)pitch < (size_t)rows_total || fstride / (size_t)pitch > 65000)
        return e.set;
    const int frame_rows = (int)(fstride / (size_t)pitch);
    // group bases must keep the texture alignment: group_frames a multiple of align / gcd(stride, align)
    size_t gcd = fstride, al = (size_t)ctx->tex_align;
    while (al) { size_t t = gcd % al; gcd = al; al = t; }
    const int unit = (int)((size_t)ctx->tex_align / gcd);
    int gf = (65000 / frame_rows) / unit * unit;
    if (gf < 1) return e.set;
    if (gf > frames) gf = frames;  // a single (possibly short) group
    const int n_groups = (frames + gf - 1) / gf;
    if (n_groups > vaw::kTexGroups) {
        // longer clips are split by the caller into launches of kTexGroups * gf frames
    }
    const int groups_here = n_groups < vaw::kTexGroups ? n_groups : vaw::kTexGroups;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 0;
    for (int k = 0; k < groups_here; ++k) {
        const int f0 = k * gf, nf = frames - f0 < gf ? frames - f0 : gf;
        cudaResourceDesc rd{};
        rd.resType = cudaResourceTypePitch2D;
        rd.res.pitch2D.devPtr = const_cast<uint8_t*>(src) + (size_t)f0 * fstride;
        rd.res.pitch2D.pitchInBytes = (size_t)pitch;
        rd.res.pitch2D.height = (size_t)(nf - 1) * frame_rows + rows_total;
        rd.res.pitch2D.width = (size_t)ctx->p.src_width;
        rd.res.pitch2D.desc = cudaCreateChannelDesc<unsigned char>();
        cudaTextureObject_t ty = 0, tc = 0;
        cudaError_t err = cudaCreateTextureObject(&ty, &rd, &td, nullptr);
        if (err == cudaSuccess) {
            rd.res.pitch2D.width = (size_t)ctx->p.src_width / 2;
            rd.res.pitch2D.desc = cudaCreateChannelDesc<uchar2>();
            err = cudaCreateTextureObject(&tc, &rd, &td, nullptr);
        }
        e.set.y[k] = ty; e.set.uv[k] = tc;
        e.n_groups = k + 1;
        if (err != cudaSuccess) {
            cudaGetLastError();
            const void* s0 = e.src; const int p0 = e.pitch; const size_t st0 = e.stride; const int fr0 = e.frames;
            destroy_tex_entry(e);
            e.src = s0; e.pitch = p0; e.stride = st0; e.frames = fr0;
            return e.set;
        }
    }
    e.set.group_frames = gf;
    e.set.frame_rows = frame_rows;
    e.set.enabled = 1;
    return e.set;
}

int launch(vaw_ctx* ctx, const uint8_t* src, int src_pitch, size_t src_stride, uint8_t* dst,
           int dst_pitch, size_t dst_stride, const float* rots, const vaw::Rot* rot0, int n_frames,
           cudaStream_t st, vaw::PieceRec* table_override = nullptr)
{
    vaw::Geom g = ctx->g;
    g.src_pitch = src_pitch;
    g.dst_pitch = dst_pitch;
    vaw::FrameBatch b{};
    b.src = src;
    b.dst = dst;
    b.src_frame_stride = src_stride;
    b.dst_frame_stride = dst_stride;
    b.rots = rots;
    if (rot0) b.rot0 = *rot0;
    const bool fused_bgr = ctx->p.format == VAW_FORMAT_NV12_TO_BGR24;
    const bool packed = ctx->p.format == VAW_FORMAT_BGR24 || ctx->p.format == VAW_FORMAT_GRAY8;
    const bool poly = fused_bgr || (ctx->p.format == VAW_FORMAT_NV12 && ctx->variant != VAW_VARIANT_GATHER) ||
                      (packed && ctx->variant == VAW_VARIANT_TILED);
    const bool tiled = ctx->variant == VAW_VARIANT_TILED || ctx->variant == VAW_VARIANT_TEX;
    // grid.z is limited to 65535 frames per launch; variant TEX to kTexGroups textures of <= 65000 rows
    int per_launch = 65535;
    if (ctx->variant == VAW_VARIANT_TEX) {
        const int rows_total = ctx->p.src_height + ctx->p.src_height / 2;
        const size_t fr = src_stride && src_pitch ? src_stride / (size_t)src_pitch : (size_t)rows_total;
        const long long cap = (long long)vaw::kTexGroups * (65000 / (long long)(fr ? fr : 1));
        // keep launches a multiple of 512 frames so that every launch's group bases stay aligned alike
        if (cap >= 512 && cap < per_launch) per_launch = (int)(cap / 512 * 512);
        else if (cap >= 1 && cap < per_launch) per_launch = (int)cap;
    }
    for (int first = 0; first < n_frames; first += per_launch) {
        vaw::FrameBatch bb = b;
        bb.n_frames = n_frames - first < per_launch ? n_frames - first : per_launch;
        bb.src = src + (size_t)first * src_stride;
        bb.dst = dst + (size_t)first * dst_stride;
        if (rots) bb.rots = rots + (size_t)first * 9;
        cudaError_t e;
        if (poly) {
            vaw::PieceRec* tab = table_override;
            if (!tab) {
                int rc = acquire_table(ctx, bb.n_frames, st);
                if (rc) return rc;
                tab = ctx->table;
            }
            cudaEvent_t* tev = ctx->time_kernels ? ctx->tev + vaw_ctx::kTev * (ctx->timed_launches % vaw_ctx::kTimeRing) : nullptr;
            if (tev) cudaEventRecord(tev[0], st);
            // Split: [builder(head) -> sampler(head)] on `st`, builder(rest) meanwhile on the side stream,
            // then sampler(rest) on `st`.  The side stream starts after everything already queued on `st`
            // (the rotations may come from there) and is ordered before the second sampler by an event.
            const bool split = ctx->split_builder && ctx->side && !table_override && ctx->variant != VAW_VARIANT_TEX &&
                               bb.n_frames >= 4 * kHeadFrames;
            const int head = split ? kHeadFrames : bb.n_frames;
            if (split) {
                VAW_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));
                VAW_CUDA(ctx, cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
                e = vaw::launch_build_pieces(ctx->gd, ctx->basis, bb.rots ? bb.rots + (size_t)head * 9 : nullptr,
                                             rot0 ? rot0->r : nullptr, bb.n_frames - head,
                                             tab + (size_t)head * ctx->pieces_per_frame, ctx->side);
                if (e != cudaSuccess) return cuda_fail(ctx, e, "piece table launch (side stream)");
                ctx->launches++;
                VAW_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->side));
            }
            e = vaw::launch_build_pieces(ctx->gd, ctx->basis, bb.rots, rot0 ? rot0->r : nullptr, head, tab, st);
            if (e != cudaSuccess) return cuda_fail(ctx, e, "piece table launch");
            ctx->launches++;
            if (tev) cudaEventRecord(tev[1], st);
            if (ctx->variant == VAW_VARIANT_TEX) {
                // certified interior pieces through the texture units; the rest through the tile kernel
                const vaw::TexSet& ts = tex_set(ctx, bb.src, src_pitch, src_stride, bb.n_frames);
                if (ts.enabled && ts.group_frames * vaw::kTexGroups >= bb.n_frames) {
                    e = vaw::launch_warp_nv12_tex(g, bb, tab, ts, st);
                    if (e != cudaSuccess) return cuda_fail(ctx, e, "texture kernel launch");
                    ctx->launches++;
                    bb.skip_interior = 1;
                }
            }
            if (tev) cudaEventRecord(tev[2], st);
            // source inside the bound slab: its tensor maps (encoded once) + the slot index of the first frame
            int clip_frame0 = -1;
            if (ctx->clip_base && bb.src >= ctx->clip_base && src_pitch == ctx->clip_pitch &&
                (bb.n_frames == 1 || src_stride == ctx->clip_stride)) {
                const size_t off = (size_t)(bb.src - ctx->clip_base);
                if (off % ctx->clip_stride == 0 && off / ctx->clip_stride + (size_t)bb.n_frames <= (size_t)ctx->clip_slots)
                    clip_frame0 = (int)(off / ctx->clip_stride);
            }
            const vaw::TileMaps* tm = nullptr;
            const vaw::PackedMaps* pm = nullptr;
            if (packed)
                pm = clip_frame0 >= 0 ? &packed_maps(ctx, ctx->clip_base, ctx->clip_pitch, ctx->clip_stride, ctx->clip_slots)
                                      : &packed_maps(ctx, bb.src, src_pitch, src_stride, bb.n_frames);
            else if (tiled)
                tm = clip_frame0 >= 0 ? &tile_maps(ctx, ctx->clip_base, ctx->clip_pitch, ctx->clip_stride, ctx->clip_slots)
                                      : &tile_maps(ctx, bb.src, src_pitch, src_stride, bb.n_frames);
            if (clip_frame0 >= 0) bb.tma_frame0 = clip_frame0;
            for (int part = 0; part < (split ? 2 : 1); ++part) {
                vaw::FrameBatch pb = bb;
                const vaw::PieceRec* ptab = tab;
                if (split) {
                    pb.n_frames = part ? bb.n_frames - head : head;
                    if (part) {
                        // everything is re-based on the first frame of the part, except the tensor maps,
                        // which stay encoded over the whole clip (frame coordinate offset by tma_frame0)
                        pb.src = bb.src + (size_t)head * src_stride;
                        pb.dst = bb.dst + (size_t)head * dst_stride;
                        if (bb.rots) pb.rots = bb.rots + (size_t)head * 9;
                        pb.tma_frame0 = bb.tma_frame0 + head;
                        ptab = tab + (size_t)head * ctx->pieces_per_frame;
                        VAW_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_join, 0));
                    }
                }
                if (packed) e = vaw::launch_warp_packed_tile(g, pb, ptab, *pm, ctx->channels, st);
                else if (fused_bgr && tiled) {
                    // cvtColor into a scratch of a few frames (<= 96 MB: it stays in the 126 MB L2), then the staged BGR
                    // kernel on the scratch -- the reference's own order of operations (FrameSourceWarp.cpp:399-401,
                    // :306-312), chunk by chunk on one stream
                    if (!ctx->bgr_scratch) {
                        ctx->bgr_pitch = ((ctx->p.src_width * 3 + 15) / 16) * 16;
                        ctx->bgr_stride = (size_t)ctx->bgr_pitch * (size_t)ctx->p.src_height;
                        long long k = (96ll << 20) / (long long)ctx->bgr_stride;
                        ctx->bgr_chunk = (int)(k < 1 ? 1 : (k > 16 ? 16 : k));
                        VAW_CUDA(ctx, cudaMalloc(&ctx->bgr_scratch, ctx->bgr_stride * (size_t)ctx->bgr_chunk));
                    }
                    const vaw::PackedMaps& sm = packed_maps(ctx, ctx->bgr_scratch, ctx->bgr_pitch, ctx->bgr_stride, ctx->bgr_chunk);
                    vaw::Geom g2 = g;
                    g2.src_pitch = ctx->bgr_pitch;
                    e = cudaSuccess;
                    for (int f = 0; f < pb.n_frames && e == cudaSuccess; f += ctx->bgr_chunk) {
                        const int k = pb.n_frames - f < ctx->bgr_chunk ? pb.n_frames - f : ctx->bgr_chunk;
                        e = vaw::launch_nv12_to_bgr(pb.src + (size_t)f * src_stride, ctx->p.src_width, ctx->p.src_height, src_pitch,
                                                    src_stride, ctx->bgr_scratch, ctx->bgr_pitch, ctx->bgr_stride, k, st);
                        if (e != cudaSuccess) break;
                        ctx->launches++;
                        vaw::FrameBatch cb = pb;
                        cb.src = ctx->bgr_scratch;
                        cb.src_frame_stride = ctx->bgr_stride;
                        cb.dst = pb.dst + (size_t)f * dst_stride;
                        if (pb.rots) cb.rots = pb.rots + (size_t)f * 9;
                        cb.n_frames = k;
                        cb.tma_frame0 = 0;
                        e = vaw::launch_warp_packed_tile(g2, cb, ptab + (size_t)f * ctx->pieces_per_frame, sm, 3, st);
                        if (e == cudaSuccess && f + k < pb.n_frames) ctx->launches++;
                    }
                }
                else if (fused_bgr) e = vaw::launch_warp_nv12_to_bgr(g, pb, ptab, st);
                else if (tiled) e = vaw::launch_warp_nv12_tile(g, pb, ptab, *tm, st);
                else e = vaw::launch_warp_nv12_poly(g, pb, ptab, st);
                if (e != cudaSuccess) return cuda_fail(ctx, e, "warp kernel launch");
                ctx->launches++;
            }
            if (!table_override) {
                ctx->table_used = true;
                ctx->table_stream = st;
                e = cudaEventRecord(ctx->table_free, st);
                if (e != cudaSuccess) return cuda_fail(ctx, e, "warp kernel launch");
            }
            if (tev) { cudaEventRecord(tev[3], st); ctx->timed_launches++; }
            continue;
        }
        switch (ctx->p.format) {
        case VAW_FORMAT_NV12: e = vaw::launch_warp_nv12_gather(g, bb, st); break;
        case VAW_FORMAT_BGR24: e = vaw::launch_warp_packed_gather(g, bb, 3, st); break;
        default: e = vaw::launch_warp_packed_gather(g, bb, 1, st); break;
        }
        if (e != cudaSuccess) return cuda_fail(ctx, e, "warp kernel launch");
        ctx->launches++;
    }
    return VAW_OK;
}

void free_host_path(vaw_ctx* ctx)
{
    for (Stage& s : ctx->stage) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        cudaFree(s.dev_in); cudaFree(s.dev_out); cudaFree(s.dev_rot); cudaFree(s.pieces);
        cudaFreeHost(s.pin_in); cudaFreeHost(s.pin_out); cudaFreeHost(s.pin_rot);
        if (s.stream) cudaStreamDestroy(s.stream);
        s = Stage{};
    }
    ctx->host_ready = false;
}

int init_host_path_impl(vaw_ctx* ctx);

int init_host_path(vaw_ctx* ctx)
{
    if (ctx->host_ready) return VAW_OK;
    const int rc = init_host_path_impl(ctx);
    if (rc != VAW_OK) {  // no half-allocated ring: a retry starts from scratch
        const std::string msg = ctx->err;
        free_host_path(ctx);
        cudaGetLastError();
        ctx->err = msg;
    }
    return rc;
}

int init_host_path_impl(vaw_ctx* ctx)
{
    size_t per = ctx->src_frame_bytes;
    int cf = (int)(ctx->host_chunk_bytes / per);
    ctx->chunk_frames = cf < 1 ? 1 : (cf > 64 ? 64 : cf);
    for (int si = 0; si < ctx->host_stages; ++si) {
        Stage& s = ctx->stage[si];
        VAW_CUDA(ctx, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        VAW_CUDA(ctx, cudaMalloc(&s.dev_in, ctx->src_frame_bytes * ctx->chunk_frames));
        VAW_CUDA(ctx, cudaMalloc(&s.dev_out, ctx->dst_frame_bytes * ctx->chunk_frames));
        VAW_CUDA(ctx, cudaMalloc(&s.dev_rot, sizeof(float) * 9 * ctx->chunk_frames));
        VAW_CUDA(ctx, cudaMallocHost(&s.pin_rot, sizeof(float) * 9 * ctx->chunk_frames));
        if (ctx->variant != VAW_VARIANT_GATHER)
            VAW_CUDA(ctx, cudaMalloc(&s.pieces, sizeof(vaw::PieceRec) * ctx->pieces_per_frame * ctx->chunk_frames));
    }
    ctx->host_ready = true;
    return VAW_OK;
}

// Tile capacity = the shared memory that the largest CTA count still fitting the biggest tile of the
// unrotated geometry (+6 % for the tilt a few degrees of rotation add) leaves each CTA: the rest of that
// budget is free head-room for larger rotations.  Returns the CTA count.
int choose_tile_cap(vaw_ctx* ctx)
{
    const int book = vaw::tile_smem_bytes(0);
    int ctas = 8;  // register limit of the 64-register instantiation
    if (const char* env = getenv("VAW_EXPERIMENT_MAX_CTAS")) {  // analysis only (two-warp CTAs: more of them fit)
        const int v = atoi(env);
        if (v >= 1 && v <= 32) ctas = v;
    }
    // INTER_CUBIC: four CTAs per SM (128 registers each); more CTAs would leave L1 too small for the 32 KB weight table
    // (C3: 16.7 k frames/s at six CTAs = no L1 to speak of, 23.9 k at five, 24.3 k at four, 23.4 k at three).
    // INTER_LANCZOS4: two, so that the 128 KB table fits L1 (4.8 k frames/s at four CTAs, 5.7 k at three, 6.1 k at two).
    if (ctx->gd.halo && !getenv("VAW_EXPERIMENT_MAX_CTAS")) ctas = ctx->gd.halo == 1 ? 4 : 2;
    while (ctas > 1 && ctx->tile_need * 106 / 100 > vaw::tile_cap_for_ctas(ctas, book)) --ctas;
    long long cap = vaw::tile_cap_for_ctas(ctas, book);
    if (ctx->gd.halo) {
        // every sample of these filters reads its 32 / 128 bytes of weights (a 32 KB / 128 KB table) through L1: the tiles
        // get what they need (+25 % for tilt) instead of the whole SM, the launcher leaves the rest to L1
        int pad = 125;
        if (const char* env = getenv("VAW_EXPERIMENT_TABLE_PAD")) pad = atoi(env) >= 100 ? atoi(env) : pad;  // analysis only
        const long long want = ((ctx->tile_need * pad / 100 + 127) / 128) * 128;
        if (want < cap && !getenv("VAW_EXPERIMENT_FULL_SMEM")) cap = want;
        ctx->table_ctas = ctas;
    }
    if (cap < vaw::kTileCapMin) cap = vaw::kTileCapMin;
    if (cap > vaw::kTileCapMax) cap = vaw::kTileCapMax;
    if (const char* env = getenv("VAW_EXPERIMENT_FORCE_CTAS")) {  // analysis only: the capacity of exactly this many CTAs, whatever the tiles need
        const int v = atoi(env);
        if (v >= 1 && v <= 16) { ctas = v; cap = vaw::tile_cap_for_ctas(v, book); }
    }
    ctx->tile_cap = (int)cap;
    ctx->gd.tile_cap = ctx->tile_cap;  // the builder picks the tile pitch of every piece against it
    if (const char* env = getenv("VAW_EXPERIMENT_PITCH64")) ctx->gd.pitch64 = atoi(env);  // analysis only
    if (getenv("VAW_EXPERIMENT_TIGHT_PITCH")) ctx->gd.tile_cap = 0;  // analysis only: always the tightest multiple of 32
    for (vaw_ctx::MapEntry& e : ctx->map_cache) e = vaw_ctx::MapEntry{};
    return ctas;
}

bool is_pinned(const void* p)
{
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

}  // namespace

extern "C" {

const char* vaw_strerror(int code)
{
    switch (code) {
    case VAW_OK: return "ok";
    case VAW_ERR_INVALID: return "invalid argument";
    case VAW_ERR_CUDA: return "CUDA error or no device";
    case VAW_ERR_UNSUPPORTED: return "unsupported";
    case VAW_ERR_NOMEM: return "out of memory";
    default: return "unknown error";
    }
}

const char* vaw_last_error(const vaw_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

size_t vaw_frame_bytes(int format, int width, int height, int pitch)
{
    (void)width;
    if (format == VAW_FORMAT_NV12_TO_BGR24) return 0;  // two formats: ask for VAW_FORMAT_NV12 (source) or VAW_FORMAT_BGR24 (output)
    if (format == VAW_FORMAT_NV12) return (size_t)pitch * (size_t)(height + height / 2);
    return (size_t)pitch * (size_t)height;
}

uint64_t vaw_launch_count(const vaw_ctx* ctx) { return ctx ? ctx->launches : 0; }

int vaw_get_variant(const vaw_ctx* ctx) { return ctx ? ctx->variant : VAW_ERR_INVALID; }

int vaw_cubic_table(int16_t* out)
{
    if (!out) return VAW_ERR_INVALID;
    vaw::build_cubic_table(out);
    return VAW_OK;
}

int vaw_lanczos4_table(int16_t* out)
{
    if (!out) return VAW_ERR_INVALID;
    vaw::build_lanczos4_table(out);
    return VAW_OK;
}

int vaw_create(const vaw_params* params, int device, vaw_ctx** out)
{
    if (!params || !out) return fail(nullptr, VAW_ERR_INVALID, "null argument");
    *out = nullptr;
    const vaw_params& p = *params;
    if (p.interpolation != VAW_INTER_LINEAR && p.interpolation != VAW_INTER_NEAREST && p.interpolation != VAW_INTER_CUBIC &&
        p.interpolation != VAW_INTER_LANCZOS4)
        return fail(nullptr, VAW_ERR_UNSUPPORTED, "only INTER_NEAREST, INTER_LINEAR, INTER_CUBIC and INTER_LANCZOS4 are implemented");
    const bool table_filter = p.interpolation == VAW_INTER_CUBIC || p.interpolation == VAW_INTER_LANCZOS4;
    // (the staged-tile kernels of every format carry every filter: vaw_tile.cu, vaw_packed_tile.cu; POLY and TEX are INTER_LINEAR only)
    if (p.interpolation != VAW_INTER_LINEAR && p.variant != VAW_VARIANT_AUTO && p.variant != VAW_VARIANT_GATHER && p.variant != VAW_VARIANT_TILED)
        return fail(nullptr, VAW_ERR_UNSUPPORTED, "INTER_NEAREST, INTER_CUBIC and INTER_LANCZOS4 run on variants GATHER and TILED (AUTO picks TILED)");
    if (p.format != VAW_FORMAT_NV12 && p.format != VAW_FORMAT_BGR24 && p.format != VAW_FORMAT_GRAY8 &&
        p.format != VAW_FORMAT_NV12_TO_BGR24)
        return fail(nullptr, VAW_ERR_INVALID, "unknown pixel format");
    // NV12 -> BGR24: INTER_LINEAR on POLY (one launch) or TILED; the other filters through TILED's chain only
    // (cvtColor into the L2-resident scratch, then the staged BGR kernel with that filter)
    if (p.format == VAW_FORMAT_NV12_TO_BGR24 &&
        ((p.variant != VAW_VARIANT_AUTO && p.variant != VAW_VARIANT_TILED && !(p.variant == VAW_VARIANT_POLY && p.interpolation == VAW_INTER_LINEAR))))
        return fail(nullptr, VAW_ERR_UNSUPPORTED, "NV12 -> BGR24: INTER_LINEAR on variant AUTO, POLY (one launch) or TILED (conversion + staged BGR kernel); the other filters on AUTO or TILED");
    if (p.variant < VAW_VARIANT_AUTO || p.variant > VAW_VARIANT_TEX || p.variant == VAW_VARIANT_PIPE)
        return fail(nullptr, VAW_ERR_UNSUPPORTED, "kernel variant not available in this build (PIPE was retired in round 2)");
    if (p.projection < 0 || p.projection > 3) return fail(nullptr, VAW_ERR_INVALID, "projection is 0..3");
    if (p.projection != 0) {
        // only createMap.cl's pair has an fp32 operation order (variant GATHER); the others exist on the
        // polynomial variants, whose coordinates come from double-precision anchors
        const bool poly_fmt = p.format == VAW_FORMAT_NV12 || p.format == VAW_FORMAT_NV12_TO_BGR24;
        if (!poly_fmt || p.variant == VAW_VARIANT_GATHER || p.interpolation != VAW_INTER_LINEAR)
            return fail(nullptr, VAW_ERR_UNSUPPORTED, "rectilinear input / fisheye output: NV12 sources, INTER_LINEAR, variants AUTO / POLY / TILED");
        if ((p.projection & 1) && (p.src_distortion[0] != 0 || p.src_distortion[1] != 0 || p.src_distortion[2] != 0 || p.src_distortion[3] != 0))
            return fail(nullptr, VAW_ERR_INVALID, "fisheye distortion coefficients with a rectilinear input camera");
    }
    if (p.variant >= VAW_VARIANT_POLY && p.variant != VAW_VARIANT_TILED && p.format != VAW_FORMAT_NV12 && p.format != VAW_FORMAT_NV12_TO_BGR24)
        return fail(nullptr, VAW_ERR_UNSUPPORTED, "variants POLY and TEX exist for NV12 only (GRAY8 / BGR24: GATHER or TILED)");
    // `short` indices in createMap.cl:10-11 and int16 taps in cv::remap cap both sizes
    if (p.src_width < 2 || p.src_height < 2 || p.out_width < 1 || p.out_height < 1 ||
        p.src_width > 32766 || p.src_height > 32766 || p.out_width > 32766 || p.out_height > 32766)
        return fail(nullptr, VAW_ERR_INVALID, "sizes must be in [2, 32766]");
    if (p.format == VAW_FORMAT_NV12 &&
        ((p.src_width | p.src_height | p.out_width | p.out_height) & 1))
        return fail(nullptr, VAW_ERR_INVALID, "NV12 sizes must be even");
    if (p.format == VAW_FORMAT_NV12_TO_BGR24 && ((p.src_width | p.src_height) & 1))
        return fail(nullptr, VAW_ERR_INVALID, "NV12 source sizes must be even");

    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(nullptr, VAW_ERR_CUDA, "no CUDA device: libvaw has no CPU fallback");
    }
    if (device < 0 || device >= n_dev) return fail(nullptr, VAW_ERR_INVALID, "bad device ordinal");

    vaw_ctx* ctx = new (std::nothrow) vaw_ctx;
    if (!ctx) return fail(nullptr, VAW_ERR_NOMEM, "out of host memory");
    ctx->p = p;
    ctx->device = device;
    ctx->src_format = p.format == VAW_FORMAT_NV12_TO_BGR24 ? VAW_FORMAT_NV12 : p.format;
    ctx->dst_format = p.format == VAW_FORMAT_NV12_TO_BGR24 ? VAW_FORMAT_BGR24 : p.format;
    ctx->channels = ctx->src_format == VAW_FORMAT_BGR24 ? 3 : 1;
    ctx->dst_channels = ctx->dst_format == VAW_FORMAT_BGR24 ? 3 : 1;
    vaw::Geom& g = ctx->g;
    g.scx = (float)p.src_center_x; g.scy = (float)p.src_center_y;   // FrameSourceWarp.cpp:283-284
    g.sfx = (float)p.src_focal_x;  g.sfy = (float)p.src_focal_y;    // :285-286
    g.mcx = (float)p.map_center_x; g.mcy = (float)p.map_center_y;   // :287-288
    g.mfx = (float)p.map_focal_x;  g.mfy = (float)p.map_focal_y;    // :289-290
    g.src_w = p.src_width; g.src_h = p.src_height;
    g.out_w = p.out_width; g.out_h = p.out_height;
    g.border = (unsigned)p.border[0] | ((unsigned)p.border[1] << 8) | ((unsigned)p.border[2] << 16) |
               ((unsigned)p.border[3] << 24);
    g.force_exact = (!centre_ok(g.scx) || !centre_ok(g.scy)) ? 1 : 0;
    g.nearest = p.interpolation == VAW_INTER_NEAREST ? 1 : 0;
    g.projection = p.projection;
    g.has_dist = 0;
    for (int i = 0; i < 4; ++i) {
        g.kd[i] = p.src_distortion[i];
        // |k| <= 10 keeps theta_d = theta (1 + k1 theta^2 + ...) far inside the fp32 range the per-pixel
        // path's shared-reciprocal sequences are certified for (real lenses: |k| < 1)
        if (!(std::fabs(g.kd[i]) <= 10.0f)) { delete ctx; return fail(nullptr, VAW_ERR_INVALID, "distortion coefficient out of range (|k| <= 10)"); }
        if (g.kd[i] != 0.0f) g.has_dist = 1;
    }
    ctx->src_frame_bytes = vaw_frame_bytes(ctx->src_format, p.src_width, p.src_height, p.src_width * ctx->channels);
    ctx->dst_frame_bytes = vaw_frame_bytes(ctx->dst_format, p.out_width, p.out_height, p.out_width * ctx->dst_channels);

    DeviceGuard dg(device);
    const int n_x = ((p.out_width + 127) / 128) * 128 + 4, n_y = ((p.out_height + 15) / 16) * 16 + 2;
    cudaError_t e = cudaMalloc(&ctx->xtab, sizeof(float) * n_x);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->ytab, sizeof(float) * n_y);
    if (e == cudaSuccess) e = vaw::launch_ray_tables(ctx->xtab, n_x, ctx->ytab, n_y, g.mcx, g.mfx, g.mcy, g.mfy, nullptr);
    if (e == cudaSuccess && (p.interpolation == VAW_INTER_CUBIC || p.interpolation == VAW_INTER_LANCZOS4)) {
        const bool cubic = p.interpolation == VAW_INTER_CUBIC;
        std::string tab(sizeof(int16_t) * (cubic ? vaw::kCubicTabEntries : vaw::kLanczosTabEntries), '\0');
        if (cubic) vaw::build_cubic_table(reinterpret_cast<int16_t*>(&tab[0]));
        else vaw::build_lanczos4_table(reinterpret_cast<int16_t*>(&tab[0]));
        g.tab_ks = cubic ? 4 : 8;
        e = cudaMalloc(&ctx->cubic_tab, tab.size());
        if (e == cudaSuccess) e = cudaMemcpy(ctx->cubic_tab, tab.data(), tab.size(), cudaMemcpyHostToDevice);
        g.cubic_tab = ctx->cubic_tab;
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        int rc = cuda_fail(nullptr, e, "vaw_create");
        vaw_destroy(ctx);  // frees whatever was allocated so far
        cudaGetLastError();
        return rc;
    }
    // AUTO: the staged-tile kernels wherever they exist (INTER_LINEAR: NV12 -> vaw_tile.cu, GRAY8 / BGR24 -> vaw_packed_tile.cu)
    ctx->variant = p.variant != VAW_VARIANT_AUTO ? p.variant
                   : ((p.interpolation == VAW_INTER_LINEAR || (p.interpolation == VAW_INTER_NEAREST && p.projection == 0) ||
                       (table_filter && p.projection == 0))
                          ? VAW_VARIANT_TILED : VAW_VARIANT_GATHER);
    // (NV12 -> BGR24: TILED = cvtColor into an L2-resident scratch + the staged BGR kernel, 30.9 k frames/s at 4K;
    //  POLY = everything in one launch with per-tap conversion, 13.8 k)
    const bool fused_tiled = p.format == VAW_FORMAT_NV12_TO_BGR24 && ctx->variant == VAW_VARIANT_TILED;
    const bool packed_fmt = p.format == VAW_FORMAT_BGR24 || p.format == VAW_FORMAT_GRAY8 || fused_tiled;
    const int tile_channels = fused_tiled ? 3 : ctx->channels;
    if (ctx->variant != VAW_VARIANT_GATHER) {
        // rows per piece: keep the cubic-in-v truncation error ~ 2.4e-3 * f_in * (PH / f_out)^4 px
        // (measured on the BASELINE geometries) below the certificate's 5e-5 px
        const double fin = std::fmax(std::fabs(p.src_focal_x), std::fabs(p.src_focal_y));
        const double fout = std::fmin(std::fabs(p.map_focal_x), std::fabs(p.map_focal_y));
        int ph = 32;
#ifndef VAW_PH_RULE
#define VAW_PH_RULE 5e-5  // = the builder's accuracy certificate; 2e-5 halved C1's pieces for nothing (204 k -> 247 k frames/s at 32 rows, same measured errors)
#endif
        while (ph > 8 && 2.4e-3 * fin * std::pow(ph / fout, 4.0) > VAW_PH_RULE) ph >>= 1;
        if (const char* env = getenv("VAW_EXPERIMENT_PH")) {  // analysis only: force the rows per piece (8, 16 or 32)
            const int v = atoi(env);
            if (v == 8 || v == 16 || v == 32) ph = v;
        }
        vaw::GeomD& d = ctx->gd;
        auto set_piece_rows = [&](int rows) {
            g.piece_h = rows;
            g.t_off = 0.5f * (float)(rows - 1);
            g.t_scale = 2.0f / (float)rows;
            d.piece_h = rows;
            make_basis(ctx->basis, rows);
            ctx->pieces_per_frame = (size_t)vaw::pieces_x(g.out_w) * vaw::pieces_y(g.out_h, rows);
        };
        d.scx = g.scx; d.scy = g.scy; d.sfx = g.sfx; d.sfy = g.sfy;  // the fp32 scalars, widened
        d.mcx = g.mcx; d.mcy = g.mcy; d.mfx = g.mfx; d.mfy = g.mfy;
        d.inv_mfx = 1.0 / d.mfx; d.inv_mfy = 1.0 / d.mfy;
        for (int i = 0; i < 4; ++i) d.kd[i] = g.kd[i];  // the fp32 coefficients, widened
        d.has_dist = g.has_dist;
        d.src_w = g.src_w; d.src_h = g.src_h; d.out_w = g.out_w; d.out_h = g.out_h;
        d.projection = p.projection;
        // INTER_CUBIC / INTER_LANCZOS4 on staged tiles (NV12): the boxes carry the filter's halo
        d.halo = (table_filter && ctx->variant == VAW_VARIANT_TILED) ? (p.interpolation == VAW_INTER_CUBIC ? 1 : 3) : 0;
        // their kernels are bound by the shared-memory / L1 data pipe: tile pitches of 64 modulo 128 bytes halve the bank
        // conflicts between lanes that share a source column in adjacent rows (NV12 cubic 28.9 k -> 30.7 k frames/s at 4K;
        // the INTER_LINEAR kernel is issue-bound and measures the same either way, so it keeps its pitches)
        d.pitch64 = d.halo ? 1 : 0;
        set_piece_rows(ph);
        e = cudaEventCreateWithFlags(&ctx->table_free, cudaEventDisableTiming);
        // (room for the same frame cut into 16-row pieces: BGR24 may fall back to them below)
        if (e == cudaSuccess) e = cudaMalloc(&ctx->dump_table, 2 * ctx->pieces_per_frame * sizeof(vaw::PieceRec));
        if (e == cudaSuccess) {
            int lo = 0, hi = 0;  // "greatest" priority is the numerically lowest
            e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
            if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->side, cudaStreamNonBlocking, hi);
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
        if (e == cudaSuccess && ctx->variant == VAW_VARIANT_TEX) {
            cudaDeviceProp prop{};
            e = cudaGetDeviceProperties(&prop, device);
            if (e == cudaSuccess) {
                ctx->tex_align = prop.textureAlignment > 0 ? (int)prop.textureAlignment : 512;
                ctx->tex_pitch_align = prop.texturePitchAlignment > 0 ? (int)prop.texturePitchAlignment : 32;
            }
        }
        if (e == cudaSuccess && (ctx->variant == VAW_VARIANT_TILED || ctx->variant == VAW_VARIANT_TEX)) {
            // size the per-CTA tile from the source boxes of the unrotated geometry, +20 % for the
            // tilt a few degrees of rotation add; more shared memory per CTA = fewer resident CTAs
            const float eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            for (int attempt = 0; attempt < 2 && e == cudaSuccess; ++attempt) {
                e = vaw::launch_build_pieces(ctx->gd, ctx->basis, nullptr, eye, 1, ctx->dump_table, nullptr);
                std::string host(ctx->pieces_per_frame * sizeof(vaw::PieceRec), '\0');
                if (e == cudaSuccess) e = cudaMemcpy(&host[0], ctx->dump_table, host.size(), cudaMemcpyDeviceToHost);
                if (e != cudaSuccess) break;
                const vaw::PieceRec* rec = reinterpret_cast<const vaw::PieceRec*>(host.data());
                long long need = 0;
                for (size_t i = 0; i < ctx->pieces_per_frame; ++i) {
                    const int nb = packed_fmt ? vaw::packed_tile_need_bytes(rec[i], tile_channels) : vaw::tile_need_bytes(rec[i], ctx->gd.halo);
                    if (nb != 0x7fffffff && nb > need) need = nb;
                }
                ctx->tile_need = need;
                const int ctas = choose_tile_cap(ctx);
                // BGR24: three bytes per source pixel.  32-row pieces win while four CTAs still share an SM (C3: 52 KB
                // tiles, 0.758 ms per 32 frames against 0.777 ms with 16-row pieces at six CTAs); larger boxes (C5: 120 KB)
                // are cut into 16-row pieces
                if (attempt == 0 && (p.format == VAW_FORMAT_BGR24 || fused_tiled) && g.piece_h > 16 && (ctas < 4 || ctx->gd.halo) && !getenv("VAW_EXPERIMENT_PH"))  // (cubic: small tiles leave L1 to the weight table)
                    set_piece_rows(16);
                else
                    break;
            }
        }
        if (e != cudaSuccess) {
            int rc = cuda_fail(nullptr, e, "vaw_create (piece tables)");
            vaw_destroy(ctx);
            cudaGetLastError();
            return rc;
        }
    }
    ctx->launches = 1;
    g.xtab = ctx->xtab;
    g.ytab = ctx->ytab;
    *out = ctx;
    return VAW_OK;
}

void vaw_destroy(vaw_ctx* ctx)
{
    if (!ctx) return;
    DeviceGuard dg(ctx->device);
    free_host_path(ctx);
    cudaDeviceSynchronize();
    for (vaw_ctx::TexEntry& te : ctx->tex_cache) destroy_tex_entry(te);
    cudaFree(ctx->table);
    cudaFree(ctx->dump_table);
    cudaFree(ctx->bgr_scratch);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->table_free) cudaEventDestroy(ctx->table_free);
    if (ctx->tev) {
        for (int i = 0; i < vaw_ctx::kTev * vaw_ctx::kTimeRing; ++i) cudaEventDestroy(ctx->tev[i]);
        delete[] ctx->tev;
    }
    cudaFree(ctx->xtab);
    cudaFree(ctx->ytab);
    cudaFree(ctx->cubic_tab);
    delete ctx;
}

int vaw_set_option(vaw_ctx* ctx, const char* name, int value)
{
    if (!ctx || !name) return VAW_ERR_INVALID;
    if (!std::strcmp(name, "force_exact")) { ctx->g.force_exact = value ? 1 : 0; return VAW_OK; }
    if (!std::strcmp(name, "split_builder")) { ctx->split_builder = value != 0; return VAW_OK; }
    if (!std::strcmp(name, "host_stages") || !std::strcmp(name, "host_chunk_mb")) {  // shape of the host-buffer pipeline
        DeviceGuard dg(ctx->device);
        if (name[5] == 's') {
            if (value < 2 || value > kStages) return fail(ctx, VAW_ERR_INVALID, "host_stages is 2..8");
            free_host_path(ctx);
            ctx->host_stages = value;
        } else {
            if (value < 1 || value > 1024) return fail(ctx, VAW_ERR_INVALID, "host_chunk_mb is 1..1024");
            free_host_path(ctx);
            ctx->host_chunk_bytes = (size_t)value << 20;
        }
        return VAW_OK;
    }
    if (!std::strcmp(name, "time_kernels")) {
        DeviceGuard dg(ctx->device);
        if (value && !ctx->tev) {
            ctx->tev = new (std::nothrow) cudaEvent_t[vaw_ctx::kTev * vaw_ctx::kTimeRing];
            if (!ctx->tev) return fail(ctx, VAW_ERR_NOMEM, "out of host memory");
            for (int i = 0; i < vaw_ctx::kTev * vaw_ctx::kTimeRing; ++i) VAW_CUDA(ctx, cudaEventCreate(&ctx->tev[i]));
        }
        ctx->time_kernels = value != 0;
        ctx->timed_launches = 0;
        return VAW_OK;
    }
    return fail(ctx, VAW_ERR_INVALID, std::string("unknown option ") + name);
}

int vaw_warp(vaw_ctx* ctx, const uint8_t* src, int src_pitch, uint8_t* dst, int dst_pitch,
             const double rotation[9], void* stream)
{
    int rc = check_buffers(ctx, src, src_pitch, dst, dst_pitch);
    if (rc) return rc;
    if (!rotation) return fail(ctx, VAW_ERR_INVALID, "null rotation");
    DeviceGuard dg(ctx->device);
    vaw::Rot R = rot_from_double(rotation);
    return launch(ctx, src, src_pitch, 0, dst, dst_pitch, 0, nullptr, &R, 1, (cudaStream_t)stream);
}

int vaw_warp_batch(vaw_ctx* ctx, const uint8_t* src, int src_pitch, size_t src_frame_stride,
                   uint8_t* dst, int dst_pitch, size_t dst_frame_stride, const float* rotations,
                   int n_frames, void* stream)
{
    int rc = check_buffers(ctx, src, src_pitch, dst, dst_pitch);
    if (rc) return rc;
    if (n_frames < 0 || (n_frames > 0 && !rotations)) return fail(ctx, VAW_ERR_INVALID, "bad batch");
    if (n_frames == 0) return VAW_OK;
    if (ctx->src_format == VAW_FORMAT_NV12 && (src_frame_stride & 1))
        return fail(ctx, VAW_ERR_INVALID, "NV12 frame stride must be even");
    if (n_frames > 1) {
        // output frames must not overlap (a zero or short stride would make frames overwrite each other);
        // source frames may repeat (stride 0 = the same frame under n rotations) but not interleave
        const size_t out_rows = ctx->dst_format == VAW_FORMAT_NV12 ? (size_t)ctx->p.out_height * 3 / 2 : (size_t)ctx->p.out_height;
        const size_t src_rows = ctx->src_format == VAW_FORMAT_NV12 ? (size_t)ctx->p.src_height * 3 / 2 : (size_t)ctx->p.src_height;
        const size_t dst_need = (out_rows - 1) * (size_t)dst_pitch + (size_t)ctx->p.out_width * ctx->dst_channels;
        const size_t src_need = (src_rows - 1) * (size_t)src_pitch + (size_t)ctx->p.src_width * ctx->channels;
        if (dst_frame_stride < dst_need)
            return fail(ctx, VAW_ERR_INVALID, "dst_frame_stride smaller than one output frame");
        if (src_frame_stride != 0 && src_frame_stride < src_need)
            return fail(ctx, VAW_ERR_INVALID, "src_frame_stride smaller than one source frame (0 = repeat the frame)");
    }
    DeviceGuard dg(ctx->device);
    return launch(ctx, src, src_pitch, src_frame_stride, dst, dst_pitch, dst_frame_stride, rotations,
                  nullptr, n_frames, (cudaStream_t)stream);
}

int vaw_bind_clip(vaw_ctx* ctx, const uint8_t* base, int pitch, size_t frame_stride, int n_slots)
{
    if (!ctx) return VAW_ERR_INVALID;
    if (!base) {  // unbind
        ctx->clip_base = nullptr; ctx->clip_pitch = 0; ctx->clip_stride = 0; ctx->clip_slots = 0;
        return VAW_OK;
    }
    if (n_slots < 1 || pitch < ctx->p.src_width * ctx->channels ||
        frame_stride < vaw_frame_bytes(ctx->src_format, ctx->p.src_width, ctx->p.src_height, pitch))
        return fail(ctx, VAW_ERR_INVALID, "bad clip layout");
    ctx->clip_base = base; ctx->clip_pitch = pitch; ctx->clip_stride = frame_stride; ctx->clip_slots = n_slots;
    return VAW_OK;
}

int vaw_upload_rotations(vaw_ctx* ctx, const double* rotations_host, int n_frames,
                         float* rotations_dev, void* stream)
{
    if (!ctx) return VAW_ERR_INVALID;
    if (!rotations_host || !rotations_dev || n_frames < 0) return fail(ctx, VAW_ERR_INVALID, "bad rotations");
    DeviceGuard dg(ctx->device);
    std::string tmp(sizeof(float) * 9 * (size_t)n_frames, '\0');
    float* f = reinterpret_cast<float*>(&tmp[0]);
    for (size_t i = 0; i < (size_t)n_frames * 9; ++i) f[i] = (float)rotations_host[i];
    VAW_CUDA(ctx, cudaMemcpyAsync(rotations_dev, f, tmp.size(), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    VAW_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    return VAW_OK;
}

int vaw_warp_batch_host(vaw_ctx* ctx, const uint8_t* src_host, uint8_t* dst_host,
                        const double* rotations_host, int n_frames)
{
    if (!ctx) return VAW_ERR_INVALID;
    if (!src_host || !dst_host || !rotations_host || n_frames < 0)
        return fail(ctx, VAW_ERR_INVALID, "null host buffer");
    if (n_frames == 0) return VAW_OK;
    DeviceGuard dg(ctx->device);
    int rc = init_host_path(ctx);
    if (rc) return rc;
    const size_t sfb = ctx->src_frame_bytes, dfb = ctx->dst_frame_bytes;
    const int src_pitch = ctx->p.src_width * ctx->channels, dst_pitch = ctx->p.out_width * ctx->dst_channels;
    // Pinned (page-locked) user buffers are copied from/to directly; pageable ones go
    // through lazily allocated pinned staging so that the copies stay asynchronous.
    const bool src_pinned = is_pinned(src_host), dst_pinned = is_pinned(dst_host);

    auto drain = [&](Stage& s) -> int {
        if (!s.busy) return VAW_OK;
        VAW_CUDA(ctx, cudaStreamSynchronize(s.stream));
        if (s.out_staged) std::memcpy(s.host_dst, s.pin_out, s.out_bytes);
        s.busy = false;
        return VAW_OK;
    };

    // On any error: wait for every stage still in flight (its DMA may be reading src_host or writing
    // dst_host / the pinned staging), forget the staged outputs and only then return, so that the
    // caller may free or reuse its buffers.  The message of the first error is kept.
    auto abort_all = [&](int code) -> int {
        const std::string msg = ctx->err;
        for (Stage& s : ctx->stage) {
            if (s.stream) cudaStreamSynchronize(s.stream);
            s.busy = false;
            s.out_staged = false;
        }
        cudaGetLastError();
        ctx->err = msg;
        return code;
    };
    auto submit = [&](Stage& s, int first, int n) -> int {
        const uint8_t* hsrc = src_host + (size_t)first * sfb;
        uint8_t* hdst = dst_host + (size_t)first * dfb;
        for (int i = 0; i < n * 9; ++i) s.pin_rot[i] = (float)rotations_host[(size_t)first * 9 + i];
        VAW_CUDA(ctx, cudaMemcpyAsync(s.dev_rot, s.pin_rot, sizeof(float) * 9 * n, cudaMemcpyHostToDevice, s.stream));
        if (!src_pinned) {
            if (!s.pin_in) VAW_CUDA(ctx, cudaMallocHost(&s.pin_in, sfb * ctx->chunk_frames));
            std::memcpy(s.pin_in, hsrc, sfb * n);
            hsrc = s.pin_in;
        }
        VAW_CUDA(ctx, cudaMemcpyAsync(s.dev_in, hsrc, sfb * n, cudaMemcpyHostToDevice, s.stream));
        s.busy = true;  // from here on the stage has work in flight that touches the caller's buffers
        int rc2 = launch(ctx, s.dev_in, src_pitch, sfb, s.dev_out, dst_pitch, dfb, s.dev_rot, nullptr, n, s.stream, s.pieces);
        if (rc2) return rc2;
        s.out_staged = !dst_pinned;
        if (s.out_staged && !s.pin_out) VAW_CUDA(ctx, cudaMallocHost(&s.pin_out, dfb * ctx->chunk_frames));
        VAW_CUDA(ctx, cudaMemcpyAsync(s.out_staged ? s.pin_out : hdst, s.dev_out, dfb * n, cudaMemcpyDeviceToHost, s.stream));
        s.host_dst = hdst;
        s.out_bytes = dfb * n;
        return VAW_OK;
    };

    int k = 0;
    for (int first = 0; first < n_frames; first += ctx->chunk_frames, ++k) {
        Stage& s = ctx->stage[k % ctx->host_stages];
        if ((rc = drain(s))) return abort_all(rc);
        const int n = n_frames - first < ctx->chunk_frames ? n_frames - first : ctx->chunk_frames;
        s.out_staged = false;
        if ((rc = submit(s, first, n))) return abort_all(rc);
    }
    // drain in submission order
    for (int i = 0; i < ctx->host_stages; ++i)
        if ((rc = drain(ctx->stage[(k + i) % ctx->host_stages]))) return abort_all(rc);
    return VAW_OK;
}

int vaw_dump_coords(vaw_ctx* ctx, const double rotation[9], int plane, float* map_x, float* map_y,
                    int map_pitch, void* stream)
{
    if (!ctx) return VAW_ERR_INVALID;
    if (!rotation || !map_x || !map_y) return fail(ctx, VAW_ERR_INVALID, "null argument");
    if (plane != 0 && !(plane == 1 && ctx->p.format == VAW_FORMAT_NV12))  // (NV12 -> BGR24 samples the converted image with the luma map only)
        return fail(ctx, VAW_ERR_INVALID, "plane 1 exists for NV12 only");
    const int need = plane ? ctx->p.out_width / 2 : ctx->p.out_width;
    if (map_pitch < need) return fail(ctx, VAW_ERR_INVALID, "map pitch smaller than a row");
    DeviceGuard dg(ctx->device);
    const vaw::Rot R = rot_from_double(rotation);
    cudaError_t e;
    if (ctx->variant != VAW_VARIANT_GATHER) {
        e = vaw::launch_build_pieces(ctx->gd, ctx->basis, nullptr, R.r, 1, ctx->dump_table, (cudaStream_t)stream);
        if (e == cudaSuccess)
            e = vaw::launch_dump_coords_poly(ctx->g, R, ctx->dump_table, plane, map_x, map_y, map_pitch,
                                             (cudaStream_t)stream);
        ctx->launches++;
    } else {
        e = vaw::launch_dump_coords(ctx->g, R, plane, map_x, map_y, map_pitch, (cudaStream_t)stream);
    }
    if (e != cudaSuccess) return cuda_fail(ctx, e, "dump_coords launch");
    ctx->launches++;
    return VAW_OK;
}

long long vaw_debug_oob_count(int device)
{
    DeviceGuard dg(device);
    const long long a = vaw::tile_oob_count(), b = vaw::packed_tile_oob_count();
    return (a < 0 || b < 0) ? (a < b ? a : b) : a + b;  // NV12 and GRAY8 / BGR24 staged kernels
}

int vaw_malloc(int device, size_t bytes, void** out)
{
    if (!out) return VAW_ERR_INVALID;
    *out = nullptr;
    DeviceGuard dg(device);
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "vaw_malloc");
    return VAW_OK;
}

int vaw_free(int device, void* ptr)
{
    DeviceGuard dg(device);
    cudaError_t e = cudaFree(ptr);
    return e == cudaSuccess ? VAW_OK : cuda_fail(nullptr, e, "vaw_free");
}

int vaw_memcpy(int device, void* dst, const void* src, size_t bytes, int to_device, void* stream)
{
    if (!dst || !src) return VAW_ERR_INVALID;
    DeviceGuard dg(device);
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost,
                                    (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    return e == cudaSuccess ? VAW_OK : cuda_fail(nullptr, e, "vaw_memcpy");
}

int vaw_sync(int device, void* stream)
{
    DeviceGuard dg(device);
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    return e == cudaSuccess ? VAW_OK : cuda_fail(nullptr, e, "vaw_sync");
}

int vaw_kernel_times(vaw_ctx* ctx, int max_launches, float* builder_ms, float* warp_ms, int* n_out)
{
    if (!ctx) return VAW_ERR_INVALID;
    if (!builder_ms || !warp_ms || !n_out || max_launches < 0) return fail(ctx, VAW_ERR_INVALID, "null argument");
    *n_out = 0;
    if (!ctx->tev) return VAW_OK;
    DeviceGuard dg(ctx->device);
    const uint64_t have = ctx->timed_launches < (uint64_t)vaw_ctx::kTimeRing ? ctx->timed_launches : vaw_ctx::kTimeRing;
    const uint64_t n = have < (uint64_t)max_launches ? have : (uint64_t)max_launches;
    for (uint64_t i = 0; i < n; ++i) {  // the most recent n launches, oldest first
        const cudaEvent_t* e = ctx->tev + vaw_ctx::kTev * ((ctx->timed_launches - n + i) % vaw_ctx::kTimeRing);
        VAW_CUDA(ctx, cudaEventSynchronize(e[3]));
        VAW_CUDA(ctx, cudaEventElapsedTime(&builder_ms[i], e[0], e[1]));
        VAW_CUDA(ctx, cudaEventElapsedTime(&warp_ms[i], e[1], e[3]));
    }
    *n_out = (int)n;
    return VAW_OK;
}

int vaw_kernel_times_split(vaw_ctx* ctx, int max_launches, float* tex_ms, float* tile_ms, int* n_out)
{
    if (!ctx) return VAW_ERR_INVALID;
    if (!tex_ms || !tile_ms || !n_out || max_launches < 0) return fail(ctx, VAW_ERR_INVALID, "null argument");
    *n_out = 0;
    if (!ctx->tev) return VAW_OK;
    DeviceGuard dg(ctx->device);
    const uint64_t have = ctx->timed_launches < (uint64_t)vaw_ctx::kTimeRing ? ctx->timed_launches : vaw_ctx::kTimeRing;
    const uint64_t n = have < (uint64_t)max_launches ? have : (uint64_t)max_launches;
    for (uint64_t i = 0; i < n; ++i) {
        const cudaEvent_t* e = ctx->tev + vaw_ctx::kTev * ((ctx->timed_launches - n + i) % vaw_ctx::kTimeRing);
        VAW_CUDA(ctx, cudaEventSynchronize(e[3]));
        VAW_CUDA(ctx, cudaEventElapsedTime(&tex_ms[i], e[1], e[2]));
        VAW_CUDA(ctx, cudaEventElapsedTime(&tile_ms[i], e[2], e[3]));
    }
    *n_out = (int)n;
    return VAW_OK;
}

// ---- placement of host-staging threads ----------------------------------------------------------------
namespace {
// "a-b,c,d-e" -> CPU set
bool parse_cpulist(const char* text, cpu_set_t* set)
{
    CPU_ZERO(set);
    int n = 0;
    const char* p = text;
    while (*p) {
        char* end = nullptr;
        long a = std::strtol(p, &end, 10);
        if (end == p) break;
        long b = a;
        p = end;
        if (*p == '-') { b = std::strtol(p + 1, &end, 10); p = end; }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET((int)c, set); ++n; }
        if (*p == ',') ++p; else break;
    }
    return n > 0;
}

bool read_small_file(const std::string& path, std::string* out)
{
    FILE* f = std::fopen(path.c_str(), "r");
    if (!f) return false;
    char buf[4096];
    const size_t n = std::fread(buf, 1, sizeof buf - 1, f);
    std::fclose(f);
    buf[n] = 0;
    *out = buf;
    return true;
}
}  // namespace

int vaw_device_numa_node(int device)
{
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char* c = bus; *c; ++c) *c = (char)std::tolower((unsigned char)*c);
    std::string text;
    if (!read_small_file(std::string("/sys/bus/pci/devices/") + bus + "/numa_node", &text)) return -1;
    return std::atoi(text.c_str());
}

int vaw_bind_thread_to_device(int device)
{
    const int node = vaw_device_numa_node(device);
    if (node < 0) return 0;  // no NUMA information (single node, VM): nothing to do
    std::string text;
    if (!read_small_file("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist", &text)) return 0;
    cpu_set_t want, have, both;
    if (!parse_cpulist(text.c_str(), &want)) return 0;
    if (sched_getaffinity(0, sizeof have, &have) != 0) return 0;
    CPU_AND(&both, &want, &have);
    if (CPU_COUNT(&both) == 0) return 0;
    return sched_setaffinity(0, sizeof both, &both) == 0 ? CPU_COUNT(&both) : 0;
}

int vaw_shard_range(int n_frames, int n_parts, int part, int* first, int* count)
{
    if (n_frames < 0 || n_parts < 1 || part < 0 || part >= n_parts || !first || !count) return VAW_ERR_INVALID;
    // contiguous ranges, sizes differing by at most one, earlier parts take the remainder
    const int base = n_frames / n_parts, rem = n_frames % n_parts;
    *first = part * base + (part < rem ? part : rem);
    *count = base + (part < rem ? 1 : 0);
    return VAW_OK;
}

int vaw_piece_stats(vaw_ctx* ctx, const double rotation[9], uint32_t counts[8], void* stream)
{
    if (!ctx) return VAW_ERR_INVALID;
    if (!rotation || !counts) return fail(ctx, VAW_ERR_INVALID, "null argument");
    for (int i = 0; i < 8; ++i) counts[i] = 0;
    if (ctx->variant == VAW_VARIANT_GATHER) return VAW_OK;
    DeviceGuard dg(ctx->device);
    const vaw::Rot R = rot_from_double(rotation);
    cudaError_t e = vaw::launch_build_pieces(ctx->gd, ctx->basis, nullptr, R.r, 1, ctx->dump_table, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "piece table launch");
    ctx->launches++;
    std::string host(ctx->pieces_per_frame * sizeof(vaw::PieceRec), '\0');
    VAW_CUDA(ctx, cudaMemcpyAsync(&host[0], ctx->dump_table, host.size(), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    VAW_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    const vaw::PieceRec* rec = reinterpret_cast<const vaw::PieceRec*>(host.data());
    for (size_t i = 0; i < ctx->pieces_per_frame; ++i) {
        counts[0]++;
        if (rec[i].flags & vaw::kPiecePoly) counts[1]++;
        if (rec[i].flags & vaw::kPieceInterior) counts[2]++;
        if (rec[i].flags & vaw::kPieceOutside) counts[3]++;
        const bool packed = ctx->p.format == VAW_FORMAT_BGR24 || ctx->p.format == VAW_FORMAT_GRAY8;
        const int nb = packed ? vaw::packed_tile_need_bytes(rec[i], ctx->channels) : vaw::tile_need_bytes(rec[i], ctx->gd.halo);
        if (nb > 0 && (nb > ctx->tile_cap)) counts[6]++;
        if (nb != 0x7fffffff && (uint32_t)nb > counts[4]) counts[4] = (uint32_t)nb;
    }
    counts[5] = (uint32_t)ctx->tile_cap;
    return VAW_OK;
}

int vaw_piece_flags(vaw_ctx* ctx, const double rotation[9], uint32_t* flags, int capacity, int* pieces_x_out,
                    int* pieces_y_out, int* piece_h_out, void* stream)
{
    if (!ctx) return VAW_ERR_INVALID;
    if (!rotation || !flags || !pieces_x_out || !pieces_y_out || !piece_h_out) return fail(ctx, VAW_ERR_INVALID, "null argument");
    *pieces_x_out = *pieces_y_out = *piece_h_out = 0;
    if (ctx->variant == VAW_VARIANT_GATHER) return VAW_OK;
    if ((size_t)capacity < ctx->pieces_per_frame) return fail(ctx, VAW_ERR_INVALID, "flags buffer too small");
    DeviceGuard dg(ctx->device);
    const vaw::Rot R = rot_from_double(rotation);
    cudaError_t e = vaw::launch_build_pieces(ctx->gd, ctx->basis, nullptr, R.r, 1, ctx->dump_table, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "piece table launch");
    ctx->launches++;
    std::string host(ctx->pieces_per_frame * sizeof(vaw::PieceRec), '\0');
    VAW_CUDA(ctx, cudaMemcpyAsync(&host[0], ctx->dump_table, host.size(), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    VAW_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    const vaw::PieceRec* rec = reinterpret_cast<const vaw::PieceRec*>(host.data());
    for (size_t i = 0; i < ctx->pieces_per_frame; ++i) flags[i] = rec[i].flags;
    *pieces_x_out = vaw::pieces_x(ctx->g.out_w);
    *pieces_y_out = vaw::pieces_y(ctx->g.out_h, ctx->g.piece_h);
    *piece_h_out = ctx->g.piece_h;
    return VAW_OK;
}

int vaw_piece_tiles(vaw_ctx* ctx, const double rotation[9], uint32_t* out, int capacity, void* stream)
{
    if (!ctx) return VAW_ERR_INVALID;
    if (!rotation || !out) return fail(ctx, VAW_ERR_INVALID, "null argument");
    if (ctx->variant == VAW_VARIANT_GATHER) return fail(ctx, VAW_ERR_UNSUPPORTED, "variant GATHER has no pieces");
    if ((size_t)capacity < ctx->pieces_per_frame) return fail(ctx, VAW_ERR_INVALID, "buffer too small");
    DeviceGuard dg(ctx->device);
    const vaw::Rot R = rot_from_double(rotation);
    cudaError_t e = vaw::launch_build_pieces(ctx->gd, ctx->basis, nullptr, R.r, 1, ctx->dump_table, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "piece table launch");
    ctx->launches++;
    std::string host(ctx->pieces_per_frame * sizeof(vaw::PieceRec), '\0');
    VAW_CUDA(ctx, cudaMemcpyAsync(&host[0], ctx->dump_table, host.size(), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    VAW_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    const vaw::PieceRec* rec = reinterpret_cast<const vaw::PieceRec*>(host.data());
    for (size_t i = 0; i < ctx->pieces_per_frame; ++i) {
        out[4 * i + 0] = rec[i].flags;
        const bool packed = ctx->p.format == VAW_FORMAT_BGR24 || ctx->p.format == VAW_FORMAT_GRAY8;
        out[4 * i + 1] = (uint32_t)(packed ? vaw::packed_tile_need_bytes(rec[i], ctx->channels) : vaw::tile_need_bytes(rec[i], ctx->gd.halo));
        out[4 * i + 2] = rec[i].stage.pl;
        out[4 * i + 3] = (uint32_t)rec[i].stage.nrows | ((uint32_t)rec[i].stage.cnrows << 16);
    }
    return VAW_OK;
}

int vaw_remap_u8(const uint8_t* src, int src_width, int src_height, int src_pitch, int channels,
                 const float* map_x, const float* map_y, int rows, int cols, int map_pitch,
                 uint8_t* dst, int dst_pitch, const uint8_t border[4], int device, void* stream)
{
    if (!src || !map_x || !map_y || !dst || !border) return fail(nullptr, VAW_ERR_INVALID, "null argument");
    if (channels < 1 || channels > 3) return fail(nullptr, VAW_ERR_UNSUPPORTED, "1 to 3 channels");
    if (src_width < 1 || src_height < 1 || src_width > 32766 || src_height > 32766 || rows < 0 || cols < 0 ||
        src_pitch < src_width * channels || dst_pitch < cols * channels || map_pitch < cols)
        return fail(nullptr, VAW_ERR_INVALID, "bad remap geometry");
    if (channels == 2 && ((src_pitch & 1) || (reinterpret_cast<uintptr_t>(src) & 1)))
        return fail(nullptr, VAW_ERR_INVALID, "2-channel source base and pitch must be even");
    if (rows == 0 || cols == 0) return VAW_OK;
    if (rows > 65535) return fail(nullptr, VAW_ERR_INVALID, "too many rows");
    DeviceGuard dg(device);
    const unsigned b = (unsigned)border[0] | ((unsigned)border[1] << 8) | ((unsigned)border[2] << 16);
    cudaError_t e = vaw::launch_remap(src, src_width, src_height, src_pitch, channels, map_x, map_y, rows,
                                      cols, map_pitch, dst, dst_pitch, b, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "remap launch");
    return VAW_OK;
}

int vaw_nv12_to_bgr(const uint8_t* src, int width, int height, int src_pitch, size_t src_frame_stride,
                    uint8_t* dst, int dst_pitch, size_t dst_frame_stride, int n_frames, int device, void* stream)
{
    if (!src || !dst) return fail(nullptr, VAW_ERR_INVALID, "null frame pointer");
    if (width < 2 || height < 2 || ((width | height) & 1) || src_pitch < width || dst_pitch < 3 * width || n_frames < 0)
        return fail(nullptr, VAW_ERR_INVALID, "bad cvtColor geometry");
    DeviceGuard dg(device);
    for (int first = 0; first < n_frames; first += 65535) {
        const int n = n_frames - first < 65535 ? n_frames - first : 65535;
        cudaError_t e = vaw::launch_nv12_to_bgr(src + (size_t)first * src_frame_stride, width, height, src_pitch,
                                                src_frame_stride, dst + (size_t)first * dst_frame_stride, dst_pitch,
                                                dst_frame_stride, n, (cudaStream_t)stream);
        if (e != cudaSuccess) return cuda_fail(nullptr, e, "nv12_to_bgr launch");
    }
    return VAW_OK;
}

int vaw_synth_nv12(uint8_t* dst, int width, int height, int pitch, size_t frame_stride,
                   int first_index, int n_frames, uint32_t seed, int white_noise, int device,
                   void* stream)
{
    if (!dst || width < 2 || height < 2 || ((width | height) & 1) || pitch < width || n_frames < 0)
        return fail(nullptr, VAW_ERR_INVALID, "bad synth arguments");
    DeviceGuard dg(device);
    for (int first = 0; first < n_frames; first += 65535) {
        int n = n_frames - first < 65535 ? n_frames - first : 65535;
        cudaError_t e = vaw::launch_synth_nv12(dst + (size_t)first * frame_stride, width, height, pitch,
                                               frame_stride, first_index + first, n, seed, white_noise,
                                               (cudaStream_t)stream);
        if (e != cudaSuccess) return cuda_fail(nullptr, e, "synth launch");
    }
    return VAW_OK;
}

int vaw_selftest_math(int device, uint32_t seed, uint64_t n_per_thread, uint64_t mismatches[4])
{
    if (!mismatches) return VAW_ERR_INVALID;
    DeviceGuard dg(device);
    unsigned long long* d = nullptr;
    VAW_CUDA(nullptr, cudaMalloc(&d, 4 * sizeof(unsigned long long)));
    cudaMemset(d, 0, 4 * sizeof(unsigned long long));
    cudaError_t e = vaw::launch_selftest_math(seed, n_per_thread, d, nullptr);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    unsigned long long h[4] = {0, 0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "selftest");
    for (int i = 0; i < 4; ++i) mismatches[i] = h[i];
    return VAW_OK;
}

}  // extern "C"
