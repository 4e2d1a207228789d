/*
 * vaw.h -- C-ABI of the B200-native warp path (libvaw.so).
 *
 * "vaw" = video-annotator warp.  This is the drop-in boundary for ONE path of
 * hedgepigdaniel/video-annotator: the per-frame fisheye->rectilinear projection with
 * per-frame camera rotation that the reference computes as a remap map
 * (opencv/createMap.cl:1-51) and applies with cv::remap inside
 * FrameSourceWarp::warp_frame (opencv/FrameSourceWarp.cpp:272-314).
 *
 * Plain C: opaque handle, POD structs, raw device/host pointers, sizes and CUDA
 * stream handles passed as void*.  No torch, OpenCV or C++ types cross the boundary.
 * There is NO CPU fallback: every entry point that computes needs a CUDA device and
 * returns VAW_ERR_CUDA otherwise.
 *
 * Each declaration cites the reference interface it replaces (paths relative to the
 * reference repository root).  INTEGRATION.md shows the reference-side binding.
 */
#ifndef VAW_H
#define VAW_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VAW_VERSION 100 /* 0.1.0 */

/* ---- error convention -------------------------------------------------------------
 * The reference throws ints (-1 on kernel failure, opencv/FrameSourceWarp.cpp:301-304;
 * EOF at end of stream, :465-467) and std::string on program build failure (:191-195).
 * The C-ABI returns 0 or a negative code; text via vaw_last_error().  The C++ shim
 * (video_annotator_b200/host/FrameSourceWarp.cpp) turns non-zero into `throw int`. */
enum {
    VAW_OK = 0,
    VAW_ERR_INVALID = -2,     /* bad argument / unsupported parameter               */
    VAW_ERR_CUDA = -3,        /* CUDA runtime or driver error, or no device          */
    VAW_ERR_UNSUPPORTED = -4, /* e.g. an interpolation flag other than the three below */
    VAW_ERR_NOMEM = -5
};

/* ---- pixel formats ----------------------------------------------------------------
 * VAW_FORMAT_NV12: the buffer FrameSourceFfmpegOpenCl produces
 *   (opencv/FrameSourceFfmpegOpenCl.cpp:58,75-85): one 8-bit plane, `pitch` bytes per
 *   row, Y rows 0..H-1 then H/2 rows of interleaved U,V.  W and H even.
 * VAW_FORMAT_BGR24: interleaved 8UC3 -- what the reference actually hands to
 *   warp_frame after cvtColor (opencv/FrameSourceWarp.cpp:401,445).
 * VAW_FORMAT_GRAY8: single 8-bit plane.
 * VAW_FORMAT_NV12_TO_BGR24: NV12 source frames, BGR24 output frames -- the reference's literal per-frame
 *   pipeline, cvtColor(COLOR_YUV2BGR_NV12) (opencv/FrameSourceWarp.cpp:399-401) followed by the 3-channel
 *   remap (:306-312), bit-exact with doing the two steps one after the other on whole frames: every
 *   bilinear tap is a source pixel converted with OpenCV's fixed-point BT.601, the border value applies to
 *   the converted image (border[0..2] = B, G, R).  Source sizes even, output size free; INTER_LINEAR.
 *   Variant TILED (what AUTO picks): a few frames at a time are converted into a scratch that stays in L2 and
 *   warped from there by the staged BGR kernel; variant POLY: ONE launch, each tap converted on the fly
 *   (no intermediate image at all, 2.2x slower). */
enum { VAW_FORMAT_NV12 = 0, VAW_FORMAT_BGR24 = 1, VAW_FORMAT_GRAY8 = 2, VAW_FORMAT_NV12_TO_BGR24 = 3 };

/* cv::InterpolationFlags values accepted for the constructor's `interpolation` parameter
 * (opencv/FrameSourceWarp.hpp:90).  All four are cv::remap's 8-bit fixed-point filters, bit for bit. */
enum { VAW_INTER_NEAREST = 0, VAW_INTER_LINEAR = 1, VAW_INTER_CUBIC = 2, VAW_INTER_LANCZOS4 = 4 };

/* Kernel variants.  Every variant applies cv::remap's integer filter exactly to its own
 * map (vaw_dump_coords returns that map); they differ in how the map is evaluated:
 *   GATHER  per pixel, op for op as createMap.cl:15-49 (3 divides, sqrt, atan): the
 *           reference-order evaluation, ~50 instructions per pixel for the coordinates;
 *   POLY    once per 128x32-pixel piece in double precision on a sparse anchor grid, then a
 *           certified tensor polynomial inside the piece (within 3e-5 px of the exact
 *           projection, i.e. the correctly rounded fp32 map up to rare last-bit
 *           differences); pieces that cannot be certified use the GATHER evaluation.
 *   TILED   POLY's coordinates; the source rectangle of every piece is first copied into shared memory by
 *           the TMA engine (32/8/4-row boxes) and the taps are read from there; one CTA per piece, each
 *           warp a 64-column x PH/2-row quadrant with two columns per lane -- 64 / 72 registers, up to eight
 *           CTAs per SM (needs a 16-byte aligned source base / pitch / frame stride, else it gathers
 *           like POLY).  Exists for NV12 (csrc/vaw_tile.cu) and for GRAY8 / BGR24 (csrc/vaw_packed_tile.cu:
 *           the reference's literal 8UC3 frames, three bytes per tap address).
 *   PIPE    (retired in round 2: a persistent producer/consumer ring pipeline over the same tiles; slower
 *           than TILED on every workload once TILED ran eight CTAs per SM.  The value is refused.)
 *   TEX     TILED, except that pieces certified interior (every tap inside the source) are filtered
 *           by the texture units: the coordinate is rounded to 1/32 px exactly as cv::remap rounds
 *           it and the filtered value is rescaled and rounded half up like (sum + 512) >> 10.  The
 *           unit's internal precision is lower than cv::remap's 10-bit weights: 5 % of the samples of
 *           white noise come out 1 LSB off (never more; tests/test_gpu_parity.py), and it is slower
 *           than TILED (tex-pipe bound).  Kept for the comparison BASELINE.json asks for,
 *           never chosen by AUTO.  Needs a texture-aligned source base (512 bytes), a pitch that is
 *           a multiple of 32 and a frame stride that is a whole number of rows; else it runs as TILED.
 * POLY and TILED produce identical bytes on NV12.  AUTO = TILED for every format and filter the staged kernels carry
 * (INTER_LINEAR / NEAREST / CUBIC / LANCZOS4 with createMap.cl's projection pair), else GATHER.  GRAY8 / BGR24 accept GATHER and TILED, NV12 -> BGR24 POLY and TILED (INTER_LINEAR). */
enum {
    VAW_VARIANT_AUTO = 0,
    VAW_VARIANT_GATHER = 1,
    VAW_VARIANT_POLY = 2,
    VAW_VARIANT_TILED = 3,
    VAW_VARIANT_PIPE = 4, /* retired: VAW_ERR_UNSUPPORTED */
    VAW_VARIANT_TEX = 5
};

/* ---- parameters -------------------------------------------------------------------
 * The 8 scalars FrameSourceWarp::warp_frame passes to createMap, in its order
 * (opencv/FrameSourceWarp.cpp:283-290), kept as doubles and cast to float inside the
 * library exactly where the reference casts to cl_float; plus sizes, format, border. */
typedef struct vaw_params {
    double src_center_x, src_center_y; /* m_input_camera.matrix(0,2), (1,2)          */
    double src_focal_x, src_focal_y;   /* m_input_camera.matrix(0,0), (1,1)          */
    double map_center_x, map_center_y; /* m_output_camera.matrix(0,2), (1,2)         */
    double map_focal_x, map_focal_y;   /* m_output_camera.matrix(0,0), (1,1)         */
    int32_t src_width, src_height;     /* input image size (luma), <= 32766          */
    int32_t out_width, out_height;     /* m_output_camera.size, <= 32766             */
    int32_t format;                    /* VAW_FORMAT_*                               */
    int32_t interpolation;             /* VAW_INTER_LINEAR; VAW_INTER_NEAREST (cv::remap's
                                          cvRound of the map); VAW_INTER_CUBIC and
                                          VAW_INTER_LANCZOS4 (its 4x4 / 8x8 fixed-point
                                          filters).  All of them run on the staged-tile
                                          kernels (variant TILED, what AUTO picks)     */
    uint8_t border[4];                 /* NV12: Y,U,V  BGR24: B,G,R  (cv::remap's
                                          borderValue; OpenCV default is 0; the NV12
                                          neutral chroma is 128)                     */
    int32_t variant;                   /* VAW_VARIANT_*                              */
    float src_distortion[4];           /* extension: k1..k4 of the input camera's cv::fisheye
                                          distortion, theta_d = theta (1 + k1 theta^2 + ... + k4 theta^8)
                                          (Camera::distortion_coefficients, FrameSourceWarp.hpp:31; the
                                          presets set zeros, FrameSourceWarp.cpp:35, and createMap.cl
                                          ignores the field).  All zero = the reference's map, bit for
                                          bit; float like every scalar the reference hands its kernel   */
    int32_t projection;                /* extension: the projection pair.  0 = fisheye input, rectilinear
                                          output: createMap.cl, the reference's only pair.  Bit 0 (value 1):
                                          RECTILINEAR input camera (no atan step).  Bit 1 (value 2): FISHEYE
                                          (equidistant) output camera -- CameraModel, FrameSourceWarp.hpp:23-26;
                                          in_p / out_p = rect | fish of the wider toolchain, src/render.ts:611-618.
                                          Non-zero values: NV12 sources, INTER_LINEAR, variants AUTO/POLY/TILED   */
    int32_t reserved[2];
} vaw_params;

/* Camera description (opencv/FrameSourceWarp.hpp:14-34). */
typedef struct vaw_camera {
    int32_t model;     /* 0 RECTILINEAR, 1 FISHEYE (CameraModel, FrameSourceWarp.hpp:23-26) */
    int32_t width, height;
    int32_t reserved;
    double matrix[9];  /* row-major 3x3 */
    double distortion[4];
} vaw_camera;

/* CameraPreset, opencv/FrameSourceWarp.hpp:14-21 (same order, same values). */
enum {
    VAW_GOPRO_H4B_WIDE43_PUBLISHED = 0,
    VAW_GOPRO_H4B_WIDE43_MEASURED = 1,
    VAW_GOPRO_H4B_WIDE43_MEASURED_STABILISATION = 2,
    VAW_GOPRO_H4B_WIDE169_PUBLISHED = 3,
    VAW_GOPRO_H4B_WIDE169_MEASURED = 4,
    VAW_GOPRO_H4B_WIDE169_MEASURED_STABILISATION = 5
};

typedef struct vaw_ctx vaw_ctx;

/* ---- camera producers (host only, double precision) -------------------------------
 * Replace get_preset_camera (opencv/FrameSourceWarp.cpp:27-86) and get_output_camera
 * (:88-165), quirks included (int FOV constants, height-scaled fx, integer diagonals,
 * truncated size). */
int vaw_get_preset_camera(int preset, int width, int height, vaw_camera *out);
int vaw_get_output_camera(const vaw_camera *input, double scale, int crop_borders, double zoom,
                          vaw_camera *out);
/* Fill the 8 scalars + sizes + distortion of `p` from two cameras (FrameSourceWarp.cpp:283-290);
 * for NV12 the output size is rounded down to even; interpolation = VAW_INTER_LINEAR (the
 * constructor's default, FrameSourceWarp.hpp:90), variant = VAW_VARIANT_AUTO and `projection` from
 * the two cameras' models (FISHEYE input + RECTILINEAR output = 0 = createMap.cl); border and
 * reserved fields are left untouched.  vaw_get_output_camera is the reference's fisheye-only
 * derivation (VAW_ERR_UNSUPPORTED for a rectilinear input) and honours input->distortion the way
 * cv::fisheye::undistortPoints does (FrameSourceWarp.cpp:93-110). */
int vaw_params_from_cameras(const vaw_camera *input, const vaw_camera *output, int format,
                            vaw_params *p);

/* ---- context ----------------------------------------------------------------------
 * Replaces the state FrameSourceWarp's constructor builds (cameras, map buffers, the
 * compiled createMap program: opencv/FrameSourceWarp.cpp:199-226).  No map buffer is
 * allocated: the map is computed on the fly inside the sampler.  One ctx per device
 * per host thread; not thread-safe. */
int vaw_create(const vaw_params *params, int device, vaw_ctx **out);
void vaw_destroy(vaw_ctx *ctx);
const char *vaw_last_error(const vaw_ctx *ctx); /* ctx may be NULL: last create error */
const char *vaw_strerror(int code);
/* Bytes of one frame buffer with row pitch `pitch` (NV12: pitch*H*3/2).  For a VAW_FORMAT_NV12_TO_BGR24
 * context ask with VAW_FORMAT_NV12 for the source and VAW_FORMAT_BGR24 for the output (the combined
 * value itself returns 0). */
size_t vaw_frame_bytes(int format, int width, int height, int pitch);
/* Number of this library's kernels launched through ctx so far. */
uint64_t vaw_launch_count(const vaw_ctx *ctx);
/* The kernel variant the context resolved to (VAW_VARIANT_*, never AUTO). */
int vaw_get_variant(const vaw_ctx *ctx);
/* Host only: the 32 x 32 x 16 fixed-point weights of cv::remap(INTER_CUBIC) as this library builds them
 * ([fraction y][fraction x][tap row][tap column], scaled by 2^15; each block of 16 sums to 2^15). */
int vaw_cubic_table(int16_t out[16384]);
/* The same for INTER_LANCZOS4: 32 x 32 x 64 weights. */
int vaw_lanczos4_table(int16_t out[65536]);

/* ---- the warp ---------------------------------------------------------------------
 * vaw_warp replaces `cv::UMat FrameSourceWarp::warp_frame(cv::UMat input, cv::Mat rotation)`
 * (opencv/FrameSourceWarp.hpp:76, opencv/FrameSourceWarp.cpp:272-314): rotation is the
 * 3x3 CV_64F matrix row-major, cast to float per element as at :291-299.  src/dst are
 * DEVICE pointers; asynchronous on `stream` (a cudaStream_t; NULL = default stream).
 * Unlike the reference the caller owns the output buffer. */
int vaw_warp(vaw_ctx *ctx, const uint8_t *src, int src_pitch, uint8_t *dst, int dst_pitch,
             const double rotation[9], void *stream);

/* Same for n_frames frames in ONE launch: frame i is read at src + i*src_frame_stride,
 * written at dst + i*dst_frame_stride, with rotation rotations[9*i .. 9*i+8] (fp32,
 * row-major, DEVICE memory).  This is how a 3-4 us frame stays off the launch-latency
 * floor. */
int vaw_warp_batch(vaw_ctx *ctx, const uint8_t *src, int src_pitch, size_t src_frame_stride,
                   uint8_t *dst, int dst_pitch, size_t dst_frame_stride,
                   const float *rotations, int n_frames, void *stream);

/* Optional hint for callers whose source frames live in ONE slab of n_slots equally spaced frames
 * (a frame pool / decoder surface ring; the C++ shim's FramePool): vaw_warp / vaw_warp_batch calls
 * whose source lies inside the slab then share the slab's TMA tensor maps (encoded once) instead of
 * encoding maps per distinct source pointer.  base = NULL unbinds.  Purely a launch-cost matter:
 * results are identical with and without it. */
int vaw_bind_clip(vaw_ctx *ctx, const uint8_t *base, int pitch, size_t frame_stride, int n_slots);

/* double[9*n] (host) -> float[9*n] (device), the (cl_float) cast of FrameSourceWarp.cpp:291-299.
 * Synchronous with respect to the host buffer. */
int vaw_upload_rotations(vaw_ctx *ctx, const double *rotations_host, int n_frames,
                         float *rotations_dev, void *stream);

/* Whole path with HOST buffers (tightly packed frames: pitch = width*channels):
 * pinned staging, host->device copy, warp, device->host copy, pipelined in chunks on
 * the ctx's own streams.  Returns when dst_host is complete.  This is the call a
 * FrameSource-style user makes when frames live in host memory. */
int vaw_warp_batch_host(vaw_ctx *ctx, const uint8_t *src_host, uint8_t *dst_host,
                        const double *rotations_host, int n_frames);

/* The map createMap.cl would have written (opencv/createMap.cl:42-49), from the same
 * device function the sampler uses.  plane 0: luma map (out_h x out_w); plane 1: NV12
 * chroma map (out_h/2 x out_w/2).  map pitch in floats.  For parity tests/debugging. */
int vaw_dump_coords(vaw_ctx *ctx, const double rotation[9], int plane, float *map_x,
                    float *map_y, int map_pitch, void *stream);

/* The reference's second pass on its own: cv::remap(src, dst, map_x, map_y, INTER_LINEAR,
 * BORDER_CONSTANT, border) as called at opencv/FrameSourceWarp.cpp:306-312, for 8-bit
 * images of 1..3 interleaved channels and CV_32FC1 maps (map pitch in floats), all in
 * DEVICE memory.  Same integer filter as the fused kernels; exists so that callers holding
 * a map can use it and so that the filter is pinned against cv::remap golden vectors. */
int vaw_remap_u8(const uint8_t *src, int src_width, int src_height, int src_pitch, int channels,
                 const float *map_x, const float *map_y, int rows, int cols, int map_pitch,
                 uint8_t *dst, int dst_pitch, const uint8_t border[4], int device, void *stream);

/* ---- device memory for hosts that do not link the CUDA runtime -----------------------------
 * The reference hands cv::UMat frames around (opencv/FrameSource.hpp:16,22); a host-language
 * binding (the C++ shim under video_annotator_b200/host, cgo, JNI ...) owns device frames
 * through these instead.  vaw_memcpy and vaw_sync block until the stream has drained. */
int vaw_malloc(int device, size_t bytes, void **out);
int vaw_free(int device, void *ptr);
int vaw_memcpy(int device, void *dst, const void *src, size_t bytes, int to_device, void *stream);
int vaw_sync(int device, void *stream);

/* ---- frame-parallel sharding ---------------------------------------------------------
 * The reference is single device, single thread; once every frame has its rotation the
 * frames are independent (warp_frame reads only its frame, its rotation and constant
 * intrinsics: opencv/FrameSourceWarp.cpp:272-314), so a clip shards into contiguous frame
 * ranges with no exchange between devices.
 * vaw_shard_range: the range [first, first + count) of part `part` of `n_parts`.
 * vaw_clip_*: one context, one host thread and one set of streams per device; every device
 * warps its own range of a host-resident clip through the pinned-staging pipeline of
 * vaw_warp_batch_host.  `devices` = CUDA ordinals (NULL: 0 .. n_devices-1). */
int vaw_shard_range(int n_frames, int n_parts, int part, int *first, int *count);
/* Placement of the host threads that stage frames for a device (Linux): the NUMA node the device's PCIe
 * root hangs off (-1: unknown / single node), and "bind the CALLING thread to that node's CPUs" (returns
 * the number of CPUs bound to, 0 if nothing was changed).  Pinned staging buffers allocated by a bound
 * thread are first-touched on the device's node, so H2D / D2H traffic does not cross the socket link.
 * vaw_clip_warp_host binds its per-device worker threads this way. */
int vaw_device_numa_node(int device);
int vaw_bind_thread_to_device(int device);
typedef struct vaw_clip vaw_clip;
int vaw_clip_create(const vaw_params *params, int n_devices, const int *devices, vaw_clip **out);
void vaw_clip_destroy(vaw_clip *clip);
int vaw_clip_warp_host(vaw_clip *clip, const uint8_t *src_host, uint8_t *dst_host,
                       const double *rotations_host, int n_frames);
const char *vaw_clip_last_error(const vaw_clip *clip);

/* ---- the step before the warp ------------------------------------------------------------
 * cv::cvtColor(frame, output, COLOR_YUV2BGR_NV12) as the reference runs it on every frame before
 * buffering and warping it (opencv/FrameSourceWarp.cpp:399-401): NV12 (width x 3*height/2 bytes,
 * even sizes) -> interleaved BGR 8UC3, bit-exact with OpenCV's fixed-point BT.601.  DEVICE
 * pointers, n_frames frames at constant strides, asynchronous on `stream`.  Feeding its output
 * to a VAW_FORMAT_BGR24 context reproduces the reference's literal pipeline. */
int vaw_nv12_to_bgr(const uint8_t *src, int width, int height, int src_pitch, size_t src_frame_stride,
                    uint8_t *dst, int dst_pitch, size_t dst_frame_stride, int n_frames, int device,
                    void *stream);

/* ---- motion measurement ---------------------------------------------------------------
 * The tracking step of FrameSourceWarp::consume_frame (opencv/FrameSourceWarp.cpp:421-427 ->
 * find_point_pairs_with_optical_flow, :242-270): cv::calcOpticalFlowPyrLK(prev_frame, current_frame,
 * prev_corners) with OpenCV's defaults (winSize 21x21, maxLevel 3, 30 iterations / eps 0.01,
 * minEigThreshold 1e-4) on the luma planes of consecutive frames.  A vaw_flow keeps the image pyramid
 * (cv::pyrDown, bit-exact) and the Scharr derivatives of the previous and the current frame in device
 * memory: vaw_flow_push_frame builds the pyramid of a new frame (a DEVICE luma plane), vaw_flow_track
 * follows HOST point arrays (x, y pairs) from the previous into the current frame; status[i] = 1 where the
 * flow was found (the reference keeps exactly those pairs, :262-268).  The arithmetic is OpenCV's
 * fixed-point scheme (oracle/lk_ref.py, pinned to cv2.calcOpticalFlowPyrLK).
 * vaw_flow_get_level: a pyramid level (which: 0 previous, 1 current) and its derivatives (dx, dy int16
 * pairs) copied to host buffers, for parity tests. */
typedef struct vaw_flow vaw_flow;
int vaw_flow_create(int width, int height, int device, vaw_flow **out);
void vaw_flow_destroy(vaw_flow *flow);
const char *vaw_flow_last_error(const vaw_flow *flow);
int vaw_flow_levels(const vaw_flow *flow);
int vaw_flow_push_frame(vaw_flow *flow, const uint8_t *luma, int pitch, void *stream);
int vaw_flow_track(vaw_flow *flow, const float *prev_pts_xy, int n, float *next_pts_xy, uint8_t *status,
                   void *stream);
int vaw_flow_get_level(vaw_flow *flow, int which, int level, uint8_t *image_host, int16_t *deriv_host,
                       int *width, int *height);
/* The corner step (opencv/FrameSourceWarp.cpp:228-240, called at :422 on the PREVIOUS frame): cv::goodFeaturesToTrack
 * (image, corners, max_corners, quality, min_distance) with OpenCV's defaults otherwise (blockSize 3, Sobel 3,
 * minimum-eigenvalue response; the reference passes 200, 0.01, 30).  which: 0 previous, 1 current frame of the
 * tracker.  Writes up to `capacity` corners as (x, y) pairs to the HOST array, *n_out = the number found.  The
 * response map and the candidate list are computed on the device; the ordered greedy distance filter runs on the
 * host.  vaw_flow_get_response: the cv::cornerMinEigenVal map of the last call (width x height floats). */
int vaw_flow_corners(vaw_flow *flow, int which, int max_corners, double quality, double min_distance,
                     float *corners_xy, int capacity, int *n_out, void *stream);
int vaw_flow_get_response(vaw_flow *flow, float *response_host);

/* guess_camera_rotation (opencv/FrameSourceWarp.cpp:316-368): the rotation of the camera between two frames from
 * tracked point pairs (pixels of the INPUT camera; x, y pairs).  Host only.  The reference undistorts both sets with
 * cv::fisheye::undistortPoints, randomises the depth of the previous rays and keeps the rotation and the inlier count
 * of cv::solvePnPRansac(100 iterations, 8 px reprojection error in the OUTPUT camera, confidence 0.99); this fits the
 * rotation directly (two-ray RANSAC, same threshold and counts, least-squares refit over the consensus set) and
 * agrees with that route to a few hundredths of a degree.  rotation: row-major 3x3, maps a ray of the previous frame
 * onto the current one; *inliers: size of the consensus set (the reference ignores a result with fewer than 40,
 * :431-438).  `seed` makes the sampling reproducible. */
int vaw_guess_rotation(const vaw_camera *input, const vaw_camera *output, const float *prev_xy, const float *cur_xy,
                       int n, uint32_t seed, double rotation[9], int *inliers);

/* ---- synthetic frames (decode is out of scope; BASELINE.json north_star) ------------
 * Fill n_frames NV12 frames in device memory with the integer test pattern
 * (frame index first_index + i). */
int vaw_synth_nv12(uint8_t *dst, int width, int height, int pitch, size_t frame_stride,
                   int first_index, int n_frames, uint32_t seed, int white_noise, int device,
                   void *stream);

/* ---- diagnostics ------------------------------------------------------------------
 * vaw_set_option(ctx, "force_exact", 1): evaluate every pixel with the IEEE library
 * division/sqrt sequences instead of the range-certified shared-reciprocal ones (the two
 * are bit-identical; tests compare them).
 * vaw_selftest_math: on-device comparison of those fast sequences with __frcp_rn /
 * __fdiv_rn / __fsqrt_rn / the k = atan(r)/r step on random operands in the certified
 * ranges; mismatches[4] = {rcp, div, sqrt, k}. */
int vaw_set_option(vaw_ctx *ctx, const char *name, int value);
/* Other options: "split_builder" (default 0): batches of >= 32 frames build the piece table of all but the
 * first 8 frames on a high-priority side stream while the sampler already works on those 8 (two sampler
 * launches per batch; same bytes).  Measured on B200: the builder leaves the critical path but its
 * instructions compete with the sampler for the same issue slots -- the step time does not change
 * (DESIGN.md), hence off by default.  "host_stages" (2..8, default 4) and "host_chunk_mb" (default 32):
 * chunks in flight and chunk size of the host-buffer pipeline (vaw_warp_batch_host). */
/* vaw_set_option(ctx, "time_kernels", 1) makes every NV12 launch record CUDA events on its
 * stream around the piece-table builder and the warp kernel (a ring of the last 512 launches);
 * (builder_ms covers the head frames' table, the rest of the table is built on a side stream under the
 * sampler, see "split_builder").  vaw_kernel_times returns the most recent launches' durations in milliseconds, oldest first
 * (it waits for them).  This is what bench.py's roofline line is computed from. */
int vaw_kernel_times(vaw_ctx *ctx, int max_launches, float *builder_ms, float *warp_ms, int *n_out);
/* The same launches' warp time split into the texture kernel (variant TEX, else ~0) and the tile kernel. */
int vaw_kernel_times_split(vaw_ctx *ctx, int max_launches, float *tex_ms, float *tile_ms, int *n_out);

/* Variant POLY: how the 128x32-pixel pieces of the output classify for `rotation`:
 * counts = {pieces, with a certified polynomial, of those fully inside the source, of those
 * fully outside (pure border fill), largest shared-memory tile a piece needs (bytes), the
 * context's tile capacity (bytes), pieces that exceed it (they gather from global memory), 0}.
 * All zero for GATHER. */
int vaw_piece_stats(vaw_ctx *ctx, const double rotation[9], uint32_t counts[8], void *stream);
/* The flags of every piece for `rotation`, row-major (pieces_y x pieces_x pieces of 128 x piece_h pixels):
 * bit 0 = certified polynomial coordinates (else the piece is evaluated per pixel, op for op), bit 1 = every
 * tap inside the source, bit 2 = every tap outside (border fill).  For the accuracy sweeps of the tests. */
int vaw_piece_flags(vaw_ctx *ctx, const double rotation[9], uint32_t *flags, int capacity, int *pieces_x_out,
                    int *pieces_y_out, int *piece_h_out, void *stream);
/* Four words per piece for `rotation`, row-major: flags, shared-memory bytes the piece's source box needs
 * (0: none, 0x7fffffff: cannot be staged), tile row pitch, luma rows | chroma rows << 16.  Analysis / tests. */
int vaw_piece_tiles(vaw_ctx *ctx, const double rotation[9], uint32_t *out, int capacity, void *stream);
int vaw_selftest_math(int device, uint32_t seed, uint64_t n_per_thread, uint64_t mismatches[4]);
/* Instrumented builds only (-DVAW_BOUNDS_CHECK): number of shared-memory tap addresses of variant
 * TILED that fell outside their staged tile since the library was loaded; -1 in a normal build. */
long long vaw_debug_oob_count(int device);

#ifdef __cplusplus
}
#endif
#endif /* VAW_H */
