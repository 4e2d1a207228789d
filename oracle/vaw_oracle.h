/*
 * vaw_oracle.h -- CPU oracle for the video-annotator per-frame warp path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (libvaw.so)
 * never links, loads or calls it and has no CPU fallback.
 *
 * What it restates (citations are relative to /root/reference/):
 *   - opencv/createMap.cl:10-50            -> vaw_oracle_create_map()
 *   - opencv/FrameSourceWarp.cpp:280-300   -> argument order of the above
 *   - cv::remap INTER_LINEAR/BORDER_CONSTANT, 8-bit, called at
 *     opencv/FrameSourceWarp.cpp:306-312   -> vaw_oracle_remap_u8()
 *     (third-party: OpenCV imgproc, not vendored in the reference; the image
 *     carries opencv-python-headless 4.13.0, the reference pins >= 4.5 in
 *     opencv/meson.build:33)
 *   - cv::cvtColor(COLOR_YUV2BGR_NV12), called at
 *     opencv/FrameSourceWarp.cpp:399-401   -> vaw_oracle_nv12_to_bgr()  (PINNED to cv2.cvtColor)
 *   - opencv/FrameSourceWarp.cpp:27-86     -> vaw_oracle_get_preset_camera()
 *   - opencv/FrameSourceWarp.cpp:88-165    -> vaw_oracle_get_output_camera()
 *
 * Pinning status:
 *   - pixel stage: PINNED.  vaw_oracle_remap_u8 is checked bit-for-bit against
 *     cv2.remap (the real cv::remap) in tests/test_oracle_remap.py and against
 *     the committed fixtures tests/golden/remap_*.npz produced by it.
 *   - coordinate stage: PINNED to the reference's own kernel.  The reference has
 *     no tests or golden vectors and there is no OpenCL runtime here, but
 *     createMap.cl compiles unmodified with gcc behind a shim header
 *     (oracle/ref_build -> oracle/_ref/libcreatemap_ref.so).  The transcription
 *     equals that library bit for bit (tests/test_oracle_ref.py: live, and via
 *     the fixture tests/golden/createmap_ref.npz generated from it); it is also
 *     cross-checked against cv2.fisheye.initUndistortRectifyMap (fp32 rounding
 *     level agreement).
 *   - camera stage: cross-checked against cv2.fisheye.undistortPoints.
 *
 * NV12-plane semantics (no reference behaviour exists: the reference warps BGR,
 * opencv/FrameSourceWarp.cpp:401,445) are defined in nv12_warp_ref.c.
 */
#ifndef VAW_ORACLE_H
#define VAW_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The 8 scalars FrameSourceWarp::warp_frame passes (FrameSourceWarp.cpp:283-290),
 * already cast to float as the reference does with (cl_float). */
typedef struct vaw_oracle_intrinsics {
    float src_center_x, src_center_y, src_focal_x, src_focal_y; /* input camera  */
    float map_center_x, map_center_y, map_focal_x, map_focal_y; /* output camera */
    /* Extension: k1..k4 of the input camera's cv::fisheye distortion,
     * theta_d = theta (1 + k1 theta^2 + k2 theta^4 + k3 theta^6 + k4 theta^8)
     * (Camera::distortion_coefficients, FrameSourceWarp.hpp:31).  The reference's presets set
     * zeros (FrameSourceWarp.cpp:35) and createMap.cl ignores the field: all zero reproduces
     * createMap.cl bit for bit.  Pinned against cv2.fisheye.initUndistortRectifyMap. */
    float dist[4];
} vaw_oracle_intrinsics;

/* createMap.cl:1-51.  map_x/map_y: rows x cols fp32, row stride `step` floats.
 * rot = rot00..rot22 row-major (createMap.cl:6-8).  n_threads splits rows. */
void vaw_oracle_create_map(float *map_x, float *map_y, int rows, int cols, int step,
                           const vaw_oracle_intrinsics *k, const float rot[9],
                           int n_threads);

/* One coordinate (same arithmetic as the loop body above). */
void vaw_oracle_create_map_point(int u, int v, const vaw_oracle_intrinsics *k,
                                 const float rot[9], float *mx, float *my);

/* Chroma map from a luma map (NV12 semantics, nv12_warp_ref.c).
 * cmap_*: (rows/2) x (cols/2), stride cstep floats. */
void vaw_oracle_chroma_map(const float *map_x, const float *map_y, int rows, int cols, int step,
                           float *cmap_x, float *cmap_y, int cstep, int n_threads);

/* cv::remap(INTER_LINEAR, BORDER_CONSTANT) on 8-bit data with cn channels (1..4).
 * src: src_h x src_w x cn, row pitch src_pitch bytes. dst: rows x cols x cn. */
void vaw_oracle_remap_u8(const uint8_t *src, int src_w, int src_h, int src_pitch, int cn,
                         const float *map_x, const float *map_y, int rows, int cols, int map_step,
                         uint8_t *dst, int dst_pitch, const uint8_t *border, int n_threads);

/* cv::remap(INTER_CUBIC, BORDER_CONSTANT), same arguments (remap_cubic_ref.c); the 32 x 32 x 16
 * fixed-point weight table it uses (weights scaled by 2^15, each block sums to 2^15). */
void vaw_oracle_remap_cubic_u8(const uint8_t *src, int src_w, int src_h, int src_pitch, int cn,
                               const float *map_x, const float *map_y, int rows, int cols, int map_step,
                               uint8_t *dst, int dst_pitch, const uint8_t *border, int n_threads);
const short *vaw_oracle_cubic_table(void);
/* cv::remap(INTER_LANCZOS4, BORDER_CONSTANT): 8 x 8 taps, table 32 x 32 x 64. */
void vaw_oracle_remap_lanczos4_u8(const uint8_t *src, int src_w, int src_h, int src_pitch, int cn,
                                  const float *map_x, const float *map_y, int rows, int cols, int map_step,
                                  uint8_t *dst, int dst_pitch, const uint8_t *border, int n_threads);
const short *vaw_oracle_lanczos4_table(void);

/* Full NV12 path: luma map -> chroma map -> remap of both planes.
 * src: (src_h*3/2) rows of src_pitch bytes; dst likewise with out_h. */
void vaw_oracle_warp_nv12(const uint8_t *src, int src_w, int src_h, int src_pitch,
                          uint8_t *dst, int out_w, int out_h, int dst_pitch,
                          const vaw_oracle_intrinsics *k, const float rot[9],
                          int border_y, int border_u, int border_v, int n_threads);

/* Literal reference behaviour: one remap of an interleaved 8UC3 frame. */
void vaw_oracle_warp_bgr(const uint8_t *src, int src_w, int src_h, int src_pitch,
                         uint8_t *dst, int out_w, int out_h, int dst_pitch,
                         const vaw_oracle_intrinsics *k, const float rot[9],
                         const uint8_t border[3], int n_threads);

/* cv::cvtColor(COLOR_YUV2BGR_NV12), FrameSourceWarp.cpp:399-401.  src: NV12 buffer (h*3/2 rows of
 * src_pitch bytes), dst: h x w x 3. */
void vaw_oracle_nv12_to_bgr(const uint8_t *src, int w, int h, int src_pitch, uint8_t *dst, int dst_pitch,
                            int n_threads);

/* Count unique source bytes touched by the in-range taps. */
int64_t vaw_oracle_touched_bytes(const float *map_x, const float *map_y, int rows, int cols,
                                 int map_step, int src_w, int src_h, int cn);

/* Camera model (FrameSourceWarp.hpp:14-34), doubles as in the reference. */
typedef struct vaw_oracle_camera {
    int model;          /* 0 RECTILINEAR, 1 FISHEYE (FrameSourceWarp.hpp:23-26) */
    double matrix[9];   /* row-major 3x3 */
    double dist[4];
    int width, height;
} vaw_oracle_camera;

void vaw_oracle_get_preset_camera(int preset, int width, int height, vaw_oracle_camera *out);
void vaw_oracle_get_output_camera(const vaw_oracle_camera *in, double scale, int crop_borders,
                                  double zoom, vaw_oracle_camera *out);

/* Synthetic content (integer recipe, identical to the device generator). */
void vaw_oracle_synth_nv12(uint8_t *dst, int w, int h, int pitch, int frame_index,
                           uint32_t seed, int white_noise);
uint32_t vaw_oracle_hash32(uint32_t seed, uint32_t n, uint32_t plane, uint32_t y, uint32_t x);

#ifdef __cplusplus
}
#endif
#endif
