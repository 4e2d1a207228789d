// FrameSourceWarp.hpp -- drop-in for the reference's stabilise-and-reproject stage.
//
// Mirrors /root/reference/opencv/FrameSourceWarp.hpp:14-94: CameraPreset, CameraModel,
// Camera, and class FrameSourceWarp with the same constructor parameter list and the same
// pull_frame / peek_frame behaviour, quirks included (frame 0 is never emitted,
// FrameSourceWarp.cpp:403-406; look-ahead of smooth_radius frames, :453; EOF padding and
// draining, :456-467; peek_frame advances like pull_frame, :478-480).
//
// What is replaced: warp_frame (:272-314) -- the createMap OpenCL kernel plus cv::remap --
// is ONE call into libvaw.so (vaw_warp), hand-written CUDA for sm_100a; no map buffers exist.
// The measurement of the inter-frame rotation (:228-270, :316-375) sits behind the RotationSource
// interface: OpticalFlowRotationSource is the reference's own chain (corners, pyramidal LK, rotation fit) on the
// GPU; tests and the demo also plug in synthetic gyro sources.  Everything downstream of the measurement
// (accumulation :441-442, the Savitzky-Golay look-ahead filter :212/:444/:471, the correction :472-475) is kept.
#ifndef VAW_FRAME_SOURCE_WARP_HPP_
#define VAW_FRAME_SOURCE_WARP_HPP_

#include <deque>
#include <memory>
#include <queue>
#include <vector>

#include "FrameSource.hpp"

enum CameraPreset {
    GOPRO_H4B_WIDE43_PUBLISHED,
    GOPRO_H4B_WIDE43_MEASURED,
    GOPRO_H4B_WIDE43_MEASURED_STABILISATION,
    GOPRO_H4B_WIDE169_PUBLISHED,
    GOPRO_H4B_WIDE169_MEASURED,
    GOPRO_H4B_WIDE169_MEASURED_STABILISATION
};

enum CameraModel { RECTILINEAR, FISHEYE };

enum InterpolationFlags { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_LANCZOS4 = 4 };  // cv::InterpolationFlags values

struct Mat33 {  // stands in for cv::Mat 3x3 CV_64F / cv::Matx33d, row-major
    double m[9];
    static Mat33 eye();
    Mat33 operator*(const Mat33& o) const;
    Mat33 t() const;    // transpose
    Mat33 inv() const;  // general 3x3 inverse (the reference calls cv::Mat::inv, :472, :475)
};

class Camera {
  public:
    CameraModel model;
    Mat33 matrix;
    double distortion_coefficients[4];
    int width, height;  // cv::Size size
};

Camera get_preset_camera(CameraPreset preset, int width, int height);                     // :27-86
Camera get_output_camera(const Camera& input, double scale, bool crop_borders, double zoom);  // :88-165

/**
 * The measured rotation of the camera between the previous frame and this one -- what
 * guess_camera_rotation (:316-375) returns and consume_frame left-multiplies (:441-442).
 * One call per consumed frame after the first; return false to reuse the previous
 * inter-frame rotation (the reference does that with fewer than 40 RANSAC inliers, :431-438).
 */
class RotationSource {
  public:
    // every consumed frame (the first one included) is shown to the source before it is asked about it
    virtual void observe_frame(const Frame& /*frame*/, long /*frame_index*/) {}
    virtual bool rotation_since_last_frame(long frame_index, Mat33& out) = 0;
    virtual ~RotationSource() = default;
};

/**
 * The reference's own measurement (consume_frame :403-438): corners of the previous frame
 * (find_corners = cv::goodFeaturesToTrack(200, 0.01, 30), re-detected when the set is older than 20 frames
 * or has shrunk below 150, :415-419), followed into the current frame with pyramidal Lucas-Kanade
 * (find_point_pairs_with_optical_flow, :242-270), then the rotation that explains the pairs
 * (guess_camera_rotation, :316-368; fewer than 40 inliers: keep the previous rotation, :431-438).
 * Corners, pyramids and tracking run on the GPU (vaw_flow_*), the rotation fit on the host (vaw_guess_rotation).
 */
class OpticalFlowRotationSource : public RotationSource {
    Camera m_input_camera, m_output_camera;
    struct vaw_flow* m_flow = nullptr;
    int m_device;
    long m_last_key_frame_index = -1;
    std::vector<float> m_last_input_frame_corners;  // x, y pairs in the previous frame
    bool m_have_rotation = false;
    Mat33 m_rotation;
    int m_last_inliers = 0, m_last_pairs = 0;
  public:
    OpticalFlowRotationSource(const Camera& input_camera, const Camera& output_camera, int device = 0);
    ~OpticalFlowRotationSource() override;
    OpticalFlowRotationSource(const OpticalFlowRotationSource&) = delete;
    OpticalFlowRotationSource& operator=(const OpticalFlowRotationSource&) = delete;
    void observe_frame(const Frame& frame, long frame_index) override;
    bool rotation_since_last_frame(long frame_index, Mat33& out) override;
    int last_inliers() const { return m_last_inliers; }  // of the most recent fit
    int last_pairs() const { return m_last_pairs; }      // tracked point pairs it was given
};

/**
 * Savitzky-Golay smoothing of a rotation sequence, evaluated at the centre of a window of
 * 2*radius+1 samples (the reference builds gram_sg::RotationFilter from
 * SavitzkyGolayFilterConfig(radius, 0, 2, 0), :212): weights applied element-wise to the
 * matrices, result projected back to SO(3) (U V^T of the SVD = the orthogonal polar factor).
 *
 * Start-up.  The reference only ever calls add() (:444, :459) and filter() (:471), and its first
 * filter() comes after radius + 1 add() calls (:453) although the library's Savitzky-Golay filter
 * needs a full window of 2*radius+1 samples: the window is therefore pre-filled by the library's
 * constructor, which cannot know any sample -- gram_sg::RotationFilter's constructor resets its
 * circular buffer to ZERO matrices (spatial_filters.cpp of the published library, restated; add()
 * is a plain push_back).  So until 2*radius+1 real samples have arrived the smoothed rotation is
 * the polar factor of a ONE-SIDED weighted sum, not a centred one.  StartUp::Zeros (default)
 * reproduces that; StartUp::FirstSample is round 1's convention (window pre-filled with the first
 * sample, "as if the camera had been still").  The library is not vendored in the reference
 * (meson.build:37, no version pin), so this stage stays formally unpinned.
 */
class RotationFilter {
  public:
    enum class StartUp { Zeros, FirstSample };
  private:
    int m_radius;
    StartUp m_start;
    std::vector<double> m_weights;
    std::deque<Mat33> m_window;
  public:
    explicit RotationFilter(int radius, StartUp start = StartUp::Zeros);
    void add(const Mat33& rotation);
    Mat33 filter() const;
};

// The stabilise-and-reproject stage: pulls frames from `source`, smooths the camera path over a
// look-ahead window and hands out each frame re-projected (fisheye -> rectilinear) under the
// rotation that cancels the shake.  Class and member names follow FrameSourceWarp.hpp:40-94.
class FrameSourceWarp : public FrameSource {
    std::shared_ptr<FrameSource> m_source;
    std::shared_ptr<RotationSource> m_rotation_source;

    Camera m_input_camera;
    Camera m_output_camera;
    struct vaw_ctx* m_ctx = nullptr;  // replaces m_map_x, m_map_y, m_remap_kernel (:47-49)
    int m_device = 0, m_format = 0;

    long m_frame_index = 0;
    Mat33 m_measured_rotation;
    Mat33 m_last_frame_rotation;

    unsigned int m_smooth_radius;
    InterpolationFlags m_interpolation;

    RotationFilter m_rotation_filter;
    std::queue<Frame> m_buffered_frames;
    std::queue<Mat33> m_buffered_rotations;

    // batched warping of look-ahead frames (set_warp_batch): outputs ready to hand out, device rotations,
    // an upstream error that surfaces after the frames that precede it
    static constexpr int kMaxWarpBatch = 64;
    int m_warp_batch = 1;
    std::deque<Frame> m_ready;
    float* m_rot_dev = nullptr;
    const uint8_t* m_bound_base = nullptr;
    long m_batch_launches = 0;
    int m_deferred_error = 0;
    bool m_deferred_error_set = false;

    void consume_frame(Frame input_frame);
    bool next_frame_and_rotation(Frame& frame, Mat33& rotation);  // pull_frame (:452-476) up to warp_frame
    void create_context();  // vaw_create from m_input_camera / m_output_camera / m_format / m_interpolation

  protected:
    // warp_frame(input, rotation), :272-314.  Virtual so that the state machine can be tested
    // without a GPU; the real one calls vaw_warp and synchronises before returning.
    virtual Frame warp_frame(Frame input, const Mat33& rotation);
    // The same for several frames whose rotations are known: one vaw_warp_batch per run of equally spaced
    // frames, one synchronisation.  Without a device context it calls warp_frame per frame.
    virtual std::vector<Frame> warp_frames(const std::vector<Frame>& inputs, const std::vector<Mat33>& rotations);

  public:
    FrameSourceWarp(
      std::shared_ptr<FrameSource> source,
      CameraPreset input_camera,
      double scale = 1,
      bool crop_borders = false,
      double zoom = 1,
      int smooth_radius = 30,
      InterpolationFlags interpolation = INTER_LINEAR,
      std::shared_ptr<RotationSource> rotation_source = nullptr,  // nullptr: the camera does not rotate
      bool create_device_context = true                           // false: CPU-only tests of the state machine
    );
    // Overload with explicit cameras: an output camera that get_output_camera cannot express
    // (BASELINE config 5: 5312x2988 -> 3840x2160 with its own focal length; the out_w / out_h / out_fx /
    // out_fy options of the wider toolchain, src/render.ts:678-681), or an input camera with distortion.
    FrameSourceWarp(
      std::shared_ptr<FrameSource> source,
      const Camera& input_camera,
      const Camera& output_camera,
      int smooth_radius = 30,
      InterpolationFlags interpolation = INTER_LINEAR,
      std::shared_ptr<RotationSource> rotation_source = nullptr,
      bool create_device_context = true
    );
    ~FrameSourceWarp() override;
    Frame pull_frame() override;
    Frame peek_frame() override;

    const Camera& input_camera() const { return m_input_camera; }
    const Camera& output_camera() const { return m_output_camera; }
    int output_width() const;   // even for NV12
    int output_height() const;
    // Warp up to `frames` look-ahead frames per launch (default 1 = the reference's pull pattern: one
    // upstream pull and one warp per call).  Frames, order and rotations are unchanged; only the number of
    // frames pulled from upstream before the first output grows by frames - 1.
    void set_warp_batch(int frames);
    long batch_launches() const { return m_batch_launches; }
};

#endif  // VAW_FRAME_SOURCE_WARP_HPP_
