// FrameSourceWarp.cpp -- host-side state machine of the warp stage over the C-ABI.
// Follows /root/reference/opencv/FrameSourceWarp.cpp:199-226 (constructor), :397-450
// (consume_frame, minus the optical-flow measurement), :452-480 (pull_frame / peek_frame).
#include "FrameSourceWarp.hpp"

#include <cmath>
#include <cstdio>

#include "../../include/vaw.h"

// ---- small 3x3 algebra (stands in for cv::Mat arithmetic) ----------------------------------
Mat33 Mat33::eye() { return Mat33{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }

Mat33 Mat33::operator*(const Mat33& o) const
{
    Mat33 r{};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[3 * i + j] = m[3 * i] * o.m[j] + m[3 * i + 1] * o.m[3 + j] + m[3 * i + 2] * o.m[6 + j];
    return r;
}

Mat33 Mat33::t() const { return Mat33{{m[0], m[3], m[6], m[1], m[4], m[7], m[2], m[5], m[8]}}; }

Mat33 Mat33::inv() const
{
    const double* a = m;
    const double c0 = a[4] * a[8] - a[5] * a[7], c1 = a[5] * a[6] - a[3] * a[8], c2 = a[3] * a[7] - a[4] * a[6];
    const double det = a[0] * c0 + a[1] * c1 + a[2] * c2;
    const double d = 1.0 / det;
    return Mat33{{c0 * d, (a[2] * a[7] - a[1] * a[8]) * d, (a[1] * a[5] - a[2] * a[4]) * d,
                  c1 * d, (a[0] * a[8] - a[2] * a[6]) * d, (a[2] * a[3] - a[0] * a[5]) * d,
                  c2 * d, (a[1] * a[6] - a[0] * a[7]) * d, (a[0] * a[4] - a[1] * a[3]) * d}};
}

// ---- device frames ---------------------------------------------------------------------------
DeviceFrame::~DeviceFrame()
{
    if (data) vaw_free(device, data);
}

Frame make_device_frame(int device, int format, int width, int height)
{
    auto f = std::make_shared<DeviceFrame>();
    const int channels = format == VAW_FORMAT_BGR24 ? 3 : 1;
    f->width = width; f->height = height; f->pitch = width * channels; f->format = format; f->device = device;
    f->bytes = vaw_frame_bytes(format, width, height, f->pitch);
    void* p = nullptr;
    const int rc = vaw_malloc(device, f->bytes, &p);
    if (rc != VAW_OK) throw rc;
    f->data = static_cast<uint8_t*>(p);
    return f;
}

// ---- cameras -----------------------------------------------------------------------------------
static Camera camera_from(const vaw_camera& c)
{
    Camera out{};
    out.model = c.model ? FISHEYE : RECTILINEAR;
    for (int i = 0; i < 9; ++i) out.matrix.m[i] = c.matrix[i];
    for (int i = 0; i < 4; ++i) out.distortion_coefficients[i] = c.distortion[i];
    out.width = c.width; out.height = c.height;
    return out;
}

static vaw_camera camera_to(const Camera& c)
{
    vaw_camera out{};
    out.model = c.model == FISHEYE;
    for (int i = 0; i < 9; ++i) out.matrix[i] = c.matrix.m[i];
    for (int i = 0; i < 4; ++i) out.distortion[i] = c.distortion_coefficients[i];
    out.width = c.width; out.height = c.height;
    return out;
}

Camera get_preset_camera(CameraPreset preset, int width, int height)
{
    vaw_camera c{};
    const int rc = vaw_get_preset_camera((int)preset, width, height, &c);
    if (rc != VAW_OK) throw rc;
    return camera_from(c);
}

Camera get_output_camera(const Camera& input, double scale, bool crop_borders, double zoom)
{
    vaw_camera in = camera_to(input), out{};
    const int rc = vaw_get_output_camera(&in, scale, crop_borders ? 1 : 0, zoom, &out);
    if (rc != VAW_OK) throw rc;
    return camera_from(out);
}

// ---- rotation smoothing --------------------------------------------------------------------------
RotationFilter::RotationFilter(int radius) : m_radius(radius < 0 ? 0 : radius)
{
    // least-squares weights of a degree-2 fit over 2m+1 samples, evaluated at the centre
    const int m = m_radius;
    m_weights.resize(2 * m + 1);
    if (m == 0) { m_weights[0] = 1.0; return; }
    const double den = (2.0 * m + 3) * (2.0 * m + 1) * (2.0 * m - 1);
    for (int i = -m; i <= m; ++i) m_weights[i + m] = 3.0 * (3.0 * m * m + 3.0 * m - 1 - 5.0 * i * i) / den;
}

void RotationFilter::add(const Mat33& rotation)
{
    const size_t n = 2 * (size_t)m_radius + 1;
    if (m_window.empty())
        for (size_t i = 0; i < n; ++i) m_window.push_back(rotation);  // start-up: as if the camera had been still
    m_window.push_back(rotation);
    while (m_window.size() > n) m_window.pop_front();
}

Mat33 RotationFilter::filter() const
{
    if (m_window.empty()) return Mat33::eye();
    Mat33 acc{};
    for (size_t k = 0; k < m_window.size(); ++k)
        for (int i = 0; i < 9; ++i) acc.m[i] += m_weights[k] * m_window[k].m[i];
    // nearest rotation = orthogonal polar factor (U V^T of the SVD): Newton iteration X <- (X + X^-T) / 2
    Mat33 x = acc;
    for (int it = 0; it < 30; ++it) {
        const Mat33 y = x.inv().t();
        double diff = 0;
        for (int i = 0; i < 9; ++i) {
            const double v = 0.5 * (x.m[i] + y.m[i]);
            diff += std::fabs(v - x.m[i]);
            x.m[i] = v;
        }
        if (diff < 1e-15) break;
    }
    return x;
}

// ---- FrameSourceWarp --------------------------------------------------------------------------------
FrameSourceWarp::FrameSourceWarp(std::shared_ptr<FrameSource> source, CameraPreset input_camera, double scale,
                                 bool crop_borders, double zoom, int smooth_radius, InterpolationFlags interpolation,
                                 std::shared_ptr<RotationSource> rotation_source, bool create_device_context)
    : m_source(source), m_rotation_source(rotation_source), m_measured_rotation(Mat33::eye()),
      m_last_frame_rotation(Mat33::eye()), m_smooth_radius((unsigned)smooth_radius), m_interpolation(interpolation),
      m_rotation_filter(smooth_radius)
{
    // :214-219 -- the size comes from the first frame (peek), the cameras from the preset
    Frame first_frame = m_source->peek_frame();
    m_device = first_frame->device;
    m_format = first_frame->format;
    m_input_camera = get_preset_camera(input_camera, first_frame->width, first_frame->height);
    m_output_camera = get_output_camera(m_input_camera, scale, crop_borders, zoom);
    if (create_device_context) create_context();
}

FrameSourceWarp::FrameSourceWarp(std::shared_ptr<FrameSource> source, const Camera& input_camera, const Camera& output_camera,
                                 int smooth_radius, InterpolationFlags interpolation,
                                 std::shared_ptr<RotationSource> rotation_source, bool create_device_context)
    : m_source(source), m_rotation_source(rotation_source), m_measured_rotation(Mat33::eye()),
      m_last_frame_rotation(Mat33::eye()), m_smooth_radius((unsigned)smooth_radius), m_interpolation(interpolation),
      m_rotation_filter(smooth_radius)
{
    Frame first_frame = m_source->peek_frame();  // :214-219: device and pixel format come from the first frame
    m_device = first_frame->device;
    m_format = first_frame->format;
    if (first_frame->width != input_camera.width || first_frame->height != input_camera.height) throw (int)VAW_ERR_INVALID;
    m_input_camera = input_camera;
    m_output_camera = output_camera;
    if (create_device_context) create_context();
}

void FrameSourceWarp::create_context()
{
    vaw_camera in = camera_to(m_input_camera), out = camera_to(m_output_camera);
    vaw_params p{};
    int rc = vaw_params_from_cameras(&in, &out, m_format, &p);
    if (rc != VAW_OK) throw rc;
    p.interpolation = (int)m_interpolation;  // NEAREST / LINEAR / CUBIC / LANCZOS4; anything else -> VAW_ERR_UNSUPPORTED
    p.border[0] = 0; p.border[1] = 128; p.border[2] = 128;  // NV12: Y 0 (cv::remap's default, :306-312), neutral chroma
    if (m_format == VAW_FORMAT_BGR24) p.border[1] = p.border[2] = 0;
    rc = vaw_create(&p, m_device, &m_ctx);
    if (rc != VAW_OK) {
        std::fprintf(stderr, "FrameSourceWarp: %s\n", vaw_last_error(nullptr));
        throw rc;  // the reference throws on kernel build failure, :191-195
    }
}

FrameSourceWarp::~FrameSourceWarp()
{
    if (m_ctx) vaw_destroy(m_ctx);
}

int FrameSourceWarp::output_width() const { return m_format == VAW_FORMAT_NV12 ? m_output_camera.width & ~1 : m_output_camera.width; }
int FrameSourceWarp::output_height() const { return m_format == VAW_FORMAT_NV12 ? m_output_camera.height & ~1 : m_output_camera.height; }

Frame FrameSourceWarp::warp_frame(Frame input, const Mat33& rotation)
{
    Frame output = make_device_frame(m_device, m_format, output_width(), output_height());
    output->index = input->index;
    int rc = vaw_warp(m_ctx, input->data, input->pitch, output->data, output->pitch, rotation.m, nullptr);
    if (rc == VAW_OK) rc = vaw_sync(m_device, nullptr);  // the reference's launch is blocking, :301
    if (rc != VAW_OK) throw -1;                           // :301-304
    return output;
}

void FrameSourceWarp::consume_frame(Frame input_frame)
{
    // :397-450 without the corner tracking: the inter-frame rotation comes from the RotationSource
    if (m_frame_index == 0) {
        // the reference only detects corners on the first frame and does not buffer it (:403-406)
    } else {
        Mat33 rotation_since_last_frame = m_last_frame_rotation;
        if (m_rotation_source) {
            Mat33 measured;
            if (m_rotation_source->rotation_since_last_frame(m_frame_index, measured)) rotation_since_last_frame = measured;
        } else {
            rotation_since_last_frame = Mat33::eye();
        }
        m_last_frame_rotation = rotation_since_last_frame;
        const Mat33 accumulated_rotation = rotation_since_last_frame * m_measured_rotation;  // :441
        m_measured_rotation = accumulated_rotation;
        m_rotation_filter.add(accumulated_rotation);
        m_buffered_frames.push(input_frame);
        m_buffered_rotations.push(accumulated_rotation);
    }
    ++m_frame_index;
}

Frame FrameSourceWarp::pull_frame()
{
    while (m_buffered_frames.size() <= m_smooth_radius) {
        try {
            consume_frame(m_source->pull_frame());
        } catch (int err) {
            if (err == EOF) {
                // Pretend the camera kept moving the same way after the last frame (:458-460)
                m_rotation_filter.add(m_measured_rotation);
                break;
            }
            throw;
        }
    }
    if (m_buffered_frames.size() == 0) throw EOF;
    // Stabilise by applying the inverse of the accumulated camera rotation (:468-475)
    Frame frame = m_buffered_frames.front();
    const Mat33 measured_rotation = m_buffered_rotations.front();
    const Mat33 corrected_rotation = m_rotation_filter.filter();
    const Mat33 rotation_correction = corrected_rotation * measured_rotation.inv();
    m_buffered_frames.pop();
    m_buffered_rotations.pop();
    return warp_frame(frame, rotation_correction.inv());
}

Frame FrameSourceWarp::peek_frame()
{
    return pull_frame();  // as the reference (:478-480): it advances
}
