// FrameSourceWarp.cpp -- host-side state machine of the warp stage over the C-ABI.
// Follows /root/reference/opencv/FrameSourceWarp.cpp:199-226 (constructor), :397-450
// (consume_frame, minus the optical-flow measurement), :452-480 (pull_frame / peek_frame).
#include "FrameSourceWarp.hpp"

#include <cmath>
#include <cstdio>
#include <map>
#include <mutex>
#include <utility>

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
FramePool::FramePool(int device, size_t frame_bytes, int slots)
    : m_device(device), m_stride((frame_bytes + 255) & ~(size_t)255), m_slots(slots < 1 ? 1 : slots), m_busy(new bool[slots < 1 ? 1 : slots]())
{
    void* p = nullptr;
    const int rc = vaw_malloc(device, m_stride * (size_t)m_slots, &p);
    if (rc != VAW_OK) throw rc;
    m_base = static_cast<uint8_t*>(p);
}

FramePool::FramePool(int device, uint8_t* base, size_t stride, int slots)
    : m_device(device), m_stride(stride), m_slots(slots < 1 ? 1 : slots), m_base(base), m_owns(false), m_busy(new bool[slots < 1 ? 1 : slots]())
{
}

FramePool::~FramePool()
{
    if (m_base && m_owns) vaw_free(m_device, m_base);
}

uint8_t* FramePool::acquire()
{
    if (m_used == m_slots) return nullptr;
    // ring order: frames acquired back to back are neighbours in the slab (until the cursor wraps)
    for (int k = 0; k < m_slots; ++k) {
        const int i = (m_cursor + k) % m_slots;
        if (!m_busy[i]) {
            m_busy[i] = true;
            ++m_used;
            m_cursor = (i + 1) % m_slots;
            return m_base + (size_t)i * m_stride;
        }
    }
    return nullptr;
}

void FramePool::release(uint8_t* slot)
{
    const size_t i = (size_t)(slot - m_base) / m_stride;
    if (slot < m_base || i >= (size_t)m_slots || !m_busy[i]) return;
    m_busy[i] = false;
    --m_used;
}

namespace {
std::mutex g_pool_mutex;
std::map<std::pair<int, size_t>, std::shared_ptr<FramePool>> g_pools;
}  // namespace

std::shared_ptr<FramePool> frame_pool(int device, size_t bytes)
{
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    std::shared_ptr<FramePool>& p = g_pools[std::make_pair(device, bytes)];
    if (!p) p = std::make_shared<FramePool>(device, bytes, kDefaultPoolSlots);
    return p;
}

void frame_pool_trim()
{
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    g_pools.clear();
}

DeviceFrame::~DeviceFrame()
{
    if (!data || alias_of) return;
    if (pool) {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        pool->release(data);
    } else {
        vaw_free(device, data);
    }
}

Frame make_device_frame(int device, int format, int width, int height)
{
    auto f = std::make_shared<DeviceFrame>();
    const int channels = format == VAW_FORMAT_BGR24 ? 3 : 1;
    f->width = width; f->height = height; f->pitch = width * channels; f->format = format; f->device = device;
    f->bytes = vaw_frame_bytes(format, width, height, f->pitch);
    std::shared_ptr<FramePool> pool = frame_pool(device, f->bytes);
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        f->data = pool->acquire();
    }
    if (f->data) {
        f->pool = pool;
        return f;
    }
    void* p = nullptr;  // more frames in flight than the pool holds: a plain allocation
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
RotationFilter::RotationFilter(int radius, StartUp start) : m_radius(radius < 0 ? 0 : radius), m_start(start)
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
    if (m_window.empty())  // start-up: the library's zero matrices, or (round 1's convention) as if the camera had been still
        for (size_t i = 0; i < n; ++i) m_window.push_back(m_start == StartUp::Zeros ? Mat33{} : rotation);
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
    if (m_rot_dev) vaw_free(m_device, m_rot_dev);
    if (m_ctx) vaw_destroy(m_ctx);
}

int FrameSourceWarp::output_width() const { return m_format == VAW_FORMAT_NV12 ? m_output_camera.width & ~1 : m_output_camera.width; }
int FrameSourceWarp::output_height() const { return m_format == VAW_FORMAT_NV12 ? m_output_camera.height & ~1 : m_output_camera.height; }

void FrameSourceWarp::set_warp_batch(int frames) { m_warp_batch = frames < 1 ? 1 : (frames > kMaxWarpBatch ? kMaxWarpBatch : frames); }

Frame FrameSourceWarp::warp_frame(Frame input, const Mat33& rotation)
{
    Frame output = make_device_frame(m_device, m_format, output_width(), output_height());
    output->index = input->index;
    int rc = vaw_warp(m_ctx, input->data, input->pitch, output->data, output->pitch, rotation.m, nullptr);
    if (rc == VAW_OK) rc = vaw_sync(m_device, nullptr);  // the reference's launch is blocking, :301
    if (rc != VAW_OK) throw -1;                           // :301-304
    return output;
}

// Several frames whose rotations are already known: runs of frames that sit at a constant stride (the
// usual case with pooled frames) go through ONE vaw_warp_batch each, with a single synchronisation at the
// end -- a 4K frame is ~12 us of GPU work, less than a launch + sync round trip.
std::vector<Frame> FrameSourceWarp::warp_frames(const std::vector<Frame>& inputs, const std::vector<Mat33>& rotations)
{
    std::vector<Frame> outputs;
    outputs.reserve(inputs.size());
    if (!m_ctx || inputs.size() == 1) {  // no device context (CPU tests) or nothing to batch
        for (size_t i = 0; i < inputs.size(); ++i) outputs.push_back(warp_frame(inputs[i], rotations[i]));
        return outputs;
    }
    const size_t n = inputs.size();
    for (size_t i = 0; i < n; ++i) {
        outputs.push_back(make_device_frame(m_device, m_format, output_width(), output_height()));
        outputs.back()->index = inputs[i]->index;
    }
    if (!m_rot_dev) {
        void* p = nullptr;
        if (vaw_malloc(m_device, sizeof(float) * 9 * kMaxWarpBatch, &p) != VAW_OK) throw -1;
        m_rot_dev = static_cast<float*>(p);
    }
    std::vector<double> rots(9 * n);
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < 9; ++k) rots[9 * i + k] = rotations[i].m[k];
    if (vaw_upload_rotations(m_ctx, rots.data(), (int)n, m_rot_dev, nullptr) != VAW_OK) throw -1;
    // the source slab's tensor maps are encoded once (a hint; any other layout still works)
    const std::shared_ptr<FramePool>& pool = inputs[0]->pool;
    if (pool && pool->base() != m_bound_base) {
        if (vaw_bind_clip(m_ctx, pool->base(), inputs[0]->pitch, pool->stride(), pool->slots()) == VAW_OK) m_bound_base = pool->base();
    }
    size_t first = 0;
    while (first < n) {
        // longest run [first, last) with constant source and output strides
        size_t last = first + 1;
        ptrdiff_t ss = 0, ds = 0;
        if (last < n) {
            ss = inputs[last]->data - inputs[first]->data;
            ds = outputs[last]->data - outputs[first]->data;
            const bool ok = ss > 0 && ds > 0 && (size_t)ss >= inputs[first]->bytes && (size_t)ds >= outputs[first]->bytes &&
                            inputs[last]->pitch == inputs[first]->pitch;
            if (ok) {
                ++last;
                while (last < n && inputs[last]->data - inputs[last - 1]->data == ss &&
                       outputs[last]->data - outputs[last - 1]->data == ds && inputs[last]->pitch == inputs[first]->pitch)
                    ++last;
            }
        }
        const int count = (int)(last - first);
        const int rc = vaw_warp_batch(m_ctx, inputs[first]->data, inputs[first]->pitch, count > 1 ? (size_t)ss : inputs[first]->bytes,
                                      outputs[first]->data, outputs[first]->pitch, count > 1 ? (size_t)ds : outputs[first]->bytes,
                                      m_rot_dev + 9 * first, count, nullptr);
        if (rc != VAW_OK) throw -1;
        ++m_batch_launches;
        first = last;
    }
    if (vaw_sync(m_device, nullptr) != VAW_OK) throw -1;
    return outputs;
}

void FrameSourceWarp::consume_frame(Frame input_frame)
{
    // :397-450; the corner tracking and the rotation fit live behind the RotationSource
    if (m_rotation_source) m_rotation_source->observe_frame(input_frame, m_frame_index);
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

// One turn of the reference's pull_frame (:452-476) up to, not including, warp_frame: top up the
// look-ahead queue, then take the oldest frame with the rotation that undoes its shake.
bool FrameSourceWarp::next_frame_and_rotation(Frame& frame, Mat33& rotation)
{
    while (m_buffered_frames.size() <= m_smooth_radius) {
        try {
            consume_frame(m_source->pull_frame());
        } catch (int err) {
            if (err == EOF) {
                // upstream has ended: the filter window is topped up with the final camera pose (:458-460)
                m_rotation_filter.add(m_measured_rotation);
                break;
            }
            throw;
        }
    }
    if (m_buffered_frames.size() == 0) return false;
    // rotation that takes the measured camera pose onto the smoothed path, inverted for the map (:468-475)
    frame = m_buffered_frames.front();
    const Mat33 measured_rotation = m_buffered_rotations.front();
    const Mat33 corrected_rotation = m_rotation_filter.filter();
    const Mat33 rotation_correction = corrected_rotation * measured_rotation.inv();
    m_buffered_frames.pop();
    m_buffered_rotations.pop();
    rotation = rotation_correction.inv();
    return true;
}

Frame FrameSourceWarp::pull_frame()
{
    if (m_warp_batch <= 1) {  // the reference's pull pattern exactly: one upstream pull, one warp per call
        Frame frame;
        Mat33 rotation;
        if (!next_frame_and_rotation(frame, rotation)) throw EOF;
        return warp_frame(frame, rotation);
    }
    // Batched: the same turns, run up to m_warp_batch times ahead -- every frame gets exactly the rotation
    // the one-at-a-time loop gives it (same filter state, same order) -- then ONE batched warp; the
    // results are handed out over the following calls.  Only the latency (frames pulled from upstream
    // before the first one comes out) grows, by m_warp_batch - 1.
    if (m_ready.empty()) {
        if (m_deferred_error_set) { m_deferred_error_set = false; throw m_deferred_error; }
        std::vector<Frame> frames;
        std::vector<Mat33> rotations;
        try {
            while ((int)frames.size() < m_warp_batch) {
                Frame frame;
                Mat33 rotation;
                if (!next_frame_and_rotation(frame, rotation)) break;
                frames.push_back(frame);
                rotations.push_back(rotation);
            }
        } catch (int err) {
            if (frames.empty()) throw;
            m_deferred_error = err;  // surfaces after the frames that precede it, as it would one at a time
            m_deferred_error_set = true;
        }
        if (frames.empty()) throw EOF;
        for (Frame& f : warp_frames(frames, rotations)) m_ready.push_back(f);
    }
    Frame out = m_ready.front();
    m_ready.pop_front();
    return out;
}

Frame FrameSourceWarp::peek_frame()
{
    return pull_frame();  // as the reference (:478-480): it advances
}

// ---- the optical-flow measurement (:228-270, :316-368, :403-438) -----------------------------------------
namespace {
vaw_camera to_c_camera(const Camera& c)
{
    vaw_camera o{};
    o.model = c.model == FISHEYE ? 1 : 0;
    o.width = c.width; o.height = c.height;
    for (int i = 0; i < 9; ++i) o.matrix[i] = c.matrix.m[i];
    for (int i = 0; i < 4; ++i) o.distortion[i] = c.distortion_coefficients[i];
    return o;
}
}  // namespace

OpticalFlowRotationSource::OpticalFlowRotationSource(const Camera& input_camera, const Camera& output_camera, int device)
    : m_input_camera(input_camera), m_output_camera(output_camera), m_device(device), m_rotation(Mat33::eye())
{
    const int rc = vaw_flow_create(input_camera.width, input_camera.height, device, &m_flow);
    if (rc != VAW_OK) throw rc;
}

OpticalFlowRotationSource::~OpticalFlowRotationSource() { vaw_flow_destroy(m_flow); }

void OpticalFlowRotationSource::observe_frame(const Frame& frame, long frame_index)
{
    // frame_gray = the luma plane of the NV12 frame (:398)
    int rc = vaw_flow_push_frame(m_flow, frame->data, frame->pitch, nullptr);
    if (rc != VAW_OK) throw rc;
    const int kMaxCorners = 200;
    std::vector<float> found(2 * kMaxCorners);
    int n = 0;
    m_have_rotation = false;
    if (m_last_key_frame_index == -1) {
        // the first frame: corners of this frame, nothing to track yet (:403-406)
        rc = vaw_flow_corners(m_flow, 1, kMaxCorners, 0.01, 30.0, found.data(), kMaxCorners, &n, nullptr);
        if (rc != VAW_OK) throw rc;
        found.resize(2 * (size_t)n);
        m_last_input_frame_corners.swap(found);
        m_last_key_frame_index = frame_index;
        return;
    }
    if (frame_index - m_last_key_frame_index > 20 || m_last_input_frame_corners.size() / 2 < 150) {
        // a fresh set, detected in the PREVIOUS frame (:415-419)
        rc = vaw_flow_corners(m_flow, 0, kMaxCorners, 0.01, 30.0, found.data(), kMaxCorners, &n, nullptr);
        if (rc != VAW_OK) throw rc;
        found.resize(2 * (size_t)n);
        m_last_input_frame_corners.swap(found);
        m_last_key_frame_index = frame_index - 1;
    }
    const int np = (int)(m_last_input_frame_corners.size() / 2);
    std::vector<float> next(2 * (size_t)np);
    std::vector<uint8_t> status((size_t)np);
    rc = vaw_flow_track(m_flow, m_last_input_frame_corners.data(), np, next.data(), status.data(), nullptr);
    if (rc != VAW_OK) throw rc;
    // the pairs for which the flow was found (:262-268)
    std::vector<float> prev_pts, cur_pts;
    for (int i = 0; i < np; ++i)
        if (status[(size_t)i]) {
            prev_pts.push_back(m_last_input_frame_corners[2 * (size_t)i]); prev_pts.push_back(m_last_input_frame_corners[2 * (size_t)i + 1]);
            cur_pts.push_back(next[2 * (size_t)i]); cur_pts.push_back(next[2 * (size_t)i + 1]);
        }
    m_last_input_frame_corners = cur_pts;  // :427
    const vaw_camera in = to_c_camera(m_input_camera), out = to_c_camera(m_output_camera);
    double R[9];
    int inliers = 0;
    m_last_pairs = (int)(cur_pts.size() / 2);
    rc = vaw_guess_rotation(&in, &out, prev_pts.data(), cur_pts.data(), m_last_pairs, (uint32_t)frame_index, R, &inliers);
    if (rc != VAW_OK) throw rc;
    m_last_inliers = inliers;
    if (inliers >= 40) {  // :431
        for (int i = 0; i < 9; ++i) m_rotation.m[i] = R[i];
        m_have_rotation = true;
    }
}

bool OpticalFlowRotationSource::rotation_since_last_frame(long /*frame_index*/, Mat33& out)
{
    if (!m_have_rotation) return false;  // the caller keeps the previous inter-frame rotation (or identity), :432-437
    out = m_rotation;
    return true;
}
