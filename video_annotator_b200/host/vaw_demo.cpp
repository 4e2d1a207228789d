// vaw_demo.cpp -- the reference's only call site of the warp path, rebuilt over the C-ABI.
//
// Mirrors /root/reference/opencv/DisplayImage.cpp:21-75: build the FrameSource chain, then
// pull warped frames until the source throws EOF, with the per-stage timing the reference's
// Profiler decorators print (/root/reference/opencv/Profiler.cpp:14-35).  Decode, VAAPI/OpenCL
// interop and imshow are out of scope (BASELINE.json north_star): frames are synthesised in
// device memory and every output frame is read back and check-summed instead of displayed.
//
//   vaw_demo <width> <height> <frames> <smooth_radius> [sigma_deg] [warp_batch]
// prints one line per emitted frame:  frame <index> crc <crc32> rot <9 doubles>
//   vaw_demo --flow [frames] [width] [height] [sigma_deg]
// runs the chain with the reference's own motion measurement (OpticalFlowRotationSource: corners + Lucas-Kanade on the
// GPU, rotation fit) on a synthetic rotating camera and prints measured against true inter-frame rotations.
//   vaw_demo --bench [frames] [warp_batch] [width] [height]
// times the pull loop of the drop-in at 4K (no read-back; the source hands out frames synthesised once
// into its slab) for the reference's pull pattern (one warp per call) and for batched look-ahead
// warping, and prints one JSON line.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "../../include/vaw.h"
#include "FrameSourceWarp.hpp"

namespace {

// Replaces AvFrameSourceFileVaapi -> AvFrameSourceMapOpenCl -> FrameSourceFfmpegOpenCl
// (DisplayImage.cpp:42-52): n NV12 frames generated on the device, then EOF.
class SyntheticFrameSource : public FrameSource {
    int m_w, m_h, m_n, m_device;
    long m_next = 0;
    Frame m_peeked;
  public:
    SyntheticFrameSource(int w, int h, int n, int device) : m_w(w), m_h(h), m_n(n), m_device(device) {}
    Frame make(long index)
    {
        if (index >= m_n) throw EOF;
        Frame f = make_device_frame(m_device, VAW_FORMAT_NV12, m_w, m_h);
        f->index = index;
        const int rc = vaw_synth_nv12(f->data, m_w, m_h, f->pitch, f->bytes, (int)index, 1, 20260001u, 0, m_device, nullptr);
        if (rc != VAW_OK || vaw_sync(m_device, nullptr) != VAW_OK) throw -1;
        return f;
    }
    Frame pull_frame() override
    {
        if (m_peeked) { Frame f = m_peeked; m_peeked.reset(); ++m_next; return f; }
        return make(m_next++);
    }
    Frame peek_frame() override
    {
        if (!m_peeked) m_peeked = make(m_next);
        return m_peeked;
    }
};

// Small per-frame rotations about all three axes from a fixed LCG.
class GyroStep {
    double m_sigma;
    unsigned m_state = 12345u;
    double next() { m_state = m_state * 1664525u + 1013904223u; return ((m_state >> 8) & 0xffff) / 32768.0 - 1.0; }
  public:
    explicit GyroStep(double sigma_deg) : m_sigma(sigma_deg * 3.14159265358979323846 / 180.0) {}
    Mat33 step()
    {
        const double ax = next() * m_sigma, ay = next() * m_sigma, az = next() * m_sigma;
        const double cx = std::cos(ax), sx = std::sin(ax), cy = std::cos(ay), sy = std::sin(ay), cz = std::cos(az), sz = std::sin(az);
        const Mat33 Rx{{1, 0, 0, 0, cx, -sx, 0, sx, cx}}, Ry{{cy, 0, sy, 0, 1, 0, -sy, 0, cy}}, Rz{{cz, -sz, 0, sz, cz, 0, 0, 0, 1}};
        return Rz * Ry * Rx;
    }
};

// A deterministic stand-in for guess_camera_rotation (FrameSourceWarp.cpp:316-375).
class GyroRotationSource : public RotationSource {
    GyroStep m_gyro;
  public:
    explicit GyroRotationSource(double sigma_deg) : m_gyro(sigma_deg) {}
    bool rotation_since_last_frame(long, Mat33& out) override
    {
        out = m_gyro.step();
        return true;
    }
};

// FrameSourceWarp that also reports the rotation each frame was warped with (by frame index: with batched
// look-ahead warping several frames are warped before the first of them is handed out).
class ReportingWarp : public FrameSourceWarp {
  public:
    using FrameSourceWarp::FrameSourceWarp;
    std::map<long, Mat33> rotation_of;
  protected:
    Frame warp_frame(Frame input, const Mat33& rotation) override
    {
        rotation_of[input->index] = rotation;
        return FrameSourceWarp::warp_frame(input, rotation);
    }
    std::vector<Frame> warp_frames(const std::vector<Frame>& inputs, const std::vector<Mat33>& rotations) override
    {
        for (size_t i = 0; i < inputs.size(); ++i) rotation_of[inputs[i]->index] = rotations[i];
        return FrameSourceWarp::warp_frames(inputs, rotations);
    }
};

// A decoder stand-in for the timing run: `ring` NV12 frames synthesised once into pooled slots, handed
// out round-robin as frames 0 .. n-1 (decode is out of scope; this keeps the source out of the timing).
class RingFrameSource : public FrameSource {
    std::vector<Frame> m_ring;
    long m_n, m_next = 0;
  public:
    RingFrameSource(int w, int h, long n, int ring, int device) : m_n(n)
    {
        for (int i = 0; i < ring; ++i) {
            Frame f = make_device_frame(device, VAW_FORMAT_NV12, w, h);
            if (vaw_synth_nv12(f->data, w, h, f->pitch, f->bytes, i, 1, 20260001u, 0, device, nullptr) != VAW_OK) throw -1;
            m_ring.push_back(f);
        }
        if (vaw_sync(device, nullptr) != VAW_OK) throw -1;
    }
    Frame view(long index)
    {
        if (index >= m_n) throw EOF;
        const Frame& slot = m_ring[(size_t)(index % (long)m_ring.size())];
        // a second handle on the slot's buffer: keeps the slot alive, frees nothing
        auto v = std::make_shared<DeviceFrame>();
        v->width = slot->width; v->height = slot->height; v->pitch = slot->pitch; v->format = slot->format;
        v->device = slot->device; v->bytes = slot->bytes; v->index = index;
        v->data = slot->data;
        v->pool = slot->pool;
        v->alias_of = slot;
        return v;
    }
    Frame pull_frame() override { return view(m_next++); }
    Frame peek_frame() override { return view(m_next); }
    const Frame& slot0() const { return m_ring[0]; }
};

// A hand-held fisheye camera looking at a fixed textured scene: frame k is the scene frame re-projected through the
// library's own fisheye -> fisheye warp for the camera pose P_k (small random rotations accumulate).  Known motion
// for the optical-flow measurement: a pixel of view k with ray r shows the scene ray P_k r, so rays move from
// view k-1 to view k by P_k^-1 P_{k-1}.
class RotatingCameraSource : public FrameSource {
    int m_w, m_h, m_n, m_device;
    long m_next = 0;
    Frame m_scene, m_peeked;
    vaw_ctx* m_ctx = nullptr;
    GyroStep m_gyro;
  public:
    std::vector<Mat33> pose;  // P_k of every frame made so far
    RotatingCameraSource(const Camera& cam, int n, double sigma_deg, int device)
        : m_w(cam.width), m_h(cam.height), m_n(n), m_device(device), m_gyro(sigma_deg)
    {
        vaw_camera in{}, out{};
        in.model = 1; in.width = m_w; in.height = m_h;
        for (int i = 0; i < 9; ++i) in.matrix[i] = cam.matrix.m[i];
        out = in;  // the same fisheye camera on the output side (vaw_params::projection bit 1)
        vaw_params p{};
        if (vaw_params_from_cameras(&in, &out, VAW_FORMAT_NV12, &p) != VAW_OK) throw -1;
        p.border[0] = 0; p.border[1] = 128; p.border[2] = 128;
        if (vaw_create(&p, device, &m_ctx) != VAW_OK) throw -1;
        // the scene: a sum of oriented sinusoids (corners for goodFeaturesToTrack, texture for Lucas-Kanade), grey chroma
        m_scene = make_device_frame(device, VAW_FORMAT_NV12, m_w, m_h);
        std::vector<uint8_t> host(m_scene->bytes, 128);
        unsigned st = 777u;
        double fx[24], fy[24], ph[24];
        for (int k = 0; k < 24; ++k) {
            auto u = [&]() { st = st * 1664525u + 1013904223u; return ((st >> 8) & 0xffff) / 65536.0; };
            fx[k] = (0.03 + 0.32 * u()) * (u() < 0.5 ? -1 : 1); fy[k] = (0.03 + 0.32 * u()) * (u() < 0.5 ? -1 : 1); ph[k] = 6.28 * u();
        }
        for (int y = 0; y < m_h; ++y)
            for (int x = 0; x < m_w; ++x) {
                double v = 0;
                for (int k = 0; k < 24; ++k) v += std::sin(x * fx[k] + y * fy[k] + ph[k]);
                const double g = 127.5 + 18.0 * v;
                host[(size_t)y * m_scene->pitch + x] = (uint8_t)(g < 0 ? 0 : (g > 255 ? 255 : g));
            }
        if (vaw_memcpy(device, m_scene->data, host.data(), m_scene->bytes, 1, nullptr) != VAW_OK) throw -1;
    }
    ~RotatingCameraSource() override { vaw_destroy(m_ctx); }
    Frame make(long index)
    {
        if (index >= m_n) throw EOF;
        while ((long)pose.size() <= index) pose.push_back(pose.empty() ? Mat33::eye() : m_gyro.step() * pose.back());
        Frame f = make_device_frame(m_device, VAW_FORMAT_NV12, m_w, m_h);
        f->index = index;
        if (vaw_warp(m_ctx, m_scene->data, m_scene->pitch, f->data, f->pitch, pose[(size_t)index].m, nullptr) != VAW_OK ||
            vaw_sync(m_device, nullptr) != VAW_OK)
            throw -1;
        return f;
    }
    Frame pull_frame() override
    {
        if (m_peeked) { Frame f = m_peeked; m_peeked.reset(); ++m_next; return f; }
        return make(m_next++);
    }
    Frame peek_frame() override
    {
        if (!m_peeked) m_peeked = make(m_next);
        return m_peeked;
    }
};

// OpticalFlowRotationSource that keeps what it measured, frame by frame
class ReportingFlow : public OpticalFlowRotationSource {
  public:
    using OpticalFlowRotationSource::OpticalFlowRotationSource;
    struct Entry { bool ok; Mat33 R; int inliers, pairs; };
    std::map<long, Entry> measured;
    bool rotation_since_last_frame(long frame_index, Mat33& out) override
    {
        const bool ok = OpticalFlowRotationSource::rotation_since_last_frame(frame_index, out);
        measured[frame_index] = Entry{ok, ok ? out : Mat33::eye(), last_inliers(), last_pairs()};
        return ok;
    }
};

double run_bench(int w, int h, long n, int batch, int radius, long* emitted_out, long* launches_out)
{
    auto source = std::make_shared<RingFrameSource>(w, h, n, 64, 0);
    FrameSourceWarp warped(source, GOPRO_H4B_WIDE169_MEASURED, 1.0, false, 1.0, radius, INTER_LINEAR,
                           std::make_shared<GyroRotationSource>(0.4));
    warped.set_warp_batch(batch);
    long emitted = 0;
    const auto t0 = std::chrono::steady_clock::now();
    while (true) {
        try {
            Frame frame = warped.pull_frame();
            ++emitted;
        } catch (int err) {
            if (err == EOF) break;
            throw;
        }
    }
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    *emitted_out = emitted;
    *launches_out = warped.batch_launches();
    return s;
}

unsigned crc32(const std::vector<uint8_t>& v)
{
    static unsigned table[256];
    static bool init = false;
    if (!init) {
        for (unsigned i = 0; i < 256; ++i) {
            unsigned c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    unsigned c = 0xFFFFFFFFu;
    for (uint8_t b : v) c = table[(c ^ b) & 0xff] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

}  // namespace

int main(int argc, char* argv[])
{
    if (argc >= 2 && std::string(argv[1]) == "--bench") {
        const long n = argc > 2 ? std::atol(argv[2]) : 1500;
        const int batch = argc > 3 ? std::atoi(argv[3]) : 30;  // = smooth_radius: the look-ahead queue holds that many frames with known rotations (8: 60.7 k, 16: 71.4 k, 24: 75.2 k, 30: 77.4 k frames/s at 4K)
        const int w = argc > 4 ? std::atoi(argv[4]) : 3840, h = argc > 5 ? std::atoi(argv[5]) : 2160;
        try {
            long e1 = 0, l1 = 0, eb = 0, lb = 0;
            run_bench(w, h, 200, batch, 30, &eb, &lb);  // warm-up: context, pools, tensor maps
            // wall-clock runs of ~20 ms each: the median of three, so that one host hiccup does not make the number
            auto median3 = [&](int b, long* e, long* l) {
                double s[3];
                for (double& v : s) v = run_bench(w, h, n, b, 30, e, l);
                std::sort(s, s + 3);
                return s[1];
            };
            const double s1 = median3(1, &e1, &l1);
            const double sb = median3(batch, &eb, &lb);
            std::printf("{\"shim\": \"FrameSourceWarp::pull_frame over a pooled ring source\", \"src\": [%d, %d], \"frames\": %ld, "
                        "\"smooth_radius\": 30, \"fps_one_warp_per_call\": %.1f, \"warp_batch\": %d, \"fps_batched\": %.1f, "
                        "\"batched_launches\": %ld, \"shim_fps\": %.1f}\n",
                        w, h, e1, e1 / s1, batch, eb / sb, lb, eb / sb);
        } catch (int err) {
            std::fprintf(stderr, "error %d: %s\n", err, vaw_last_error(nullptr));
            return 1;
        }
        frame_pool_trim();
        return 0;
    }
    if (argc >= 2 && std::string(argv[1]) == "--flow") {
        // vaw_demo --flow [frames] [width] [height] [sigma_deg]: the reference's whole chain with its own measurement --
        // a rotating camera, OpticalFlowRotationSource (corners + LK on the GPU, rotation fit), smoothing, warp --
        // and the measured inter-frame rotations next to the true ones.  One line per frame, then a JSON summary.
        const int n = argc > 2 ? std::atoi(argv[2]) : 40;
        const int w = argc > 3 ? std::atoi(argv[3]) : 1920, h = argc > 4 ? std::atoi(argv[4]) : 1080;
        const double sigma = argc > 5 ? std::atof(argv[5]) : 0.5;
        try {
            const Camera cam = get_preset_camera(GOPRO_H4B_WIDE169_MEASURED, w, h);
            const Camera out = get_output_camera(cam, 1.0, false, 1.0);
            auto source = std::make_shared<RotatingCameraSource>(cam, n, sigma, 0);
            auto flow = std::make_shared<ReportingFlow>(cam, out, 0);
            FrameSourceWarp warped(source, cam, out, 5, INTER_LINEAR, flow);
            long emitted = 0;
            const auto t0 = std::chrono::steady_clock::now();
            while (true) {
                try { Frame f = warped.pull_frame(); ++emitted; }
                catch (int err) { if (err == EOF) break; throw; }
            }
            if (vaw_sync(0, nullptr) != VAW_OK) throw -1;
            const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            double worst = 0, sum = 0;
            int fits = 0, min_inliers = 1 << 30;
            for (const auto& kv : flow->measured) {
                const long k = kv.first;
                const Mat33 want = source->pose[(size_t)k].t() * source->pose[(size_t)k - 1];
                const Mat33 d = kv.second.R * want.t();
                double c = (d.m[0] + d.m[4] + d.m[8] - 1) / 2;
                c = c > 1 ? 1 : (c < -1 ? -1 : c);
                const double err = std::acos(c) * 180.0 / 3.14159265358979323846;
                std::printf("frame %ld fit %d pairs %d inliers %d error_deg %.5f\n", k, (int)kv.second.ok, kv.second.pairs, kv.second.inliers, err);
                if (kv.second.ok) { ++fits; worst = err > worst ? err : worst; sum += err; min_inliers = kv.second.inliers < min_inliers ? kv.second.inliers : min_inliers; }
            }
            std::printf("{\"flow_demo\": \"RotatingCameraSource -> FrameSourceWarp + OpticalFlowRotationSource\", \"src\": [%d, %d], \"frames\": %d, "
                        "\"emitted\": %ld, \"fits\": %d, \"worst_error_deg\": %.5f, \"mean_error_deg\": %.5f, \"min_inliers\": %d, "
                        "\"fps_incl_synthesis\": %.1f}\n",
                        w, h, n, emitted, fits, worst, fits ? sum / fits : 0.0, fits ? min_inliers : 0, emitted / secs);
        } catch (int err) {
            std::fprintf(stderr, "error %d: %s | %s\n", err, vaw_last_error(nullptr), vaw_flow_last_error(nullptr));
            return 1;
        }
        frame_pool_trim();
        return 0;
    }
    if (argc < 5) {
        std::fprintf(stderr, "Usage: %s <width> <height> <frames> <smooth_radius> [sigma_deg]\n", argv[0]);
        return -1;
    }
    const int w = std::atoi(argv[1]), h = std::atoi(argv[2]), n = std::atoi(argv[3]), radius = std::atoi(argv[4]);
    const double sigma = argc > 5 ? std::atof(argv[5]) : 0.4;
    const int warp_batch = argc > 6 ? std::atoi(argv[6]) : 1;
    try {
        auto source = std::make_shared<SyntheticFrameSource>(w, h, n, 0);
        auto warped = std::make_shared<ReportingWarp>(source, GOPRO_H4B_WIDE169_MEASURED, 1.0, false, 1.0, radius,
                                                      INTER_LINEAR, std::make_shared<GyroRotationSource>(sigma));
        warped->set_warp_batch(warp_batch);
        std::printf("output %d %d\n", warped->output_width(), warped->output_height());
        long emitted = 0;
        const auto t0 = std::chrono::steady_clock::now();
        while (true) {
            try {
                Frame frame = warped->pull_frame();
                std::vector<uint8_t> host(frame->bytes);
                if (vaw_memcpy(frame->device, host.data(), frame->data, frame->bytes, 0, nullptr) != VAW_OK) throw -1;
                std::printf("frame %ld crc %08x rot", frame->index, crc32(host));
                for (double v : warped->rotation_of[frame->index].m) std::printf(" %.17g", v);
                warped->rotation_of.erase(frame->index);
                std::printf("\n");
                ++emitted;
            } catch (int err) {
                if (err == EOF) break;  // DisplayImage.cpp:66-70
                throw;
            }
        }
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::fprintf(stderr, "warp (incl. synthesis and readback): %ld frames, %.3f ms/frame, %.1f fps\n", emitted,
                     emitted ? s * 1e3 / emitted : 0.0, emitted ? emitted / s : 0.0);
    } catch (int err) {
        std::fprintf(stderr, "error %d: %s\n", err, vaw_last_error(nullptr));
        return 1;
    }
    return 0;
}
