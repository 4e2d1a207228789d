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
//   vaw_demo --bench [frames] [warp_batch] [width] [height]
// times the pull loop of the drop-in at 4K (no read-back; the source hands out frames synthesised once
// into its slab) for the reference's pull pattern (one warp per call) and for batched look-ahead
// warping, and prints one JSON line.
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

// A deterministic stand-in for guess_camera_rotation (FrameSourceWarp.cpp:316-375): small
// per-frame rotations about all three axes from a fixed LCG.
class GyroRotationSource : public RotationSource {
    double m_sigma;
    unsigned m_state = 12345u;
    double next() { m_state = m_state * 1664525u + 1013904223u; return ((m_state >> 8) & 0xffff) / 32768.0 - 1.0; }
  public:
    explicit GyroRotationSource(double sigma_deg) : m_sigma(sigma_deg * 3.14159265358979323846 / 180.0) {}
    bool rotation_since_last_frame(long, Mat33& out) override
    {
        const double ax = next() * m_sigma, ay = next() * m_sigma, az = next() * m_sigma;
        const double cx = std::cos(ax), sx = std::sin(ax), cy = std::cos(ay), sy = std::sin(ay), cz = std::cos(az), sz = std::sin(az);
        const Mat33 Rx{{1, 0, 0, 0, cx, -sx, 0, sx, cx}}, Ry{{cy, 0, sy, 0, 1, 0, -sy, 0, cy}}, Rz{{cz, -sz, 0, sz, cz, 0, 0, 0, 1}};
        out = Rz * Ry * Rx;
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
        const int batch = argc > 3 ? std::atoi(argv[3]) : 16;
        const int w = argc > 4 ? std::atoi(argv[4]) : 3840, h = argc > 5 ? std::atoi(argv[5]) : 2160;
        try {
            long e1 = 0, l1 = 0, eb = 0, lb = 0;
            run_bench(w, h, 200, batch, 30, &eb, &lb);  // warm-up: context, pools, tensor maps
            const double s1 = run_bench(w, h, n, 1, 30, &e1, &l1);
            const double sb = run_bench(w, h, n, batch, 30, &eb, &lb);
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
