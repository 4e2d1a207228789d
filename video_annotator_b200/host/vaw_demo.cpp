// vaw_demo.cpp -- the reference's only call site of the warp path, rebuilt over the C-ABI.
//
// Mirrors /root/reference/opencv/DisplayImage.cpp:21-75: build the FrameSource chain, then
// pull warped frames until the source throws EOF, with the per-stage timing the reference's
// Profiler decorators print (/root/reference/opencv/Profiler.cpp:14-35).  Decode, VAAPI/OpenCL
// interop and imshow are out of scope (BASELINE.json north_star): frames are synthesised in
// device memory and every output frame is read back and check-summed instead of displayed.
//
//   vaw_demo <width> <height> <frames> <smooth_radius> [sigma_deg]
// prints one line per emitted frame:  frame <index> crc <crc32> rot <9 doubles>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
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

// FrameSourceWarp that also reports the rotation each frame was warped with.
class ReportingWarp : public FrameSourceWarp {
  public:
    using FrameSourceWarp::FrameSourceWarp;
    Mat33 last_rotation = Mat33::eye();
  protected:
    Frame warp_frame(Frame input, const Mat33& rotation) override
    {
        last_rotation = rotation;
        return FrameSourceWarp::warp_frame(input, rotation);
    }
};

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
    if (argc < 5) {
        std::fprintf(stderr, "Usage: %s <width> <height> <frames> <smooth_radius> [sigma_deg]\n", argv[0]);
        return -1;
    }
    const int w = std::atoi(argv[1]), h = std::atoi(argv[2]), n = std::atoi(argv[3]), radius = std::atoi(argv[4]);
    const double sigma = argc > 5 ? std::atof(argv[5]) : 0.4;
    try {
        auto source = std::make_shared<SyntheticFrameSource>(w, h, n, 0);
        auto warped = std::make_shared<ReportingWarp>(source, GOPRO_H4B_WIDE169_MEASURED, 1.0, false, 1.0, radius,
                                                      INTER_LINEAR, std::make_shared<GyroRotationSource>(sigma));
        std::printf("output %d %d\n", warped->output_width(), warped->output_height());
        long emitted = 0;
        const auto t0 = std::chrono::steady_clock::now();
        while (true) {
            try {
                Frame frame = warped->pull_frame();
                std::vector<uint8_t> host(frame->bytes);
                if (vaw_memcpy(frame->device, host.data(), frame->data, frame->bytes, 0, nullptr) != VAW_OK) throw -1;
                std::printf("frame %ld crc %08x rot", frame->index, crc32(host));
                for (double v : warped->last_rotation.m) std::printf(" %.17g", v);
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
