// CPU test of the FrameSourceWarp state machine (video_annotator_b200/host), no GPU needed:
// warp_frame is overridden with a pass-through that records the rotation it was given.
// Behaviours checked are the reference's (opencv/FrameSourceWarp.cpp:397-480).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/vaw.h"
#include "../../video_annotator_b200/host/FrameSourceWarp.hpp"

static int g_checks = 0, g_failed = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        ++g_checks;                                                        \
        if (!(cond)) { ++g_failed; std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)

struct FakeSource : FrameSource {
    int n, pulls = 0;
    explicit FakeSource(int n_) : n(n_) {}
    Frame make(long i)
    {
        if (i >= n) throw EOF;
        auto f = std::make_shared<DeviceFrame>();
        f->width = 1920; f->height = 1080; f->pitch = 1920; f->format = VAW_FORMAT_NV12; f->index = i;
        return f;
    }
    Frame pull_frame() override { return make(pulls++); }
    Frame peek_frame() override { return make(pulls); }
};

struct SpinSource : RotationSource {  // constant angular rate about z
    double step;
    explicit SpinSource(double deg) : step(deg * 3.14159265358979323846 / 180.0) {}
    bool rotation_since_last_frame(long, Mat33& out) override
    {
        out = Mat33{{std::cos(step), -std::sin(step), 0, std::sin(step), std::cos(step), 0, 0, 0, 1}};
        return true;
    }
};

struct RecordingWarp : FrameSourceWarp {
    using FrameSourceWarp::FrameSourceWarp;
    std::vector<Mat33> rotations;
    std::vector<long> indices;
  protected:
    Frame warp_frame(Frame input, const Mat33& rotation) override
    {
        rotations.push_back(rotation);
        indices.push_back(input->index);
        return input;
    }
};

static double dist_to_identity(const Mat33& r)
{
    const Mat33 e = Mat33::eye();
    double d = 0;
    for (int i = 0; i < 9; ++i) d = std::fmax(d, std::fabs(r.m[i] - e.m[i]));
    return d;
}

int main()
{
    {  // algebra
        const Mat33 a{{2, 1, 0, 0, 3, 1, 1, 0, 4}};
        CHECK(dist_to_identity(a * a.inv()) < 1e-14);
        const Mat33 r{{0, -1, 0, 1, 0, 0, 0, 0, 1}};
        RotationFilter f(30);                            // the library's start-up: a window of zero matrices
        for (int i = 0; i < 31; ++i) f.add(r);           // the reference's first filter() comes after radius + 1 samples (:453)
        const Mat33 s = f.filter();
        CHECK(dist_to_identity(s * r.t()) < 1e-12);      // a constant sequence is its own smoothing (one-sided window included)
        CHECK(dist_to_identity(s * s.t()) < 1e-12);      // result is orthonormal
        RotationFilter g(30, RotationFilter::StartUp::FirstSample);  // round 1's convention
        g.add(r);
        CHECK(dist_to_identity(g.filter() * r.t()) < 1e-12);
    }
    {  // no rotation: frame 0 dropped, order kept, look-ahead latency, EOF drain
        const int n = 10, radius = 3;
        auto src = std::make_shared<FakeSource>(n);
        RecordingWarp w(src, GOPRO_H4B_WIDE169_MEASURED, 1, false, 1, radius, INTER_LINEAR, nullptr, false);
        CHECK(w.output_camera().width == 1759 && w.output_camera().height == 998);
        CHECK(w.output_width() == 1758 && w.output_height() == 998);                 // NV12: even
        Frame first = w.pull_frame();
        CHECK(first->index == 1);                       // frame 0 is never emitted (:403-406)
        CHECK(src->pulls == radius + 2);                // frame 0 + radius+1 buffered frames (:453)
        int emitted = 1;
        bool eof = false;
        for (int i = 0; i < 50 && !eof; ++i) {
            try { Frame f = w.pull_frame(); CHECK(f->index == 1 + emitted); ++emitted; }
            catch (int err) { CHECK(err == EOF); eof = true; }
        }
        CHECK(eof && emitted == n - 1);
        for (const Mat33& r : w.rotations) CHECK(dist_to_identity(r) < 1e-12);
        bool again = false;
        try { w.pull_frame(); } catch (int err) { again = err == EOF; }
        CHECK(again);                                   // stays at EOF
    }
    {  // steady spin: the smoothed path equals the measured one mid-stream -> correction ~ identity
        const int n = 40, radius = 5;
        auto src = std::make_shared<FakeSource>(n);
        RecordingWarp w(src, GOPRO_H4B_WIDE169_MEASURED, 1, false, 1, radius, INTER_LINEAR,
                        std::make_shared<SpinSource>(1.0), false);
        try { while (true) w.pull_frame(); } catch (int err) { CHECK(err == EOF); }
        CHECK((int)w.rotations.size() == n - 1);
        for (size_t i = 2 * radius; i + radius + 1 < w.rotations.size(); ++i) CHECK(dist_to_identity(w.rotations[i]) < 2e-4);
        // at the start the window is one-sided (zero matrices on the past side): a real correction
        CHECK(dist_to_identity(w.rotations[0]) > 1e-3);
        for (const Mat33& r : w.rotations) CHECK(std::fabs((r * r.t()).m[0] - 1.0) < 1e-9);
    }
    {  // peek_frame advances like pull_frame (:478-480)
        auto src = std::make_shared<FakeSource>(6);
        RecordingWarp w(src, GOPRO_H4B_WIDE169_MEASURED, 1, false, 1, 1, INTER_LINEAR, nullptr, false);
        CHECK(w.peek_frame()->index == 1);
        CHECK(w.pull_frame()->index == 2);
    }
    {  // camera producers agree with the C-ABI
        const Camera c = get_preset_camera(GOPRO_H4B_WIDE43_MEASURED, 1920, 1440);
        const Camera o = get_output_camera(c, 0.5, false, 1);          // the DisplayImage.cpp:55 setting
        CHECK(o.width == 988 && o.height == 744);
        CHECK(std::fabs(o.matrix.m[0] - 187.299) < 1e-3);
        bool threw = false;
        try { get_preset_camera((CameraPreset)42, 10, 10); } catch (int err) { threw = err == VAW_ERR_INVALID; }
        CHECK(threw);
    }
    {  // the overload with explicit cameras: same state machine, cameras taken as given
        auto src = std::make_shared<FakeSource>(6);
        Camera in = get_preset_camera(GOPRO_H4B_WIDE169_MEASURED, 1920, 1080);
        in.distortion_coefficients[0] = 0.02;
        Camera out = get_output_camera(in, 1, false, 1);
        out.width = 1280; out.height = 720;
        out.matrix.m[0] = out.matrix.m[4] = 400.0; out.matrix.m[2] = 639.5; out.matrix.m[5] = 359.5;
        RecordingWarp w(src, in, out, 1, INTER_CUBIC, nullptr, false);
        CHECK(w.output_width() == 1280 && w.output_height() == 720);
        CHECK(w.input_camera().distortion_coefficients[0] == 0.02);
        CHECK(w.pull_frame()->index == 1);
        Camera wrong = in;
        wrong.width = 1280;
        bool threw = false;
        try { RecordingWarp bad(std::make_shared<FakeSource>(6), wrong, out, 1, INTER_LINEAR, nullptr, false); }
        catch (int err) { threw = err == VAW_ERR_INVALID; }
        CHECK(threw);                                   // the input camera must describe the frames of the source
    }
    {  // batched look-ahead warping: same frames, same order, same rotations as one warp per call
        const int n = 41, radius = 4;
        std::vector<Mat33> ref_rot;
        std::vector<long> ref_idx;
        for (int batch : {1, 5, 64}) {
            auto src = std::make_shared<FakeSource>(n);
            RecordingWarp w(src, GOPRO_H4B_WIDE169_MEASURED, 1, false, 1, radius, INTER_LINEAR,
                            std::make_shared<SpinSource>(0.7), false);
            w.set_warp_batch(batch);
            std::vector<long> out_idx;
            if (batch == 5) {  // latency: batch - 1 more upstream pulls before the first frame comes out
                out_idx.push_back(w.pull_frame()->index);
                CHECK(src->pulls == radius + 2 + (batch - 1));
            }
            try { while (true) out_idx.push_back(w.pull_frame()->index); } catch (int err) { CHECK(err == EOF); }
            CHECK((int)out_idx.size() == n - 1);
            for (size_t i = 0; i < out_idx.size(); ++i) CHECK(out_idx[i] == (long)i + 1);
            if (batch == 1) { ref_rot = w.rotations; ref_idx = w.indices; continue; }
            CHECK(w.indices == ref_idx);
            CHECK(w.rotations.size() == ref_rot.size());
            for (size_t i = 0; i < ref_rot.size() && i < w.rotations.size(); ++i)
                for (int k = 0; k < 9; ++k) CHECK(w.rotations[i].m[k] == ref_rot[i].m[k]);  // bit for bit
            bool again = false;
            try { w.pull_frame(); } catch (int err) { again = err == EOF; }
            CHECK(again);
        }
    }
    {  // an upstream failure surfaces after the frames that precede it, also when batching
        struct FailingSource : FakeSource {
            using FakeSource::FakeSource;
            Frame pull_frame() override { if (pulls == 9) { ++pulls; throw -7; } return FakeSource::pull_frame(); }
        };
        for (int batch : {1, 4}) {
            auto src = std::make_shared<FailingSource>(30);
            RecordingWarp w(src, GOPRO_H4B_WIDE169_MEASURED, 1, false, 1, 2, INTER_LINEAR, nullptr, false);
            w.set_warp_batch(batch);
            int got = 0, err_code = 0;
            try { while (true) { w.pull_frame(); ++got; } } catch (int err) { err_code = err; }
            CHECK(err_code == -7);
            CHECK(got == 6);  // frames 1..8 buffered; emitting frame k needs frame k + radius: 1..6 come out, then the error
        }
    }
    {  // frame pool bookkeeping (over caller-owned memory: no device needed)
        static uint8_t slab[8 * 256];
        FramePool pool(0, slab, 256, 8);
        uint8_t* a = pool.acquire();
        uint8_t* b = pool.acquire();
        uint8_t* c = pool.acquire();
        CHECK(a == slab && b == slab + 256 && c == slab + 512 && pool.used() == 3);  // ring order: neighbours
        pool.release(b);
        CHECK(pool.acquire() == slab + 768);       // the cursor moves on, the hole is reused only after a lap
        for (int i = 0; i < 4; ++i) CHECK(pool.acquire() != nullptr);
        CHECK(pool.used() == 7 && pool.acquire() == slab + 256 && pool.acquire() == nullptr);
        pool.release(slab + 5 * 256 + 1);          // not a slot start of a busy slot: slot 5 is released (address inside it)
        pool.release(nullptr);                      // ignored
        CHECK(pool.used() == 7);
    }
    if (std::getenv("VAW_ROT_INCREMENTS")) {  // rotation chain against video_annotator_b200/rotations.py (tests/test_host_shim.py)
        // file: n lines of axis-angle increments; prints the warp rotation of every emitted frame
        struct Replay : RotationSource {
            std::vector<Mat33> inc;
            bool rotation_since_last_frame(long frame_index, Mat33& out) override
            {
                if (frame_index < 1 || (size_t)frame_index > inc.size()) return false;
                out = inc[(size_t)frame_index - 1];
                return true;
            }
        };
        auto replay = std::make_shared<Replay>();
        FILE* f = std::fopen(std::getenv("VAW_ROT_INCREMENTS"), "r");
        double vx, vy, vz;
        while (f && std::fscanf(f, "%lf %lf %lf", &vx, &vy, &vz) == 3) {
            const double th = std::sqrt(vx * vx + vy * vy + vz * vz);
            Mat33 r = Mat33::eye();
            if (th >= 1e-15) {
                const double kx = vx / th, ky = vy / th, kz = vz / th, s = std::sin(th), c1 = 1 - std::cos(th);
                const Mat33 K{{0, -kz, ky, kz, 0, -kx, -ky, kx, 0}};
                const Mat33 K2 = K * K;
                for (int i = 0; i < 9; ++i) r.m[i] += s * K.m[i] + c1 * K2.m[i];
            }
            replay->inc.push_back(r);
        }
        if (f) std::fclose(f);
        const int n = (int)replay->inc.size() + 1, radius = std::atoi(std::getenv("VAW_ROT_RADIUS") ? std::getenv("VAW_ROT_RADIUS") : "30");
        auto src = std::make_shared<FakeSource>(n);
        RecordingWarp w(src, GOPRO_H4B_WIDE169_MEASURED, 1, false, 1, radius, INTER_LINEAR, replay, false);
        w.set_warp_batch(7);
        try { while (true) w.pull_frame(); } catch (int err) { CHECK(err == EOF); }
        for (const Mat33& r : w.rotations) {
            std::printf("ROT");
            for (double v : r.m) std::printf(" %.17g", v);
            std::printf("\n");
        }
    }
    std::printf("%s: %d checks, %d failed\n", g_failed ? "FAILED" : "OK", g_checks, g_failed);
    return g_failed ? 1 : 0;
}
