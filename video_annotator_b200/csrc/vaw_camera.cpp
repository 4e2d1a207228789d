// vaw_camera.cpp -- camera parameter producers of the warp path (host, double precision).
//
// Replaces get_preset_camera (/root/reference/opencv/FrameSourceWarp.cpp:27-86) and
// get_output_camera (:88-165).  They run once per stream on the host; kept bug-for-bug
// because they decide the output size and intrinsics the kernel is fed:
//   - the "PUBLISHED" field-of-view constants are declared `const int` in the reference
//     (:22-25), so 122.6/94.4/118.2/69.5 degrees act as 122/94/118/69;
//   - the "MEASURED" presets scale cx by width but fx, fy AND cy by height (:50-77);
//   - both diagonals of the scale estimate pass through integer cv::Point (:142-150);
//   - the output size is truncated, not rounded (:163).
// cv::fisheye::undistortPoints (OpenCV calib3d) is called with the camera's
// distortion_coefficients and identity R/P (:93-110).  The presets carry zeros (:35), which
// reduces to p * tan(theta)/theta with theta = |p| clamped to pi/2; non-zero k1..k4 (the
// explicit-camera constructor) go through the same Newton inversion OpenCV runs (10 iterations,
// epsilon 1e-8), checked against cv2.fisheye.undistortPoints in tests/test_oracle_camera.py.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../include/vaw.h"

namespace {

constexpr double kPi = 3.1415926535897932384626433832795;

struct Pt { double x, y; };

// cv::fisheye::undistortPoints for one point, R = P = identity: theta_d = |p| (clamped to the
// model's 180-degree field), theta from theta_d = theta (1 + k1 theta^2 + ... + k4 theta^8) by
// Newton's method started at theta_d, result p * tan(theta) / theta_d.  Points whose iteration
// does not converge or changes sign come back as (-1e6, -1e6), as in OpenCV.
Pt undistort_fisheye(Pt d, const double K[9], const double kd[4])
{
    Pt pw{(d.x - K[2]) / K[0], (d.y - K[5]) / K[4]};
    double theta_d = std::sqrt(pw.x * pw.x + pw.y * pw.y);
    theta_d = std::fmin(std::fmax(-kPi / 2, theta_d), kPi / 2);
    double theta = theta_d, scale = 0.0;
    bool converged = false;
    if (std::fabs(theta_d) > 1e-8) {
        for (int j = 0; j < 10; ++j) {
            const double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t6 * t2;
            const double k0 = kd[0] * t2, k1 = kd[1] * t4, k2 = kd[2] * t6, k3 = kd[3] * t8;
            const double fix = (theta * (1 + k0 + k1 + k2 + k3) - theta_d) / (1 + 3 * k0 + 5 * k1 + 7 * k2 + 9 * k3);
            theta -= fix;
            if (std::fabs(fix) < 1e-8) { converged = true; break; }
        }
        scale = std::tan(theta) / theta_d;
    } else {
        converged = true;
    }
    const bool flipped = (theta_d < 0 && theta > 0) || (theta_d > 0 && theta < 0);
    if (converged && !flipped) return {pw.x * scale, pw.y * scale};
    return {-1000000.0, -1000000.0};
}

inline int round_half_even(double v) { return (int)std::lrint(v); }  // cv::saturate_cast<int>(double)

}  // namespace

extern "C" int vaw_get_preset_camera(int preset, int width, int height, vaw_camera* out)
{
    if (!out || width <= 0 || height <= 0) return VAW_ERR_INVALID;
    if (preset < VAW_GOPRO_H4B_WIDE43_PUBLISHED || preset > VAW_GOPRO_H4B_WIDE169_MEASURED_STABILISATION)
        return VAW_ERR_INVALID;
    const int fov_h_43 = (int)122.6, fov_v_43 = (int)94.4, fov_h_169 = (int)118.2, fov_v_169 = (int)69.5;
    double fx = 1, fy = 1, cx = (width - 1.) / 2, cy = (height - 1.) / 2;
    struct Measured { double cx, cy, fx, fy, ref_w, ref_h; };
    static const Measured kMeasured[6] = {
        {}, {967.37, 711.07, 942.96, 942.53, 1920, 1440}, {965.90, 712.94, 1045.58, 1045.64, 1920, 1440},
        {}, {1361.80, 745.19, 1392.49, 1383.47, 2704, 1520}, {1357.49, 736.74, 1626.67, 1619.46, 2704, 1520}};
    switch (preset) {
    case VAW_GOPRO_H4B_WIDE43_PUBLISHED:
        fx = width / (fov_h_43 * kPi / 180);
        fy = height / (fov_v_43 * kPi / 180);
        break;
    case VAW_GOPRO_H4B_WIDE169_PUBLISHED:
        fx = width / (fov_h_169 * kPi / 180);
        fy = height / (fov_v_169 * kPi / 180);
        break;
    default: {
        const Measured& m = kMeasured[preset];
        cx = m.cx * width / m.ref_w;
        cy = m.cy * height / m.ref_h;
        fx = m.fx * height / m.ref_h;
        fy = m.fy * height / m.ref_h;
    }
    }
    std::memset(out, 0, sizeof *out);
    out->model = 1;  // FISHEYE
    out->width = width;
    out->height = height;
    const double K[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
    std::memcpy(out->matrix, K, sizeof K);
    return VAW_OK;
}

extern "C" int vaw_get_output_camera(const vaw_camera* input, double scale, int crop_borders,
                                     double zoom, vaw_camera* out)
{
    if (!input || !out || zoom == 0) return VAW_ERR_INVALID;
    // get_output_camera always runs fisheye::undistortPoints (:93-110), whatever Camera::model says;
    // this library's kernels implement the fisheye input model only, so anything else is refused
    // here instead of being warped with the wrong projection
    if (input->model != 1) return VAW_ERR_UNSUPPORTED;
    const double* K = input->matrix;
    const double w1 = input->width - 1, h1 = input->height - 1;
    // four corners, then the four edge midpoints through the principal point
    const Pt probes[8] = {{0, 0}, {0, h1}, {w1, 0}, {w1, h1}, {K[2], 0}, {w1, K[5]}, {K[2], h1}, {0, K[5]}};
    Pt e[8];
    for (int i = 0; i < 8; ++i) e[i] = undistort_fisheye(probes[i], K, input->distortion);

    const int first = crop_borders ? 4 : 0;
    double min_x = e[first].x, max_x = e[first].x, min_y = e[first].y, max_y = e[first].y;
    for (int i = first + 1; i < 8; ++i) {
        min_x = std::fmin(min_x, e[i].x);
        max_x = std::fmax(max_x, e[i].x);
        min_y = std::fmin(min_y, e[i].y);
        max_y = std::fmax(max_y, e[i].y);
    }
    const double in_dx = round_half_even(w1), in_dy = round_half_even(h1);
    const double out_dx = round_half_even(e[3].x - e[0].x), out_dy = round_half_even(e[3].y - e[0].y);
    const double f = scale * std::sqrt(in_dx * in_dx + in_dy * in_dy) / std::sqrt(out_dx * out_dx + out_dy * out_dy);

    std::memset(out, 0, sizeof *out);
    out->model = 0;  // RECTILINEAR
    const double M[9] = {f, 0, f * -min_x / zoom, 0, f, f * -min_y / zoom, 0, 0, 1};
    std::memcpy(out->matrix, M, sizeof M);
    out->width = (int)(f * (max_x - min_x) / zoom);
    out->height = (int)(f * (max_y - min_y) / zoom);
    return VAW_OK;
}

extern "C" int vaw_params_from_cameras(const vaw_camera* input, const vaw_camera* output, int format,
                                       vaw_params* p)
{
    if (!input || !output || !p) return VAW_ERR_INVALID;
    if (input->model < 0 || input->model > 1 || output->model < 0 || output->model > 1) return VAW_ERR_INVALID;
    p->src_center_x = input->matrix[2];
    p->src_center_y = input->matrix[5];
    p->src_focal_x = input->matrix[0];
    p->src_focal_y = input->matrix[4];
    p->map_center_x = output->matrix[2];
    p->map_center_y = output->matrix[5];
    p->map_focal_x = output->matrix[0];
    p->map_focal_y = output->matrix[4];
    p->src_width = input->width;
    p->src_height = input->height;
    p->out_width = output->width;
    p->out_height = output->height;
    if (format == VAW_FORMAT_NV12) {
        p->out_width &= ~1;
        p->out_height &= ~1;
    }
    p->format = format;
    p->interpolation = VAW_INTER_LINEAR;  // the constructor's default (FrameSourceWarp.hpp:90)
    p->variant = VAW_VARIANT_AUTO;
    // CameraModel (FrameSourceWarp.hpp:23-26): FISHEYE in + RECTILINEAR out is createMap.cl's pair (0)
    p->projection = (input->model == 0 ? 1 : 0) | (output->model == 1 ? 2 : 0);
    for (int i = 0; i < 4; ++i) p->src_distortion[i] = (float)input->distortion[i];  // zeros for the presets (:35)
    return VAW_OK;
}

// ---- guess_camera_rotation (/root/reference/opencv/FrameSourceWarp.cpp:316-368) -------------------------------
// The reference undistorts the tracked points of the previous frame to normalised rays and those of the current
// frame to pixels of the output camera (cv::fisheye::undistortPoints, :321-337), gives every previous ray a RANDOM
// depth ("prevents the detection of translations, but doesn't affect rotations", :341-350) and calls
// cv::solvePnPRansac(100 iterations, 8 px, 0.99) (:353-365), of which it keeps the rotation and the inlier count.
// With random depths the translation of that pose is meaningless and the model that explains the inliers is a pure
// rotation of the rays; this restatement fits exactly that: RANSAC over two-ray samples (TRIAD), consensus by the
// same 8-pixel reprojection error in the output camera, least-squares rotation (Wahba / Kabsch, polar factor) over
// the consensus set, re-scored and re-fitted twice.  cv::solvePnPRansac lives in OpenCV's calib3d (third-party) and
// draws from rand() / cv::theRNG(), so equality is to a tolerance: tests pin the result to the real cv2 route to
// 0.05 degrees on synthetic motions (tests/test_oracle_rotation.py, tests/golden/rotation_cases.npz).
namespace {

struct V3 { double x, y, z; };
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 unit(V3 a) { const double n = std::sqrt(dot(a, a)); return n > 0 ? V3{a.x / n, a.y / n, a.z / n} : V3{0, 0, 1}; }
inline V3 mul(const double R[9], V3 a) { return {R[0] * a.x + R[1] * a.y + R[2] * a.z, R[3] * a.x + R[4] * a.y + R[5] * a.z, R[6] * a.x + R[7] * a.y + R[8] * a.z}; }

// rotation with R a1 = b1 exactly and a2 brought as close to b2 as a rotation about b1 allows
bool triad(V3 a1, V3 a2, V3 b1, V3 b2, double R[9])
{
    const V3 an = cross(a1, a2), bn = cross(b1, b2);
    if (dot(an, an) < 1e-12 || dot(bn, bn) < 1e-12) return false;
    const V3 ta2 = unit(an), tb2 = unit(bn), ta3 = cross(a1, ta2), tb3 = cross(b1, tb2);
    const V3 ta[3] = {a1, ta2, ta3}, tb[3] = {b1, tb2, tb3};
    for (int i = 0; i < 9; ++i) R[i] = 0;
    for (int k = 0; k < 3; ++k) {
        const double bv[3] = {tb[k].x, tb[k].y, tb[k].z}, av[3] = {ta[k].x, ta[k].y, ta[k].z};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) R[3 * i + j] += bv[i] * av[j];
    }
    return true;
}

// nearest rotation to M (polar factor, Newton iteration X <- (X + X^-T) / 2); false when M is (near) singular
bool polar_rotation(const double M[9], double R[9])
{
    double X[9], n = 0;
    for (int i = 0; i < 9; ++i) n += M[i] * M[i];
    if (!(n > 0)) return false;
    n = std::sqrt(n);
    for (int i = 0; i < 9; ++i) X[i] = M[i] / n;
    for (int it = 0; it < 60; ++it) {
        const double c0 = X[4] * X[8] - X[5] * X[7], c1 = X[5] * X[6] - X[3] * X[8], c2 = X[3] * X[7] - X[4] * X[6];
        const double det = X[0] * c0 + X[1] * c1 + X[2] * c2;
        if (!(std::fabs(det) > 1e-14)) return false;
        const double d = 1.0 / det;
        // inverse transposed = cofactor matrix / det
        const double C[9] = {c0 * d, c1 * d, c2 * d,
                             (X[2] * X[7] - X[1] * X[8]) * d, (X[0] * X[8] - X[2] * X[6]) * d, (X[1] * X[6] - X[0] * X[7]) * d,
                             (X[1] * X[5] - X[2] * X[4]) * d, (X[2] * X[3] - X[0] * X[5]) * d, (X[0] * X[4] - X[1] * X[3]) * d};
        double delta = 0;
        for (int i = 0; i < 9; ++i) {
            const double v = 0.5 * (X[i] + C[i]);
            delta += (v - X[i]) * (v - X[i]);
            X[i] = v;
        }
        if (delta < 1e-30) break;
    }
    const double det = X[0] * (X[4] * X[8] - X[5] * X[7]) + X[1] * (X[5] * X[6] - X[3] * X[8]) + X[2] * (X[3] * X[7] - X[4] * X[6]);
    if (!(det > 0.5)) return false;  // a reflection: the correspondences do not describe a rotation
    for (int i = 0; i < 9; ++i) R[i] = X[i];
    return true;
}

struct Xorshift {
    uint64_t s;
    explicit Xorshift(uint32_t seed) : s(0x9E3779B97F4A7C15ull ^ ((uint64_t)seed * 0xD1342543DE82EF95ull + 1)) {}
    uint32_t next() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 32); }
};

}  // namespace

extern "C" int vaw_guess_rotation(const vaw_camera* input, const vaw_camera* output, const float* prev_xy, const float* cur_xy,
                                  int n, uint32_t seed, double rotation[9], int* inliers_out)
{
    if (!input || !output || !rotation || !inliers_out || n < 0 || (n > 0 && (!prev_xy || !cur_xy))) return VAW_ERR_INVALID;
    for (int i = 0; i < 9; ++i) rotation[i] = (i % 4 == 0) ? 1.0 : 0.0;
    *inliers_out = 0;
    const double fx = output->matrix[0], fy = output->matrix[4], cx = output->matrix[2], cy = output->matrix[5];
    std::vector<V3> a, b;  // rays of the previous / the current frame
    std::vector<Pt> pix;   // the current points in output-camera pixels
    a.reserve(n); b.reserve(n); pix.reserve(n);
    for (int i = 0; i < n; ++i) {
        const Pt p = undistort_fisheye({prev_xy[2 * i], prev_xy[2 * i + 1]}, input->matrix, input->distortion);
        const Pt c = undistort_fisheye({cur_xy[2 * i], cur_xy[2 * i + 1]}, input->matrix, input->distortion);
        if (p.x <= -999999.0 || c.x <= -999999.0 || !std::isfinite(p.x + p.y + c.x + c.y)) continue;
        a.push_back(unit({p.x, p.y, 1.0}));
        b.push_back(unit({c.x, c.y, 1.0}));
        pix.push_back({fx * c.x + cx, fy * c.y + cy});
    }
    const int m = (int)a.size();
    if (m < 4) return VAW_OK;  // cv::solvePnPRansac needs four points: the caller sees 0 inliers (and < 40 keeps the last rotation)
    const double thr2 = 8.0 * 8.0;
    const auto score = [&](const double R[9], std::vector<int>* set) {
        int cnt = 0;
        if (set) set->clear();
        for (int i = 0; i < m; ++i) {
            const V3 r = mul(R, a[i]);
            if (!(r.z > 1e-9)) continue;
            const double ex = fx * r.x / r.z + cx - pix[i].x, ey = fy * r.y / r.z + cy - pix[i].y;
            if (ex * ex + ey * ey < thr2) { ++cnt; if (set) set->push_back(i); }
        }
        return cnt;
    };
    Xorshift rng(seed);
    double best[9];
    for (int i = 0; i < 9; ++i) best[i] = rotation[i];
    int best_cnt = score(best, nullptr);  // the camera did not move: a valid hypothesis
    int iters = 100;
    for (int it = 0; it < iters; ++it) {
        const int i = (int)(rng.next() % (uint32_t)m);
        int j = (int)(rng.next() % (uint32_t)(m - 1));
        if (j >= i) ++j;
        double R[9];
        if (!triad(a[i], a[j], b[i], b[j], R)) continue;
        const int cnt = score(R, nullptr);
        if (cnt > best_cnt) {
            best_cnt = cnt;
            for (int k = 0; k < 9; ++k) best[k] = R[k];
            // cv::RANSACUpdateNumIters with confidence 0.99 and a two-point model
            const double w = (double)cnt / m, denom = std::log(1.0 - w * w);
            if (denom < -1e-12) {
                const double need = std::log(1.0 - 0.99) / denom;
                if (need < iters) iters = (int)std::ceil(need) > it + 1 ? (int)std::ceil(need) : it + 1;
            } else {
                iters = it + 1;
            }
        }
    }
    std::vector<int> set;
    for (int round = 0; round < 3; ++round) {
        const int cnt = score(best, &set);
        if (cnt < 3) break;
        double M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int idx : set) {
            const double bv[3] = {b[idx].x, b[idx].y, b[idx].z}, av[3] = {a[idx].x, a[idx].y, a[idx].z};
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) M[3 * r + c] += bv[r] * av[c];
        }
        double R[9];
        if (!polar_rotation(M, R)) break;
        // the fit over the whole consensus set replaces the two-ray hypothesis (cv::solvePnPRansac likewise ends
        // with a solve over its inliers); a sample or two may drop out of the 8-pixel band, that is expected
        if (score(R, nullptr) < cnt / 2) break;
        for (int k = 0; k < 9; ++k) best[k] = R[k];
    }
    best_cnt = score(best, nullptr);
    for (int i = 0; i < 9; ++i) rotation[i] = best[i];
    *inliers_out = best_cnt;
    return VAW_OK;
}
