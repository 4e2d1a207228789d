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
#include <cstring>
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
