/*
 * camera_ref.c -- restatement of the reference's camera parameter producers.
 * TEST INFRASTRUCTURE ONLY (see vaw_oracle.h).
 *
 *   get_preset_camera   /root/reference/opencv/FrameSourceWarp.cpp:27-86
 *   get_output_camera   /root/reference/opencv/FrameSourceWarp.cpp:88-165
 *   types               /root/reference/opencv/FrameSourceWarp.hpp:14-34
 *
 * cv::fisheye::undistortPoints (third-party, OpenCV calib3d, call site
 * FrameSourceWarp.cpp:93-110) is restated for the only case the reference uses:
 * all four distortion coefficients zero (FrameSourceWarp.cpp:35), R = P = identity.
 * Published algorithm: pw = ((x-cx)/fx, (y-cy)/fy); theta_d = |pw| clamped to
 * [-pi/2, pi/2]; with k = 0 the Newton solve returns theta = theta_d;
 * scale = tan(theta)/theta_d (1 when theta_d <= 1e-8); pu = pw*scale.
 * Cross-checked against cv2.fisheye.undistortPoints in tests/test_oracle_camera.py.
 */
#include <math.h>
#include <string.h>
#include "vaw_oracle.h"

#define CV_PI 3.1415926535897932384626433832795

/* FrameSourceWarp.cpp:22-25 -- declared `const int`, so the published FOVs truncate. */
static const int GOPRO_H5B_FOV_H_43W_NOSTAB = (int)122.6;
static const int GOPRO_H5B_FOV_V_43W_NOSTAB = (int)94.4;
static const int GOPRO_H5B_FOV_H_169W_NOSTAB = (int)118.2;
static const int GOPRO_H5B_FOV_V_169W_NOSTAB = (int)69.5;

enum { /* FrameSourceWarp.hpp:14-21 */
    GOPRO_H4B_WIDE43_PUBLISHED,
    GOPRO_H4B_WIDE43_MEASURED,
    GOPRO_H4B_WIDE43_MEASURED_STABILISATION,
    GOPRO_H4B_WIDE169_PUBLISHED,
    GOPRO_H4B_WIDE169_MEASURED,
    GOPRO_H4B_WIDE169_MEASURED_STABILISATION
};

void vaw_oracle_get_preset_camera(int preset, int width, int height, vaw_oracle_camera *out)
{
    double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    /* :31-32 default principal point at the centre */
    m[2] = (width - 1.) / 2;
    m[5] = (height - 1.) / 2;
    switch (preset) {
    case GOPRO_H4B_WIDE43_PUBLISHED: /* :38-43 */
        m[0] = width / (GOPRO_H5B_FOV_H_43W_NOSTAB * CV_PI / 180);
        m[4] = height / (GOPRO_H5B_FOV_V_43W_NOSTAB * CV_PI / 180);
        break;
    case GOPRO_H4B_WIDE169_PUBLISHED: /* :44-49 */
        m[0] = width / (GOPRO_H5B_FOV_H_169W_NOSTAB * CV_PI / 180);
        m[4] = height / (GOPRO_H5B_FOV_V_169W_NOSTAB * CV_PI / 180);
        break;
    case GOPRO_H4B_WIDE43_MEASURED: /* :50-56; fx, fy, cy scale by HEIGHT */
        m[2] = 967.37 * width / 1920;
        m[5] = 711.07 * height / 1440;
        m[0] = 942.96 * height / 1440;
        m[4] = 942.53 * height / 1440;
        break;
    case GOPRO_H4B_WIDE43_MEASURED_STABILISATION: /* :57-63 */
        m[2] = 965.90 * width / 1920;
        m[5] = 712.94 * height / 1440;
        m[0] = 1045.58 * height / 1440;
        m[4] = 1045.64 * height / 1440;
        break;
    case GOPRO_H4B_WIDE169_MEASURED: /* :64-70 */
        m[2] = 1361.80 * width / 2704;
        m[5] = 745.19 * height / 1520;
        m[0] = 1392.49 * height / 1520;
        m[4] = 1383.47 * height / 1520;
        break;
    case GOPRO_H4B_WIDE169_MEASURED_STABILISATION: /* :71-77 */
        m[2] = 1357.49 * width / 2704;
        m[5] = 736.74 * height / 1520;
        m[0] = 1626.67 * height / 1520;
        m[4] = 1619.46 * height / 1520;
        break;
    default:
        break;
    }
    out->model = 1; /* FISHEYE, :81 */
    memcpy(out->matrix, m, sizeof m);
    memset(out->dist, 0, sizeof out->dist); /* :35 */
    out->width = width;
    out->height = height;
}

/* fisheye::undistortPoints, D = 0, R = P = I */
static void undistort_point_zero_dist(double x, double y, const double K[9], double *ox, double *oy)
{
    double pwx = (x - K[2]) / K[0], pwy = (y - K[5]) / K[4];
    double theta_d = sqrt(pwx * pwx + pwy * pwy);
    if (theta_d > CV_PI / 2) theta_d = CV_PI / 2;
    double scale = 1.0;
    if (theta_d > 1e-8) scale = tan(theta_d) / theta_d;
    *ox = pwx * scale;
    *oy = pwy * scale;
}

/* cv::Point(const Point2d&) conversion = saturate_cast<int> = cvRound (half-even) */
static int cv_round_d(double v) { return (int)lrint(v); }

void vaw_oracle_get_output_camera(const vaw_oracle_camera *in, double scale, int crop_borders,
                                  double zoom, vaw_oracle_camera *out)
{
    int w = in->width, h = in->height;
    const double *K = in->matrix;
    /* :94-106 corners then edge midpoints */
    double px[8] = {0, 0, w - 1, w - 1, K[2], w - 1, K[2], 0};
    double py[8] = {0, h - 1, 0, h - 1, 0, K[5], h - 1, K[5]};
    double ex[8], ey[8];
    for (int i = 0; i < 8; ++i) undistort_point_zero_dist(px[i], py[i], K, &ex[i], &ey[i]);

    /* :119-139 bounding rectangle over all 8 points, or the midpoints only */
    int start = crop_borders ? 4 : 0;
    double max_x = ex[start], min_x = ex[start], max_y = ey[start], min_y = ey[start];
    for (int i = start + 1; i < 8; ++i) {
        if (max_x < ex[i]) max_x = ex[i];
        if (ex[i] < min_x) min_x = ex[i];
        if (max_y < ey[i]) max_y = ey[i];
        if (ey[i] < min_y) min_y = ey[i];
    }

    /* :142-150 both diagonals pass through integer cv::Point */
    int idx = cv_round_d(w - 1), idy = cv_round_d(h - 1);
    double in_len = sqrt(1. * idx * idx + idy * idy);
    int odx = cv_round_d(ex[3] - ex[0]), ody = cv_round_d(ey[3] - ey[0]);
    double out_len = sqrt(1. * odx * odx + ody * ody);
    scale *= in_len / out_len;

    /* :153-157 */
    double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    m[0] = scale;
    m[4] = scale;
    m[2] = scale * -min_x / zoom;
    m[5] = scale * -min_y / zoom;

    out->model = 0; /* RECTILINEAR, :160 */
    memcpy(out->matrix, m, sizeof m);
    memset(out->dist, 0, sizeof out->dist);
    /* :163 Size(double, double) truncates toward zero */
    out->width = (int)(scale * (max_x - min_x) / zoom);
    out->height = (int)(scale * (max_y - min_y) / zoom);
}
