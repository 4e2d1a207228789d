// vaw_flow_api.cu -- C-ABI of the motion-measurement kernels (include/vaw.h: vaw_flow_*).
//
// Replaces the per-frame tracking step of FrameSourceWarp::consume_frame
// (/root/reference/opencv/FrameSourceWarp.cpp:421-427 -> find_point_pairs_with_optical_flow, :242-270):
// the reference keeps the previous frame's luma plane (m_last_input_frame, :447) and runs
// cv::calcOpticalFlowPyrLK(prev, current, prev_corners) on every frame.  Here a vaw_flow holds the pyramids of
// the previous and the current frame in device memory; pushing a frame builds its pyramid once (the reference
// rebuilds both pyramids on every call), tracking reads both.  No CPU fallback.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <string>
#include <vector>
#include "../../include/vaw.h"
#include "vaw_flow.cuh"

struct vaw_flow {
    int device = 0, width = 0, height = 0, levels = 0;
    int lw[vaw::kFlowMaxLevels] = {}, lh[vaw::kFlowMaxLevels] = {};
    uint8_t* image[2][vaw::kFlowMaxLevels] = {};  // [slot][level]; level 0 is a private copy of the luma plane
    short2* deriv[2][vaw::kFlowMaxLevels] = {};
    int cur = 0, frames = 0;  // slot of the current frame; frames pushed so far
    float2 *d_prev = nullptr, *d_next = nullptr;
    uint8_t* d_status = nullptr;
    int capacity = 0;
    // corner detection (allocated by the first vaw_flow_corners call)
    float* d_eig = nullptr;
    uint2* d_cand = nullptr;
    unsigned* d_scalars = nullptr;  // [0] maximum of the response (float bits), [1] candidate count
    unsigned cand_capacity = 0;
    std::string err;
};

namespace {

thread_local std::string g_flow_error;

struct Guard {
    int prev = -1;
    explicit Guard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~Guard() { if (prev >= 0) cudaSetDevice(prev); }
};

int flow_fail(vaw_flow* f, int code, const std::string& msg)
{
    if (f) f->err = msg; else g_flow_error = msg;
    return code;
}

vaw::FlowPyramid pyramid_of(const vaw_flow* f, int slot)
{
    vaw::FlowPyramid p{};
    p.levels = f->levels;
    for (int l = 0; l < f->levels; ++l) {
        p.level[l].image = f->image[slot][l];
        p.level[l].deriv = f->deriv[slot][l];
        p.level[l].w = f->lw[l]; p.level[l].h = f->lh[l];
        p.level[l].pitch = f->lw[l]; p.level[l].dpitch = f->lw[l];
    }
    return p;
}

}  // namespace

extern "C" {

const char* vaw_flow_last_error(const vaw_flow* f) { return f ? f->err.c_str() : g_flow_error.c_str(); }

int vaw_flow_create(int width, int height, int device, vaw_flow** out)
{
    if (!out) return VAW_ERR_INVALID;
    *out = nullptr;
    if (width < 32 || height < 32 || width > 32766 || height > 32766) return flow_fail(nullptr, VAW_ERR_INVALID, "plane size must be in [32, 32766]");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return flow_fail(nullptr, VAW_ERR_CUDA, "no CUDA device: libvaw has no CPU fallback");
    }
    if (device < 0 || device >= n_dev) return flow_fail(nullptr, VAW_ERR_INVALID, "bad device ordinal");
    vaw_flow* f = new (std::nothrow) vaw_flow;
    if (!f) return VAW_ERR_NOMEM;
    f->device = device; f->width = width; f->height = height;
    // cv::buildOpticalFlowPyramid with maxLevel 3 and winSize 21: a level is added while the NEXT size stays above the window
    int w = width, h = height;
    for (int l = 0; l < vaw::kFlowMaxLevels; ++l) {
        f->lw[l] = w; f->lh[l] = h; f->levels = l + 1;
        w = (w + 1) / 2; h = (h + 1) / 2;
        if (w <= 21 || h <= 21) break;
    }
    Guard g(device);
    cudaError_t e = cudaSuccess;
    for (int s = 0; s < 2 && e == cudaSuccess; ++s)
        for (int l = 0; l < f->levels && e == cudaSuccess; ++l) {
            e = cudaMalloc(&f->image[s][l], (size_t)f->lw[l] * f->lh[l]);
            if (e == cudaSuccess) e = cudaMalloc(&f->deriv[s][l], sizeof(short2) * (size_t)f->lw[l] * f->lh[l]);
        }
    if (e != cudaSuccess) {
        const std::string msg = std::string("vaw_flow_create: ") + cudaGetErrorString(e);
        vaw_flow_destroy(f);
        cudaGetLastError();
        return flow_fail(nullptr, VAW_ERR_CUDA, msg);
    }
    *out = f;
    return VAW_OK;
}

void vaw_flow_destroy(vaw_flow* f)
{
    if (!f) return;
    Guard g(f->device);
    cudaDeviceSynchronize();
    for (int s = 0; s < 2; ++s)
        for (int l = 0; l < vaw::kFlowMaxLevels; ++l) { cudaFree(f->image[s][l]); cudaFree(f->deriv[s][l]); }
    cudaFree(f->d_prev); cudaFree(f->d_next); cudaFree(f->d_status);
    cudaFree(f->d_eig); cudaFree(f->d_cand); cudaFree(f->d_scalars);
    delete f;
}

int vaw_flow_levels(const vaw_flow* f) { return f ? f->levels : VAW_ERR_INVALID; }

int vaw_flow_push_frame(vaw_flow* f, const uint8_t* luma, int pitch, void* stream)
{
    if (!f) return VAW_ERR_INVALID;
    if (!luma || pitch < f->width) return flow_fail(f, VAW_ERR_INVALID, "bad luma plane");
    Guard g(f->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int s = f->frames == 0 ? f->cur : 1 - f->cur;
    cudaError_t e = cudaMemcpy2DAsync(f->image[s][0], (size_t)f->lw[0], luma, (size_t)pitch, (size_t)f->width, (size_t)f->height,
                                      cudaMemcpyDeviceToDevice, st);
    for (int l = 0; l < f->levels && e == cudaSuccess; ++l) {
        if (l > 0) e = vaw::launch_pyr_down(f->image[s][l - 1], f->lw[l - 1], f->lh[l - 1], f->lw[l - 1], f->image[s][l], f->lw[l], st);
        if (e == cudaSuccess) e = vaw::launch_scharr(f->image[s][l], f->lw[l], f->lh[l], f->lw[l], f->deriv[s][l], f->lw[l], st);
    }
    if (e != cudaSuccess) return flow_fail(f, VAW_ERR_CUDA, std::string("vaw_flow_push_frame: ") + cudaGetErrorString(e));
    f->cur = s;
    f->frames++;
    return VAW_OK;
}

int vaw_flow_track(vaw_flow* f, const float* prev_pts_xy, int n, float* next_pts_xy, uint8_t* status, void* stream)
{
    if (!f) return VAW_ERR_INVALID;
    if (n < 0 || (n > 0 && (!prev_pts_xy || !next_pts_xy || !status))) return flow_fail(f, VAW_ERR_INVALID, "null point arrays");
    if (f->frames < 2) return flow_fail(f, VAW_ERR_INVALID, "tracking needs two pushed frames");
    if (n == 0) return VAW_OK;
    Guard g(f->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (n > f->capacity) {
        cudaFree(f->d_prev); cudaFree(f->d_next); cudaFree(f->d_status);
        f->d_prev = f->d_next = nullptr; f->d_status = nullptr; f->capacity = 0;
        const int cap = n < 1024 ? 1024 : n;
        cudaError_t e = cudaMalloc(&f->d_prev, sizeof(float2) * cap);
        if (e == cudaSuccess) e = cudaMalloc(&f->d_next, sizeof(float2) * cap);
        if (e == cudaSuccess) e = cudaMalloc(&f->d_status, cap);
        if (e != cudaSuccess) return flow_fail(f, VAW_ERR_CUDA, std::string("vaw_flow_track: ") + cudaGetErrorString(e));
        f->capacity = cap;
    }
    cudaError_t e = cudaMemcpyAsync(f->d_prev, prev_pts_xy, sizeof(float2) * n, cudaMemcpyHostToDevice, st);
    // cv::calcOpticalFlowPyrLK defaults: TermCriteria(COUNT + EPS, 30, 0.01), minEigThreshold 1e-4
    if (e == cudaSuccess)
        e = vaw::launch_lk_track(pyramid_of(f, 1 - f->cur), pyramid_of(f, f->cur), f->d_prev, n, f->d_next, f->d_status, 30, 0.01f, 1e-4f, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(next_pts_xy, f->d_next, sizeof(float2) * n, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(status, f->d_status, n, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return flow_fail(f, VAW_ERR_CUDA, std::string("vaw_flow_track: ") + cudaGetErrorString(e));
    return VAW_OK;
}

// find_corners of the reference (FrameSourceWarp.cpp:228-240): cv::goodFeaturesToTrack(image, corners, max_corners,
// quality, min_distance) with blockSize 3 and the minimum-eigenvalue response.  Device: response map, its maximum,
// candidate list (vaw_corners.cu).  Host: the inherently sequential tail, exactly as OpenCV orders it -- candidates by
// decreasing response (ties: higher address first), greedy minimum-distance filter on a grid of min_distance cells.
// Only as many candidates are sorted as the selection needs (the strongest 4096 first, then four times more ...).
int vaw_flow_corners(vaw_flow* f, int which, int max_corners, double quality, double min_distance, float* corners_xy,
                     int capacity, int* n_out, void* stream)
{
    if (!f) return VAW_ERR_INVALID;
    if (!n_out || (capacity > 0 && !corners_xy) || capacity < 0 || which < 0 || which > 1 || !(quality > 0.0) || min_distance < 0.0)
        return flow_fail(f, VAW_ERR_INVALID, "bad corner arguments");
    *n_out = 0;
    if (f->frames < 1 || (which == 0 && f->frames < 2)) return flow_fail(f, VAW_ERR_INVALID, "no such frame has been pushed");
    Guard g(f->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int w = f->width, h = f->height;
    cudaError_t e = cudaSuccess;
    if (!f->d_eig) {
        const size_t px = (size_t)w * h;
        const unsigned cap = (unsigned)std::min<size_t>(px / 6 + 4096, 0x7fffffffu);
        e = cudaMalloc(&f->d_eig, px * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&f->d_cand, (size_t)cap * sizeof(uint2));
        if (e == cudaSuccess) e = cudaMalloc(&f->d_scalars, 2 * sizeof(unsigned));
        if (e != cudaSuccess) {
            cudaFree(f->d_eig); cudaFree(f->d_cand); cudaFree(f->d_scalars);
            f->d_eig = nullptr; f->d_cand = nullptr; f->d_scalars = nullptr;
            return flow_fail(f, VAW_ERR_CUDA, std::string("vaw_flow_corners: ") + cudaGetErrorString(e));
        }
        f->cand_capacity = cap;
    }
    const int s = which ? f->cur : 1 - f->cur;
    e = vaw::launch_corner_response(f->image[s][0], w, h, f->lw[0], f->d_eig, f->d_scalars, st);
    if (e == cudaSuccess)
        e = vaw::launch_corner_candidates(f->d_eig, w, h, f->d_scalars, quality, f->d_cand, f->cand_capacity, f->d_scalars + 1, st);
    unsigned scal[2] = {0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(scal, f->d_scalars, sizeof scal, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return flow_fail(f, VAW_ERR_CUDA, std::string("vaw_flow_corners: ") + cudaGetErrorString(e));
    if (scal[1] > f->cand_capacity) return flow_fail(f, VAW_ERR_UNSUPPORTED, "more corner candidates than one pixel in six: not a natural image");
    std::vector<uint2> cand(scal[1]);
    if (!cand.empty()) {
        e = cudaMemcpy(cand.data(), f->d_cand, cand.size() * sizeof(uint2), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return flow_fail(f, VAW_ERR_CUDA, std::string("vaw_flow_corners: ") + cudaGetErrorString(e));
    }
    // decreasing response; equal responses: the higher address (= larger y * w + x) first
    const auto stronger = [](const uint2& a, const uint2& b) { return a.x != b.x ? a.x > b.x : a.y > b.y; };
    const int cell = (int)std::lrint(min_distance);
    const bool filter = min_distance >= 1.0;
    const int gw = filter ? (w + cell - 1) / cell : 0, gh = filter ? (h + cell - 1) / cell : 0;
    std::vector<std::vector<std::pair<int, int>>> grid((size_t)gw * gh);
    const double md2 = min_distance * min_distance;
    int n = 0;
    size_t sorted = 0;
    bool done = false;
    while (sorted < cand.size() && !done) {
        // extend the sorted prefix: the strongest `want` candidates of the rest, in order
        const size_t want = std::min(cand.size() - sorted, sorted == 0 ? (size_t)4096 : 3 * sorted);
        std::partial_sort(cand.begin() + sorted, cand.begin() + sorted + want, cand.end(), stronger);
        for (size_t i = sorted; i < sorted + want; ++i) {
            const int y = (int)(cand[i].y / (unsigned)w), x = (int)(cand[i].y - (unsigned)y * (unsigned)w);
            bool good = true;
            if (filter) {
                const int xc = x / cell, yc = y / cell;
                for (int yy = std::max(0, yc - 1); yy <= std::min(gh - 1, yc + 1) && good; ++yy)
                    for (int xx = std::max(0, xc - 1); xx <= std::min(gw - 1, xc + 1) && good; ++xx)
                        for (const auto& m : grid[(size_t)yy * gw + xx]) {
                            const float dx = (float)(x - m.first), dy = (float)(y - m.second);
                            if (dx * dx + dy * dy < md2) { good = false; break; }
                        }
                if (good) grid[(size_t)yc * gw + xc].emplace_back(x, y);
            }
            if (!good) continue;
            if (n < capacity) { corners_xy[2 * n] = (float)x; corners_xy[2 * n + 1] = (float)y; }
            ++n;
            if ((max_corners > 0 && n == max_corners) || n == capacity) { done = true; break; }
        }
        sorted += want;
    }
    *n_out = n;
    return VAW_OK;
}

// The response map of the last vaw_flow_corners call (w x h floats, host buffer), for parity tests.
int vaw_flow_get_response(vaw_flow* f, float* response_host)
{
    if (!f) return VAW_ERR_INVALID;
    if (!response_host || !f->d_eig) return flow_fail(f, VAW_ERR_INVALID, "no response map yet");
    Guard g(f->device);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(response_host, f->d_eig, (size_t)f->width * f->height * sizeof(float), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return flow_fail(f, VAW_ERR_CUDA, std::string("vaw_flow_get_response: ") + cudaGetErrorString(e));
    return VAW_OK;
}

int vaw_flow_get_level(vaw_flow* f, int which, int level, uint8_t* image_host, int16_t* deriv_host, int* width, int* height)
{
    if (!f) return VAW_ERR_INVALID;
    if (which < 0 || which > 1 || level < 0 || level >= f->levels || !width || !height) return flow_fail(f, VAW_ERR_INVALID, "bad level");
    Guard g(f->device);
    const int s = which ? f->cur : 1 - f->cur;
    *width = f->lw[level]; *height = f->lh[level];
    const size_t px = (size_t)f->lw[level] * f->lh[level];
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess && image_host) e = cudaMemcpy(image_host, f->image[s][level], px, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && deriv_host) e = cudaMemcpy(deriv_host, f->deriv[s][level], px * sizeof(short2), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return flow_fail(f, VAW_ERR_CUDA, std::string("vaw_flow_get_level: ") + cudaGetErrorString(e));
    return VAW_OK;
}

}  // extern "C"
