// vaw_flow_api.cu -- C-ABI of the motion-measurement kernels (include/vaw.h: vaw_flow_*).
//
// Replaces the per-frame tracking step of FrameSourceWarp::consume_frame
// (/root/reference/opencv/FrameSourceWarp.cpp:421-427 -> find_point_pairs_with_optical_flow, :242-270):
// the reference keeps the previous frame's luma plane (m_last_input_frame, :447) and runs
// cv::calcOpticalFlowPyrLK(prev, current, prev_corners) on every frame.  Here a vaw_flow holds the pyramids of
// the previous and the current frame in device memory; pushing a frame builds its pyramid once (the reference
// rebuilds both pyramids on every call), tracking reads both.  No CPU fallback.
#include <cuda_runtime.h>
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
