// vaw_clip.cpp -- frame-parallel scheduler over the GPUs of one box (north_star item 3).
//
// No reference counterpart: the reference is single device and single thread
// (/root/reference/opencv/DisplayImage.cpp:60-72 pulls one frame at a time).  Once each frame
// has its rotation the frames are independent (FrameSourceWarp.cpp:272-314), so the clip is cut
// into contiguous ranges, one per device, and every device runs the host-buffer pipeline of
// vaw_warp_batch_host (pinned staging, H2D -> warp -> D2H on its own streams) from its own
// host thread, bound to the CPUs of that device's NUMA node (so the pinned staging it allocates is
// local to the device's PCIe root).  No collective, no peer traffic: NCCL has nothing to do on this path.
#include <cstdint>
#include <new>
#include <string>
#include <thread>
#include <vector>
#include "../../include/vaw.h"

struct vaw_clip {
    std::vector<vaw_ctx*> ctx;
    std::vector<int> device;
    vaw_params p{};
    std::string err;
};

namespace {
thread_local std::string g_clip_create_error;
}

extern "C" {

const char* vaw_clip_last_error(const vaw_clip* clip) { return clip ? clip->err.c_str() : g_clip_create_error.c_str(); }

int vaw_clip_create(const vaw_params* params, int n_devices, const int* devices, vaw_clip** out)
{
    if (!params || !out || n_devices < 1) {
        g_clip_create_error = "bad argument";
        return VAW_ERR_INVALID;
    }
    *out = nullptr;
    vaw_clip* clip = new (std::nothrow) vaw_clip;
    if (!clip) return VAW_ERR_NOMEM;
    clip->p = *params;
    for (int i = 0; i < n_devices; ++i) {
        const int dev = devices ? devices[i] : i;
        vaw_ctx* c = nullptr;
        const int rc = vaw_create(params, dev, &c);
        if (rc != VAW_OK) {
            g_clip_create_error = std::string("device ") + std::to_string(dev) + ": " + vaw_last_error(nullptr);
            vaw_clip_destroy(clip);
            return rc;
        }
        clip->ctx.push_back(c);
        clip->device.push_back(dev);
    }
    *out = clip;
    return VAW_OK;
}

void vaw_clip_destroy(vaw_clip* clip)
{
    if (!clip) return;
    for (vaw_ctx* c : clip->ctx) vaw_destroy(c);
    delete clip;
}

int vaw_clip_warp_host(vaw_clip* clip, const uint8_t* src_host, uint8_t* dst_host, const double* rotations_host,
                       int n_frames)
{
    if (!clip) return VAW_ERR_INVALID;
    if (!src_host || !dst_host || !rotations_host || n_frames < 0) {
        clip->err = "null host buffer";
        return VAW_ERR_INVALID;
    }
    const int n = (int)clip->ctx.size();
    const int sfmt = clip->p.format == VAW_FORMAT_NV12_TO_BGR24 ? VAW_FORMAT_NV12 : clip->p.format;
    const int dfmt = clip->p.format == VAW_FORMAT_NV12_TO_BGR24 ? VAW_FORMAT_BGR24 : clip->p.format;
    const int sch = sfmt == VAW_FORMAT_BGR24 ? 3 : 1, dch = dfmt == VAW_FORMAT_BGR24 ? 3 : 1;
    const size_t sfb = vaw_frame_bytes(sfmt, clip->p.src_width, clip->p.src_height, clip->p.src_width * sch);
    const size_t dfb = vaw_frame_bytes(dfmt, clip->p.out_width, clip->p.out_height, clip->p.out_width * dch);
    std::vector<int> rc(n, VAW_OK);
    std::vector<std::thread> th;
    for (int i = 0; i < n; ++i) {
        int first = 0, count = 0;
        vaw_shard_range(n_frames, n, i, &first, &count);
        if (count == 0) continue;
        th.emplace_back([=, &rc]() {
            vaw_bind_thread_to_device(clip->device[i]);  // before the first pinned allocation of this context
            rc[i] = vaw_warp_batch_host(clip->ctx[i], src_host + (size_t)first * sfb, dst_host + (size_t)first * dfb,
                                        rotations_host + (size_t)first * 9, count);
        });
    }
    for (std::thread& t : th) t.join();
    for (int i = 0; i < n; ++i)
        if (rc[i] != VAW_OK) {
            clip->err = std::string("device ") + std::to_string(clip->device[i]) + ": " + vaw_last_error(clip->ctx[i]);
            return rc[i];
        }
    return VAW_OK;
}

}  // extern "C"
