// FrameSource.hpp -- the pull interface of the frame pipeline, over device frames.
//
// Mirrors /root/reference/opencv/FrameSource.hpp:9-24: same class name, same two methods,
// same contract (return the next frame; raise when none is ready; end of stream is
// `throw EOF`, /root/reference/opencv/FrameSourceWarp.cpp:465-467).  The reference passes
// cv::UMat (an OpenCL buffer); here a frame is a ref-counted handle to an NV12 (or packed)
// buffer in CUDA device memory, owned through the C-ABI (include/vaw.h) -- OpenCV's C++
// headers are not part of this build.
#ifndef VAW_FRAME_SOURCE_HPP_
#define VAW_FRAME_SOURCE_HPP_

#include <cstddef>
#include <cstdint>
#include <memory>

class FramePool;

struct DeviceFrame {
    uint8_t* data = nullptr;  // device pointer
    int width = 0, height = 0, pitch = 0;
    int format = 0;           // VAW_FORMAT_*: NV12 = one plane of width x 3*height/2 bytes
                              // (opencv/FrameSourceFfmpegOpenCl.cpp:58,75-85)
    int device = 0;
    long index = -1;          // position in the stream, for bookkeeping
    size_t bytes = 0;
    std::shared_ptr<FramePool> pool;  // the slab this frame's buffer is a slot of (null: a plain allocation)
    std::shared_ptr<DeviceFrame> alias_of;  // set: a second handle on that frame's buffer (keeps it alive, frees nothing)
    ~DeviceFrame();           // returns the slot to its pool (or the buffer to the C-ABI allocator)
};

// A slab of equally spaced frame buffers in device memory: ONE allocation for `slots` frames, handed out
// in ring order, so that frames requested one after the other sit at a constant stride -- which is what
// lets a run of buffered frames go through one batched launch (vaw_warp_batch) -- and so that the steady
// state allocates nothing (the reference gets a fresh cv::UMat per frame from OpenCV's own buffer pool).
class FramePool {
    int m_device;
    size_t m_stride;
    int m_slots, m_cursor = 0, m_used = 0;
    uint8_t* m_base = nullptr;
    bool m_owns = true;
    std::unique_ptr<bool[]> m_busy;
  public:
    FramePool(int device, size_t frame_bytes, int slots);  // throws int on failure
    // over memory the caller owns (a decoder's surface ring; bookkeeping tests): `slots` frames `stride` bytes apart
    FramePool(int device, uint8_t* base, size_t stride, int slots);
    ~FramePool();
    FramePool(const FramePool&) = delete;
    FramePool& operator=(const FramePool&) = delete;
    uint8_t* acquire();          // next free slot in ring order, nullptr when every slot is out
    void release(uint8_t* slot);
    uint8_t* base() const { return m_base; }
    size_t stride() const { return m_stride; }
    int slots() const { return m_slots; }
    int used() const { return m_used; }
    int device() const { return m_device; }
};
// The shared pool for frames of `bytes` bytes on `device` (created on first use, kDefaultSlots slots).
std::shared_ptr<FramePool> frame_pool(int device, size_t bytes);
void frame_pool_trim();  // drop the shared pools (frames still out keep their slab alive)
constexpr int kDefaultPoolSlots = 96;

using Frame = std::shared_ptr<DeviceFrame>;  // stands in for cv::UMat (ref-counted, returned by value)

// A frame buffer from the shared pool (a plain allocation when the pool is exhausted); throws int on failure.
Frame make_device_frame(int device, int format, int width, int height);

// One stage of the pull pipeline.  Contract as in the reference (FrameSource.hpp:9-24):
//   pull_frame  hands out the next frame and moves on;
//   peek_frame  hands out the next frame and stays where it is;
// both throw when nothing can be delivered -- `throw EOF` (an int) at the end of the stream.
class FrameSource {
  public:
    virtual Frame pull_frame() = 0;
    virtual Frame peek_frame() = 0;
    virtual ~FrameSource() = default;
};

#endif  // VAW_FRAME_SOURCE_HPP_
