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

struct DeviceFrame {
    uint8_t* data = nullptr;  // device pointer
    int width = 0, height = 0, pitch = 0;
    int format = 0;           // VAW_FORMAT_*: NV12 = one plane of width x 3*height/2 bytes
                              // (opencv/FrameSourceFfmpegOpenCl.cpp:58,75-85)
    int device = 0;
    long index = -1;          // position in the stream, for bookkeeping
    size_t bytes = 0;
    ~DeviceFrame();           // returns the buffer to the C-ABI allocator
};

using Frame = std::shared_ptr<DeviceFrame>;  // stands in for cv::UMat (ref-counted, returned by value)

Frame make_device_frame(int device, int format, int width, int height);  // throws int on failure

/**
 * A source of frames
 */
class FrameSource {
  public:
    /**
     * Return the next frame and advance the current position
     * Raises an exception if no frames are ready
     */
    virtual Frame pull_frame() = 0;

    /**
     * Return the next frame but do not advance the current position
     * Raises an exception if no frames are ready
     */
    virtual Frame peek_frame() = 0;

    virtual ~FrameSource() = default;
};

#endif  // VAW_FRAME_SOURCE_HPP_
