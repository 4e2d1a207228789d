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
