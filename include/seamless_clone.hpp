// seamless_clone.hpp -- C++ host API over the C ABI (scb.h): OpenCV's signature
//
//     void seamlessClone(InputArray src, InputArray dst, InputArray mask, Point p, OutputArray blend, int flags);
//
// kept as   scb::seamlessClone(const Mat& src, const Mat& dst, const Mat& mask, Point p, Mat& blend, int flags)
// so that a call site switches by changing the namespace.  When <opencv2/core.hpp> is available define
// SCB_WITH_OPENCV before including this header and the overload taking cv::Mat is enabled too; the
// authoring image has no OpenCV C++ headers, so a minimal Mat / Point pair with the same member names
// (rows, cols, data, step, channels(), create(), empty()) is provided.
//
// Replaces, in the reference: SeamlessClone::seamlessCloneGPU(Mat dst, Mat patch, Mat mask, Point, Mat& blend, int)
// (/root/reference/seamlessClone-CUDA/seamlessClone_imp.cpp:430-486) -- note the reference's argument order
// (dst first) and its aliasing of blend onto the caller's dst; this API follows OpenCV on both counts.
#pragma once

#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "scb.h"

namespace scb {

enum {
    NORMAL_CLONE = SCB_NORMAL_CLONE, MIXED_CLONE = SCB_MIXED_CLONE, MONOCHROME_TRANSFER = SCB_MONOCHROME_TRANSFER,
    NORMAL_CLONE_WIDE = SCB_NORMAL_CLONE_WIDE, MIXED_CLONE_WIDE = SCB_MIXED_CLONE_WIDE, MONOCHROME_TRANSFER_WIDE = SCB_MONOCHROME_TRANSFER_WIDE
};

struct Point {
    int x = 0, y = 0;
    Point() = default;
    Point(int x_, int y_) : x(x_), y(y_) {}
};

// 8-bit image, interleaved channels, optionally owning its pixels (cv::Mat subset)
class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;
    size_t step = 0;

    Mat() = default;
    Mat(int r, int c, int ch) { create(r, c, ch); }
    Mat(int r, int c, int ch, void* external, size_t step_bytes = 0) : rows(r), cols(c), data((unsigned char*)external), step(step_bytes ? step_bytes : (size_t)c * ch), ch_(ch) {}
    void create(int r, int c, int ch) {
        if (r == rows && c == cols && ch == ch_ && data) return;  // like cv::Mat::create: keep a buffer of the right shape (owned or external)
        own_ = std::shared_ptr<unsigned char>(new unsigned char[(size_t)r * c * ch], std::default_delete<unsigned char[]>());
        data = own_.get();
        rows = r;
        cols = c;
        ch_ = ch;
        step = (size_t)c * ch;
    }
    int channels() const { return ch_; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    Mat clone() const {
        Mat m(rows, cols, ch_);
        for (int y = 0; y < rows; ++y) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * ch_);
        return m;
    }
    scb_image view() const { return scb_image{data, rows, cols, ch_, (int64_t)step}; }

private:
    int ch_ = 0;
    std::shared_ptr<unsigned char> own_;
};

// cv::Exception-shaped error: what() reads like OpenCV's "(-215:Assertion failed) ..." lines
class Exception : public std::runtime_error {
public:
    int code;
    Exception(int c, const std::string& msg) : std::runtime_error("scb::seamlessClone error " + std::to_string(c) + ": " + msg), code(c) {}
};

class Context {
public:
    explicit Context(int device = 0, void* stream = nullptr) {
        int rc = scb_create(device, stream, &ctx_);
        if (rc != SCB_OK) throw Exception(rc, scb_last_error(nullptr));
    }
    ~Context() { scb_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    scb_context* handle() const { return ctx_; }
    void sync() { check(scb_sync(ctx_)); }
    void check(int rc) const {
        if (rc != SCB_OK) throw Exception(rc, scb_last_error(ctx_));
    }

    // OpenCV semantics on host images: blend is (re)allocated to dst's size, dst and mask are never written.
    void seamlessClone(const Mat& src, const Mat& dst, const Mat& mask, Point p, Mat& blend, int flags = NORMAL_CLONE) {
        if (src.empty() || dst.empty()) throw Exception(SCB_ERR_INVALID_ARGUMENT, "src/dst empty");
        Mat src3 = src, grey;
        if (src.channels() == 1) {  // OpenCV replicates a grey source
            src3.create(src.rows, src.cols, 3);
            for (int y = 0; y < src.rows; ++y)
                for (int x = 0; x < src.cols; ++x) {
                    unsigned char v = src.data[(size_t)y * src.step + x];
                    unsigned char* o = src3.data + (size_t)y * src3.step + 3 * x;
                    o[0] = o[1] = o[2] = v;
                }
        }
        if (mask.empty()) {  // "no mask" == all 255
            grey.create(src.rows, src.cols, 1);
            std::memset(grey.data, 255, (size_t)src.rows * src.cols);
        } else if (mask.channels() == 1) {
            grey = mask;
        } else {  // cvtColor BGR2GRAY, OpenCV 4.x fixed point
            grey.create(mask.rows, mask.cols, 1);
            const int ch = mask.channels();
            for (int y = 0; y < mask.rows; ++y)
                for (int x = 0; x < mask.cols; ++x) {
                    const unsigned char* m = mask.data + (size_t)y * mask.step + (size_t)ch * x;
                    grey.data[(size_t)y * grey.step + x] = (unsigned char)((m[0] * 3735u + m[1] * 19235u + m[2] * 9798u + 16384u) >> 15);
                }
        }
        blend.create(dst.rows, dst.cols, 3);
        scb_image vs = src3.view(), vd = dst.view(), vm = grey.view(), vb = blend.view();
        check(scb_seamless_clone(ctx_, &vs, &vd, &vm, p.x, p.y, &vb, flags, SCB_MEM_HOST));
    }

private:
    scb_context* ctx_ = nullptr;
};

// the free function, OpenCV's signature; one lazily created context per device
inline void seamlessClone(const Mat& src, const Mat& dst, const Mat& mask, Point p, Mat& blend, int flags = NORMAL_CLONE, int device = 0) {
    static std::vector<std::unique_ptr<Context>> ctxs;
    if ((int)ctxs.size() <= device) ctxs.resize(device + 1);
    if (!ctxs[device]) ctxs[device].reset(new Context(device));
    ctxs[device]->seamlessClone(src, dst, mask, p, blend, flags);
}

#ifdef SCB_WITH_OPENCV
}  // namespace scb
#include <opencv2/core.hpp>
namespace scb {
inline void seamlessClone(const cv::Mat& src, const cv::Mat& dst, const cv::Mat& mask, cv::Point p, cv::Mat& blend, int flags = NORMAL_CLONE, int device = 0) {
    CV_Assert(src.depth() == CV_8U && dst.type() == CV_8UC3);
    Mat s(src.rows, src.cols, src.channels(), src.data, src.step), d(dst.rows, dst.cols, 3, dst.data, dst.step);
    Mat m = mask.empty() ? Mat() : Mat(mask.rows, mask.cols, mask.channels(), mask.data, mask.step);
    blend.create(dst.size(), CV_8UC3);
    Mat b(blend.rows, blend.cols, 3, blend.data, blend.step);
    seamlessClone(s, d, m, Point(p.x, p.y), b, flags, device);
}
#endif

}  // namespace scb
