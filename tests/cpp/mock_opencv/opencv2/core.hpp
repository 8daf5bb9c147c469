// MOCK of <opencv2/core.hpp> -- TEST INFRASTRUCTURE ONLY.
// This image holds no OpenCV C++ headers, so the cv::Mat overload of include/seamless_clone.hpp
// (enabled with -DSCB_WITH_OPENCV) could never be compiled here.  This header declares just the slice of cv::Mat / cv::Point /
// cv::Size the overload touches, with OpenCV's names, signatures and type codes, so that the overload is at least compiled,
// linked and run (tests/test_cpp_api.py).  It is NOT OpenCV and holds no OpenCV code.
#pragma once
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#define CV_8U 0
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) (((depth) & 7) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_Assert(expr) \
    do {                \
        if (!(expr)) throw std::runtime_error("CV_Assert failed: " #expr); \
    } while (0)

namespace cv {
struct Point {
    int x = 0, y = 0;
    Point() = default;
    Point(int x_, int y_) : x(x_), y(y_) {}
};
struct Size {
    int width = 0, height = 0;
    Size() = default;
    Size(int w, int h) : width(w), height(h) {}
};
class Mat {
  public:
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;
    size_t step = 0;
    Mat() = default;
    Mat(int r, int c, int type, void* d, size_t s = 0) : rows(r), cols(c), data((unsigned char*)d), type_(type) { step = s ? s : (size_t)c * channels(); }
    Mat(const Mat&) = delete;
    Mat& operator=(const Mat&) = delete;
    ~Mat() {
        if (owned_) std::free(data);
    }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    Size size() const { return Size(cols, rows); }
    void create(Size sz, int type) {
        if (owned_) std::free(data);
        rows = sz.height;
        cols = sz.width;
        type_ = type;
        step = (size_t)cols * channels();
        data = (unsigned char*)std::malloc(step * (size_t)rows);
        owned_ = true;
    }

  private:
    int type_ = CV_8UC1;
    bool owned_ = false;
};
}  // namespace cv
