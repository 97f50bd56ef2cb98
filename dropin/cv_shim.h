// cv_shim.h - the small subset of OpenCV that the two drop-in classes touch, for compile-checking them in an
// image that has no OpenCV headers.  In the reference tree (catkin package orb_slam3_ros) the real
// <opencv2/opencv.hpp> is used instead: build with -DORBX_USE_REAL_OPENCV.
#pragma once
#ifdef ORBX_USE_REAL_OPENCV
#include <opencv2/core/core.hpp>
#else
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
#define CV_8U 0
#define CV_8UC1 0
namespace cv {
struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float a, float b) : x(a), y(b) {} };
struct KeyPoint {            // same 28-byte layout as cv::KeyPoint
    Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1;
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");
class Mat {
public:
    int rows = 0, cols = 0; size_t step = 0; unsigned char* data = nullptr;
    Mat() {}
    Mat(int r, int c, int /*type*/) { create(r, c, CV_8U); }
    Mat(int r, int c, int /*type*/, void* p, size_t s = 0) : rows(r), cols(c), step(s ? s : (size_t)c), data((unsigned char*)p) {}
    void create(int r, int c, int /*type*/) {
        if (r == rows && c == cols && own_) return;
        rows = r; cols = c; step = (size_t)c; own_.reset(new unsigned char[(size_t)r * c + 1]); data = own_.get();
    }
    void release() { own_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    bool isContinuous() const { return step == (size_t)cols; }
    unsigned char* ptr(int r = 0) { return data + (size_t)r * step; }
    const unsigned char* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    Mat row(int r) const { return Mat(1, cols, CV_8U, data + (size_t)r * step, step); }
    Mat getMat() const { return *this; }
private:
    std::shared_ptr<unsigned char[]> own_;
};
typedef const Mat& InputArray;
typedef Mat& OutputArray;
}  // namespace cv
#endif
