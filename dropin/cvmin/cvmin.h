// cvmin.h - a minimal stand-in for the part of OpenCV's C++ API that the ORB front-end of the reference touches.
//
// This image has no OpenCV headers or libraries.  Two things are built against this header instead:
//   * dropin/*.cc (compile check and GPU tests of the drop-in classes; in the reference tree the real OpenCV is used);
//   * oracle/_ref: the reference's OWN sources (R/src/ORBextractor.cc, R/src/ORBmatcher.cc, R/src/CameraModels/Pinhole.cpp,
//     R/Thirdparty/DBoW2/DBoW2/*.cpp, function-level extracts of Frame.cc / KeyFrame.cc / MapPoint.cc), compiled UNMODIFIED
//     from /root/reference by oracle/ref/Makefile, as the parity oracle of record (test infrastructure).
// The types are layout-compatible where the reference relies on it (cv::KeyPoint = 28 bytes, cv::Point = 2 ints).  The image
// primitives (resize, copyMakeBorder, GaussianBlur, FAST, fastAtan2, undistortPoints) are only DECLARED here; their definitions
// (oracle/ref/cvmin_impl.cc) call the cv2-4.13-pinned C routines of oracle/orb_oracle.c, so they exist in the oracle builds only.
// Small dense algebra on CV_32F / CV_64F matrices (pose arithmetic of the matcher) is defined inline: products and dot
// products accumulate in double as OpenCV's generic gemm / dotProd do; OpenCV's fused small-matrix fast paths are not
// reproduced (host-side arithmetic, identical on both sides of every parity test).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cfloat>
#include <climits>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <numeric>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#define CVMIN 1
#define CV_MAJOR_VERSION 4
#define CV_MINOR_VERSION 13
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 CV_8U
#define CV_32SC1 CV_32S
#define CV_32FC1 CV_32F
#define CV_64FC1 CV_64F
#define CV_PI 3.1415926535897932384626433832795

typedef unsigned char uchar;
typedef unsigned short ushort;

// cvRound = round half to even (SSE cvtsd2si / cvtss2si under the default rounding mode)
static inline int cvRound(double v) { return (int)lrint(v); }
static inline int cvRound(float v) { return (int)lrintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(int v) { return v; }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(int v) { return v; }

namespace cv {

using ::uchar;
using std::string;
typedef std::string String;

[[noreturn]] static inline void cvmin_fail(const char* what)
{
    fprintf(stderr, "cvmin: unsupported use: %s\n", what);
    abort();
}

template <typename T> static inline T saturate_cast(double v) { return (T)v; }
template <> inline uchar saturate_cast<uchar>(double v) { int i = cvRound(v); return (uchar)(i < 0 ? 0 : i > 255 ? 255 : i); }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }

// ---- points, sizes, rectangles ----
template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
    template <typename U> Point_(const Point_<U>& p) : x((T)p.x), y((T)p.y) {}
    T dot(const Point_& o) const { return x * o.x + y * o.y; }
};
template <typename T> static inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> static inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <typename T> static inline Point_<T>& operator+=(Point_<T>& a, const Point_<T>& b) { a.x += b.x; a.y += b.y; return a; }
template <typename T> static inline Point_<T>& operator-=(Point_<T>& a, const Point_<T>& b) { a.x -= b.x; a.y -= b.y; return a; }
template <typename T> static inline Point_<T>& operator*=(Point_<T>& a, float b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
template <typename T> static inline Point_<T>& operator*=(Point_<T>& a, double b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
template <typename T> static inline Point_<T>& operator*=(Point_<T>& a, int b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
template <typename T> static inline Point_<T> operator*(const Point_<T>& a, float b) { return Point_<T>((T)(a.x * b), (T)(a.y * b)); }
template <typename T> static inline bool operator==(const Point_<T>& a, const Point_<T>& b) { return a.x == b.x && a.y == b.y; }
template <typename T> static inline bool operator!=(const Point_<T>& a, const Point_<T>& b) { return !(a == b); }
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <typename T> static inline double norm(const Point_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }

template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
    template <typename U> Point3_(const Point3_<U>& p) : x((T)p.x), y((T)p.y), z((T)p.z) {}
};
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;
typedef Point3_<int> Point3i;

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    T area() const { return width * height; }
};
typedef Size_<int> Size;
typedef Size_<int> Size2i;
static inline bool operator==(const Size& a, const Size& b) { return a.width == b.width && a.height == b.height; }
static inline bool operator!=(const Size& a, const Size& b) { return !(a == b); }

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T a, T b, T w, T h) : x(a), y(b), width(w), height(h) {}
};
typedef Rect_<int> Rect;

struct Range {
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
    static Range all() { return Range(INT32_MIN, INT32_MAX); }
};

template <typename T> struct Scalar_ {
    T val[4];
    Scalar_() { val[0] = val[1] = val[2] = val[3] = 0; }
    Scalar_(T a, T b = 0, T c = 0, T d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    T operator[](int i) const { return val[i]; }
};
typedef Scalar_<double> Scalar;

// ---- cv::KeyPoint (28 bytes, the layout the C ABI's orbx_keypoint mirrors) ----
class KeyPoint {
public:
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f p, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(p), size(s), angle(a), response(r), octave(o), class_id(c) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");
static_assert(sizeof(Point) == 8, "cv::Point layout");

struct KeyPointsFilter {
    static void retainBest(std::vector<KeyPoint>& keypoints, int npoints);       // oracle/ref/cvmin_impl.cc (dead code path of the reference)
};

// ---- cv::Mat: 2-D, single channel, reference-counted storage, row/column views ----
template <typename T> struct DataType;
template <> struct DataType<uchar> { enum { type = CV_8U }; };
template <> struct DataType<signed char> { enum { type = CV_8S }; };
template <> struct DataType<unsigned short> { enum { type = CV_16U }; };
template <> struct DataType<short> { enum { type = CV_16S }; };
template <> struct DataType<int> { enum { type = CV_32S }; };
template <> struct DataType<float> { enum { type = CV_32F }; };
template <> struct DataType<double> { enum { type = CV_64F }; };

class Mat;
template <typename T> class Mat_;
template <typename T> class MatCommaInitializer_;

class Mat {
public:
    enum { AUTO_STEP = 0 };
    int flags;            // element type (CV_8U .. CV_64F)
    int rows, cols;
    uchar* data;
    size_t step;          // bytes per row

    Mat() : flags(CV_8U), rows(0), cols(0), data(nullptr), step(0) {}
    Mat(int r, int c, int type) : flags(CV_8U), rows(0), cols(0), data(nullptr), step(0) { create(r, c, type); }
    Mat(Size s, int type) : flags(CV_8U), rows(0), cols(0), data(nullptr), step(0) { create(s.height, s.width, type); }
    Mat(int r, int c, int type, const Scalar& s) : flags(CV_8U), rows(0), cols(0), data(nullptr), step(0) { create(r, c, type); setTo(s.val[0]); }
    Mat(int r, int c, int type, void* p, size_t st = AUTO_STEP) : flags(type), rows(r), cols(c), data((uchar*)p), step(st ? st : (size_t)c * esz(type)) {}
    Mat(const Mat& m, const Rect& roi) : flags(m.flags), rows(roi.height), cols(roi.width),
        data(m.data + (size_t)roi.y * m.step + (size_t)roi.x * esz(m.flags)), step(m.step), hold_(m.hold_)
    {
        if (roi.x < 0 || roi.y < 0 || roi.x + roi.width > m.cols || roi.y + roi.height > m.rows) cvmin_fail("Mat ROI out of range");
    }
    template <typename T> explicit Mat(const std::vector<T>& v) : flags(DataType<T>::type), rows((int)v.size()), cols(1), data((uchar*)v.data()), step(sizeof(T)) {}
    template <typename T> explicit Mat(const Point3_<T>& p) : flags(CV_8U), rows(0), cols(0), data(nullptr), step(0)
    {
        create(3, 1, DataType<T>::type); at<T>(0) = p.x; at<T>(1) = p.y; at<T>(2) = p.z;
    }

    static size_t esz(int type) { static const size_t s[7] = {1, 1, 2, 2, 4, 4, 8}; return s[type & 7]; }
    void create(int r, int c, int type)
    {
        type &= 7;
        if (data && r == rows && c == cols && type == flags) return;
        flags = type; rows = r; cols = c; step = (size_t)c * esz(type);
        const size_t bytes = step * (size_t)r;
        if (bytes == 0) { hold_.reset(); data = nullptr; return; }
        hold_.reset((uchar*)malloc(bytes + 64), free);       // malloc, not operator new: see oracle/ref/ref_alloc.cc
        data = hold_.get();
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    void release() { hold_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return flags; }
    int depth() const { return flags; }
    int channels() const { return 1; }
    size_t elemSize() const { return esz(flags); }
    size_t elemSize1() const { return esz(flags); }
    size_t step1() const { return step / esz(flags); }
    size_t total() const { return (size_t)rows * cols; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return rows <= 1 || step == (size_t)cols * esz(flags); }

    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    template <typename T> T& at(int r, int c) { return reinterpret_cast<T*>(data + (size_t)r * step)[c]; }
    template <typename T> const T& at(int r, int c) const { return reinterpret_cast<const T*>(data + (size_t)r * step)[c]; }
    template <typename T> T& at(int i) { return const_cast<T&>(static_cast<const Mat*>(this)->at<T>(i)); }
    template <typename T> const T& at(int i) const
    {
        if (rows == 1 || isContinuous()) return reinterpret_cast<const T*>(data)[i];
        if (cols == 1) return *reinterpret_cast<const T*>(data + (size_t)i * step);
        const int r = i / cols; return reinterpret_cast<const T*>(data + (size_t)r * step)[i - r * cols];
    }
    template <typename T> T& at(Point p) { return at<T>(p.y, p.x); }
    template <typename T> const T& at(Point p) const { return at<T>(p.y, p.x); }

    Mat operator()(const Rect& roi) const { return Mat(*this, roi); }
    Mat operator()(Range rr, Range cr) const
    {
        const int r0 = rr.start == INT32_MIN ? 0 : rr.start, r1 = rr.end == INT32_MAX ? rows : rr.end;
        const int c0 = cr.start == INT32_MIN ? 0 : cr.start, c1 = cr.end == INT32_MAX ? cols : cr.end;
        return Mat(*this, Rect(c0, r0, c1 - c0, r1 - r0));
    }
    Mat row(int r) const { return Mat(*this, Rect(0, r, cols, 1)); }
    Mat col(int c) const { return Mat(*this, Rect(c, 0, 1, rows)); }
    Mat rowRange(int a, int b) const { return Mat(*this, Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return Mat(*this, Rect(a, 0, b - a, rows)); }
    Mat rowRange(const Range& r) const { return rowRange(r.start, r.end); }
    Mat colRange(const Range& r) const { return colRange(r.start, r.end); }

    // channels do not exist here: an N x 2 CV_32F matrix stands for both its 1- and its 2-channel view (Frame::UndistortKeyPoints
    // reshapes between them around cv::undistortPoints, which takes the N x 2 layout directly in this stand-in)
    Mat reshape(int /*cn*/, int /*rows*/ = 0) const { return *this; }
    Mat clone() const { Mat m; copyTo(m); return m; }
    void copyTo(Mat& dst) const
    {
        if (dst.data == data && dst.rows == rows && dst.cols == cols && dst.step == step) return;
        dst.create(rows, cols, flags);
        const size_t rb = (size_t)cols * esz(flags);
        for (int r = 0; r < rows; r++) memmove(dst.data + (size_t)r * dst.step, data + (size_t)r * step, rb);
    }
    inline void copyTo(const class _OutputArray& dst) const;
    void convertTo(Mat& dst, int type, double alpha = 1, double beta = 0) const
    {
        Mat out(rows, cols, type);
        for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) out.setd(r, c, getd(r, c) * alpha + beta);
        dst = out;
    }
    Mat& setTo(double v) { for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) setd(r, c, v); return *this; }
    Mat& operator=(const Scalar& s) { return setTo(s.val[0]); }

    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); for (int i = 0; i < r; i++) memset(m.ptr(i), 0, (size_t)c * esz(type & 7)); return m; }
    static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }
    static Mat ones(int r, int c, int type) { Mat m(r, c, type); m.setTo(1); return m; }
    static Mat eye(int r, int c, int type) { Mat m = zeros(r, c, type); for (int i = 0; i < r && i < c; i++) m.setd(i, i, 1); return m; }

    // element access as double (CV_8U, CV_32S, CV_32F, CV_64F)
    double getd(int r, int c) const
    {
        switch (flags) {
        case CV_8U: return at<uchar>(r, c);
        case CV_8S: return at<signed char>(r, c);
        case CV_16U: return at<unsigned short>(r, c);
        case CV_16S: return at<short>(r, c);
        case CV_32S: return at<int>(r, c);
        case CV_32F: return at<float>(r, c);
        case CV_64F: return at<double>(r, c);
        default: cvmin_fail("Mat element type");
        }
    }
    void setd(int r, int c, double v)
    {
        switch (flags) {
        case CV_8U: at<uchar>(r, c) = saturate_cast<uchar>(v); break;
        case CV_8S: { const int i = cvRound(v); at<signed char>(r, c) = (signed char)(i < -128 ? -128 : i > 127 ? 127 : i); break; }
        case CV_16U: { const int i = cvRound(v); at<unsigned short>(r, c) = (unsigned short)(i < 0 ? 0 : i > 65535 ? 65535 : i); break; }
        case CV_16S: { const int i = cvRound(v); at<short>(r, c) = (short)(i < -32768 ? -32768 : i > 32767 ? 32767 : i); break; }
        case CV_32S: at<int>(r, c) = cvRound(v); break;
        case CV_32F: at<float>(r, c) = (float)v; break;
        case CV_64F: at<double>(r, c) = v; break;
        default: cvmin_fail("Mat element type");
        }
    }

    Mat t() const
    {
        Mat m(cols, rows, flags);
        const size_t e = esz(flags);
        for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) memcpy(m.data + (size_t)c * m.step + (size_t)r * e, data + (size_t)r * step + (size_t)c * e, e);
        return m;
    }
    double dot(const Mat& o) const
    {
        if (total() != o.total() || flags != o.flags) cvmin_fail("Mat::dot shapes");
        double s = 0;
        const int n = (int)total();
        // both operands in row-major element order
        for (int i = 0; i < n; i++) s += elem_linear(i) * o.elem_linear(i);
        return s;
    }
    Mat mul(const Mat& o, double scale = 1) const
    {
        Mat m(rows, cols, flags);
        for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) m.setd(r, c, getd(r, c) * o.getd(r, c) * scale);
        return m;
    }
    Mat cross(const Mat& o) const
    {
        Mat m(rows, cols, flags);
        const double a0 = elem_linear(0), a1 = elem_linear(1), a2 = elem_linear(2), b0 = o.elem_linear(0), b1 = o.elem_linear(1), b2 = o.elem_linear(2);
        const double v[3] = {a1 * b2 - a2 * b1, a2 * b0 - a0 * b2, a0 * b1 - a1 * b0};
        for (int i = 0; i < 3; i++) { if (cols == 1) m.setd(i, 0, v[i]); else m.setd(0, i, v[i]); }
        return m;
    }
    // inverse of a 2x2 / 3x3 matrix by cofactors with a double determinant (OpenCV's closed form for n <= 3, DECOMP_LU)
    Mat inv() const
    {
        if (rows != cols || (rows != 2 && rows != 3)) cvmin_fail("Mat::inv supports 2x2 and 3x3");
        Mat m = zeros(rows, cols, flags);
        if (rows == 2) {
            const double a = getd(0, 0), b = getd(0, 1), c = getd(1, 0), d = getd(1, 1);
            double det = a * d - b * c;
            if (det != 0) { det = 1. / det; m.setd(0, 0, d * det); m.setd(0, 1, -b * det); m.setd(1, 0, -c * det); m.setd(1, 1, a * det); }
            return m;
        }
        double s[3][3];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) s[i][j] = getd(i, j);
        double det = s[0][0] * (s[1][1] * s[2][2] - s[1][2] * s[2][1]) - s[0][1] * (s[1][0] * s[2][2] - s[1][2] * s[2][0]) +
                     s[0][2] * (s[1][0] * s[2][1] - s[1][1] * s[2][0]);
        if (det != 0) {
            det = 1. / det;
            m.setd(0, 0, (s[1][1] * s[2][2] - s[1][2] * s[2][1]) * det);
            m.setd(0, 1, (s[0][2] * s[2][1] - s[0][1] * s[2][2]) * det);
            m.setd(0, 2, (s[0][1] * s[1][2] - s[0][2] * s[1][1]) * det);
            m.setd(1, 0, (s[1][2] * s[2][0] - s[1][0] * s[2][2]) * det);
            m.setd(1, 1, (s[0][0] * s[2][2] - s[0][2] * s[2][0]) * det);
            m.setd(1, 2, (s[0][2] * s[1][0] - s[0][0] * s[1][2]) * det);
            m.setd(2, 0, (s[1][0] * s[2][1] - s[1][1] * s[2][0]) * det);
            m.setd(2, 1, (s[0][1] * s[2][0] - s[0][0] * s[2][1]) * det);
            m.setd(2, 2, (s[0][0] * s[1][1] - s[0][1] * s[1][0]) * det);
        }
        return m;
    }
    double elem_linear(int i) const { const int r = cols ? i / cols : 0; return getd(r, i - r * cols); }

protected:
    std::shared_ptr<uchar> hold_;
};

static inline void cvmin_same_shape(const Mat& a, const Mat& b)
{
    if (a.rows != b.rows || a.cols != b.cols || a.flags != b.flags) cvmin_fail("matrix operands of different shape or type");
}
static inline Mat operator+(const Mat& a, const Mat& b)
{
    cvmin_same_shape(a, b);
    Mat m(a.rows, a.cols, a.flags);
    if (a.flags == CV_32F) { for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) m.at<float>(r, c) = a.at<float>(r, c) + b.at<float>(r, c); }
    else for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) m.setd(r, c, a.getd(r, c) + b.getd(r, c));
    return m;
}
static inline Mat operator-(const Mat& a, const Mat& b)
{
    cvmin_same_shape(a, b);
    Mat m(a.rows, a.cols, a.flags);
    if (a.flags == CV_32F) { for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) m.at<float>(r, c) = a.at<float>(r, c) - b.at<float>(r, c); }
    else for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) m.setd(r, c, a.getd(r, c) - b.getd(r, c));
    return m;
}
static inline Mat operator*(const Mat& a, double s)
{
    Mat m(a.rows, a.cols, a.flags);
    // cv::Mat * scalar = convertTo with a scale: in float for CV_32F (cvtScale 32f -> 32f), in double otherwise
    if (a.flags == CV_32F) { const float f = (float)s; for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) m.at<float>(r, c) = a.at<float>(r, c) * f; }
    else for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) m.setd(r, c, a.getd(r, c) * s);
    return m;
}
static inline Mat operator*(double s, const Mat& a) { return a * s; }
static inline Mat operator/(const Mat& a, double s) { return a * (1. / s); }
static inline Mat operator-(const Mat& a) { return a * -1.0; }
static inline Mat operator*(const Mat& a, const Mat& b)
{
    if (a.cols != b.rows || a.flags != b.flags || (a.flags != CV_32F && a.flags != CV_64F)) cvmin_fail("matrix product operands");
    Mat m(a.rows, b.cols, a.flags);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < b.cols; j++) {
            double s = 0;
            for (int k = 0; k < a.cols; k++) s += a.getd(i, k) * b.getd(k, j);
            m.setd(i, j, s);
        }
    return m;
}
static inline Mat operator+(const Mat& a, double s) { Mat m(a.rows, a.cols, a.flags); for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) m.setd(r, c, a.getd(r, c) + s); return m; }
static inline Mat operator-(const Mat& a, double s) { return a + (-s); }
static inline void hconcat(const Mat& a, const Mat& b, Mat& dst)
{
    if (a.rows != b.rows || a.flags != b.flags) cvmin_fail("hconcat operands");
    Mat m(a.rows, a.cols + b.cols, a.flags);
    const size_t e = Mat::esz(a.flags);
    for (int r = 0; r < a.rows; r++) {
        memcpy(m.ptr(r), a.ptr(r), (size_t)a.cols * e);
        memcpy(m.ptr(r) + (size_t)a.cols * e, b.ptr(r), (size_t)b.cols * e);
    }
    dst = m;
}
static inline Mat& operator+=(Mat& a, const Mat& b) { a = a + b; return a; }
static inline Mat& operator-=(Mat& a, const Mat& b) { a = a - b; return a; }
static inline Mat& operator*=(Mat& a, double s) { a = a * s; return a; }
static inline Mat& operator/=(Mat& a, double s) { a = a / s; return a; }

enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6 };
static inline double norm(const Mat& a, int type = NORM_L2)
{
    double s = 0;
    const int n = (int)a.total();
    if (type == NORM_L2) { for (int i = 0; i < n; i++) { const double v = a.elem_linear(i); s += v * v; } return std::sqrt(s); }
    if (type == NORM_L1) { for (int i = 0; i < n; i++) s += std::fabs(a.elem_linear(i)); return s; }
    if (type == NORM_INF) { for (int i = 0; i < n; i++) s = std::max(s, std::fabs(a.elem_linear(i))); return s; }
    cvmin_fail("norm type");
}
static inline double norm(const Mat& a, const Mat& b, int type = NORM_L2)
{
    cvmin_same_shape(a, b);
    double s = 0;
    const int n = (int)a.total();
    if (type == NORM_L1) { for (int i = 0; i < n; i++) s += std::fabs(a.elem_linear(i) - b.elem_linear(i)); return s; }
    if (type == NORM_L2) { for (int i = 0; i < n; i++) { const double v = a.elem_linear(i) - b.elem_linear(i); s += v * v; } return std::sqrt(s); }
    cvmin_fail("norm type");
}
template <typename T> static inline double norm(const Point3_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y + (double)p.z * p.z); }

template <typename T> class Mat_ : public Mat {
public:
    Mat_() : Mat() { flags = DataType<T>::type; }
    Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
    Mat_(const Mat& m) : Mat(m) { if (m.flags != DataType<T>::type && !m.empty()) cvmin_fail("Mat_ type conversion"); }
    T& operator()(int r, int c) { return at<T>(r, c); }
    const T& operator()(int r, int c) const { return at<T>(r, c); }
    T& operator()(int i) { return at<T>(i); }
    const T& operator()(int i) const { return at<T>(i); }
};
template <typename T> class MatCommaInitializer_ {
public:
    MatCommaInitializer_(const Mat_<T>& m) : m_(m), i_(0) {}
    template <typename U> MatCommaInitializer_& operator,(U v)
    {
        if (i_ >= (int)m_.total()) cvmin_fail("comma initialiser overflow");
        m_.template at<T>(i_ / m_.cols, i_ % m_.cols) = (T)v; i_++; return *this;
    }
    operator Mat_<T>() const { return m_; }
    operator Mat() const { return m_; }
    Mat_<T> m_; int i_;
};
template <typename T, typename U> static inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, U v)
{
    MatCommaInitializer_<T> ci(m);
    return (ci, v);
}
typedef Mat_<float> Mat1f;
typedef Mat_<double> Mat1d;

// ---- argument proxies ----
class _InputArray {
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat& m) : m_(&m) {}
    Mat getMat() const { return m_ ? *m_ : Mat(); }
    bool empty() const { return !m_ || m_->empty(); }
private:
    const Mat* m_;
};
class _OutputArray {
public:
    _OutputArray() : m_(nullptr) {}
    _OutputArray(Mat& m) : m_(&m) {}
    _OutputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}          // fixed-size destination (e.g. a row view)
    void create(int r, int c, int type) const { if (!m_) cvmin_fail("OutputArray without matrix"); m_->create(r, c, type); }
    void create(Size s, int type) const { create(s.height, s.width, type); }
    void release() const { if (m_) m_->release(); }
    Mat getMat() const { return m_ ? *m_ : Mat(); }
    Mat& getMatRef() const { return *m_; }
    bool needed() const { return m_ != nullptr; }
private:
    Mat* m_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& InputOutputArray;
static inline InputArray noArray() { static _InputArray none; return none; }

inline void Mat::copyTo(const _OutputArray& dst) const
{
    Mat& d = dst.getMatRef();
    if (d.data && d.rows == rows && d.cols == cols && d.flags == flags) {          // copy into an existing (view) matrix
        const size_t rb = (size_t)cols * esz(flags);
        for (int r = 0; r < rows; r++) memmove(d.data + (size_t)r * d.step, data + (size_t)r * step, rb);
    } else copyTo(d);
}

// ---- image primitives: declared here, defined in oracle/ref/cvmin_impl.cc on top of oracle/orb_oracle.c ----
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType, const Scalar& value = Scalar());
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
float fastAtan2(float y, float x);
void undistortPoints(InputArray src, OutputArray dst, InputArray cameraMatrix, InputArray distCoeffs, InputArray R = noArray(), InputArray P = noArray());

struct DMatch {
    int queryIdx, trainIdx, imgIdx; float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.4e38f) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
};
class BFMatcher {
public:
    BFMatcher(int normType = NORM_L2, bool crossCheck = false) : norm_(normType) { (void)crossCheck; }
    void knnMatch(InputArray q, InputArray t, std::vector<std::vector<DMatch> >& matches, int k) const;   // cvmin_impl.cc
private:
    int norm_;
};

// ---- persistence: only so that DBoW2's YAML save / load members compile; never executed ----
class FileNode {
public:
    FileNode operator[](const char*) const { cvmin_fail("cv::FileStorage"); }
    FileNode operator[](const std::string&) const { cvmin_fail("cv::FileStorage"); }
    FileNode operator[](int) const { cvmin_fail("cv::FileStorage"); }
    size_t size() const { return 0; }
    operator int() const { cvmin_fail("cv::FileStorage"); }
    operator float() const { cvmin_fail("cv::FileStorage"); }
    operator double() const { cvmin_fail("cv::FileStorage"); }
    operator std::string() const { cvmin_fail("cv::FileStorage"); }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string&, int) { cvmin_fail("cv::FileStorage"); }
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char*) const { cvmin_fail("cv::FileStorage"); }
    FileNode operator[](const std::string&) const { cvmin_fail("cv::FileStorage"); }
};
template <typename T> static inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }

}  // namespace cv
