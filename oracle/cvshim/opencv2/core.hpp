// cvshim/opencv2/core.hpp - TEST INFRASTRUCTURE ONLY.
// A minimal stand-in for the handful of OpenCV core types the reference's own helper sources use
// (/root/reference/image_stitching/{quaternion.h, euler.h, serializer.cpp, cropper.cpp}), so that those files can be
// compiled UNMODIFIED from where they lie into oracle/_ref/libisb_ref.so (recipe: oracle/Makefile, target _ref).
// OpenCV's C++ headers are not in this image; only what those four files touch is modelled:
//   cv::Mat (refcounted, 2-D, row-major, types 8U/8S/16S/32F/64F x 1..4 channels), Mat_<T>, Vec<T,n>, Size, Point, Rect,
//   Scalar, Mat::eye/zeros, at<T>(r,c), ROI operator(), convertTo (saturating), operator> (Mat, scalar).
// Nothing here is part of the product (libisb.so never includes it).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <string_view>
#include <vector>

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16SC3 CV_MAKETYPE(CV_16S, 3)

namespace cv {

template <typename T> struct Point_ {
    T x{}, y{};
    Point_() = default;
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
using Point = Point_<int>;

template <typename T> struct Size_ {
    T width{}, height{};
    Size_() = default;
    Size_(T w, T h) : width(w), height(h) {}
};
using Size = Size_<int>;

template <typename T> struct Rect_ {
    T x{}, y{}, width{}, height{};
    Rect_() = default;
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
using Rect = Rect_<int>;

template <typename T, int N> struct Vec {
    T val[N]{};
    Vec() = default;
    Vec(T a, T b, T c) { static_assert(N == 3, "3-element ctor"); val[0] = a; val[1] = b; val[2] = c; }
    Vec(T a, T b, T c, T d) { static_assert(N == 4, "4-element ctor"); val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
using Vec3d = Vec<double, 3>;
using Vec3f = Vec<float, 3>;
using Vec4i = Vec<int, 4>;

struct Scalar {
    double val[4]{};
    Scalar() = default;
    Scalar(double a, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    double operator[](int i) const { return val[i]; }
};

inline int cvshim_depth(int type) { return type & 7; }
inline int cvshim_channels(int type) { return (type >> CV_CN_SHIFT) + 1; }
inline size_t cvshim_elem1(int type)
{
    static const size_t s[7] = {1, 1, 2, 2, 4, 4, 8};
    return s[cvshim_depth(type)];
}

class Mat {
public:
    int rows = 0, cols = 0;
    Mat() = default;
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type) { create(s.height, s.width, type); }
    template <typename T, int N> explicit Mat(const Vec<T, N>& v)
    {   // cv::Mat(Vec) is an N x 1 single-channel matrix
        create(N, 1, sizeof(T) == 8 ? CV_64F : CV_32F);
        for (int i = 0; i < N; ++i) at<T>(i, 0) = v[i];
    }
    void create(int r, int c, int type)
    {
        rows = r; cols = c; type_ = type;
        step_ = (size_t)c * elemSize();
        buf_ = std::make_shared<std::vector<uint8_t>>(step_ * (size_t)r + 16, (uint8_t)0);
        data_ = buf_->data();
    }
    int type() const { return type_; }
    int channels() const { return cvshim_channels(type_); }
    size_t elemSize() const { return cvshim_elem1(type_) * cvshim_channels(type_); }
    size_t step() const { return step_; }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return rows == 0 || cols == 0; }
    uint8_t* ptr(int r = 0) { return data_ + (size_t)r * step_; }
    const uint8_t* ptr(int r = 0) const { return data_ + (size_t)r * step_; }
    template <typename T> T& at(int r, int c) { return *reinterpret_cast<T*>(data_ + (size_t)r * step_ + (size_t)c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const
    {
        return *reinterpret_cast<const T*>(data_ + (size_t)r * step_ + (size_t)c * sizeof(T));
    }
    static Mat zeros(Size s, int type) { return Mat(s, type); }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    static Mat eye(Size s, int type)
    {
        Mat m(s, type);
        for (int i = 0; i < std::min(m.rows, m.cols); ++i) {
            if (cvshim_depth(type) == CV_32F) m.at<float>(i, i) = 1.f;
            else if (cvshim_depth(type) == CV_64F) m.at<double>(i, i) = 1.0;
            else m.ptr(i)[(size_t)i * m.elemSize()] = 1;
        }
        return m;
    }
    Mat operator()(const Rect& r) const
    {   // ROI view sharing the buffer
        assert(r.x >= 0 && r.y >= 0 && r.width >= 0 && r.height >= 0 && r.x + r.width <= cols && r.y + r.height <= rows);
        Mat m = *this;
        m.data_ = data_ + (size_t)r.y * step_ + (size_t)r.x * elemSize();
        m.rows = r.height;
        m.cols = r.width;
        return m;
    }
    Mat clone() const
    {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; ++r) std::memcpy(m.ptr(r), ptr(r), (size_t)cols * elemSize());
        return m;
    }
    size_t view_offset() const { return buf_ ? (size_t)(data_ - buf_->data()) : 0; }
    // saturating depth conversion (only the pairs the reference's helpers hit: 16S/8U -> 8U, float <-> double)
    void convertTo(Mat& dst, int rtype) const
    {
        const int ddepth = cvshim_depth(rtype), cn = channels();
        Mat out(rows, cols, CV_MAKETYPE(ddepth, cn));
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols * cn; ++c) {
                double v = 0;
                switch (cvshim_depth(type_)) {
                case CV_8U: v = ptr(r)[c]; break;
                case CV_16S: v = reinterpret_cast<const int16_t*>(ptr(r))[c]; break;
                case CV_32F: v = reinterpret_cast<const float*>(ptr(r))[c]; break;
                case CV_64F: v = reinterpret_cast<const double*>(ptr(r))[c]; break;
                default: assert(false);
                }
                switch (ddepth) {
                case CV_8U: out.ptr(r)[c] = (uint8_t)std::min(255.0, std::max(0.0, std::nearbyint(v))); break;
                case CV_32F: reinterpret_cast<float*>(out.ptr(r))[c] = (float)v; break;
                case CV_64F: reinterpret_cast<double*>(out.ptr(r))[c] = v; break;
                default: assert(false);
                }
            }
        dst = out;
    }
protected:
    int type_ = 0;
    size_t step_ = 0;
    std::shared_ptr<std::vector<uint8_t>> buf_;
    uint8_t* data_ = nullptr;
};

// mask = gray > 0  (8UC1 in, 8UC1 0/255 out)
inline Mat operator>(const Mat& a, double s)
{
    assert(a.type() == CV_8UC1);
    Mat m(a.rows, a.cols, CV_8UC1);
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) m.ptr(r)[c] = a.ptr(r)[c] > s ? 255 : 0;
    return m;
}

template <typename T> class Mat_ : public Mat {
public:
    Mat_() = default;
    Mat_(int r, int c) : Mat(r, c, sizeof(T) == 8 ? CV_64F : CV_32F) {}
    Mat_(const Mat& m) : Mat(m) {}
    static Mat_ eye(Size s) { return Mat_(Mat::eye(s, sizeof(T) == 8 ? CV_64F : CV_32F)); }
};

}  // namespace cv
