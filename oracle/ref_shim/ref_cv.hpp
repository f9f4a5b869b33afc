// TEST INFRASTRUCTURE - stand-in for the few OpenCV types and calls the reference's CAPE sources use (cv::Mat_<T> as a dense
// row-major image, 3x3 erode / dilate with the reference's anchors and borders - the same semantics tests/test_oracle_cape.py
// checks against the real cv2 - element-wise compare / subtract, forEach, countNonZero, minMaxLoc, the tick counter).
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <vector>

using uchar = unsigned char;
using ushort = unsigned short;
using uint = unsigned int;
using int64 = long long;

namespace cv {

constexpr int BORDER_CONSTANT = 0;
constexpr int INTER_NEAREST = 0;
constexpr int CV_8UC3 = 16;

struct Point {
    int x, y;
    Point(int x_ = 0, int y_ = 0) : x(x_), y(y_) {}
};
struct Size {
    int width, height;
    Size(int w = 0, int h = 0) : width(w), height(h) {}
};
struct Scalar {
    double v[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {}
};

inline int64 getTickCount() { return std::chrono::steady_clock::now().time_since_epoch().count(); }
inline double getTickFrequency() { return 1e9; }

template <class T>
class Mat_ {
  public:
    int rows = 0, cols = 0;
    std::shared_ptr<std::vector<T>> buf;   // cv::Mat copies share their pixels
    T* data = nullptr;

    Mat_() = default;
    Mat_(int r, int c) : rows(r), cols(c), buf(std::make_shared<std::vector<T>>(size_t(r) * c)), data(buf->data()) {}
    Mat_(int r, int c, T v) : Mat_(r, c) { std::fill(buf->begin(), buf->end(), v); }
    static Mat_ ones(int r, int c) { return Mat_(r, c, T(1)); }
    static Mat_ zeros(int r, int c) { return Mat_(r, c, T(0)); }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return rows == 0 || cols == 0; }
    template <class U = T>
    U& at(int r, int c) { return data[size_t(r) * cols + c]; }
    template <class U = T>
    const U& at(int r, int c) const { return data[size_t(r) * cols + c]; }
    T& operator()(int r, int c) { return data[size_t(r) * cols + c]; }
    const T& operator()(int r, int c) const { return data[size_t(r) * cols + c]; }
    template <class U = T>
    U* ptr(int r) { return data + size_t(r) * cols; }
    template <class U = T>
    const U* ptr(int r) const { return data + size_t(r) * cols; }
    Mat_ clone() const
    {
        Mat_ m(rows, cols);
        std::copy(data, data + size_t(rows) * cols, m.data);
        return m;
    }
    Mat_& operator=(const Scalar& s)
    {
        std::fill(data, data + size_t(rows) * cols, T(s.v[0]));
        return *this;
    }
    Mat_& operator=(const T v)
    {
        std::fill(data, data + size_t(rows) * cols, v);
        return *this;
    }
    // setTo(value, mask): pixels whose mask is non-zero
    void setTo(const T v, const Mat_<uchar>& mask)
    {
        for (size_t i = 0; i < size_t(rows) * cols; ++i)
            if (mask.data[i]) data[i] = v;
    }
    template <class F>
    void forEach(const F& f) const
    {
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) {
                const int position[2] = {r, c};
                f(data[size_t(r) * cols + c], position);
            }
    }
};

// mat == value -> 255 / 0 mask (cv::MatExpr of a compare)
template <class T>
Mat_<uchar> operator==(const Mat_<T>& m, const int v)
{
    Mat_<uchar> out(m.rows, m.cols);
    for (size_t i = 0; i < size_t(m.rows) * m.cols; ++i) out.data[i] = m.data[i] == T(v) ? 255 : 0;
    return out;
}
// saturating subtraction of 8-bit images
inline Mat_<uchar> operator-(const Mat_<uchar>& a, const Mat_<uchar>& b)
{
    Mat_<uchar> out(a.rows, a.cols);
    for (size_t i = 0; i < size_t(a.rows) * a.cols; ++i) out.data[i] = uchar(std::max(0, int(a.data[i]) - int(b.data[i])));
    return out;
}

inline int countNonZero(const Mat_<uchar>& m)
{
    int n = 0;
    for (size_t i = 0; i < size_t(m.rows) * m.cols; ++i) n += m.data[i] != 0;
    return n;
}
inline void minMaxLoc(const Mat_<uchar>& m, double* mn, double* mx)
{
    double lo = 255, hi = 0;
    for (size_t i = 0; i < size_t(m.rows) * m.cols; ++i) lo = std::min(lo, double(m.data[i])), hi = std::max(hi, double(m.data[i]));
    if (mn) *mn = lo;
    if (mx) *mx = hi;
}

// 3x3 morphology, anchor at the kernel centre. erode: out-of-image neighbours count as `border` when a constant border is
// given (the reference passes BORDER_CONSTANT, Scalar(0)), else they are ignored (cv::erode's default border = +inf);
// dilate: out-of-image neighbours are ignored (default border = -inf). In-place calls are allowed (cv works on a copy).
// (cv's output arrays may be const objects - the reference writes its preallocated member masks from a const method - so the
// destination is taken by const reference and written through its shared pixel buffer; it must have the source's size.)
inline void morph3x3(const Mat_<uchar>& src, const Mat_<uchar>& dst_, const Mat_<uchar>& kernel, const bool erode,
                     const bool constant_border, const uchar border)
{
    Mat_<uchar> in = src.clone();
    Mat_<uchar> dst = dst_;   // shares the pixels
    if (dst.rows != src.rows || dst.cols != src.cols) throw std::logic_error("morph3x3: destination not allocated");
    for (int r = 0; r < in.rows; ++r)
        for (int c = 0; c < in.cols; ++c) {
            int acc = erode ? 255 : 0;
            for (int dr = -1; dr <= 1; ++dr)
                for (int dc = -1; dc <= 1; ++dc) {
                    if (!kernel.at<uchar>(dr + 1, dc + 1)) continue;
                    const int rr = r + dr, cc = c + dc;
                    int v;
                    if (rr < 0 || rr >= in.rows || cc < 0 || cc >= in.cols) {
                        if (!(erode && constant_border)) continue;
                        v = border;
                    }
                    else
                        v = in.at<uchar>(rr, cc);
                    acc = erode ? std::min(acc, v) : std::max(acc, v);
                }
            dst.at<uchar>(r, c) = uchar(acc);
        }
}
inline void erode(const Mat_<uchar>& src, const Mat_<uchar>& dst, const Mat_<uchar>& kernel)
{
    morph3x3(src, dst, kernel, true, false, 0);
}
inline void erode(const Mat_<uchar>& src, const Mat_<uchar>& dst, const Mat_<uchar>& kernel, Point, int, int, const Scalar& border)
{
    morph3x3(src, dst, kernel, true, true, uchar(border.v[0]));
}
inline void dilate(const Mat_<uchar>& src, const Mat_<uchar>& dst, const Mat_<uchar>& kernel) { morph3x3(src, dst, kernel, false, false, 0); }

using Mat = Mat_<uchar>;

}  // namespace cv
