// TEST INFRASTRUCTURE - a stand-in for the part of Eigen the reference's CAPE sources use, so that those sources
// (/root/reference/src/features/primitives/*.cpp, compiled where they lie, unmodified) build on a machine without Eigen.
// Dense column-major matrices, evaluated eagerly; just the members and operators those translation units call. The numeric
// kernels whose rounding matters are the ones oracle/linalg.hpp restates from Eigen's published algorithms (3x3 self-adjoint
// eigen-solver, 3x3 inverse / determinant) and plain left-to-right dot products - the same assumptions the oracle documents,
// each pinned separately in tests/test_oracle_cape.py (numpy / LAPACK cross-checks). NOT Eigen: what this build pins is the
// reference's own control flow and arithmetic around those kernels.
#pragma once
// Eigen/Core includes <emmintrin.h> / <xmmintrin.h> on x86-64 (src/Core/util/ConfigureVectorization.h), and through them
// <mm_malloc.h> -> <stdlib.h>, whose libstdc++ wrapper brings std::abs's floating-point overloads into the global namespace.
// The reference calls an UNQUALIFIED abs() on floats (plane_segment.cpp:51, the depth-continuity test): with Eigen's headers
// that resolves to abs(float); without them only ::abs(int) is visible and the difference would be truncated to an integer -
// a stand-in must reproduce that part of the environment too (found by comparing this build with the oracle).
#include <emmintrin.h>
#include <xmmintrin.h>

#include <array>
#include <atomic>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <memory>
#include <mutex>
#include <ostream>
#include <stdexcept>
#include <type_traits>
#include <vector>

#include "../linalg.hpp"

namespace Eigen {

constexpr int Dynamic = -1;
using Index = std::ptrdiff_t;

template <class T, int R, int C>
class Matrix;
template <class T>
using DynMat = Matrix<T, Dynamic, Dynamic>;

// coefficient-wise context (.array()): *, /, comparisons and reductions act per coefficient
template <class T>
struct Arr;

// a rectangular window into a matrix (row(), col(), block()): assignable, convertible to a matrix
template <class T>
class View {
  public:
    T* p;
    Index r, c, ld;   // element (i, j) at p[i + j * ld]
    View(T* p_, Index r_, Index c_, Index ld_) : p(p_), r(r_), c(c_), ld(ld_) {}
    Index rows() const { return r; }
    Index cols() const { return c; }
    Index size() const { return r * c; }
    T& operator()(Index i, Index j) const { return p[i + j * ld]; }
    T& operator()(Index i) const { return r == 1 ? p[i * ld] : p[i]; }
    T& operator[](Index i) const { return (*this)(i); }
    T x() const { return (*this)(0); }
    T y() const { return (*this)(1); }
    T z() const { return (*this)(2); }
    DynMat<T> eval() const;
    template <class M>
    const View& assign(const M& m) const
    {
        assert(m.rows() == r && m.cols() == c);
        for (Index j = 0; j < c; ++j)
            for (Index i = 0; i < r; ++i) (*this)(i, j) = m(i, j);
        return *this;
    }
    template <int R2, int C2>
    const View& operator=(const Matrix<T, R2, C2>& m) const { return assign(m); }
    const View& operator=(const View& v) const { return assign(v.eval()); }
    const View& operator=(const Arr<T>& a) const;
    template <int R2, int C2>
    const View& operator-=(const Matrix<T, R2, C2>& m) const
    {
        for (Index j = 0; j < c; ++j)
            for (Index i = 0; i < r; ++i) (*this)(i, j) -= m(i, j);
        return *this;
    }
    template <int R2, int C2>
    const View& operator+=(const Matrix<T, R2, C2>& m) const
    {
        for (Index j = 0; j < c; ++j)
            for (Index i = 0; i < r; ++i) (*this)(i, j) += m(i, j);
        return *this;
    }
    Arr<T> array() const;
    DynMat<T> transpose() const;
    T norm() const;
    T squaredNorm() const;
    T dot(const DynMat<T>& o) const;
    DynMat<T> normalized() const;
    template <int N>
    DynMat<T> head() const;
};

// CRTP base: constructors of classes DERIVED from a matrix type (the reference's coordinate structs inherit theirs with
// `using vector3::vector3`) must accept a plain matrix; an inherited constructor whose parameter is the base class itself is
// excluded from overload resolution, one taking this base is not (which is how Eigen's own EigenBase makes the same code work)
template <class D>
struct MatBase {
    const D& derived() const { return static_cast<const D&>(*this); }
};

template <class T, int R, int C>
class Matrix : public MatBase<Matrix<T, R, C>> {
    Index r_, c_;
    std::vector<T> d_;

  public:
    using Scalar = T;
    static constexpr bool kFixed = R >= 0 && C >= 0;
    static constexpr bool kVector = R == 1 || C == 1;

    Matrix() : r_(R < 0 ? 0 : R), c_(C < 0 ? 0 : C), d_(size_t(r_ * c_), T()) {}
    // one argument: a dynamic vector's size
    explicit Matrix(Index n) : r_(C == 1 ? n : (R < 0 ? 1 : R)), c_(C == 1 ? 1 : n), d_(size_t(r_ * c_), T())
    {
        static_assert(!kFixed, "size constructor on a fixed-size matrix");
    }
    // two arguments: (rows, cols) of a dynamic matrix, or the two coefficients of a fixed 2-vector
    template <class A, class B, class = std::enable_if_t<std::is_arithmetic_v<A> && std::is_arithmetic_v<B>>>
    Matrix(A a, B b)
    {
        if constexpr (kFixed) {
            static_assert(R * C == 2, "two coefficients for a 2-vector");
            r_ = R, c_ = C, d_ = {T(a), T(b)};
        }
        else {
            r_ = R < 0 ? Index(a) : R, c_ = C < 0 ? Index(b) : C;
            if (R >= 0 && C < 0 && false) c_ = Index(b);
            d_.assign(size_t(r_ * c_), T());
        }
    }
    Matrix(T a, T b, T c) : r_(R), c_(C), d_{a, b, c} { static_assert(kFixed && R * C == 3, "three coefficients"); }
    Matrix(T a, T b, T c, T d) : r_(R), c_(C), d_{a, b, c, d} { static_assert(kFixed && R * C == 4, "four coefficients"); }
    // rows of coefficients: matrix33({{..}, {..}, {..}})
    Matrix(std::initializer_list<std::initializer_list<T>> rows) : r_(Index(rows.size())), c_(Index(rows.begin()->size()))
    {
        d_.assign(size_t(r_ * c_), T());
        Index i = 0;
        for (const auto& row : rows) {
            Index j = 0;
            for (const T v : row) (*this)(i, j++) = v;
            ++i;
        }
    }
    template <class T2, int R2, int C2>
    Matrix(const MatBase<Matrix<T2, R2, C2>>& base) : r_(base.derived().rows()), c_(base.derived().cols()), d_(size_t(r_ * c_))
    {
        const Matrix<T2, R2, C2>& o = base.derived();
        // a vector may be given as a row or a column (Eigen transposes vectors on assignment)
        if (kVector && kFixed && o.rows() == C && o.cols() == R) r_ = R, c_ = C;
        assert((R < 0 || r_ == R) && (C < 0 || c_ == C));
        for (size_t k = 0; k < d_.size(); ++k) d_[k] = T(o.data()[k]);
    }
    Matrix(const View<T>& v) : Matrix(v.eval()) {}
    Matrix(const View<const T>& v) : Matrix(v.eval()) {}

    static Matrix Zero()
    {
        static_assert(kFixed, "Zero() without sizes");
        return Matrix();
    }
    static Matrix Zero(Index n)
    {
        Matrix m(n);
        return m;
    }
    static Matrix Zero(Index r, Index c) { return Matrix(r, c); }
    static Matrix Identity()
    {
        Matrix m;
        for (Index i = 0; i < m.rows() && i < m.cols(); ++i) m(i, i) = T(1);
        return m;
    }
    static Matrix Constant(T v)
    {
        static_assert(kFixed, "Constant(value) on a fixed-size matrix");
        Matrix m;
        m.setConstant(v);
        return m;
    }
    static Matrix Constant(Index n, T v)
    {
        Matrix m(n);
        m.setConstant(v);
        return m;
    }
    static Matrix Constant(Index r, Index c, T v)
    {
        Matrix m(r, c);
        m.setConstant(v);
        return m;
    }

    Index rows() const { return r_; }
    Index cols() const { return c_; }
    Index size() const { return r_ * c_; }
    T* data() { return d_.data(); }
    const T* data() const { return d_.data(); }
    void resize(Index r, Index c) { r_ = r, c_ = c, d_.assign(size_t(r * c), T()); }
    void resize(Index n) { resize(C == 1 ? n : 1, C == 1 ? 1 : n); }
    void setZero() { std::fill(d_.begin(), d_.end(), T()); }
    void setConstant(T v) { std::fill(d_.begin(), d_.end(), v); }
    void fill(T v) { setConstant(v); }

    // std::vector<bool> has no references to elements: bool matrices store unsigned char
    auto& operator()(Index i, Index j) { return d_[size_t(i + j * r_)]; }
    const auto& operator()(Index i, Index j) const { return d_[size_t(i + j * r_)]; }
    auto& operator()(Index i) { return d_[size_t(i)]; }
    const auto& operator()(Index i) const { return d_[size_t(i)]; }
    auto& operator[](Index i) { return d_[size_t(i)]; }
    const auto& operator[](Index i) const { return d_[size_t(i)]; }
    T x() const { return d_[0]; }
    T y() const { return d_[1]; }
    T z() const { return d_[2]; }
    T w() const { return d_[3]; }
    T& x() { return d_[0]; }
    T& y() { return d_[1]; }
    T& z() { return d_[2]; }
    T& w() { return d_[3]; }

    View<T> row(Index i) { return View<T>(d_.data() + i, 1, c_, r_); }
    View<T> col(Index j) { return View<T>(d_.data() + j * r_, r_, 1, r_); }
    View<T> block(Index i, Index j, Index r, Index c) { return View<T>(d_.data() + i + j * r_, r, c, r_); }
    template <int BR, int BC>
    View<T> block(Index i, Index j) { return block(i, j, BR, BC); }
    DynMat<T> row(Index i) const { return View<T>(const_cast<T*>(d_.data()) + i, 1, c_, r_).eval(); }
    Matrix<T, R, 1> col(Index j) const
    {
        Matrix<T, R, 1> v = View<T>(const_cast<T*>(d_.data()) + j * r_, r_, 1, r_).eval();
        return v;
    }
    DynMat<T> block(Index i, Index j, Index r, Index c) const { return View<T>(const_cast<T*>(d_.data()) + i + j * r_, r, c, r_).eval(); }
    template <int BR, int BC>
    Matrix<T, BR, BC> block(Index i, Index j) const
    {
        Matrix<T, BR, BC> m = block(i, j, BR, BC);
        return m;
    }
    template <int N>
    View<T> head() { return View<T>(d_.data(), N, 1, r_ * c_); }
    template <int N>
    View<T> tail() { return View<T>(d_.data() + (size() - N), N, 1, r_ * c_); }
    template <int N>
    Matrix<T, N, 1> tail() const
    {
        Matrix<T, N, 1> v;
        for (int i = 0; i < N; ++i) v(i) = d_[size_t(size() - N + i)];
        return v;
    }
    template <int N>
    Matrix<T, N, 1> head() const
    {
        Matrix<T, N, 1> v;
        for (int i = 0; i < N; ++i) v(i) = d_[size_t(i)];
        return v;
    }

    // the comma initialiser on vectors, as the sources use it: `v << a;`, `v6 << position, angles;`, `v << x, y, z;`
    // Eigen's CommaInitializer: items are placed left to right, a row of blocks at a time (a scalar is a 1 x 1 block)
    struct CommaInit {
        Matrix& m;
        Index row = 0, col = 0, block_rows = 0;
        void place(const T* src, Index br, Index bc, Index ld)
        {
            if (col >= m.c_) row += block_rows, col = 0, block_rows = 0;
            if (col == 0) block_rows = br;
            for (Index j = 0; j < bc; ++j)
                for (Index i = 0; i < br; ++i) m(row + i, col + j) = src[i + j * ld];
            col += bc;
        }
        template <int R2, int C2>
        CommaInit& operator,(const Matrix<T, R2, C2>& o)
        {
            place(o.data(), o.rows(), o.cols(), o.rows());
            return *this;
        }
        CommaInit& operator,(const T v)
        {
            place(&v, 1, 1, 1);
            return *this;
        }
    };
    template <int R2, int C2>
    CommaInit operator<<(const Matrix<T, R2, C2>& o)
    {
        CommaInit c{*this};
        c, o;
        return c;
    }
    CommaInit operator<<(const T v)
    {
        CommaInit c{*this};
        c, v;
        return c;
    }

    Matrix<T, C, R> transpose() const
    {
        Matrix<T, C, R> t;
        t.resize_for(c_, r_);
        for (Index j = 0; j < c_; ++j)
            for (Index i = 0; i < r_; ++i) t(j, i) = (*this)(i, j);
        return t;
    }
    Matrix<T, C, R> adjoint() const { return transpose(); }
    void resize_for(Index r, Index c) { r_ = r, c_ = c, d_.assign(size_t(r * c), T()); }

    T squaredNorm() const
    {
        T s = T();
        for (const T v : d_) s += v * v;
        return s;
    }
    T norm() const { return std::sqrt(squaredNorm()); }
    template <int P>
    T lpNorm() const
    {
        static_assert(P == 1, "only the 1-norm");
        T s = T();
        for (const T v : d_) s += std::abs(v);
        return s;
    }
    Matrix& base() { return *this; }
    const Matrix& base() const { return *this; }
    bool allFinite() const
    {
        for (const T v : d_)
            if (!std::isfinite(v)) return false;
        return true;
    }
    template <class S, class = std::enable_if_t<std::is_arithmetic_v<S>>>
    Matrix<unsigned char, R, C> operator>(const S s) const
    {
        Matrix<unsigned char, R, C> m;
        m.resize_for(r_, c_);
        for (size_t k = 0; k < d_.size(); ++k) m.data()[k] = d_[k] > T(s) ? 1 : 0;
        return m;
    }
    Matrix normalized() const
    {
        // Eigen: v / sqrt(squaredNorm) when the squared norm is positive
        const T n2 = squaredNorm();
        if (n2 > T(0)) return *this / std::sqrt(n2);
        return *this;
    }
    void normalize() { *this = normalized(); }
    template <int R2, int C2>
    T dot(const Matrix<T, R2, C2>& o) const
    {
        assert(o.size() == size());
        T s = T();
        for (size_t k = 0; k < d_.size(); ++k) s = k == 0 ? d_[0] * o.data()[0] : s + d_[k] * o.data()[k];
        return s;
    }
    template <int R2, int C2>
    Matrix<T, 3, 1> cross(const Matrix<T, R2, C2>& o) const
    {
        return Matrix<T, 3, 1>(d_[1] * o(2) - d_[2] * o(1), d_[2] * o(0) - d_[0] * o(2), d_[0] * o(1) - d_[1] * o(0));
    }
    T sum() const
    {
        T s = T();
        for (const T v : d_) s += v;
        return s;
    }
    Index count() const
    {
        Index n = 0;
        for (const T v : d_) n += v ? 1 : 0;
        return n;
    }
    template <int R2, int C2>
    Matrix cwiseProduct(const Matrix<T, R2, C2>& o) const
    {
        assert(o.size() == size());
        Matrix m = *this;
        for (size_t k = 0; k < d_.size(); ++k) m.d_[k] = d_[k] * o.data()[k];
        return m;
    }
    Matrix cwiseSqrt() const
    {
        Matrix m = *this;
        for (T& v : m.d_) v = std::sqrt(v);
        return m;
    }
    bool isApproxToConstant(const T value, const T prec = T(1e-12)) const
    {
        // Eigen: isApprox(Constant(value)) = |this - c|^2 <= prec^2 min(|this|^2, |c|^2)
        T diff = T(), a = T(), b = T();
        for (const T v : d_) diff += (v - value) * (v - value), a += v * v, b += value * value;
        return diff <= prec * prec * std::min(a, b);
    }
    View<T> segment(Index start, Index n) { return View<T>(d_.data() + start, n, 1, r_ * c_); }
    Matrix cwiseAbs() const
    {
        Matrix m = *this;
        for (T& v : m.d_) v = std::abs(v);
        return m;
    }
    Arr<T> array() const;
    Matrix matrix() const { return *this; }
    // homogeneous coordinates of a fixed-size column vector
    Matrix<T, (R > 0 ? R + 1 : Dynamic), 1> homogeneous() const
    {
        static_assert(R > 0 && C == 1, "homogeneous() of a fixed-size column vector");
        Matrix<T, R + 1, 1> h;
        for (int i = 0; i < R; ++i) h(i) = d_[size_t(i)];
        h(R) = T(1);
        return h;
    }
    // diagonal() as an assignable view (`cov.diagonal() += v`) and as a value
    struct Diagonal {
        Matrix& m;
        template <int R2, int C2>
        Diagonal& operator+=(const Matrix<T, R2, C2>& v)
        {
            for (Index i = 0; i < std::min(m.rows(), m.cols()); ++i) m(i, i) += v(i);
            return *this;
        }
        operator Matrix<T, R, 1>() const
        {
            Matrix<T, R, 1> d;
            d.resize_for(std::min(m.rows(), m.cols()), 1);
            for (Index i = 0; i < d.rows(); ++i) d(i) = m(i, i);
            return d;
        }
    };
    Diagonal diagonal() { return Diagonal{*this}; }
    Matrix<T, R, 1> diagonal() const
    {
        Matrix<T, R, 1> d;
        d.resize_for(std::min(r_, c_), 1);
        for (Index i = 0; i < d.rows(); ++i) d(i) = (*this)(i, i);
        return d;
    }
    // colwise().norm(): the Euclidean norm of every column, as a row
    struct Colwise {
        const Matrix& m;
        DynMat<T> norm() const
        {
            DynMat<T> out(1, m.cols());
            for (Index j = 0; j < m.cols(); ++j) {
                T s = T();
                for (Index i = 0; i < m.rows(); ++i) s += m(i, j) * m(i, j);
                out(0, j) = std::sqrt(s);
            }
            return out;
        }
    };
    Colwise colwise() const { return Colwise{*this}; }
    bool hasNaN() const
    {
        for (const T v : d_)
            if (v != v) return true;
        return false;
    }
    bool isApprox(const Matrix& o, const T prec = T(1e-12)) const
    {
        T diff = T(), a = T(), b = T();
        for (size_t k = 0; k < d_.size(); ++k) {
            diff += (d_[k] - o.d_[k]) * (d_[k] - o.d_[k]);
            a += d_[k] * d_[k];
            b += o.d_[k] * o.d_[k];
        }
        return diff <= prec * prec * std::min(a, b);
    }

    // 3x3 only (oracle/linalg.hpp: Eigen's cofactor formulas)
    T determinant() const
    {
        static_assert(R == 3 && C == 3, "determinant of a 3x3");
        return oracle::det3(to_mat3());
    }
    Matrix inverse() const
    {
        static_assert((R == 3 && C == 3) || (R == 4 && C == 4), "inverse of a 3x3 or a 4x4");
        Matrix m;
        if constexpr (R == 3) {
            const oracle::Mat3 inv = oracle::inverse3(to_mat3());
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) m(i, j) = inv.m[i][j];
        }
        else {
            oracle::Mat4 a;
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j) a.m[i][j] = (*this)(i, j);
            const oracle::Mat4 inv = oracle::inverse4(a);
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j) m(i, j) = inv.m[i][j];
        }
        return m;
    }
    // MatrixBase::eulerAngles(0, 1, 2) (oracle/linalg.hpp: euler_angles_012, Eigen's algorithm)
    Matrix<T, 3, 1> eulerAngles(int a0, int a1, int a2) const
    {
        static_assert(R == 3 && C == 3, "eulerAngles of a rotation matrix");
        if (a0 != 0 || a1 != 1 || a2 != 2) throw std::logic_error("eulerAngles: only (0, 1, 2)");
        double res[3];
        oracle::euler_angles_012(to_mat3(), res);
        return Matrix<T, 3, 1>(res[0], res[1], res[2]);
    }
    oracle::Mat3 to_mat3() const
    {
        oracle::Mat3 a;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) a.m[i][j] = (*this)(i, j);
        return a;
    }

    Matrix operator-() const
    {
        Matrix m = *this;
        for (T& v : m.d_) v = -v;
        return m;
    }
    Matrix& operator+=(const Matrix& o)
    {
        for (size_t k = 0; k < d_.size(); ++k) d_[k] += o.d_[k];
        return *this;
    }
    Matrix& operator-=(const Matrix& o)
    {
        for (size_t k = 0; k < d_.size(); ++k) d_[k] -= o.d_[k];
        return *this;
    }
    Matrix& operator*=(const T s)
    {
        for (T& v : d_) v *= s;
        return *this;
    }
    Matrix& operator/=(const T s)
    {
        for (T& v : d_) v /= s;
        return *this;
    }
};

template <class T, int R1, int C1, int R2, int C2>
Matrix<T, (R1 >= 0 ? R1 : R2), (C1 >= 0 ? C1 : C2)> operator+(const Matrix<T, R1, C1>& a, const Matrix<T, R2, C2>& b)
{
    assert(a.size() == b.size());
    Matrix<T, (R1 >= 0 ? R1 : R2), (C1 >= 0 ? C1 : C2)> m;
    m.resize_for(a.rows(), a.cols());
    for (Index k = 0; k < a.size(); ++k) m.data()[k] = a.data()[k] + b.data()[k];
    return m;
}
template <class T, int R1, int C1, int R2, int C2>
Matrix<T, (R1 >= 0 ? R1 : R2), (C1 >= 0 ? C1 : C2)> operator-(const Matrix<T, R1, C1>& a, const Matrix<T, R2, C2>& b)
{
    assert(a.size() == b.size());
    Matrix<T, (R1 >= 0 ? R1 : R2), (C1 >= 0 ? C1 : C2)> m;
    m.resize_for(a.rows(), a.cols());
    for (Index k = 0; k < a.size(); ++k) m.data()[k] = a.data()[k] - b.data()[k];
    return m;
}
template <class T, int R, int C>
DynMat<T> operator-(const View<T>& a, const Matrix<T, R, C>& b) { return a.eval() - b; }
template <class T, int R, int C>
DynMat<T> operator+(const View<T>& a, const Matrix<T, R, C>& b) { return a.eval() + b; }
template <class T, int R, int C, class S, class = std::enable_if_t<std::is_arithmetic_v<S>>>
Matrix<T, R, C> operator*(Matrix<T, R, C> a, const S s) { return a *= T(s); }
template <class T, int R, int C, class S, class = std::enable_if_t<std::is_arithmetic_v<S>>>
Matrix<T, R, C> operator*(const S s, Matrix<T, R, C> a) { return a *= T(s); }
template <class T, int R, int C, class S, class = std::enable_if_t<std::is_arithmetic_v<S>>>
Matrix<T, R, C> operator/(Matrix<T, R, C> a, const S s) { return a /= T(s); }

// matrix product: every coefficient is a left-to-right sum over the inner index
template <class T, int R, int K, int K2, int C>
Matrix<T, R, C> operator*(const Matrix<T, R, K>& a, const Matrix<T, K2, C>& b)
{
    assert(a.cols() == b.rows());
    Matrix<T, R, C> m;
    m.resize_for(a.rows(), b.cols());
    for (Index i = 0; i < a.rows(); ++i)
        for (Index j = 0; j < b.cols(); ++j) {
            T s = T();
            for (Index k = 0; k < a.cols(); ++k) s = k == 0 ? a(i, 0) * b(0, j) : s + a(i, k) * b(k, j);
            m(i, j) = s;
        }
    return m;
}

template <class T>
DynMat<T> View<T>::eval() const
{
    DynMat<T> m(r, c);
    for (Index j = 0; j < c; ++j)
        for (Index i = 0; i < r; ++i) m(i, j) = (*this)(i, j);
    return m;
}
template <class T>
Arr<T> View<T>::array() const { return Arr<T>{eval()}; }
template <class T>
const View<T>& View<T>::operator=(const Arr<T>& a) const { return assign(a.m); }
template <class T>
DynMat<T> View<T>::transpose() const { return eval().transpose(); }
template <class T>
T View<T>::norm() const { return eval().norm(); }
template <class T>
T View<T>::squaredNorm() const { return eval().squaredNorm(); }
template <class T>
T View<T>::dot(const DynMat<T>& o) const { return eval().dot(o); }
template <class T>
DynMat<T> View<T>::normalized() const { return eval().normalized(); }
template <class T>
template <int N>
DynMat<T> View<T>::head() const
{
    DynMat<T> v(N, 1);
    for (int i = 0; i < N; ++i) v(i) = (*this)(i);
    return v;
}

template <class T>
struct Arr {
    DynMat<T> m;
    T sum() const { return m.sum(); }
    Index count() const { return m.count(); }
    Arr<unsigned char> operator>(const T s) const { return Arr<unsigned char>{m > s}; }
    Arr<unsigned char> operator>=(const T s) const
    {
        Arr<unsigned char> out{DynMat<unsigned char>(m.rows(), m.cols())};
        for (Index k = 0; k < m.size(); ++k) out.m.data()[k] = m.data()[k] >= s ? 1 : 0;
        return out;
    }
    Arr<unsigned char> operator<=(const T s) const
    {
        Arr<unsigned char> out{DynMat<unsigned char>(m.rows(), m.cols())};
        for (Index k = 0; k < m.size(); ++k) out.m.data()[k] = m.data()[k] <= s ? 1 : 0;
        return out;
    }
    Arr<unsigned char> operator<=(const Arr& o) const
    {
        Arr<unsigned char> out{DynMat<unsigned char>(m.rows(), m.cols())};
        for (Index k = 0; k < m.size(); ++k) out.m.data()[k] = m.data()[k] <= o.m.data()[k] ? 1 : 0;
        return out;
    }
    bool all() const
    {
        for (Index k = 0; k < m.size(); ++k)
            if (!m.data()[k]) return false;
        return true;
    }
    bool any() const
    {
        for (Index k = 0; k < m.size(); ++k)
            if (m.data()[k]) return true;
        return false;
    }
    Arr abs() const { return Arr{m.cwiseAbs()}; }
    DynMat<T> matrix() const { return m; }
    template <int R, int C>
    operator Matrix<T, R, C>() const { return Matrix<T, R, C>(m); }
};
template <class T>
Arr<T> operator*(const Arr<T>& a, const Arr<T>& b)
{
    assert(a.m.size() == b.m.size());
    Arr<T> out{a.m};
    for (Index k = 0; k < a.m.size(); ++k) out.m.data()[k] = a.m.data()[k] * b.m.data()[k];
    return out;
}
template <class T>
Arr<T> operator+(const Arr<T>& a, const Arr<T>& b)
{
    assert(a.m.size() == b.m.size());
    Arr<T> out{a.m};
    for (Index k = 0; k < a.m.size(); ++k) out.m.data()[k] = a.m.data()[k] + b.m.data()[k];
    return out;
}
template <class T>
Arr<T> operator/(const Arr<T>& a, const Arr<T>& b)
{
    assert(a.m.size() == b.m.size());
    Arr<T> out{a.m};
    for (Index k = 0; k < a.m.size(); ++k) out.m.data()[k] = a.m.data()[k] / b.m.data()[k];
    return out;
}
template <class T, int R, int C>
Arr<T> Matrix<T, R, C>::array() const { return Arr<T>{DynMat<T>(*this)}; }

template <class T, int N>
using Vector = Matrix<T, N, 1>;
using MatrixXf = Matrix<float, Dynamic, Dynamic>;
using MatrixXd = Matrix<double, Dynamic, Dynamic>;
using VectorXd = Matrix<double, Dynamic, 1>;
using Vector2d = Matrix<double, 2, 1>;
using Vector3d = Matrix<double, 3, 1>;
using Vector4d = Matrix<double, 4, 1>;
using Matrix2d = Matrix<double, 2, 2>;
using Matrix3d = Matrix<double, 3, 3>;
using Matrix4d = Matrix<double, 4, 4>;

template <class T, int R, int C>
std::ostream& operator<<(std::ostream& os, const Matrix<T, R, C>& m)
{
    for (Index i = 0; i < m.rows(); ++i) {
        for (Index j = 0; j < m.cols(); ++j) os << (j ? " " : "") << m(i, j);
        if (i + 1 < m.rows()) os << "\n";
    }
    return os;
}

template <class T>
using aligned_allocator = std::allocator<T>;

// Quaternion<double> with Eigen's conventions: coefficients stored (x, y, z, w), constructor (w, x, y, z), Hamilton product,
// toRotationMatrix() by Eigen's formula (oracle/linalg.hpp: quat_to_rot)
template <class T>
class Quaternion {
    T x_ = 0, y_ = 0, z_ = 0, w_ = 1;

  public:
    Quaternion() = default;
    Quaternion(T w, T x, T y, T z) : x_(x), y_(y), z_(z), w_(w) {}
    static Quaternion Identity() { return Quaternion(1, 0, 0, 0); }
    void setIdentity() { *this = Identity(); }
    Quaternion& operator*=(const Quaternion& b) { return *this = *this * b; }
    T w() const { return w_; }
    T x() const { return x_; }
    T y() const { return y_; }
    T z() const { return z_; }
    T& w() { return w_; }
    T& x() { return x_; }
    T& y() { return y_; }
    T& z() { return z_; }
    Matrix<T, 4, 1> coeffs() const { return Matrix<T, 4, 1>(x_, y_, z_, w_); }
    // summed w, x, y, z: the order oracle/pose.cpp documents for PoseBase::set_parameters (Eigen's own packet reduction is
    // (x^2 + z^2) + (y^2 + w^2) with SSE2; a <= 1 ulp effect on q either way, third-party arithmetic)
    T squaredNorm() const { return ((w_ * w_ + x_ * x_) + y_ * y_) + z_ * z_; }
    T norm() const { return std::sqrt(squaredNorm()); }
    void normalize()
    {
        // Eigen: coeffs().normalize() -> divide by sqrt(squaredNorm) when positive
        const T n2 = squaredNorm();
        if (n2 > T(0)) {
            const T n = std::sqrt(n2);
            x_ /= n, y_ /= n, z_ /= n, w_ /= n;
        }
    }
    Quaternion normalized() const
    {
        Quaternion q = *this;
        q.normalize();
        return q;
    }
    Quaternion conjugate() const { return Quaternion(w_, -x_, -y_, -z_); }
    Quaternion inverse() const
    {
        const T n2 = squaredNorm();
        return Quaternion(w_ / n2, -x_ / n2, -y_ / n2, -z_ / n2);
    }
    Quaternion operator*(const Quaternion& b) const
    {
        const Quaternion& a = *this;
        return Quaternion(a.w_ * b.w_ - a.x_ * b.x_ - a.y_ * b.y_ - a.z_ * b.z_, a.w_ * b.x_ + a.x_ * b.w_ + a.y_ * b.z_ - a.z_ * b.y_,
                          a.w_ * b.y_ + a.y_ * b.w_ + a.z_ * b.x_ - a.x_ * b.z_, a.w_ * b.z_ + a.z_ * b.w_ + a.x_ * b.y_ - a.y_ * b.x_);
    }
    Matrix<T, 3, 3> toRotationMatrix() const
    {
        const double q[4] = {w_, x_, y_, z_};
        const oracle::Mat3 r = oracle::quat_to_rot(q);
        Matrix<T, 3, 3> m;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) m(i, j) = r.m[i][j];
        return m;
    }
    Matrix<T, 3, 3> matrix() const { return toRotationMatrix(); }
    T angularDistance(const Quaternion& o) const
    {
        const Quaternion d = *this * o.conjugate();
        const T vn = std::sqrt(d.x_ * d.x_ + d.y_ * d.y_ + d.z_ * d.z_);
        return T(2) * std::atan2(vn, std::abs(d.w_));
    }
    bool isApprox(const Quaternion& o, const T prec = T(1e-12)) const { return coeffs().isApprox(o.coeffs(), prec); }
};
using Quaterniond = Quaternion<double>;

// Affine3d as coordinates/point_coordinates.cpp::get_transformation_matrix uses it (a function outside the CAPE path, present
// in a translation unit the CAPE path needs): `T.linear() << a, b, c;` fills the columns, `T.translation() << t;`
struct Affine3d {
    Matrix3d lin = Matrix3d::Identity();
    Vector3d trans;
    static Affine3d Identity() { return Affine3d(); }
    struct Columns {
        Matrix3d& m;
        int next = 0;
        Columns& operator<<(const Vector3d& v) { return *this, v; }
        Columns& operator,(const Vector3d& v)
        {
            for (int i = 0; i < 3; ++i) m(i, next) = v(i);
            ++next;
            return *this;
        }
    };
    Columns linear() { return Columns{lin}; }
    Vector3d& translation() { return trans; }
    const Matrix3d& rotation() const { return lin; }
};

// SelfAdjointEigenSolver<Matrix3d>: the tridiagonal QL of Eigen's compute(), as restated in oracle/linalg.hpp
template <class M>
class SelfAdjointEigenSolver {
    Matrix<double, 3, 1> vals_;
    Matrix<double, 3, 3> vecs_;

  public:
    template <class T, int R, int C>
    explicit SelfAdjointEigenSolver(const Matrix<T, R, C>& a)
    {
        assert(a.rows() == 3 && a.cols() == 3);
        oracle::Mat3 m, q;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) m.m[i][j] = a(i, j);
        double ev[3];
        oracle::self_adjoint_eigen3(m, ev, q);
        for (int i = 0; i < 3; ++i) {
            vals_(i) = ev[i];
            for (int j = 0; j < 3; ++j) vecs_(i, j) = q.m[i][j];
        }
    }
    const Matrix<double, 3, 1>& eigenvalues() const { return vals_; }
    const Matrix<double, 3, 3>& eigenvectors() const { return vecs_; }
};

}  // namespace Eigen
