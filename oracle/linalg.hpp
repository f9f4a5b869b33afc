// TEST INFRASTRUCTURE — CPU oracle. Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
// legs may use anything under oracle/. Never linked into the product library.
//
// Tiny fixed-size FP64 linear algebra restating the Eigen (>= 3.4, un-vendored: CMakeLists.txt:35) routines
// the reference calls on the hot path. Eigen's sources are NOT on this machine; the algorithms below are
// restated from the published Eigen 3.4 algorithms (file names given per function). "parity unpinned" at the
// last-ulp level: cross-checked against numpy.linalg / scipy only (tests/test_oracle_linalg.py).
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace oracle {

struct Vec3 {
    double x = 0, y = 0, z = 0;
    double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline Vec3 operator+(const Vec3& a, const Vec3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator-(const Vec3& a) { return {-a.x, -a.y, -a.z}; }
inline Vec3 operator*(const Vec3& a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator*(double s, const Vec3& a) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator/(const Vec3& a, double s) { return {a.x / s, a.y / s, a.z / s}; }
// Eigen's 3-vector dot/squaredNorm reduce as (x*x' + y*y') + z*z' (redux, no vectorisation for size 3 doubles
// with SSE2 unaligned: linear traversal).
inline double dot(const Vec3& a, const Vec3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline double sqnorm(const Vec3& a) { return dot(a, a); }
inline double norm(const Vec3& a) { return std::sqrt(sqnorm(a)); }
inline Vec3 cross(const Vec3& a, const Vec3& b)
{
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// Eigen normalized()/normalize(): divides by the norm when the squared norm is > 0.
inline Vec3 normalized(const Vec3& a)
{
    const double z = sqnorm(a);
    if (z > 0.0) return a / std::sqrt(z);
    return a;
}

struct Mat3 {
    double m[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    double& operator()(int r, int c) { return m[r][c]; }
    double operator()(int r, int c) const { return m[r][c]; }
};
struct Mat4 {
    double m[4][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
    double& operator()(int r, int c) { return m[r][c]; }
    double operator()(int r, int c) const { return m[r][c]; }
};

inline Mat3 matmul(const Mat3& a, const Mat3& b)
{
    Mat3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r(i, j) = (a(i, 0) * b(0, j) + a(i, 1) * b(1, j)) + a(i, 2) * b(2, j);
    return r;
}
inline Vec3 matvec(const Mat3& a, const Vec3& v)
{
    return {(a(0, 0) * v.x + a(0, 1) * v.y) + a(0, 2) * v.z, (a(1, 0) * v.x + a(1, 1) * v.y) + a(1, 2) * v.z,
            (a(2, 0) * v.x + a(2, 1) * v.y) + a(2, 2) * v.z};
}
inline Mat3 transpose(const Mat3& a)
{
    Mat3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r(i, j) = a(j, i);
    return r;
}
inline Mat4 matmul(const Mat4& a, const Mat4& b)
{
    Mat4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = a(i, 0) * b(0, j);
            for (int k = 1; k < 4; ++k) s += a(i, k) * b(k, j);
            r(i, j) = s;
        }
    return r;
}

// Eigen/src/LU/Determinant.h, 3x3: bruteforce_det3_helper(m,0,1,2) - (m,1,0,2) + (m,2,0,1)
inline double det3(const Mat3& a)
{
    auto h = [&](int i, int j, int k) { return a(0, i) * (a(1, j) * a(2, k) - a(1, k) * a(2, j)); };
    return h(0, 1, 2) - h(1, 0, 2) + h(2, 0, 1);
}

// Eigen/src/LU/InverseImpl.h, size 3: cofactors; result(j,i) = cofactor<i,j> * invdet, det from column 0.
inline Mat3 inverse3(const Mat3& a)
{
    auto cof = [&](int i, int j) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return a(i1, j1) * a(i2, j2) - a(i1, j2) * a(i2, j1);
    };
    const double c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
    const double det = (c00 * a(0, 0) + c10 * a(1, 0)) + c20 * a(2, 0);
    const double invdet = 1.0 / det;
    Mat3 r;
    r(0, 0) = c00 * invdet;
    r(0, 1) = c10 * invdet;
    r(0, 2) = c20 * invdet;
    r(1, 0) = cof(0, 1) * invdet;
    r(1, 1) = cof(1, 1) * invdet;
    r(1, 2) = cof(2, 1) * invdet;
    r(2, 0) = cof(0, 2) * invdet;
    r(2, 1) = cof(1, 2) * invdet;
    r(2, 2) = cof(2, 2) * invdet;
    return r;
}

// General 4x4 inverse by cofactors (Eigen/src/LU/InverseImpl.h size-4 path is a SIMD cofactor scheme; the
// association of the products is not reproduced — differences are last-ulp).
inline Mat4 inverse4(const Mat4& a)
{
    const double* m = &a.m[0][0];
    double inv[16];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] +
             m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] -
             m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] +
             m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] -
              m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] -
             m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] +
             m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] -
             m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] +
              m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] +
             m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] -
             m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] +
              m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] -
              m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] -
             m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] +
             m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] -
              m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] +
              m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    const double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    const double invdet = 1.0 / det;
    Mat4 r;
    for (int i = 0; i < 16; ++i) (&r.m[0][0])[i] = inv[i] * invdet;
    return r;
}

// Eigen/src/Jacobi/Jacobi.h: JacobiRotation<double>::makeGivens(p, q) (real case).
struct Givens {
    double c, s;
};
inline Givens make_givens(double p, double q)
{
    Givens g;
    if (q == 0.0) {
        g.c = p < 0.0 ? -1.0 : 1.0;
        g.s = 0.0;
    }
    else if (p == 0.0) {
        g.c = 0.0;
        g.s = q < 0.0 ? 1.0 : -1.0;
    }
    else if (std::fabs(p) > std::fabs(q)) {
        const double t = q / p;
        double u = std::sqrt(1.0 + t * t);
        if (p < 0.0) u = -u;
        g.c = 1.0 / u;
        g.s = -t * g.c;
    }
    else {
        const double t = p / q;
        double u = std::sqrt(1.0 + t * t);
        if (q < 0.0) u = -u;
        g.s = -1.0 / u;
        g.c = -t * g.s;
    }
    return g;
}

// Eigen numext::hypot for reals (MathFunctionsImpl.h: positive_real_hypot).
inline double eigen_hypot(double x, double y)
{
    x = std::fabs(x);
    y = std::fabs(y);
    const double p = std::max(x, y);
    if (p == 0.0) return 0.0;
    const double qp = std::min(y, x) / p;
    return p * std::sqrt(1.0 + qp * qp);
}

// Eigen/src/Eigenvalues/SelfAdjointEigenSolver.h: compute() (iterative path: scale, 3x3 closed-form
// tridiagonalisation, implicit symmetric QR steps with Wilkinson shift, ascending sort).
// Only the LOWER triangle of `a` is read. Returns false on no convergence.
inline bool self_adjoint_eigen3(const Mat3& a, double evals[3], Mat3& evecs)
{
    const int n = 3;
    double mat[3][3] = {{a(0, 0), 0, 0}, {a(1, 0), a(1, 1), 0}, {a(2, 0), a(2, 1), a(2, 2)}};
    double scale = 0.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) scale = std::max(scale, std::fabs(mat[i][j]));
    if (scale == 0.0) scale = 1.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j <= i; ++j) mat[i][j] /= scale;

    // Tridiagonalization.h: tridiagonalization_inplace_selector<MatrixType,3,false>
    double diag[3], subdiag[2];
    double q[3][3];
    const double tol = DBL_MIN;
    diag[0] = mat[0][0];
    const double v1norm2 = mat[2][0] * mat[2][0];
    if (v1norm2 <= tol) {
        diag[1] = mat[1][1];
        diag[2] = mat[2][2];
        subdiag[0] = mat[1][0];
        subdiag[1] = mat[2][1];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) q[i][j] = (i == j) ? 1.0 : 0.0;
    }
    else {
        const double beta = std::sqrt(mat[1][0] * mat[1][0] + v1norm2);
        const double invBeta = 1.0 / beta;
        const double m01 = mat[1][0] * invBeta;
        const double m02 = mat[2][0] * invBeta;
        const double qq = 2.0 * m01 * mat[2][1] + m02 * (mat[2][2] - mat[1][1]);
        diag[1] = mat[1][1] + m02 * qq;
        diag[2] = mat[2][2] - m02 * qq;
        subdiag[0] = beta;
        subdiag[1] = mat[2][1] - m01 * qq;
        const double qi[3][3] = {{1, 0, 0}, {0, m01, m02}, {0, m02, -m01}};
        std::memcpy(q, qi, sizeof(q));
    }

    // computeFromTridiagonal_impl
    const int maxIterations = 30;
    int end = n - 1, start = 0, iter = 0;
    const double considerAsZero = DBL_MIN;
    const double precision_inv = 1.0 / DBL_EPSILON;
    while (end > 0) {
        for (int i = start; i < end; ++i) {
            if (std::fabs(subdiag[i]) < considerAsZero) {
                subdiag[i] = 0.0;
            }
            else {
                const double scaled_subdiag = precision_inv * subdiag[i];
                if (scaled_subdiag * scaled_subdiag <= (std::fabs(diag[i]) + std::fabs(diag[i + 1]))) subdiag[i] = 0.0;
            }
        }
        while (end > 0 && subdiag[end - 1] == 0.0) end--;
        if (end <= 0) break;
        iter++;
        if (iter > maxIterations * n) break;
        start = end - 1;
        while (start > 0 && subdiag[start - 1] != 0.0) start--;

        // tridiagonal_qr_step
        const double td = (diag[end - 1] - diag[end]) * 0.5;
        const double e = subdiag[end - 1];
        double mu = diag[end];
        if (td == 0.0) {
            mu -= std::fabs(e);
        }
        else if (e != 0.0) {
            const double e2 = e * e;
            const double h = eigen_hypot(td, e);
            if (e2 == 0.0)
                mu -= e / ((td + (td > 0.0 ? h : -h)) / e);
            else
                mu -= e2 / (td + (td > 0.0 ? h : -h));
        }
        double x = diag[start] - mu;
        double z = subdiag[start];
        for (int k = start; k < end && z != 0.0; ++k) {
            const Givens rot = make_givens(x, z);
            const double c = rot.c, s = rot.s;
            const double sdk = s * diag[k] + c * subdiag[k];
            const double dkp1 = s * subdiag[k] + c * diag[k + 1];
            diag[k] = c * (c * diag[k] - s * subdiag[k]) - s * (c * subdiag[k] - s * diag[k + 1]);
            diag[k + 1] = s * sdk + c * dkp1;
            subdiag[k] = c * sdk - s * dkp1;
            if (k > start) subdiag[k - 1] = c * subdiag[k - 1] - s * z;
            x = subdiag[k];
            if (k < end - 1) {
                z = -s * subdiag[k + 1];
                subdiag[k + 1] = c * subdiag[k + 1];
            }
            // Q = Q * G : applyOnTheRight(k, k+1, rot) -> x' = c x - s y ; y' = s x + c y
            for (int i = 0; i < n; ++i) {
                const double xi = q[i][k], yi = q[i][k + 1];
                q[i][k] = c * xi - s * yi;
                q[i][k + 1] = s * xi + c * yi;
            }
        }
    }
    const bool ok = iter <= maxIterations * n;
    if (ok) {
        for (int i = 0; i < n - 1; ++i) {
            int k = 0;
            double mn = diag[i];
            for (int j = 1; j < n - i; ++j)
                if (diag[i + j] < mn) {
                    mn = diag[i + j];
                    k = j;
                }
            if (k > 0) {
                std::swap(diag[i], diag[k + i]);
                for (int r = 0; r < n; ++r) std::swap(q[r][i], q[r][k + i]);
            }
        }
    }
    for (int i = 0; i < 3; ++i) {
        evals[i] = diag[i] * scale;
        for (int j = 0; j < 3; ++j) evecs(i, j) = q[i][j];
    }
    return ok;
}

// Quaternion (w,x,y,z) -> rotation matrix. Eigen/src/Geometry/Quaternion.h toRotationMatrix().
inline Mat3 quat_to_rot(const double q[4])
{
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    Mat3 r;
    r(0, 0) = 1.0 - (tyy + tzz);
    r(0, 1) = txy - twz;
    r(0, 2) = txz + twy;
    r(1, 0) = txy + twz;
    r(1, 1) = 1.0 - (txx + tzz);
    r(1, 2) = tyz - twx;
    r(2, 0) = txz - twy;
    r(2, 1) = tyz + twx;
    r(2, 2) = 1.0 - (txx + tyy);
    return r;
}

// Quaternion product a*b, (w,x,y,z).
inline void quat_mul(const double a[4], const double b[4], double out[4])
{
    out[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    out[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    out[2] = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
    out[3] = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
}

// Eigen/src/Geometry/EulerAngles.h: MatrixBase::eulerAngles(0,1,2) (Eigen 3.4: first angle in [0,pi]).
inline void euler_angles_012(const Mat3& m, double res[3])
{
    const int i = 0, j = 1, k = 2;  // odd = 0
    res[0] = std::atan2(m(j, k), m(k, k));
    const double c2 = std::sqrt(m(i, i) * m(i, i) + m(i, j) * m(i, j));
    if (res[0] > 0.0) {
        res[0] -= M_PI;
        res[1] = std::atan2(-m(i, k), -c2);
    }
    else {
        res[1] = std::atan2(-m(i, k), c2);
    }
    const double s1 = std::sin(res[0]);
    const double c1 = std::cos(res[0]);
    res[2] = std::atan2(s1 * m(k, i) - c1 * m(j, i), c1 * m(j, j) - s1 * m(k, j));
    res[0] = -res[0];
    res[1] = -res[1];
    res[2] = -res[2];
}

}  // namespace oracle
