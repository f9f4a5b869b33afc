// Device-side plane model of CAPE: 9 FP64 sums + count -> (centroid, normal, d, MSE, score).
// Replaces Plane_Segment::fit_plane / get_point_cloud_Huygen_covariance / can_be_merged
// (src/features/primitives/plane_segment.cpp:205-284,322-326) and Eigen::SelfAdjointEigenSolver<Matrix3d>
// (iterative path: scaling, closed-form 3x3 tridiagonalisation, implicit symmetric QR with Wilkinson shift).
// Compiled with -fmad=false so every operation rounds exactly like the reference's x86-64 (SSE2, no FMA) build.
#pragma once
#include <float.h>

#include "common.cuh"

namespace rs {

struct PlaneModel {
    int count;
    int planar;
    double S[9];  // Sx Sy Sz Sxs Sys Szs Sxy Syz Szx
    double c[3];  // centroid
    double n[3];  // normal
    double d;
    double mse;
    double score;
};

__device__ __forceinline__ void plane_clear(PlaneModel& p)
{
    p.count = 0;
    p.planar = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) p.S[i] = 0.0;
    p.c[0] = p.c[1] = p.c[2] = 0.0;
    p.n[0] = p.n[1] = p.n[2] = 0.0;
    p.d = 0.0;
    p.mse = DBL_MAX;
    p.score = 0.0;
}

__device__ __forceinline__ double dot3(const double* a, const double* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

// Eigen normalized(): divide by sqrt(squaredNorm) when squaredNorm > 0
__device__ __forceinline__ void normalize3(double* v)
{
    const double z = dot3(v, v);
    if (z > 0.0) {
        const double s = sqrt(z);
        v[0] = v[0] / s;
        v[1] = v[1] / s;
        v[2] = v[2] / s;
    }
}

// JacobiRotation<double>::makeGivens (Eigen/src/Jacobi/Jacobi.h)
__device__ __forceinline__ void make_givens(const double p, const double q, double& c, double& s)
{
    if (q == 0.0) {
        c = p < 0.0 ? -1.0 : 1.0;
        s = 0.0;
    }
    else if (p == 0.0) {
        c = 0.0;
        s = q < 0.0 ? 1.0 : -1.0;
    }
    else if (fabs(p) > fabs(q)) {
        const double t = q / p;
        double u = sqrt(1.0 + t * t);
        if (p < 0.0) u = -u;
        c = 1.0 / u;
        s = -t * c;
    }
    else {
        const double t = p / q;
        double u = sqrt(1.0 + t * t);
        if (q < 0.0) u = -u;
        s = -1.0 / u;
        c = -t * s;
    }
}

__device__ __forceinline__ double eigen_hypot(double x, double y)
{
    x = fabs(x);
    y = fabs(y);
    const double p = x > y ? x : y;
    if (p == 0.0) return 0.0;
    const double qp = (y < x ? y : x) / p;
    return p * sqrt(1.0 + qp * qp);
}

// Symmetric 3x3 eigen-decomposition, ascending eigenvalues; only the lower triangle is read.
// a = {a00, a10, a11, a20, a21, a22}. q[r][c] = component r of eigenvector c.
__device__ inline void self_adjoint_eigen3(const double a00, const double a10, const double a11, const double a20,
                                           const double a21, const double a22, double ev[3], double q[3][3])
{
    double scale = fabs(a00);
    scale = fmax(scale, fabs(a10));
    scale = fmax(scale, fabs(a11));
    scale = fmax(scale, fabs(a20));
    scale = fmax(scale, fabs(a21));
    scale = fmax(scale, fabs(a22));
    if (scale == 0.0) scale = 1.0;
    const double m00 = a00 / scale, m10 = a10 / scale, m11 = a11 / scale, m20 = a20 / scale, m21 = a21 / scale,
                 m22 = a22 / scale;

    // The tridiagonal QL below is Eigen's loop (SelfAdjointEigenSolver.h: computeFromTridiagonal_impl +
    // tridiagonal_qr_step) written out for n = 3 with every index a compile-time constant: diag / subdiag / q stay in
    // registers (the dynamically indexed version lived in local memory), the operations and their order are unchanged.
    double d0 = m00, d1, d2, e0, e1;
    const double v1norm2 = m20 * m20;
    if (v1norm2 <= DBL_MIN) {
        d1 = m11;
        d2 = m22;
        e0 = m10;
        e1 = m21;
        q[0][0] = 1, q[0][1] = 0, q[0][2] = 0;
        q[1][0] = 0, q[1][1] = 1, q[1][2] = 0;
        q[2][0] = 0, q[2][1] = 0, q[2][2] = 1;
    }
    else {
        const double beta = sqrt(m10 * m10 + v1norm2);
        const double invBeta = 1.0 / beta;
        const double m01 = m10 * invBeta;
        const double m02 = m20 * invBeta;
        const double qq = 2.0 * m01 * m21 + m02 * (m22 - m11);
        d1 = m11 + m02 * qq;
        d2 = m22 - m02 * qq;
        e0 = beta;
        e1 = m21 - m01 * qq;
        q[0][0] = 1, q[0][1] = 0, q[0][2] = 0;
        q[1][0] = 0, q[1][1] = m01, q[1][2] = m02;
        q[2][0] = 0, q[2][1] = m02, q[2][2] = -m01;
    }

    const int n = 3;
    const int maxIterations = 30;
    int end = n - 1, start = 0, iter = 0;
    const double precision_inv = 1.0 / DBL_EPSILON;
// subdiag[i] -> 0 when negligible against its diagonal neighbours
#define RS_DEFLATE(E, DA, DB)                                                         \
    {                                                                                 \
        if (fabs(E) < DBL_MIN) {                                                      \
            E = 0.0;                                                                  \
        }                                                                             \
        else {                                                                        \
            const double scaled_subdiag = precision_inv * E;                          \
            if (scaled_subdiag * scaled_subdiag <= (fabs(DA) + fabs(DB))) E = 0.0;    \
        }                                                                             \
    }
// one Givens step of the implicit QR sweep on rows / columns K, K+1 (DK = diag[K], DK1 = diag[K+1], EK = subdiag[K])
#define RS_GIVENS_STEP(K, DK, DK1, EK, PREV_STMT, NEXT_STMT)                          \
    {                                                                                 \
        double c, s;                                                                  \
        make_givens(x, z, c, s);                                                      \
        const double sdk = s * DK + c * EK;                                           \
        const double dkp1 = s * EK + c * DK1;                                         \
        DK = c * (c * DK - s * EK) - s * (c * EK - s * DK1);                          \
        DK1 = s * sdk + c * dkp1;                                                     \
        EK = c * sdk - s * dkp1;                                                      \
        PREV_STMT;                                                                    \
        x = EK;                                                                       \
        NEXT_STMT;                                                                    \
        _Pragma("unroll") for (int i = 0; i < 3; ++i)                                 \
        {                                                                             \
            const double xi = q[i][K], yi = q[i][K + 1];                              \
            q[i][K] = c * xi - s * yi;                                                \
            q[i][K + 1] = s * xi + c * yi;                                            \
        }                                                                             \
    }
    while (end > 0) {
        if (start <= 0 && 0 < end) RS_DEFLATE(e0, d0, d1)
        if (start <= 1 && 1 < end) RS_DEFLATE(e1, d1, d2)
        if (end == 2 && e1 == 0.0) end = 1;
        if (end == 1 && e0 == 0.0) end = 0;
        if (end <= 0) break;
        iter++;
        if (iter > maxIterations * n) break;
        start = end - 1;
        if (start == 1 && e0 != 0.0) start = 0;

        const double dem1 = end == 2 ? d1 : d0, de = end == 2 ? d2 : d1;
        const double td = (dem1 - de) * 0.5;
        const double e = end == 2 ? e1 : e0;
        double mu = de;
        if (td == 0.0) {
            mu -= fabs(e);
        }
        else if (e != 0.0) {
            const double e2 = e * e;
            const double h = eigen_hypot(td, e);
            if (e2 == 0.0)
                mu -= e / ((td + (td > 0.0 ? h : -h)) / e);
            else
                mu -= e2 / (td + (td > 0.0 ? h : -h));
        }
        double x = (start == 0 ? d0 : d1) - mu;
        double z = start == 0 ? e0 : e1;
        if (start == 0) {
            // k = 0 (k > start: no; k < end - 1 iff end == 2), then k = 1 when end == 2 and the bulge is still there
            if (z != 0.0) RS_GIVENS_STEP(0, d0, d1, e0, (void)0, if (end == 2) { z = -s * e1; e1 = c * e1; })
            if (end == 2 && z != 0.0) RS_GIVENS_STEP(1, d1, d2, e1, e0 = c * e0 - s * z, (void)0)
        }
        else if (z != 0.0) {
            // start == 1, end == 2: the single step k = 1
            RS_GIVENS_STEP(1, d1, d2, e1, (void)0, (void)0)
        }
    }
#undef RS_DEFLATE
#undef RS_GIVENS_STEP
    if (iter <= maxIterations * n) {
        // ascending selection sort (Eigen: for i, k = argmin of the tail, strict <), eigenvector columns follow
        int k = 0;
        double mn = d0;
        if (d1 < mn) {
            mn = d1;
            k = 1;
        }
        if (d2 < mn) k = 2;
        if (k == 1) {
            const double t = d0;
            d0 = d1, d1 = t;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double u = q[r][0];
                q[r][0] = q[r][1], q[r][1] = u;
            }
        }
        else if (k == 2) {
            const double t = d0;
            d0 = d2, d2 = t;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double u = q[r][0];
                q[r][0] = q[r][2], q[r][2] = u;
            }
        }
        if (d2 < d1) {
            const double t = d1;
            d1 = d2, d2 = t;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double u = q[r][1];
                q[r][1] = q[r][2], q[r][2] = u;
            }
        }
    }
    const double diag[3] = {d0, d1, d2};
    ev[0] = diag[0] * scale;
    ev[1] = diag[1] * scale;
    ev[2] = diag[2] * scale;
}

// Plane_Segment::fit_plane (plane_segment.cpp:232-284): updates c, and (unless degenerate) n, d, mse, score.
__device__ inline void plane_fit(PlaneModel& p)
{
    p.planar = 0;
    const double o = 1.0 / static_cast<double>(p.count);
    const double Sx = p.S[0], Sy = p.S[1], Sz = p.S[2], Sxs = p.S[3], Sys = p.S[4], Szs = p.S[5], Sxy = p.S[6],
                 Syz = p.S[7], Szx = p.S[8];
    p.c[0] = Sx * o;
    p.c[1] = Sy * o;
    p.c[2] = Sz * o;
    const double xx = fmax(0.0, Sxs - (Sx * Sx) * o);
    const double yy = fmax(0.0, Sys - (Sy * Sy) * o);
    const double zz = fmax(0.0, Szs - (Sz * Sz) * o);
    const double xy = Sxy - Sx * Sy * o;
    const double xz = Szx - Sx * Sz * o;
    const double yz = Syz - Sy * Sz * o;
    // Eigen 3x3 determinant: bruteforce_det3_helper(0,1,2) - (1,0,2) + (2,0,1)
    const double det = xx * (yy * zz - yz * yz) - xy * (xy * zz - yz * xz) + xz * (xy * yz - yy * xz);
    if (fabs(det - 0.0) <= DBL_EPSILON) return;

    double ev[3], q[3][3];
    self_adjoint_eigen3(xx, xy, yy, xz, yz, zz, ev, q);
    const double l0 = fabs(ev[0]), l1 = fabs(ev[1]);
    double n[3] = {q[0][0], q[1][0], q[2][0]};
    normalize3(n);
    const double dd = -dot3(n, p.c);
    if (dd <= 0) {
        n[0] = -n[0], n[1] = -n[1], n[2] = -n[2];
        p.d = -dd;
    }
    else {
        p.d = dd;
    }
    // PlaneCoordinates ctor + operator= both re-normalise (plane_coordinates.hpp:23,30-40)
    normalize3(n);
    normalize3(n);
    p.n[0] = n[0], p.n[1] = n[1], p.n[2] = n[2];
    p.mse = l0 * o;
    p.score = l1 / fmax(l0, 1e-6);
    p.planar = 1;
}

// Plane_Segment::can_be_merged(a -> b). cosMerge = cos(18 deg) evaluated on the host in FP64.
__device__ __forceinline__ bool plane_can_merge(const double* an, const double ad, const double* bn, const double* bc,
                                                const double maxDist, const double cosMerge)
{
    return dot3(an, bn) > cosMerge && fabs(dot3(an, bc) + ad) < maxDist;
}

}  // namespace rs
