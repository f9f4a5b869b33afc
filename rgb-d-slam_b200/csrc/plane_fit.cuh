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

    double diag[3], subdiag[2];
    diag[0] = m00;
    const double v1norm2 = m20 * m20;
    if (v1norm2 <= DBL_MIN) {
        diag[1] = m11;
        diag[2] = m22;
        subdiag[0] = m10;
        subdiag[1] = m21;
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
        diag[1] = m11 + m02 * qq;
        diag[2] = m22 - m02 * qq;
        subdiag[0] = beta;
        subdiag[1] = m21 - m01 * qq;
        q[0][0] = 1, q[0][1] = 0, q[0][2] = 0;
        q[1][0] = 0, q[1][1] = m01, q[1][2] = m02;
        q[2][0] = 0, q[2][1] = m02, q[2][2] = -m01;
    }

    const int n = 3;
    const int maxIterations = 30;
    int end = n - 1, start = 0, iter = 0;
    const double precision_inv = 1.0 / DBL_EPSILON;
    while (end > 0) {
        for (int i = start; i < end; ++i) {
            if (fabs(subdiag[i]) < DBL_MIN) {
                subdiag[i] = 0.0;
            }
            else {
                const double scaled_subdiag = precision_inv * subdiag[i];
                if (scaled_subdiag * scaled_subdiag <= (fabs(diag[i]) + fabs(diag[i + 1]))) subdiag[i] = 0.0;
            }
        }
        while (end > 0 && subdiag[end - 1] == 0.0) end--;
        if (end <= 0) break;
        iter++;
        if (iter > maxIterations * n) break;
        start = end - 1;
        while (start > 0 && subdiag[start - 1] != 0.0) start--;

        const double td = (diag[end - 1] - diag[end]) * 0.5;
        const double e = subdiag[end - 1];
        double mu = diag[end];
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
        double x = diag[start] - mu;
        double z = subdiag[start];
        for (int k = start; k < end && z != 0.0; ++k) {
            double c, s;
            make_givens(x, z, c, s);
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
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double xi = q[i][k], yi = q[i][k + 1];
                q[i][k] = c * xi - s * yi;
                q[i][k + 1] = s * xi + c * yi;
            }
        }
    }
    if (iter <= maxIterations * n) {
        for (int i = 0; i < n - 1; ++i) {
            int k = 0;
            double mn = diag[i];
            for (int j = 1; j < n - i; ++j)
                if (diag[i + j] < mn) {
                    mn = diag[i + j];
                    k = j;
                }
            if (k > 0) {
                const double t = diag[i];
                diag[i] = diag[k + i];
                diag[k + i] = t;
                for (int r = 0; r < 3; ++r) {
                    const double u = q[r][i];
                    q[r][i] = q[r][k + i];
                    q[r][k + i] = u;
                }
            }
        }
    }
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
