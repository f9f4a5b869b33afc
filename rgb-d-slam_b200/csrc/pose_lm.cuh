// Device code shared by the pose kernels (pose_solve.cu: preparation + fused solve kernel; pose_chain.cu: the three-launch
// chain): feature residuals / Jacobians, the one-warp Levenberg-Marquardt (MINPACK lmder / lmpar on the normal equations),
// the counter-based generator of RS_RNG_DEVICE and the covariance reduction. Included inside an anonymous namespace of each
// translation unit.
#pragma once
#include <float.h>

#include <algorithm>

#include "plane_fit.cuh"  // normalize3
#include "pose_internal.cuh"

namespace rs {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr double kSqrtEps = 1.4901161193847656e-08;  // sqrt(DBL_EPSILON): ftol, xtol and the difference step
constexpr int kRunning = -100;

// parameters.hpp:23-44
constexpr double kPointInlierPx = 3.0;                       // float 3.0f
constexpr double kPlaneInlierMm = 50.0;                      // float 50.0f
constexpr double kPlaneInlierNormal = 0.20000000298023224;   // float 0.2f
constexpr double kEarlyStopProportion = 0.800000011920929;   // double initialised from 0.80f
constexpr double kPointScore = 1.0 / 5.0;                    // 1 / minimumPointForOptimization
constexpr double kPlaneScore = 1.0 / 3.0;                    // 1 / minimumPlanesForOptimization
constexpr double kPoint2dScore = 1.0 / 5.0;                  // 1 / minimumPoint2dForOptimization
constexpr double kPoint2dInlierPx = 3.0;                     // float 3.0f
constexpr double kPoint2dWeight = 0.3 / 2.0;                 // get_alpha_reduction() / parts (map_point2d.cpp:25,47)

// IOptimizationFeature::get_score / get_feature_part_count per feature type
__device__ __forceinline__ double score_of(const int type)
{
    return type == RS_FEAT_PLANE ? kPlaneScore : (type == RS_FEAT_POINT2D ? kPoint2dScore : kPointScore);
}
__device__ __forceinline__ int parts_of(const int type) { return type == RS_FEAT_PLANE ? 3 : 2; }

// world <- camera rotation of the pose: R' = C * R(q), t' = C * t with C = [[0,0,1],[-1,0,0],[0,-1,0]]
// (camera_transformation.cpp:11-23). world->camera is then p_c = R'^T (P - t') and the plane world->camera map
// (inverse of [[R',0],[-t'^T R',1]], :52-71) is n_c = R'^T n_w, d_c = t'.n_w + d_w.
struct Xform {
    double R[9];
    double t[3];
};

struct WarpLM {
    double x[6], xt[6], diag[6], p[6], wa2[6], sc[6], g[6], xs[6];
    double A[36];     // J^T J (full, symmetric)
    double C[21];     // packed lower triangle of the column-scaled S A S
    Xform T;          // transform at the point being evaluated (x, then the trial points)
    double Rk[3][9];  // R'(x + h_k e_k) for the three rotation coefficients (scratch of the Jacobian set-up)
    double dR[27];    // forward-difference derivative of R' along the three rotation coefficients
    double ih[3];     // 1 / h_k of those differences
    double fnorm, par, delta, xnorm, gnorm, pnorm;
    double ss1;       // |f|^2 at the trial point (hand-over between the warp that evaluates it and the lane that judges it)
    int status, nfev, iter, again;
};

struct Problem {
    int n;
    const short* idx;
    const int32_t* type;  // [M]
    const double* obs;    // [4][M]
    const double* map;    // [4][M]
    const double* aux;    // [4][M] global: first observation + inverse depth of the inverse-depth (point2d) features
    int M;
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// levenberg_marquardt_functors.cpp:29-38,82-86 + PoseBase normalisation (pose.cpp:18-22)
__device__ inline void quaternion_from_coefficients(const double* x, double q[4])
{
    const double alpha = x[3] * x[3] + x[4] * x[4] + x[5] * x[5];
    const double divider = 1.0 / (alpha + 1.0);
    q[0] = 2.0 * x[3] * divider;
    q[1] = 2.0 * x[4] * divider;
    q[2] = 2.0 * x[5] * divider;
    q[3] = (1.0 - alpha) * divider;
    // PoseBase normalises the quaternion it is given (pose.cpp:18-22); one reciprocal square root instead of a square
    // root and four divisions (this runs on the serial lane-0 path of every LM iteration)
    const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    if (n2 > 0.0) {
        const double inn = rsqrt(n2);
        q[0] *= inn, q[1] *= inn, q[2] *= inn, q[3] *= inn;
    }
}

// Quaternion (w,x,y,z) -> rotation (Eigen toRotationMatrix), row-major
__device__ inline void quat_to_rot(const double q[4], double r[9])
{
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    r[0] = 1.0 - (tyy + tzz), r[1] = txy - twz, r[2] = txz + twy;
    r[3] = txy + twz, r[4] = 1.0 - (txx + tzz), r[5] = tyz - twx;
    r[6] = txz - twy, r[7] = tyz + twx, r[8] = 1.0 - (txx + tyy);
}

__device__ inline void make_xform(const double* x, Xform& T)
{
    double q[4], r[9];
    quaternion_from_coefficients(x, q);
    quat_to_rot(q, r);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        T.R[0 + j] = r[6 + j];
        T.R[3 + j] = -r[0 + j];
        T.R[6 + j] = -r[3 + j];
    }
    T.t[0] = x[2], T.t[1] = -x[0], T.t[2] = -x[1];
}

// levenberg_marquardt_functors.cpp:14-27,74-80
__device__ inline void coefficients_from_pose(const double* pose7, double x[6])
{
    x[0] = pose7[0], x[1] = pose7[1], x[2] = pose7[2];
    const double divider = 1.0 / fmax(1.0 + pose7[6], 0.001);
    x[3] = pose7[3] * divider;
    x[4] = pose7[4] * divider;
    x[5] = pose7[5] * divider;
}

// MatrixBase::eulerAngles(0,1,2) of R(q) (PoseBase::get_vector, pose.hpp:30-35)
__device__ inline void pose_vector6(const double* x, double v[6])
{
    double q[4], m[9];
    quaternion_from_coefficients(x, q);
    quat_to_rot(q, m);
    v[0] = x[0], v[1] = x[1], v[2] = x[2];
    double r0 = atan2(m[5], m[8]), r1;
    const double c2 = sqrt(m[0] * m[0] + m[1] * m[1]);
    if (r0 > 0.0) {
        r0 -= kPi;
        r1 = atan2(-m[2], -c2);
    }
    else {
        r1 = atan2(-m[2], c2);
    }
    const double s1 = sin(r0), c1 = cos(r0);
    const double r2 = atan2(s1 * m[6] - c1 * m[3], c1 * m[4] - s1 * m[7]);
    v[3] = -r0, v[4] = -r1, v[5] = -r2;
}

// WorldCoordinate::get_signed_distance_2D_px (point_coordinates.cpp:245-260)
__device__ __forceinline__ void point_distance(const double o0, const double o1, const double X, const double Y,
                                               const double Z, const Xform& T, const PoseIntrinsics& K, double& du,
                                               double& dv)
{
    const double d0 = X - T.t[0], d1 = Y - T.t[1], d2 = Z - T.t[2];
    const double xc = (T.R[0] * d0 + T.R[3] * d1) + T.R[6] * d2;
    const double yc = (T.R[1] * d0 + T.R[4] * d1) + T.R[7] * d2;
    const double zc = (T.R[2] * d0 + T.R[5] * d1) + T.R[8] * d2;
    const double inv = 1.0 / zc;
    const double u = inv * (K.fx * xc + K.cx * zc);
    const double v = inv * (K.fy * yc + K.cy * zc);
    if (u != u || v != v) {
        du = DBL_MAX, dv = DBL_MAX;
        return;
    }
    du = o0 - u;
    dv = o1 - v;
}

// PlaneWorldCoordinates::to_camera_coordinates (plane_coordinates.cpp:20-24): n renormalised, d kept
__device__ __forceinline__ void plane_to_camera(const double n0, const double n1, const double n2, const double dw,
                                                const Xform& T, double np[3], double& dp)
{
    double v0 = (T.R[0] * n0 + T.R[3] * n1) + T.R[6] * n2;
    double v1 = (T.R[1] * n0 + T.R[4] * n1) + T.R[7] * n2;
    double v2 = (T.R[2] * n0 + T.R[5] * n1) + T.R[8] * n2;
    const double z = (v0 * v0 + v1 * v1) + v2 * v2;
    if (z > 0.0) {
        const double is = rsqrt(z);
        v0 *= is, v1 *= is, v2 *= is;
    }
    np[0] = v0, np[1] = v1, np[2] = v2;
    dp = ((T.t[0] * n0 + T.t[1] * n1) + T.t[2] * n2) + dw;
}

// Point2dOptimizationFeature::get_distance (map_point2d.cpp:40-45) -> InverseDepthWorldPoint::compute_signed_screen_distance
// (inverse_depth_coordinates.cpp:58-68,142-173) -> Segment<2>::distance (line.hpp:27-41,95-99): the matched pixel against
// the screen line through the projections of the point's furthest / closest depth estimates. Device layout of the feature:
// o = (u, v, dFar, dNear) with d* = min(inverse depth -/+ 3 sqrt(sigma), 1e-9) formed once by the preparation step,
// m = (theta, phi) (the part the Monte-Carlo variation perturbs), ax = (first observation X, Y, Z).
__device__ __forceinline__ void point2d_distance(const double o[4], const double m[4], const double ax[3], const Xform& T,
                                                 const PoseIntrinsics& K, double& du, double& dv)
{
    double st, ct, sp, cp;
    sincos(m[0], &st, &ct);
    sincos(m[1], &sp, &cp);
    const double b0 = 1.0 * st * cp, b1 = 1.0 * st * sp, b2 = 1.0 * ct;
    double s[2], e[2];
    bool ok = true;
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        const double den = o[2 + w];
        const double d0 = (ax[0] + b0 / den) - T.t[0], d1 = (ax[1] + b1 / den) - T.t[1], d2 = (ax[2] + b2 / den) - T.t[2];
        const double xc = (T.R[0] * d0 + T.R[3] * d1) + T.R[6] * d2;
        const double yc = (T.R[1] * d0 + T.R[4] * d1) + T.R[7] * d2;
        const double zc = (T.R[2] * d0 + T.R[5] * d1) + T.R[8] * d2;
        const double inv = 1.0 / zc;
        const double u = inv * (K.fx * xc + K.cx * zc), v = inv * (K.fy * yc + K.cy * zc);
        ok = ok && !(u != u || v != v);
        if (w == 0)
            s[0] = u, s[1] = v;
        else
            e[0] = u, e[1] = v;
    }
    if (!ok) {
        du = DBL_MAX, dv = DBL_MAX;
        return;
    }
    double n0 = e[0] - s[0], n1 = e[1] - s[1];
    const double z = n0 * n0 + n1 * n1;
    if (z > 0.0) {
        const double l = sqrt(z);
        n0 /= l, n1 /= l;
    }
    const double along = (o[0] - s[0]) * n0 + (o[1] - s[1]) * n1;
    du = o[0] - (s[0] + n0 * along);
    dv = o[1] - (s[1] + n1 * along);
}

// One feature's residual entries (Global_Pose_Estimator::operator(), levenberg_marquardt_functors.cpp:128-169):
// point -> 1/2 (du, dv); plane -> 1/3 (d_c n_c - d_p n_p). Returns the entry count.
// `aux` / `M` / `gi`: where an inverse-depth feature finds its first observation (global memory, component-major).
template <bool P2D>
__device__ __forceinline__ int feature_residual(const int type, const double o[4], const double m[4], const Xform& T,
                                                const PoseIntrinsics& K, double r[3], const double* aux, const int M,
                                                const int gi)
{
    if (type == RS_FEAT_POINT) {
        double du, dv;
        point_distance(o[0], o[1], m[0], m[1], m[2], T, K, du, dv);
        r[0] = du * 1.0 / 2.0;
        r[1] = dv * 1.0 / 2.0;
        r[2] = 0.0;
        return 2;
    }
    if (P2D && type == RS_FEAT_POINT2D) {
        const double ax[3] = {aux[gi], aux[M + gi], aux[2 * M + gi]};
        double du, dv;
        point2d_distance(o, m, ax, T, K, du, dv);
        r[0] = du * kPoint2dWeight;
        r[1] = dv * kPoint2dWeight;
        r[2] = 0.0;
        return 2;
    }
    double np[3], dp;
    plane_to_camera(m[0], m[1], m[2], m[3], T, np, dp);
    r[0] = (o[3] * o[0] - dp * np[0]) * 1.0 / 3.0;
    r[1] = (o[3] * o[1] - dp * np[1]) * 1.0 / 3.0;
    r[2] = (o[3] * o[2] - dp * np[2]) * 1.0 / 3.0;
    return 3;
}

// distance_utils.cpp:6-9: atan2(sin(a - b), cos(a - b)). The arguments are components of unit normals, so |a - b| <= 2 < pi
// and the expression is a - b up to a few ulp; the libm chain (three FP64 transcendentals per component) is only
// evaluated when that could decide the comparison with the threshold.
__device__ __noinline__ double angle_distance_exact(const double a, const double b) { return atan2(sin(a - b), cos(a - b)); }
__device__ __forceinline__ bool angle_within(const double a, const double b, const double thr)
{
    const double d = fabs(a - b);
    if (fabs(d - thr) > 1e-9) return d <= thr;
    return fabs(angle_distance_exact(a, b)) <= thr;
}

// IOptimizationFeature::is_inlier (map_point.cpp:34-38, map_primitive.cpp:33-49)
template <bool P2D>
__device__ __forceinline__ bool feature_is_inlier(const int type, const double o[4], const double m[4], const Xform& T,
                                                  const PoseIntrinsics& K, const double* aux, const int M, const int gi)
{
    if (type == RS_FEAT_POINT) {
        double du, dv;
        point_distance(o[0], o[1], m[0], m[1], m[2], T, K, du, dv);
        const double dist = (du >= DBL_MAX || dv >= DBL_MAX) ? DBL_MAX : fabs(du) + fabs(dv);
        return dist <= kPointInlierPx;
    }
    if (P2D && type == RS_FEAT_POINT2D) {
        // map_point2d.cpp:33-38: (get_distance().array() <= threshold).all() - on the SIGNED distance
        const double ax[3] = {aux[gi], aux[M + gi], aux[2 * M + gi]};
        double du, dv;
        point2d_distance(o, m, ax, T, K, du, dv);
        return du <= kPoint2dInlierPx && dv <= kPoint2dInlierPx;
    }
    double np[3], dp;
    plane_to_camera(m[0], m[1], m[2], m[3], T, np, dp);
    return angle_within(o[0], np[0], kPlaneInlierNormal) && angle_within(o[1], np[1], kPlaneInlierNormal) &&
           angle_within(o[2], np[2], kPlaneInlierNormal) && fabs(o[3] - dp) <= kPlaneInlierMm;
}

__device__ __forceinline__ int load_feature(const Problem& P, const int k, int& type, double o[4], double m[4])
{
    const int i = P.idx ? int(P.idx[k]) : k;
    type = P.type[i];
    o[0] = P.obs[i], o[1] = P.obs[P.M + i];
    m[0] = P.map[i], m[1] = P.map[P.M + i], m[2] = P.map[2 * P.M + i];
    if (type == RS_FEAT_POINT) {   // a point uses (u, v) and (X, Y, Z) only
        o[2] = 0.0, o[3] = 0.0, m[3] = 0.0;
    }
    else {
        o[2] = P.obs[2 * P.M + i], o[3] = P.obs[3 * P.M + i];
        m[3] = P.map[3 * P.M + i];
    }
    return i;
}

// |f(x)|^2 over the problem's features with transform T (warp-wide result)
template <bool P2D>
__device__ inline double eval_sumsq(const Problem& P, const Xform& T, const PoseIntrinsics& K, const int lane)
{
    double ss = 0.0;
    for (int k = lane; k < P.n; k += 32) {
        int type;
        double o[4], m[4], r[3];
        const int gi = load_feature(P, k, type, o, m);
        feature_residual<P2D>(type, o, m, T, K, r, P.aux, P.M, gi);
        ss += r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    }
    return warp_sum(ss);
}

// ---- Jacobian -------------------------------------------------------------------------------------------------------
// Eigen::NumericalDiff<F, Forward> differentiates the residual vector: column j = (f(x + h_j e_j) - f(x)) / h_j with
// h_j = sqrt(eps) |x_j|. The residuals depend on x only through the transform (R', t'), t' is LINEAR in x0..x2 and R'
// depends on x3..x5 only, so the same forward difference is taken here one level down - on the transform instead of on
// every residual:  dR'_k = (R'(x + h_k e_k) - R'(x)) / h_k  (three 3x3 matrices per Jacobian, built once per iteration by
// three lanes) - and carried to the residual rows by the chain rule per feature. The translation columns are exact; the
// rotation columns differ from the reference's by its O(h) truncation term (1e-8 relative) which is two orders below
// the rounding noise (1e-6 relative: residuals of ~100 px known to 1e-14, divided by h ~ 1e-8) that ANY evaluation order
// of the reference's own difference quotient carries. Cost per point: one projection + ~60 FMA instead of seven
// projections. nfev still counts the 7 evaluations NumericalDiff would have made, so stop code 5 fires at the same place.
__device__ __forceinline__ void accumulate_row(const double (&J)[6], const double r, double (&a)[32])
{
    int t = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        a[21 + i] += J[i] * r;
#pragma unroll
        for (int j = i; j < 6; ++j) a[t++] += J[i] * J[j];
    }
}

// rows of one feature at S.T / S.dR, accumulated into a[0..20] (upper triangle of J^T J, row-major) and a[21..26] (J^T r)
// Inverse-depth features keep NumericalDiff's own scheme (seven residual evaluations): their residual is a point-to-line
// distance through two projections, they are rare (the live pipeline never produces them), and S still holds what the
// perturbed transforms are made of.
__device__ __noinline__ void point2d_jacobian(const double (&o)[4], const double (&m)[4], const double* aux, const int M,
                                              const int gi, const WarpLM& S, const PoseIntrinsics& K, double* c /* [27] contribution */)
{
    const double ax[3] = {aux[gi], aux[M + gi], aux[2 * M + gi]};
    double du, dv;
    point2d_distance(o, m, ax, S.T, K, du, dv);
    const double r0 = du * kPoint2dWeight, r1 = dv * kPoint2dWeight;
    double J0[6], J1[6];
    for (int j = 0; j < 6; ++j) {
        Xform Tj = S.T;
        double ih;
        if (j < 3) {
            double h = kSqrtEps * fabs(S.x[j]);
            if (h == 0.0) h = kSqrtEps;
            const double xj = S.x[j] + h;
            if (j == 0) Tj.t[1] = -xj;       // t' = (x2, -x0, -x1)
            else if (j == 1) Tj.t[2] = -xj;
            else Tj.t[0] = xj;
            ih = 1.0 / h;
        }
        else {
            for (int i = 0; i < 9; ++i) Tj.R[i] = S.Rk[j - 3][i];
            ih = S.ih[j - 3];
        }
        double eu, ev;
        point2d_distance(o, m, ax, Tj, K, eu, ev);
        J0[j] = (eu * kPoint2dWeight - r0) * ih;
        J1[j] = (ev * kPoint2dWeight - r1) * ih;
    }
    int t = 0;
    for (int i = 0; i < 6; ++i) {
        c[21 + i] = J0[i] * r0 + J1[i] * r1;
        for (int j = i; j < 6; ++j) c[t++] = J0[i] * J0[j] + J1[i] * J1[j];
    }
}

// dR' entry i of rotation coefficient k, wherever the caller keeps the 27 of them (an array; per-lane strided shared memory)
struct DRArray {
    const double* p;
    __device__ __forceinline__ double at(const int k, const int i) const { return p[9 * k + i]; }
};

template <class DR>
__device__ __forceinline__ void feature_jacobian(const int type, const double (&o)[4], const double (&m)[4], const Xform& T,
                                                 const DR& dR, const PoseIntrinsics& K, double (&a)[32])
{
    if (type == RS_FEAT_POINT) {
        const double d0 = m[0] - T.t[0], d1 = m[1] - T.t[1], d2 = m[2] - T.t[2];
        const double xc = (T.R[0] * d0 + T.R[3] * d1) + T.R[6] * d2;
        const double yc = (T.R[1] * d0 + T.R[4] * d1) + T.R[7] * d2;
        const double zc = (T.R[2] * d0 + T.R[5] * d1) + T.R[8] * d2;
        const double inv = 1.0 / zc;
        const double u = inv * (K.fx * xc + K.cx * zc);
        const double v = inv * (K.fy * yc + K.cy * zc);
        if (u != u || v != v) return;  // residual DBL_MAX at x and at every x + h: a zero row in the reference too
        const double r0 = (o[0] - u) * 0.5, r1 = (o[1] - v) * 0.5;
        // d r0 = A0 dxc + B0 dzc, d r1 = A1 dyc + B1 dzc
        const double A0 = -0.5 * K.fx * inv, B0 = (0.5 * K.fx * xc) * (inv * inv);
        const double A1 = -0.5 * K.fy * inv, B1 = (0.5 * K.fy * yc) * (inv * inv);
        double J0[6], J1[6];
        // x0 -> d1 += 1, x1 -> d2 += 1, x2 -> d0 -= 1  (t' = (x2, -x0, -x1))
        J0[0] = A0 * T.R[3] + B0 * T.R[5], J1[0] = A1 * T.R[4] + B1 * T.R[5];
        J0[1] = A0 * T.R[6] + B0 * T.R[8], J1[1] = A1 * T.R[7] + B1 * T.R[8];
        J0[2] = -(A0 * T.R[0] + B0 * T.R[2]), J1[2] = -(A1 * T.R[1] + B1 * T.R[2]);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double dx = (dR.at(k, 0) * d0 + dR.at(k, 3) * d1) + dR.at(k, 6) * d2;
            const double dy = (dR.at(k, 1) * d0 + dR.at(k, 4) * d1) + dR.at(k, 7) * d2;
            const double dz = (dR.at(k, 2) * d0 + dR.at(k, 5) * d1) + dR.at(k, 8) * d2;
            J0[3 + k] = A0 * dx + B0 * dz;
            J1[3 + k] = A1 * dy + B1 * dz;
        }
        accumulate_row(J0, r0, a);
        accumulate_row(J1, r1, a);
        return;
    }
    // plane: r = (d_o n_o - d_p n_p) / 3, n_p = R'^T n / |R'^T n|, d_p = t'.n + d_w
    const double v0 = (T.R[0] * m[0] + T.R[3] * m[1]) + T.R[6] * m[2];
    const double v1 = (T.R[1] * m[0] + T.R[4] * m[1]) + T.R[7] * m[2];
    const double v2 = (T.R[2] * m[0] + T.R[5] * m[1]) + T.R[8] * m[2];
    const double z = (v0 * v0 + v1 * v1) + v2 * v2;
    const double is = z > 0.0 ? rsqrt(z) : 1.0;
    const double np[3] = {v0 * is, v1 * is, v2 * is};
    const double dp = ((T.t[0] * m[0] + T.t[1] * m[1]) + T.t[2] * m[2]) + m[3];
    const double third = 1.0 / 3.0;
    double J[3][6], r[3];
    // d d_p / d x0 = -n1, / d x1 = -n2, / d x2 = +n0
    const double c0 = m[1] * third, c1 = m[2] * third, c2 = -m[0] * third;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        r[c] = (o[3] * o[c] - dp * np[c]) * third;
        J[c][0] = c0 * np[c];
        J[c][1] = c1 * np[c];
        J[c][2] = c2 * np[c];
    }
    const double f = -dp * is * third;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double w0 = (dR.at(k, 0) * m[0] + dR.at(k, 3) * m[1]) + dR.at(k, 6) * m[2];
        const double w1 = (dR.at(k, 1) * m[0] + dR.at(k, 4) * m[1]) + dR.at(k, 7) * m[2];
        const double w2 = (dR.at(k, 2) * m[0] + dR.at(k, 5) * m[1]) + dR.at(k, 8) * m[2];
        const double along = (np[0] * w0 + np[1] * w1) + np[2] * w2;  // removed by the renormalisation
        J[0][3 + k] = f * (w0 - along * np[0]);
        J[1][3 + k] = f * (w1 - along * np[1]);
        J[2][3 + k] = f * (w2 - along * np[2]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) accumulate_row(J[c], r[c], a);
}

// Sum over the warp of 32 per-lane values each, by recursive halving: after the five exchange levels lane L holds the
// warp total of entry L in v[0] (31 exchanges instead of the 32 x 5 of a butterfly all-reduce).
__device__ __forceinline__ double reduce_scatter32(double (&v)[32], const int lane)
{
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const double send = up ? v[i] : v[i + half];
            const double keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(FULL, send, half);
        }
    }
    return v[0];
}

// ---- lane-0 algebra on the shared 6x6 state ---------------------------------------------------------------------
// MINPACK's lmder/lmpar work on the triangular factor R of J P = Q R and on Q^T r. Every quantity they need is a
// function of A = J^T J = P R^T R P^T and g = J^T r = P R^T (Q^T r):   the Gauss-Newton step solves A x = g, the
// damped step solves (A + par D^2) x = g, |J p|^2 = p^T A p, (R^T Q^T r)_j = g_perm(j), and the Newton correction of
// lmpar is w^T (A + par D^2)^-1 w. They are evaluated here with LDL^T factorisations (no square roots, six
// reciprocals) of the column-scaled matrix C = S A S, S = diag(1/|J_j|), which removes the mm-vs-quaternion scale
// disparity of the columns before the squared condition number can hurt. Same iterates as lmpar/qrsolv up to rounding.
__device__ __forceinline__ double norm6(const double* v)
{
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) s += v[i] * v[i];
    return sqrt(s);
}

// Packed lower-triangular index; with fully unrolled loops every index is a compile-time constant, so the 6x6
// working set below lives in registers (no local-memory round trips on the serial lane-0 path).
#define RS_T(i, j) ((i) * ((i) + 1) / 2 + (j))

// In-place LDL^T of a packed symmetric positive definite 6x6: m(i,j), j < i, becomes l_ij; dinv = 1 / pivots.
// Returns false when a pivot is not above `tiny`.
__device__ __forceinline__ bool ldl6_packed(double (&m)[21], double (&dinv)[6], const double tiny)
{
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double dk = m[RS_T(k, k)];
        ok = ok && (dk > tiny);
        const double inv = 1.0 / dk;
        dinv[k] = inv;
        double col[6];
#pragma unroll
        for (int i = k + 1; i < 6; ++i) col[i] = m[RS_T(i, k)];
#pragma unroll
        for (int i = k + 1; i < 6; ++i) {
            const double lik = col[i] * inv;
#pragma unroll
            for (int j = k + 1; j <= i; ++j) m[RS_T(i, j)] -= lik * col[j];
            m[RS_T(i, k)] = lik;
        }
    }
    return ok;
}

// z <- L^-1 z (unit lower factor, packed)
__device__ __forceinline__ void forward6_packed(const double (&m)[21], double (&z)[6])
{
#pragma unroll
    for (int i = 1; i < 6; ++i) {
        double sum = z[i];
#pragma unroll
        for (int j = 0; j < i; ++j) sum -= m[RS_T(i, j)] * z[j];
        z[i] = sum;
    }
}

// z <- L^-T D^-1 z
__device__ __forceinline__ void backward6_packed(const double (&m)[21], const double (&dinv)[6], double (&z)[6])
{
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double sum = z[i] * dinv[i];
#pragma unroll
        for (int j = i + 1; j < 6; ++j) sum -= m[RS_T(j, i)] * z[j];
        z[i] = sum;
    }
}

// Rank-deficient / ill-conditioned J (a pivot of the unpivoted factorisation collapsed): MINPACK's pivoted path.
// Diagonally pivoted LDL^T of C gives the rank and the basic Gauss-Newton solution (zeros on the dependent columns);
// parl = 0 when the rank is deficient. Rare (degenerate subsets, a frozen coordinate), so plain loops on local arrays.
struct LmparIO {
    double C[21], sc[6], g[6], diag[6], xs[6];
    double delta, par;
};
__device__ __noinline__ void lmpar_deficient(LmparIO& S)
{
    const double dwarf = DBL_MIN, delta = S.delta;
    double M[36], dinv[6], sg[6], e2[6], z[6], x[6], wa2[6];
    int perm[6];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j <= i; ++j) M[i * 6 + j] = M[j * 6 + i] = S.C[RS_T(i, j)];
    for (int j = 0; j < 6; ++j) {
        perm[j] = j;
        sg[j] = S.sc[j] * S.g[j];
        const double e = S.diag[j] * S.sc[j];
        e2[j] = e * e;
    }
    int rank = 6;
    for (int k = 0; k < 6; ++k) {
        int piv = k;
        double best = M[k * 6 + k];
        for (int i = k + 1; i < 6; ++i)
            if (M[i * 6 + i] > best) best = M[i * 6 + i], piv = i;
        if (piv != k) {
            for (int j = 0; j < 6; ++j) {
                const double t = M[k * 6 + j];
                M[k * 6 + j] = M[piv * 6 + j], M[piv * 6 + j] = t;
            }
            for (int i = 0; i < 6; ++i) {
                const double t = M[i * 6 + k];
                M[i * 6 + k] = M[i * 6 + piv], M[i * 6 + piv] = t;
            }
            const int t = perm[k];
            perm[k] = perm[piv], perm[piv] = t;
        }
        // C has a unit diagonal: a pivot at rounding level means a column that depends on the previous ones
        if (!(best > 64.0 * DBL_EPSILON)) {
            rank = k;
            break;
        }
        dinv[k] = 1.0 / best;
        for (int i = k + 1; i < 6; ++i) {
            const double lik = M[i * 6 + k] * dinv[k];
            for (int j = k + 1; j <= i; ++j) M[i * 6 + j] -= lik * M[j * 6 + k];
        }
        for (int i = k + 1; i < 6; ++i) M[i * 6 + k] *= dinv[k];
    }
    // basic solution on the leading `rank` pivots
    for (int i = 0; i < rank; ++i) {
        double sum = sg[perm[i]];
        for (int j = 0; j < i; ++j) sum -= M[i * 6 + j] * z[j];
        z[i] = sum;
    }
    for (int i = rank - 1; i >= 0; --i) {
        double sum = z[i] * dinv[i];
        for (int j = i + 1; j < rank; ++j) sum -= M[j * 6 + i] * z[j];
        z[i] = sum;
    }
    for (int i = 0; i < 6; ++i) x[perm[i]] = i < rank ? S.sc[perm[i]] * z[i] : 0.0;
    for (int j = 0; j < 6; ++j) wa2[j] = S.diag[j] * x[j];
    double dxnorm = norm6(wa2);
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) {
        S.par = 0.0;
        for (int j = 0; j < 6; ++j) S.xs[j] = x[j];
        return;
    }
    double parl = 0.0;
    if (rank == 6) {
        double q = 0.0;
        for (int i = 0; i < 6; ++i) {
            const int pi = perm[i];
            double sum = S.sc[pi] * (S.diag[pi] * wa2[pi] / dxnorm);
            for (int j = 0; j < i; ++j) sum -= M[i * 6 + j] * z[j];
            z[i] = sum;
            q += sum * sum * dinv[i];
        }
        parl = fp / delta / q;
    }
    double gn = 0.0;
    for (int j = 0; j < 6; ++j) {
        const double t = S.g[j] / S.diag[j];
        gn += t * t;
    }
    const double gnorm = sqrt(gn);
    double paru = gnorm / delta;
    if (paru == 0.0) paru = dwarf / fmin(delta, 0.1);
    double par = fmin(fmax(S.par, parl), paru);
    if (par == 0.0) par = gnorm / dxnorm;
    for (int iter = 1;; ++iter) {
        if (par == 0.0) par = fmax(dwarf, 0.001 * paru);
        // unpivoted LDL^T of C + par E^2 (positive definite for par > 0)
        for (int i = 0; i < 6; ++i) {
            for (int j = 0; j < i; ++j) M[i * 6 + j] = S.C[RS_T(i, j)];
            M[i * 6 + i] = S.C[RS_T(i, i)] + par * e2[i];
        }
        for (int k = 0; k < 6; ++k) {
            dinv[k] = 1.0 / M[k * 6 + k];
            for (int i = k + 1; i < 6; ++i) {
                const double lik = M[i * 6 + k] * dinv[k];
                for (int j = k + 1; j <= i; ++j) M[i * 6 + j] -= lik * M[j * 6 + k];
            }
            for (int i = k + 1; i < 6; ++i) M[i * 6 + k] *= dinv[k];
        }
        for (int i = 0; i < 6; ++i) {
            double sum = sg[i];
            for (int j = 0; j < i; ++j) sum -= M[i * 6 + j] * z[j];
            z[i] = sum;
        }
        for (int i = 5; i >= 0; --i) {
            double sum = z[i] * dinv[i];
            for (int j = i + 1; j < 6; ++j) sum -= M[j * 6 + i] * z[j];
            z[i] = sum;
        }
        for (int j = 0; j < 6; ++j) {
            x[j] = S.sc[j] * z[j];
            wa2[j] = S.diag[j] * x[j];
        }
        dxnorm = norm6(wa2);
        const double temp = fp;
        fp = dxnorm - delta;
        if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
        double q = 0.0;
        for (int i = 0; i < 6; ++i) {
            double sum = S.sc[i] * (S.diag[i] * (wa2[i] / dxnorm));
            for (int j = 0; j < i; ++j) sum -= M[i * 6 + j] * z[j];
            z[i] = sum;
            q += sum * sum * dinv[i];
        }
        const double parc = fp / delta / q;
        if (fp > 0.0) parl = fmax(parl, par);
        if (fp < 0.0) paru = fmin(paru, par);
        par = fmax(parl, par + parc);
    }
    S.par = par;
    for (int j = 0; j < 6; ++j) S.xs[j] = x[j];
}

// unsupported/Eigen/src/NonLinearOptimization/lmpar.h (lmpar2): trust-region parameter S.par and step S.xs.
// One factorise-and-solve body serves the Gauss-Newton step (pass 0, par = 0) and the damped steps (passes 1..10).
template <class St>
__device__ __forceinline__ void lmpar(St& S)
{
    const double dwarf = DBL_MIN;
    const double delta = S.delta;
    double sg[6], e2[6], x[6], wa2[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        sg[j] = S.sc[j] * S.g[j];
        const double e = S.diag[j] * S.sc[j];
        e2[j] = e * e;
    }
    double par = 0.0, parl = 0.0, paru = 0.0, fp = 0.0;
    int iter = 0;
#pragma unroll 1
    while (true) {
        double M[21], dinv[6], z[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int j = 0; j < i; ++j) M[RS_T(i, j)] = S.C[RS_T(i, j)];
            M[RS_T(i, i)] = S.C[RS_T(i, i)] + par * e2[i];
            z[i] = sg[i];
        }
        const bool ok = ldl6_packed(M, dinv, 64.0 * DBL_EPSILON);
        if (iter == 0 && !ok) {
            // out of line, on a copy of what it needs: the caller's state may live in registers (one LM per lane)
            LmparIO io;
            for (int i = 0; i < 21; ++i) io.C[i] = S.C[i];
            for (int j = 0; j < 6; ++j) io.sc[j] = S.sc[j], io.g[j] = S.g[j], io.diag[j] = S.diag[j];
            io.delta = S.delta, io.par = S.par;
            lmpar_deficient(io);
            S.par = io.par;
            for (int j = 0; j < 6; ++j) S.xs[j] = io.xs[j];
            return;
        }
        forward6_packed(M, z);
        backward6_packed(M, dinv, z);
        double dx2 = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            x[j] = S.sc[j] * z[j];
            wa2[j] = S.diag[j] * x[j];
            dx2 += wa2[j] * wa2[j];
        }
        const double idx = dx2 > 0.0 ? rsqrt(dx2) : 0.0;   // 1 / |D x|
        const double dxnorm = dx2 * idx;
        const double temp = fp;
        fp = dxnorm - delta;
        if (iter == 0) {
            if (fp <= 0.1 * delta) break;  // the Gauss-Newton step is inside the trust region: par = 0
        }
        else if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10)
            break;
        // Newton correction fp / delta / (w^T (C + par E^2)^-1 w), w = S D^2 x / |D x| (at par = 0 this is parl)
        double q = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) z[i] = S.sc[i] * (S.diag[i] * (wa2[i] * idx));
        forward6_packed(M, z);
#pragma unroll
        for (int i = 0; i < 6; ++i) q += z[i] * z[i] * dinv[i];
        const double parc = fp / (delta * q);
        if (iter == 0) {
            parl = parc;
            double gn = 0.0;
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const double t = S.g[j] / S.diag[j];
                gn += t * t;
            }
            const double gnorm = sqrt(gn);
            paru = gnorm / delta;
            if (paru == 0.0) paru = dwarf / fmin(delta, 0.1);
            par = fmin(fmax(S.par, parl), paru);
            if (par == 0.0) par = gnorm / dxnorm;
        }
        else {
            if (fp > 0.0) parl = fmax(parl, par);
            if (fp < 0.0) paru = fmin(paru, par);
            par = fmax(parl, par + parc);
        }
        if (par == 0.0) par = fmax(dwarf, 0.001 * paru);
        ++iter;
    }
    S.par = par;
#pragma unroll
    for (int j = 0; j < 6; ++j) S.xs[j] = x[j];
}

// ---- the serial part of one LM iteration, on whatever holds the 6x6 state (shared memory for a warp-wide LM, registers for
// one LM per lane): same member names, same arithmetic.
// After the Jacobian pass (S.A = J^T J, S.g = J^T r): column scaling, the scaled matrix C, MINPACK's diag / gnorm bookkeeping.
template <class St>
__device__ __forceinline__ void lm_after_jacobian(St& S)
{
    S.nfev += 7;  // NumericalDiff re-evaluates f(x) and then one evaluation per column
    double sc[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const double ajj = S.A[j * 6 + j];
        sc[j] = ajj > 0.0 ? rsqrt(ajj) : 1.0;   // 1 / |J_j|
        S.wa2[j] = ajj > 0.0 ? ajj * sc[j] : 0.0;
        S.sc[j] = sc[j];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) S.C[RS_T(i, j)] = S.A[i * 6 + j] * sc[i] * sc[j];
    if (S.iter == 1) {
        double tt[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            S.diag[j] = (S.wa2[j] == 0.0) ? 1.0 : S.wa2[j];
            tt[j] = S.diag[j] * S.x[j];
        }
        S.xnorm = norm6(tt);
        S.delta = 100.0 * S.xnorm;
        if (S.delta == 0.0) S.delta = 100.0;
    }
    // gnorm = max_j |J_j . r| / (|J_j| |r|)
    double gnorm = 0.0;
    if (S.fnorm != 0.0) {
        const double ifn = 1.0 / S.fnorm;
#pragma unroll
        for (int j = 0; j < 6; ++j)
            if (S.wa2[j] != 0.0) gnorm = fmax(gnorm, fabs((S.g[j] * ifn) * sc[j]));
    }
    S.gnorm = gnorm;
    if (gnorm <= 0.0) S.status = 4;  // CosinusTooSmall (gtol = 0)
#pragma unroll
    for (int j = 0; j < 6; ++j) S.diag[j] = fmax(S.diag[j], S.wa2[j]);
}

// Trust-region step: lmpar, the trial point S.xt and its transform S.T.
template <class St>
__device__ __forceinline__ void lm_propose(St& S)
{
    lmpar(S);
    double tt[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        S.p[j] = -S.xs[j];
        S.xt[j] = S.x[j] + S.p[j];
        tt[j] = S.diag[j] * S.p[j];
    }
    S.pnorm = norm6(tt);
    if (S.iter == 1) S.delta = fmin(S.delta, S.pnorm);
    make_xform(S.xt, S.T);
}

// The trial point's |f|^2 is in: actual / predicted reduction, trust-region update, acceptance, stop codes (lmder's tail).
template <class St>
__device__ __forceinline__ void lm_judge(St& S, const double ss1, const int maxfev)
{
    ++S.nfev;
    const double fnorm = S.fnorm, fnorm1 = sqrt(ss1), pnorm = S.pnorm;
    const double inv_fnorm = 1.0 / fnorm;
    double actred = -1.0;
    if (0.1 * fnorm1 < fnorm) actred = 1.0 - (fnorm1 * inv_fnorm) * (fnorm1 * inv_fnorm);
    // |J p|^2 = p^T A p
    double pAp = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double sum = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) sum += S.A[i * 6 + j] * S.p[j];
        pAp += sum * S.p[i];
    }
    const double temp1 = fmax(pAp, 0.0) * inv_fnorm * inv_fnorm;
    const double temp2 = S.par * (pnorm * inv_fnorm) * (pnorm * inv_fnorm);
    const double prered = temp1 + temp2 * 2.0;
    const double dirder = -(temp1 + temp2);
    double ratio = 0.0;
    if (prered != 0.0) ratio = actred / prered;
    if (ratio <= 0.25) {
        double temp = 0.0;
        if (actred >= 0.0) temp = 0.5;
        if (actred < 0.0) temp = 0.5 * dirder / (dirder + 0.5 * actred);
        if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
        S.delta = temp * fmin(S.delta, pnorm * 10.0);
        S.par /= temp;
    }
    else if (!(S.par != 0.0 && ratio < 0.75)) {
        S.delta = pnorm * 2.0;
        S.par = 0.5 * S.par;
    }
    if (ratio >= 1e-4) {
        double tt[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            S.x[j] = S.xt[j];
            tt[j] = S.diag[j] * S.x[j];
        }
        S.xnorm = norm6(tt);
        S.fnorm = fnorm1;
        ++S.iter;
    }
    const double ftol = kSqrtEps, xtol = kSqrtEps;
    int status = kRunning;
    if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0 && S.delta <= xtol * S.xnorm)
        status = 3;
    else if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0)
        status = 1;
    else if (S.delta <= xtol * S.xnorm)
        status = 2;
    else if (S.nfev >= maxfev)
        status = 5;
    else if (fabs(actred) <= DBL_EPSILON && prered <= DBL_EPSILON && 0.5 * ratio <= 1.0)
        status = 6;
    else if (S.delta <= DBL_EPSILON * S.xnorm)
        status = 7;
    else if (S.gnorm <= DBL_EPSILON)
        status = 8;
    S.status = status;
    S.again = (status == kRunning && ratio < 1e-4) ? 1 : 0;
}

// Eigen::LevenbergMarquardt<NumericalDiff<F,Forward>>::minimize on S.x (in/out). Whole warp must call; returns the
// Eigen status (<= 0 failure, 1..8 MINPACK info). m = residual count of the problem. Inlined at exactly ONE call site per
// kernel: the body is ~4k instructions and the serial lane-0 chains are latency bound, so instruction-cache residency
// matters, and inlining lets the compiler see that S and the feature arrays live in shared memory (LDS, not generic LD).
template <bool P2D>
__device__ __forceinline__ int lm_minimize_warp(WarpLM& S, const Problem& P, const PoseIntrinsics& K, const int m,
                                             const int maxfev, const int lane, const volatile int* abort = nullptr)
{
    if (m < 6 || maxfev <= 0) return 0;  // ImproperInputParameters
    if (lane == 0) make_xform(S.x, S.T);
    __syncwarp();
    {
        const double ss = eval_sumsq<P2D>(P, S.T, K, lane);
        if (lane == 0) {
            S.fnorm = sqrt(ss);
            S.par = 0.0, S.delta = 0.0, S.xnorm = 0.0;
            S.iter = 1, S.nfev = 1, S.status = kRunning;
        }
    }
    __syncwarp();

#pragma unroll 1
    while (true) {
        // a speculative RANSAC hypothesis is dropped as soon as the serial rule has stopped the loop before it; the flag
        // lives in global memory (several CTAs may work on a frame), so it is read here and looked at after the set-up
#ifndef RS_ABORT_AT_TOP
        const int aborted = abort ? *abort : 0;
#else
        if (abort && *abort) return 0;
        const int aborted = 0;
#endif
        // ---- Jacobian set-up: R'(x) on lane 0, R'(x + h_k e_k) on lanes 1..3, then the 27 difference quotients ----
        if (lane < 4) {
            double xx[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) xx[j] = S.x[j];
            double h = 0.0;
#pragma unroll
            for (int j = 3; j < 6; ++j)
                if (lane == j - 2) {
                    h = kSqrtEps * fabs(xx[j]);  // NumericalDiff: h = sqrt(eps) |x_j|, or sqrt(eps) when x_j == 0
                    if (h == 0.0) h = kSqrtEps;
                    xx[j] += h;
                }
            Xform Tk;
            make_xform(xx, Tk);
            if (lane == 0)
                S.T = Tk;
            else {
#pragma unroll
                for (int i = 0; i < 9; ++i) S.Rk[lane - 1][i] = Tk.R[i];
                S.ih[lane - 1] = 1.0 / h;
            }
        }
        __syncwarp();
        if (aborted) return 0;
        if (lane < 27) {
            const int k = lane / 9, i = lane - 9 * k;
            S.dR[lane] = (S.Rk[k][i] - S.T.R[i]) * S.ih[k];
        }
        __syncwarp();
        double a[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = 0.0;
#pragma unroll 1
        for (int k = lane; k < P.n; k += 32) {
            int type;
            double o[4], mm[4];
            const int gi = load_feature(P, k, type, o, mm);
            if (P2D && type == RS_FEAT_POINT2D) {
                double c[27];   // out of line and through local memory, so that a[] stays in registers
                point2d_jacobian(o, mm, P.aux, P.M, gi, S, K, c);
#pragma unroll
                for (int i = 0; i < 27; ++i) a[i] += c[i];
            }
            else
                feature_jacobian(type, o, mm, S.T, DRArray{S.dR}, K, a);
        }
        const double mine = reduce_scatter32(a, lane);
        if (lane < 21) {
            int i = 0, rem = lane;
            while (rem >= 6 - i) rem -= 6 - i, ++i;
            const int j = i + rem;
            S.A[i * 6 + j] = mine;
            S.A[j * 6 + i] = mine;
        }
        else if (lane < 27)
            S.g[lane - 21] = mine;
        __syncwarp();

        if (lane == 0) lm_after_jacobian(S);
        __syncwarp();
        if (S.status != kRunning) break;

        // ---- inner loop: trust-region step until the ratio is acceptable ----
#pragma unroll 1
        while (true) {
            if (lane == 0) lm_propose(S);
            __syncwarp();
            const double ss1 = eval_sumsq<P2D>(P, S.T, K, lane);
            if (lane == 0) lm_judge(S, ss1, maxfev);
            __syncwarp();
            if (S.status != kRunning || !S.again) break;
        }
        if (S.status != kRunning) break;
    }
    const int status = S.status;
    __syncwarp();
    return status;
}

// compute_optimized_global_pose (pose_optimization.cpp:302-359) for the whole warp. x0 -> S.x; returns success and
// leaves the optimised coefficients in S.x.
template <bool P2D>
__device__ __forceinline__ bool optimize_pose_warp(WarpLM& S, const Problem& P, const PoseIntrinsics& K, const double* x0,
                                                   const int m, const double score, const int maxfev, const int lane,
                                                   const volatile int* abort = nullptr)
{
    bool finite = true;
#pragma unroll
    for (int j = 0; j < 6; ++j) finite = finite && isfinite(x0[j]);
    if (!finite || m <= 1 || score < 1.0) return false;
    if (lane == 0)
        for (int j = 0; j < 6; ++j) S.x[j] = x0[j];
    __syncwarp();
    const int status = lm_minimize_warp<P2D>(S, P, K, m, maxfev, lane, abort);
    if (status <= 0) return false;
    // the reference rejects a pose whose [position, Euler angles] vector has a NaN: that vector is finite exactly
    // when the coefficients and the quaternion built from them are
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; ++j) ok = ok && isfinite(S.x[j]);
    return ok && isfinite(S.x[3] * S.x[3] + S.x[4] * S.x[4] + S.x[5] * S.x[5]);
}

// ---- the same solve for the W warps of a CTA at once, serial parts batched -----------------------------------------------
// Every warp of the CTA owns one LM problem (the Monte-Carlo kernel: eight samples of one frame); the passes over the
// features stay warp-wide, but the serial 6x6 parts of ALL the CTA's problems run together on lanes 0..W-1 of warp 0 -
// one instruction stream for W problems instead of W streams at 1/32 lane utilisation (a third of the instructions the
// one-warp version issues). The price is a CTA barrier around each serial part, which the other CTA resident on the SM
// fills. Same arithmetic per problem, bit for bit: only the lane that executes it changes.
// All threads of the CTA must call; `mine` = this warp has a problem (P, lm[warp].x = start point). Status in lm[warp].status.
template <bool P2D>
__device__ __forceinline__ void lm_minimize_cta(WarpLM* lm, const int nwarps, const Problem& P, const PoseIntrinsics& K,
                                                const int m, const int maxfev, const int warp, const int lane, const bool mine)
{
    WarpLM& S = lm[warp];
    const bool leader = warp == 0 && lane < nwarps;   // lane l of warp 0 runs the serial parts of problem l
    if (lane == 0) {
        if (!mine || m < 6 || maxfev <= 0)
            S.status = 0;   // ImproperInputParameters (or no problem)
        else {
            make_xform(S.x, S.T);
            S.status = kRunning;
        }
    }
    __syncwarp();
    if (S.status == kRunning) {
        const double ss = eval_sumsq<P2D>(P, S.T, K, lane);
        if (lane == 0) {
            S.fnorm = sqrt(ss);
            S.par = 0.0, S.delta = 0.0, S.xnorm = 0.0;
            S.iter = 1, S.nfev = 1;
        }
    }
    __syncthreads();
#pragma unroll 1
    while (true) {
        if (S.status == kRunning) {
            // ---- Jacobian set-up and pass, as in lm_minimize_warp ----
            if (lane < 4) {
                double xx[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) xx[j] = S.x[j];
                double h = 0.0;
#pragma unroll
                for (int j = 3; j < 6; ++j)
                    if (lane == j - 2) {
                        h = kSqrtEps * fabs(xx[j]);
                        if (h == 0.0) h = kSqrtEps;
                        xx[j] += h;
                    }
                Xform Tk;
                make_xform(xx, Tk);
                if (lane == 0)
                    S.T = Tk;
                else {
#pragma unroll
                    for (int i = 0; i < 9; ++i) S.Rk[lane - 1][i] = Tk.R[i];
                    S.ih[lane - 1] = 1.0 / h;
                }
            }
            __syncwarp();
            if (lane < 27) {
                const int k = lane / 9, i = lane - 9 * k;
                S.dR[lane] = (S.Rk[k][i] - S.T.R[i]) * S.ih[k];
            }
            __syncwarp();
            double a[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = 0.0;
#pragma unroll 1
            for (int k = lane; k < P.n; k += 32) {
                int type;
                double o[4], mm[4];
                const int gi = load_feature(P, k, type, o, mm);
                if (P2D && type == RS_FEAT_POINT2D) {
                    double c[27];
                    point2d_jacobian(o, mm, P.aux, P.M, gi, S, K, c);
#pragma unroll
                    for (int i = 0; i < 27; ++i) a[i] += c[i];
                }
                else
                    feature_jacobian(type, o, mm, S.T, DRArray{S.dR}, K, a);
            }
            const double mine_v = reduce_scatter32(a, lane);
            if (lane < 21) {
                int i = 0, rem = lane;
                while (rem >= 6 - i) rem -= 6 - i, ++i;
                const int j = i + rem;
                S.A[i * 6 + j] = mine_v;
                S.A[j * 6 + i] = mine_v;
            }
            else if (lane < 27)
                S.g[lane - 21] = mine_v;
        }
        __syncthreads();
        if (leader) {
            WarpLM& Q = lm[lane];
            if (Q.status == kRunning) lm_after_jacobian(Q);
            Q.again = Q.status == kRunning ? 1 : 0;   // takes part in the trust-region steps below
        }
        __syncthreads();
        // ---- trust-region steps until every problem of the CTA has an acceptable ratio (or has stopped) ----
#pragma unroll 1
        while (true) {
            if (leader && lm[lane].again) lm_propose(lm[lane]);
            __syncthreads();
            if (S.again) {
                const double ss1 = eval_sumsq<P2D>(P, S.T, K, lane);
                if (lane == 0) S.ss1 = ss1;
            }
            __syncthreads();
            int more = 0;
            if (leader && lm[lane].again) {
                lm_judge(lm[lane], lm[lane].ss1, maxfev);
                more = lm[lane].again;
            }
            if (!__syncthreads_or(more)) break;
        }
        if (!__syncthreads_or(S.status == kRunning)) break;
    }
}

// optimize_pose_warp for the warps of a CTA working on problems of the SAME frame (x0, m and score are the frame's: the
// guards below take every warp the same way). Success per warp; the optimised coefficients are left in lm[warp].x.
template <bool P2D>
__device__ __forceinline__ bool optimize_pose_cta(WarpLM* lm, const int nwarps, const Problem& P, const PoseIntrinsics& K,
                                                  const double* x0, const int m, const double score, const int maxfev,
                                                  const int warp, const int lane, const bool mine)
{
    bool finite = true;
#pragma unroll
    for (int j = 0; j < 6; ++j) finite = finite && isfinite(x0[j]);
    if (!finite || m <= 1 || score < 1.0) return false;
    WarpLM& S = lm[warp];
    if (lane == 0)
        for (int j = 0; j < 6; ++j) S.x[j] = x0[j];
    __syncwarp();
    lm_minimize_cta<P2D>(lm, nwarps, P, K, m, maxfev, warp, lane, mine);
    if (!mine || S.status <= 0) return false;
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; ++j) ok = ok && isfinite(S.x[j]);
    return ok && isfinite(S.x[3] * S.x[3] + S.x[4] * S.x[4] + S.x[5] * S.x[5]);
}

// ---- counter-based generator of the RS_RNG_DEVICE mode --------------------------------------------------------------
__host__ __device__ inline uint64_t mix64(uint64_t z)
{
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__device__ inline uint64_t rng_key(const uint32_t seed, const uint32_t domain, const uint32_t frame)
{
    return mix64((uint64_t(seed) << 32) ^ (uint64_t(domain) << 24) ^ frame);
}
// Four standard normals for (frame, sample, feature): two Box-Muller pairs, each from one 64-bit counter hash. The pair
// is evaluated in FP32 with the MUFU intrinsics (24-bit uniforms, |g| <= 5.8) and widened: a Monte-Carlo perturbation
// needs the distribution, not 53 bits, and the FP64 log / sqrt / sincospi chain was 10 % of the variance kernel.
// pose_export_normals_kernel returns exactly these values, which is how the oracle is fed the same draws.
__device__ inline void device_normals(const uint32_t seed, const int frame, const int sample, const int feature,
                                      double g[4])
{
    const uint64_t key = rng_key(seed, 2u, uint32_t(frame));
    const uint64_t ctr = (uint64_t(uint32_t(sample)) << 32) | (uint64_t(uint32_t(feature)) << 1);
#pragma unroll
    for (int pair = 0; pair < 2; ++pair) {
        const uint64_t a = mix64(key ^ mix64(ctr + uint64_t(pair)));
        const float u1 = (float(uint32_t(a >> 40)) + 1.0f) * 5.9604644775390625e-08f;        // (0, 1]
        const float u2 = float(uint32_t(a >> 8) & 0xffffffu) * 5.9604644775390625e-08f;     // [0, 1)
        const float rad = sqrtf(-2.0f * __logf(u1));
        float sn, cs;
        __sincosf(6.2831853071795865f * u2, &sn, &cs);
        g[2 * pair] = double(rad * cs);
        g[2 * pair + 1] = double(rad * sn);
    }
}


// is_covariance_valid (covariances.hpp:13-44): finite, isApprox-symmetric, LDLT without a negative pivot
__device__ bool covariance_valid(const double* c)
{
    for (int i = 0; i < 36; ++i)
        if (!isfinite(c[i])) return false;
    double diff2 = 0.0, n2 = 0.0;
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            const double dd = c[i * 6 + j] - c[j * 6 + i];
            diff2 += dd * dd;
            n2 += c[i * 6 + j] * c[i * 6 + j];
        }
    if (!(diff2 <= 1e-12 * 1e-12 * n2)) return false;
    double a[36];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) a[i * 6 + j] = c[(i < j ? i : j) * 6 + (i < j ? j : i)];
    bool neg = false;
    for (int k = 0; k < 6; ++k) {
        int p = k;
        double best = fabs(a[k * 6 + k]);
        for (int i = k + 1; i < 6; ++i)
            if (fabs(a[i * 6 + i]) > best) {
                best = fabs(a[i * 6 + i]);
                p = i;
            }
        if (p != k) {
            for (int j = 0; j < 6; ++j) {
                const double t = a[k * 6 + j];
                a[k * 6 + j] = a[p * 6 + j];
                a[p * 6 + j] = t;
            }
            for (int i = 0; i < 6; ++i) {
                const double t = a[i * 6 + k];
                a[i * 6 + k] = a[i * 6 + p];
                a[i * 6 + p] = t;
            }
        }
        const double dkk = a[k * 6 + k];
        if (dkk < 0.0) neg = true;
        if (fabs(dkk) <= DBL_MIN) break;
        for (int i = k + 1; i < 6; ++i) {
            const double l = a[i * 6 + k] / dkk;
            for (int j = k + 1; j < 6; ++j) a[i * 6 + j] -= l * a[k * 6 + j];
        }
    }
    return !neg;
}

// compute_pose_variance's reduction (pose_optimization.cpp:414-437) for one frame, by one warp. Lanes split the samples
// for the mean, then lane e < 21 owns one entry of the upper triangle and sums it over the samples in sample order.
// v6 / v_ok were written by other CTAs of the same launch: read past L1.
__device__ void frame_covariance_warp(const PoseBuffers& buf, const PoseLaunch& prm, const int b, const int lane)
{
    const double* v6 = buf.v6 + size_t(b) * buf.max_variance * 6;
    const int32_t* vok = buf.v_ok + size_t(b) * buf.max_variance;
    rs_pose_out* out = buf.out + b;
    double medium[6] = {0, 0, 0, 0, 0, 0};
    int cnt = 0;
    for (int s = lane; s < prm.n_variance; s += 32)
        if (__ldcg(vok + s)) {
#pragma unroll
            for (int j = 0; j < 6; ++j) medium[j] += __ldcg(v6 + s * 6 + j);
            ++cnt;
        }
    cnt = __reduce_add_sync(FULL, cnt);
#pragma unroll
    for (int j = 0; j < 6; ++j) medium[j] = warp_sum(medium[j]);
    if (lane == 0) out->n_variance_ok = cnt;
    if (unsigned(cnt) < unsigned(prm.n_variance) / 2u) {
        if (lane == 0) out->status = -2;
        return;
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) medium[j] /= double(cnt);
    int ei = 0, ej = 0;
    {
        int rem = lane < 21 ? lane : 0;
        while (rem >= 6 - ei) rem -= 6 - ei, ++ei;
        ej = ei + rem;
    }
    double mi = 0.0, mj = 0.0;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        if (j == ei) mi = medium[j];
        if (j == ej) mj = medium[j];
    }
    double acc = 0.0;
    for (int s = 0; s < prm.n_variance; ++s)
        if (__ldcg(vok + s)) acc += (__ldcg(v6 + s * 6 + ei) - mi) * (__ldcg(v6 + s * 6 + ej) - mj);
    acc /= double(cnt - 1);
    if (ei == ej) acc += 0.001;
    if (lane < 21) {
        out->cov[ei * 6 + ej] = acc;
        out->cov[ej * 6 + ei] = acc;
    }
    __syncwarp();
    if (lane == 0) {
        double cov[36];
        for (int i = 0; i < 36; ++i) cov[i] = *reinterpret_cast<volatile double*>(&out->cov[i]);
        out->status = covariance_valid(cov) ? 1 : -2;
    }
}

}  // namespace

}  // namespace rs
