// One Levenberg-Marquardt solve per LANE: the state lives in the lane's registers / local memory, the passes over the features
// are serial loops (no shuffles), the trust-region algebra is pose_lm.cuh's (lm_after_jacobian / lm_propose / lm_judge / lmpar on
// the lane's own 6x6 state). Used where many SMALL problems run side by side: the hypotheses of a RANSAC with hundreds of
// iterations (pose_wide.cu: problems of 3-16 features, 65 536 of them per step). The caller supplies where the features come
// from (Source) and where the 27 entries of dR' are kept (DRStore).
// Measured and not kept: the Monte-Carlo solves one sample per lane (the frame's ~290 inliers as a serial loop per lane, normals
// in global memory as [feature][component][sample], dR' in strided shared memory, 168 registers): 3.4 ms per 256 frames
// against the one-warp-per-sample kernel's 0.81 - 25 600 samples are only 800 warps of work, each a chain of ~225 k dependent
// instructions with an L2 load per feature and nothing to hide it behind. A lane per problem pays when problems are tiny and
// plentiful, not when they are few and long.
#pragma once
#include "pose_lm.cuh"

namespace rs {

namespace {

// The LM state of one lane: the members lm_after_jacobian / lm_propose / lm_judge / lmpar use.
struct LaneLM {
    double x[6], xt[6], diag[6], p[6], wa2[6], sc[6], g[6], xs[6];
    double A[36];
    double C[21];
    Xform T;
    double fnorm, par, delta, xnorm, gnorm, pnorm;
    int status, nfev, iter, again;
};

struct DRLocal {   // dR' in the lane's registers
    double v[27];
    __device__ __forceinline__ double at(const int k, const int i) const { return v[9 * k + i]; }
    __device__ __forceinline__ void set(const int k, const int i, const double x) { v[9 * k + i] = x; }
};

// Source: int count() const; template load(k, type, o[4], m[4]) -> feature index (all features are points or planes)
template <class Source>
__device__ __forceinline__ double lane_sumsq(const Source& src, const Xform& T, const PoseIntrinsics& K)
{
    double ss = 0.0;
#pragma unroll 1
    for (int k = 0; k < src.count(); ++k) {
        int type;
        double o[4], m[4], r[3];
        const int i = src.load(k, type, o, m);
        feature_residual<false>(type, o, m, T, K, r, nullptr, 0, i);
        ss += r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    }
    return ss;
}

// LevenbergMarquardt::minimizeInit + the first residual evaluation
template <class Source>
__device__ __forceinline__ void lane_lm_begin(LaneLM& S, const Source& src, const PoseIntrinsics& K, const double* x0, const int m,
                                              const int maxfev)
{
#pragma unroll
    for (int j = 0; j < 6; ++j) S.x[j] = x0[j];
    if (m < 6 || maxfev <= 0) {
        S.status = 0;   // ImproperInputParameters
        return;
    }
    make_xform(S.x, S.T);
    S.fnorm = sqrt(lane_sumsq(src, S.T, K));
    S.par = 0.0, S.delta = 0.0, S.xnorm = 0.0;
    S.iter = 1, S.nfev = 1, S.status = kRunning;
}

// One outer iteration of minimizeOneStep: Jacobian (forward differences of the transform, chain rule per feature, as the
// warp-wide LM does), then trust-region steps until one is accepted or the solve stops.
template <class Source, class DRStore>
__device__ __forceinline__ void lane_lm_step(LaneLM& S, const Source& src, DRStore& dR, const PoseIntrinsics& K, const int maxfev)
{
    make_xform(S.x, S.T);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double xx[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) xx[j] = S.x[j];
        double h = kSqrtEps * fabs(xx[3 + k]);   // NumericalDiff: h = sqrt(eps) |x_j|, or sqrt(eps) when x_j == 0
        if (h == 0.0) h = kSqrtEps;
        xx[3 + k] += h;
        Xform Tk;
        make_xform(xx, Tk);
        const double ih = 1.0 / h;
#pragma unroll
        for (int i = 0; i < 9; ++i) dR.set(k, i, (Tk.R[i] - S.T.R[i]) * ih);
    }
    double a[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i] = 0.0;
#pragma unroll 1
    for (int k = 0; k < src.count(); ++k) {
        int type;
        double o[4], mm[4];
        src.load(k, type, o, mm);
        feature_jacobian(type, o, mm, S.T, dR, K, a);
    }
    {
        int t = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            S.g[i] = a[21 + i];
#pragma unroll
            for (int j = i; j < 6; ++j) {
                S.A[i * 6 + j] = a[t];
                S.A[j * 6 + i] = a[t];
                ++t;
            }
        }
    }
    lm_after_jacobian(S);
    if (S.status != kRunning) return;
#pragma unroll 1
    while (true) {
        lm_propose(S);
        const double ss1 = lane_sumsq(src, S.T, K);
        lm_judge(S, ss1, maxfev);
        if (S.status != kRunning || !S.again) break;
    }
}

}  // namespace

}  // namespace rs
