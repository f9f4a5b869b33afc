// TEST INFRASTRUCTURE - stand-in for unsupported/Eigen/NonLinearOptimization as pose_optimization.cpp uses it:
// LevenbergMarquardt<NumericalDiff<Functor>>::minimize(x) = MINPACK lmdif with Eigen's forward differences. The algorithm is
// third-party code; it is the oracle's restatement (oracle/pose.cpp: lm_minimize, pinned against scipy's MINPACK in
// tests/test_oracle_pose.py) that runs here, so this build pins what the REFERENCE wrote around it - the RANSAC loop, the
// subset draws, the residual functor, inlier tests, the Monte-Carlo variance - not the LM iterate path itself.
#pragma once
#include "../pose.hpp"
#include "ref_eigen.hpp"

namespace Eigen {

namespace LevenbergMarquardtSpace {
enum Status {
    NotStarted = -2,
    Running = -1,
    ImproperInputParameters = 0,
    RelativeReductionTooSmall = 1,
    RelativeErrorTooSmall = 2,
    RelativeErrorAndReductionTooSmall = 3,
    CosinusTooSmall = 4,
    TooManyFunctionEvaluation = 5,
    FtolTooSmall = 6,
    XtolTooSmall = 7,
    GtolTooSmall = 8,
    UserAsked = 9
};
}

enum NumericalDiffMode { Forward, Central };

template <class Functor, NumericalDiffMode mode = Forward>
struct NumericalDiff : public Functor {
    using Functor::Functor;
    NumericalDiff(const Functor& f) : Functor(f) {}
};

template <class FunctorType, class Scalar = double>
class LevenbergMarquardt {
    FunctorType& functor_;

  public:
    explicit LevenbergMarquardt(FunctorType& f) : functor_(f) {}
    template <class Vec>
    LevenbergMarquardtSpace::Status minimize(Vec& x)
    {
        const int m = int(functor_.values());
        Matrix<double, Dynamic, 1> fvec(m);
        const oracle::ResidualFn fn = [&](const double xx[6], double* out) {
            Matrix<double, 6, 1> xv;
            for (int i = 0; i < 6; ++i) xv(i) = xx[i];
            // Global_Pose_Estimator leaves a feature's entries untouched when its distance has a NaN: they keep what the LM's
            // buffer held (Eigen hands the functor its own work vectors), so the buffer goes in as well as out
            for (int i = 0; i < m; ++i) fvec(i) = out[i];
            functor_(xv, fvec);
            for (int i = 0; i < m; ++i) out[i] = fvec(i);
        };
        double xx[6];
        for (int i = 0; i < 6; ++i) xx[i] = x(i);
        const oracle::LMResult r = oracle::lm_minimize(fn, m, xx, 400);   // Eigen's default maxfev
        for (int i = 0; i < 6; ++i) x(i) = xx[i];
        return LevenbergMarquardtSpace::Status(r.status);
    }
};
template <class F>
LevenbergMarquardt(F&) -> LevenbergMarquardt<F, double>;

}  // namespace Eigen
