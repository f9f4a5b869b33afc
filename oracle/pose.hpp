// TEST INFRASTRUCTURE — CPU oracle (see linalg.hpp header). Restates the reference's pose solve:
//   src/pose_optimization/{pose_optimization.cpp,levenberg_marquardt_functors.cpp,ransac.hpp},
//   src/map_management/map_features/{map_point.cpp:16-65,map_primitive.cpp:15-85},
//   src/coordinates/{point,plane}_coordinates.*, src/utils/{camera_transformation,pose,angle_utils,distance_utils}.cpp
// plus Eigen 3.4's unsupported LevenbergMarquardt / NumericalDiff / lmpar2 / qrsolv and ColPivHouseholderQR,
// which are NOT vendored by the reference (find_package(Eigen3), CMakeLists.txt:35) and not on this machine:
// restated from the published algorithms (MINPACK lmdif/lmpar/qrsolv). PINS: (1) the 40 scenarios of
// tests/test_pose_optimization.cpp re-run against this oracle with the reference's own tolerances
// (tests/test_oracle_pose.py); (2) the LM cross-checked against scipy.optimize.leastsq (MINPACK); (3) the reference's own
// pose-solve translation units compiled against stand-in Eigen headers (oracle/ref_shim -> oracle/_ref/libref_pose.so): this
// restatement equals that build BIT FOR BIT - status, inlier mask, pose, Monte-Carlo covariance - on the 40 scenarios and on
// random mixed point / plane / point2d problems (tests/test_reference_pose_build.py).
// The arithmetic inside Eigen's LM / QR itself stays "parity unpinned" beyond (2).
#pragma once
#include <cstdint>
#include <functional>
#include <random>
#include <vector>

#include "../include/rgbdslam_b200.h"
#include "linalg.hpp"

namespace oracle {

struct Intrinsics {
    double fx = 550, fy = 550, cx = 320, cy = 240;
};

struct Pose7 {
    double t[3] = {0, 0, 0};
    double q[4] = {1, 0, 0, 0};  // w x y z
};

// Eigen::LevenbergMarquardt<NumericalDiff<F, Forward>>::minimize. Returns the Eigen status code
// (LevenbergMarquardtSpace::Status: <=0 failure, 1..8 as MINPACK info). f(x, fvec) fills m residuals.
struct LMResult {
    int status = 0;
    int nfev = 0;
    int iterations = 0;
    double fnorm = 0;
};
using ResidualFn = std::function<void(const double x[6], double* fvec)>;
LMResult lm_minimize(const ResidualFn& f, int m, double x[6], int maxfev = 400);

// utils::compute_world_to_camera_transform(quaternion, position), camera_transformation.cpp:11-44
Mat4 world_to_camera(const double q[4], const double t[3]);
// utils::compute_plane_world_to_camera_matrix(w2c), camera_transformation.cpp:52-71
Mat4 plane_world_to_camera(const Mat4& w2c);

// levenberg_marquardt_functors.cpp:14-38,74-86
void coefficients_from_pose(const Pose7& p, double x[6]);
Pose7 pose_from_coefficients(const double x[6]);
// PoseBase::get_vector (pose.hpp:30-35): position + eulerAngles(0,1,2)
void pose_vector6(const Pose7& p, double v[6]);
// utils::get_quaternion_from_euler_angles (angle_utils.cpp:6-11); EulerAngles(yaw,pitch,roll)
void quaternion_from_euler(double yaw, double pitch, double roll, double q[4]);

// Residual vector of Global_Pose_Estimator::operator() (levenberg_marquardt_functors.cpp:128-169)
void pose_residuals(const Intrinsics& K, const std::vector<rs_match>& feats, const double x[6], double* fvec);
int residual_count(const std::vector<rs_match>& feats);
// IOptimizationFeature::is_inlier for one feature
bool feature_is_inlier(const Intrinsics& K, const rs_match& f, const Mat4& w2c, const Mat4& planeW2c);
double feature_score(const rs_match& f);
int feature_parts(const rs_match& f);

// Random source: the reference's thread-local engine + distributions (random.hpp:17-39). When the explicit
// lists are set they are consumed instead (parity against the library's RS_RNG_DEVICE mode).
struct PoseRandom {
    std::mt19937 engine;
    std::normal_distribution<double> normal{0.0, 1.0};
    const int32_t* subsets = nullptr;  // [max_iterations][RS_MAX_SUBSET], -1 padded, in residual order
    const double* normals = nullptr;   // [n_variance][max_matches][4]
    int max_matches = 0;
    explicit PoseRandom(uint32_t seed) : engine(seed) {}
};

struct PoseSolveResult {
    rs_pose_out out{};
    std::vector<uint8_t> inlier_mask;
    // taps for parity debugging
    std::vector<std::vector<int>> subsets;      // subset per started iteration (residual order)
    std::vector<Pose7> candidate_poses;         // LM result per iteration (identity quaternion+0 when failed)
    std::vector<int> candidate_ok;
    std::vector<double> candidate_scores;
    Pose7 ransac_best;
};

// Pose_Optimization::compute_optimized_pose (pose_optimization.cpp:264-300). Plane normals of the features are
// normalised on entry as the PlaneCoordinates ctors do. max_iterations<=0 -> 119; n_variance<0 -> 100.
PoseSolveResult pose_solve(const Intrinsics& K, const Pose7& cur, const std::vector<rs_match>& feats, int max_iterations,
                           int n_variance, PoseRandom& rnd, int lm_max_fev = 400);

// compute_optimized_global_pose (:302-359) on a feature list; returns false when the reference would.
bool optimized_global_pose(const Intrinsics& K, const Pose7& cur, const std::vector<rs_match>& feats, Pose7& out,
                           int lm_max_fev = 400, LMResult* info = nullptr);

int ransac_default_iterations();  // 119, pose_optimization.cpp:129-132

}  // namespace oracle
