// TEST INFRASTRUCTURE - C entry point of oracle/_ref/libref_pose.so: the reference's own pose-solve translation units
// (pose_optimization.cpp, levenberg_marquardt_functors.cpp, utils/pose.cpp, utils/camera_transformation.cpp, the coordinate
// classes, ransac.hpp, random.hpp - compiled where they lie, unmodified) plus the three optimisation-feature classes, whose
// declarations and member definitions are cut out of map_management/map_features/map_{point,primitive,point2d}.{hpp,cpp} AT BUILD
// TIME (oracle/ref_shim/Makefile, into scratch files under /tmp deleted after the compile - those files also define the map classes, which drag the
// whole local map in; nothing of them is committed here). Third-party stand-ins: ref_eigen.hpp, and for
// Eigen::LevenbergMarquardt<NumericalDiff<...>> the oracle's restated MINPACK lmdif (ref_eigen_lm.hpp) - so this build pins what the
// REFERENCE wrote (RANSAC loop and early stop, std::shuffle subset draws, the residual functor, inlier tests, the per-feature random
// variations and the Monte-Carlo covariance), not the LM iterate path.
#define private public
#define protected public
#include "coordinates/inverse_depth_coordinates.hpp"
#include "coordinates/plane_coordinates.hpp"
#include "coordinates/point_coordinates.hpp"
#include "matches_containers.hpp"
#include "outputs/logger.hpp"
#include "parameters.hpp"
#include "pose_optimization/pose_optimization.hpp"
#include "pose_optimization/levenberg_marquardt_functors.hpp"
#include "utils/camera_transformation.hpp"
#include "utils/random.hpp"
#undef private
#undef protected

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>

#include "../../include/rgbdslam_b200.h"
#include "../pose.hpp"
#include "ref_common.inc"

namespace rgbd_slam::map_management {
#include "gen_features_decl.inc"
#include "gen_features_impl.inc"
}  // namespace rgbd_slam::map_management

// utils/angle_utils.cpp is Eigen::AngleAxisd / Eigen::EulerAngles (unsupported module) and nothing else: third-party algebra.
// Stand-ins with the oracle's restatement of the same products (oracle/pose.cpp: quaternion_from_euler); the second one only
// feeds PoseBase::display.
namespace rgbd_slam::utils {
quaternion get_quaternion_from_euler_angles(const EulerAngles& e) noexcept
{
    double q[4];
    oracle::quaternion_from_euler(e.yaw, e.pitch, e.roll, q);
    return quaternion(q[0], q[1], q[2], q[3]);
}
EulerAngles get_euler_angles_from_quaternion(const quaternion& quat) noexcept
{
    const vector3 a = quat.toRotationMatrix().eulerAngles(0, 1, 2);
    return EulerAngles(a.z(), a.y(), a.x());
}
}  // namespace rgbd_slam::utils

using namespace rgbd_slam;

namespace {
// rs_match (include/rgbdslam_b200.h) -> the reference's optimisation features, ids = index in the list
matches_containers::match_container build_features(const rs_match* m, int n)
{
    matches_containers::match_container matches;
    for (int i = 0; i < n; ++i) {
        const rs_match& f = m[i];
        if (f.type == RS_FEAT_POINT)
            matches.push_back(std::make_shared<map_management::PointOptimizationFeature>(
                    ScreenCoordinate2D(f.obs[0], f.obs[1]), WorldCoordinate(f.map[0], f.map[1], f.map[2]),
                    vector3(f.sigma[0], f.sigma[1], f.sigma[2]), size_t(i), size_t(i)));
        else if (f.type == RS_FEAT_PLANE)
            matches.push_back(std::make_shared<map_management::PlaneOptimizationFeature>(
                    PlaneCameraCoordinates(vector4(f.obs[0], f.obs[1], f.obs[2], f.obs[3])),
                    PlaneWorldCoordinates(vector4(f.map[0], f.map[1], f.map[2], f.map[3])),
                    vector4(f.sigma[0], f.sigma[1], f.sigma[2], f.sigma[3]), size_t(i), size_t(i)));
        else {
            vector6 dev;
            dev.setZero();
            dev(InverseDepthWorldPoint::inverseDepthIndex) = f.sigma[0];
            dev(InverseDepthWorldPoint::thetaIndex) = f.sigma[1];
            dev(InverseDepthWorldPoint::phiIndex) = f.sigma[2];
            matches.push_back(std::make_shared<map_management::Point2dOptimizationFeature>(
                    ScreenCoordinate2D(f.obs[0], f.obs[1]),
                    InverseDepthWorldPoint(WorldCoordinate(f.map[0], f.map[1], f.map[2]), f.map[3], f.obs[2], f.obs[3]), dev,
                    size_t(i), size_t(i)));
        }
    }
    return matches;
}
void store_pose(const utils::PoseBase& p, double* pose7)
{
    const vector3 t = p.get_position();
    const quaternion q = p.get_orientation_quaternion();
    pose7[0] = t.x(), pose7[1] = t.y(), pose7[2] = t.z();
    pose7[3] = q.w(), pose7[4] = q.x(), pose7[5] = q.y(), pose7[6] = q.z();
}
void load_parameters()
{
    static std::once_flag once;
    std::call_once(once, []() { Parameters::load_defaut(); });
}
}  // namespace

extern "C" {

// Pose_Optimization::compute_optimized_pose on one frame's matches, on a fresh thread so that the thread-local engine of
// utils/random.hpp starts from its MAKE_DETERMINISTIC seed (0). Returns 1 when the reference returned true.
// inlier_mask[i] = match i ended in featureSets._inliers.
int ref_pose_solve(const double* cur_pose7, const rs_match* m, int n, double* out_pose7, double* out_cov36, uint8_t* inlier_mask)
{
    load_parameters();
    int ok = 0;
    std::thread worker([&]() {
        const matches_containers::match_container matches = build_features(m, n);
        const utils::PoseBase current(vector3(cur_pose7[0], cur_pose7[1], cur_pose7[2]),
                                      quaternion(cur_pose7[3], cur_pose7[4], cur_pose7[5], cur_pose7[6]));
        utils::Pose optimized;
        matches_containers::match_sets sets;
        ok = pose_optimization::Pose_Optimization::compute_optimized_pose(current, matches, optimized, sets) ? 1 : 0;
        std::memset(inlier_mask, 0, size_t(n));
        for (const auto& f : sets._inliers) inlier_mask[f->_detectedFeatureId] = 1;
        store_pose(optimized, out_pose7);
        const matrix66 cov = optimized.get_pose_variance();
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) out_cov36[6 * r + c] = cov(r, c);
    });
    worker.join();
    return ok;
}

// utils::PoseBase(position, orientation) as the reference stores it (set_parameters normalises the quaternion): the pose a caller
// of compute_optimized_pose holds, hence the pose to hand to the oracle / the library for the same problem.
void ref_pose_base(const double* pose7_in, double* pose7_out)
{
    const utils::PoseBase p(vector3(pose7_in[0], pose7_in[1], pose7_in[2]),
                            quaternion(pose7_in[3], pose7_in[4], pose7_in[5], pose7_in[6]));
    store_pose(p, pose7_out);
}

// Global_Pose_Estimator::operator() (levenberg_marquardt_functors.cpp:128-169) at the coefficient vector x; fvec holds
// sum(get_feature_part_count) values. Returns that count.
int ref_pose_residuals(const rs_match* m, int n, const double* x6, double* fvec)
{
    load_parameters();
    const matches_containers::match_container matches = build_features(m, n);
    size_t parts = 0;
    for (const auto& f : matches) parts += f->get_feature_part_count();
    const pose_optimization::Global_Pose_Estimator estimator(parts, matches);
    Eigen::Matrix<double, 6, 1> x;
    for (int k = 0; k < 6; ++k) x(k) = x6[k];
    vectorxd values(parts);
    values.setZero();
    estimator(x, values);
    for (size_t k = 0; k < parts; ++k) fvec[k] = values(k);
    return int(parts);
}

// Pose_Optimization::compute_optimized_global_pose: one LM over all matches from cur_pose7. Returns 1 on success.
int ref_pose_lm(const double* cur_pose7, const rs_match* m, int n, double* out_pose7)
{
    load_parameters();
    const matches_containers::match_container matches = build_features(m, n);
    const utils::PoseBase current(vector3(cur_pose7[0], cur_pose7[1], cur_pose7[2]),
                                  quaternion(cur_pose7[3], cur_pose7[4], cur_pose7[5], cur_pose7[6]));
    utils::PoseBase result;
    const bool ok = pose_optimization::Pose_Optimization::compute_optimized_global_pose(current, matches, result);
    store_pose(result, out_pose7);
    return ok ? 1 : 0;
}

// get_optimization_coefficient_from_pose / get_pose_from_optimization_coefficients (levenberg_marquardt_functors.cpp:74-86)
void ref_pose_coefficients(const double* pose7, double* x6)
{
    const utils::PoseBase p(vector3(pose7[0], pose7[1], pose7[2]), quaternion(pose7[3], pose7[4], pose7[5], pose7[6]));
    const vector6 x = pose_optimization::get_optimization_coefficient_from_pose(p);
    for (int k = 0; k < 6; ++k) x6[k] = x(k);
}
void ref_pose_from_coefficients(const double* x6, double* pose7)
{
    vector6 x;
    for (int k = 0; k < 6; ++k) x(k) = x6[k];
    store_pose(pose_optimization::get_pose_from_optimization_coefficients(x), pose7);
}

// Diagnostic: the oracle's LM driven by the REFERENCE functor from cur_pose7, with the oracle's own residual function evaluated
// beside it at every point the LM visits. Returns the number of evaluations whose residual vectors differ in any bit;
// first_x6 / first_pair = the first such point and (oracle value, reference value) of its first differing residual.
int ref_pose_trace_lm(const double* cur_pose7, const rs_match* m, int n, double* first_x6, double* first_pair, int* first_index)
{
    load_parameters();
    const matches_containers::match_container matches = build_features(m, n);
    size_t parts = 0;
    for (const auto& f : matches) parts += f->get_feature_part_count();
    const pose_optimization::Global_Pose_Estimator estimator(parts, matches);
    std::vector<rs_match> feats(m, m + n);
    std::vector<double> mine(parts);
    vectorxd theirs(parts);
    int differing = 0;
    const oracle::ResidualFn fn = [&](const double* xx, double* out) {
        Eigen::Matrix<double, 6, 1> x;
        for (int k = 0; k < 6; ++k) x(k) = xx[k];
        estimator(x, theirs);
        oracle::pose_residuals(oracle::Intrinsics{}, feats, xx, mine.data());
        bool bad = false;
        for (size_t k = 0; k < parts; ++k) {
            out[k] = theirs(k);
            if (mine[k] != theirs(k) && !bad) {
                bad = true;
                if (differing == 0) {
                    for (int j = 0; j < 6; ++j) first_x6[j] = xx[j];
                    first_pair[0] = mine[k], first_pair[1] = theirs(k);
                    *first_index = int(k);
                }
            }
        }
        differing += bad;
    };
    oracle::Pose7 start;
    for (int k = 0; k < 3; ++k) start.t[k] = cur_pose7[k];
    for (int k = 0; k < 4; ++k) start.q[k] = cur_pose7[3 + k];
    double x[6];
    oracle::coefficients_from_pose(start, x);
    oracle::lm_minimize(fn, int(parts), x, 400);
    return differing;
}

// IOptimizationFeature::is_inlier of every match under pose7 (get_features_inliers_outliers, pose_optimization.cpp:33-72)
void ref_pose_inliers(const double* pose7, const rs_match* m, int n, uint8_t* inlier_mask)
{
    load_parameters();
    const matches_containers::match_container matches = build_features(m, n);
    const WorldToCameraMatrix w2c = utils::compute_world_to_camera_transform(quaternion(pose7[3], pose7[4], pose7[5], pose7[6]),
                                                                             vector3(pose7[0], pose7[1], pose7[2]));
    int i = 0;
    for (const auto& f : matches) inlier_mask[i++] = f->is_inlier(w2c) ? 1 : 0;
}

}  // extern "C"
