// TEST INFRASTRUCTURE - force-included in front of the reference's CAPE translation units (oracle/ref_shim/Makefile).
// Stands in for the reference headers that only exist to pull third-party code in, under their own include guards:
//   src/types.hpp (Eigen typedefs + Eigen-internal functor traits)  ->  the same type names on oracle/ref_shim/ref_eigen.hpp
#pragma once
#include <cmath>
#include <string>
#include <vector>

#include "ref_cv.hpp"
#include "ref_eigen.hpp"

#ifndef RGBDSLAM_TYPES_HPP
#define RGBDSLAM_TYPES_HPP
#endif

namespace rgbd_slam {

constexpr double EulerToRadian = M_PI / 180.0;

using Matrixb = Eigen::Matrix<unsigned char, Eigen::Dynamic, Eigen::Dynamic>;   // bool coefficients, addressable storage
using matrixf = Eigen::MatrixXf;
using matrixd = Eigen::MatrixXd;
using vector2 = Eigen::Vector2d;
using vectorxd = Eigen::VectorXd;
using vectorb = Eigen::Matrix<unsigned char, Eigen::Dynamic, 1>;
using vector3 = Eigen::Matrix<double, 3, 1>;
using vector4 = Eigen::Vector4d;
using matrix22 = Eigen::Matrix2d;
using matrix33 = Eigen::Matrix3d;
using matrix34 = Eigen::Matrix<double, 3, 4>;
using matrix43 = Eigen::Matrix<double, 4, 3>;
using matrix44 = Eigen::Matrix4d;
using quaternion = Eigen::Quaternion<double>;
using vector6 = Eigen::Matrix<double, 6, 1>;
using vector7 = Eigen::Matrix<double, 7, 1>;
using matrix66 = Eigen::Matrix<double, 6, 6>;
using matrix77 = Eigen::Matrix<double, 7, 7>;

struct ScreenCoordinate2DCovariance : public matrix22 {};
struct ScreenCoordinateCovariance : public matrix33 {};
struct CameraCoordinateCovariance : public matrix33 {};
struct WorldCoordinateCovariance : public matrix33 {};

struct TransitionMatrix : public matrix44 {
    using matrix44::matrix44;
    [[nodiscard]] matrix33 rotation() const noexcept { return this->block<3, 3>(0, 0); }
    [[nodiscard]] vector3 translation() const noexcept { return vector3((*this)(0, 3), (*this)(1, 3), (*this)(2, 3)); }
};
struct WorldToCameraMatrix : public TransitionMatrix {};
struct CameraToWorldMatrix : public TransitionMatrix {};
struct PlaneWorldToCameraMatrix : public TransitionMatrix {};
struct PlaneCameraToWorldMatrix : public TransitionMatrix {};

struct EulerAngles {
    double yaw = 0, pitch = 0, roll = 0;
};

template <class T>
T constexpr inline SQR(const T x)
{
    return x * x;
}

using vector3_vector = std::vector<vector3>;

}  // namespace rgbd_slam

//   src/utils/covariances.hpp (Eigen LDLT / self-adjoint views)  ->  the one predicate the CAPE sources reference
#ifndef RGBDSLAM_UTILS_COVARIANCES_HPP
#define RGBDSLAM_UTILS_COVARIANCES_HPP
#endif
namespace rgbd_slam::utils {
[[nodiscard]] double get_depth_quantization(const double depht) noexcept;   // defined from the reference's own text: see Makefile
// is_covariance_valid<N> (covariances.hpp:13-50): finite; N == 1: non-negative; symmetric in Eigen's isApprox sense
// (|C - C^T|^2 <= 1e-24 min(|C|^2, |C^T|^2)); selfadjointView<Upper>().ldlt() without a negative pivot (Eigen's diagonally
// pivoted LDL^T, as oracle/kalman.cpp restates it)
template <class M>
bool is_covariance_valid(const M& c, std::string& reason)
{
    const int N = int(c.rows());
    if (c.hasNaN() || !c.allFinite()) {
        reason = "invalid values";
        return false;
    }
    if (N == 1) return c(0, 0) >= 0;
    double diff2 = 0, n2 = 0;
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            const double dd = c(i, j) - c(j, i);
            diff2 += dd * dd;
            n2 += c(i, j) * c(i, j);
        }
    if (!(diff2 <= 1e-12 * 1e-12 * n2)) {
        reason = "not symetrical";
        return false;
    }
    std::vector<double> a(size_t(N) * N);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) a[size_t(i) * N + j] = c(std::min(i, j), std::max(i, j));
    bool neg = false;
    for (int k = 0; k < N; ++k) {
        int p = k;
        double best = std::fabs(a[size_t(k) * N + k]);
        for (int i = k + 1; i < N; ++i)
            if (std::fabs(a[size_t(i) * N + i]) > best) best = std::fabs(a[size_t(i) * N + i]), p = i;
        if (p != k) {
            for (int j = 0; j < N; ++j) std::swap(a[size_t(k) * N + j], a[size_t(p) * N + j]);
            for (int i = 0; i < N; ++i) std::swap(a[size_t(i) * N + k], a[size_t(i) * N + p]);
        }
        const double dkk = a[size_t(k) * N + k];
        if (dkk < 0) neg = true;
        if (std::fabs(dkk) <= 2.2250738585072014e-308) break;
        for (int i = k + 1; i < N; ++i) {
            const double l = a[size_t(i) * N + k] / dkk;
            for (int j = k + 1; j < N; ++j) a[size_t(i) * N + j] -= l * a[size_t(k) * N + j];
        }
    }
    if (neg) {
        reason = "not positive semi definite";
        return false;
    }
    return true;
}
template <class M>
bool is_covariance_valid(const M& c)
{
    std::string reason;
    return is_covariance_valid(c, reason);
}
}  // namespace rgbd_slam::utils


//   src/coordinates/polygon_coordinates.hpp, src/utils/polygon.hpp (boost::geometry, flann concave hull)  ->  a boundary holder.
//   The reference fits a polygon to the ordered boundary points and DROPS the plane when boost calls the fit invalid; this
//   stand-in keeps the points and accepts every boundary of three or more points, so plane_container can hold planes the
//   reference would have dropped. The label grids and the plane / cylinder segments - what the fixtures pin - are unaffected.
#ifndef RGBDSLAM_COORD_POLYGON_HPP
#define RGBDSLAM_COORD_POLYGON_HPP
#endif
#ifndef RGBDSLAM_UTILS_POLYGON_UTILS_HPP
#define RGBDSLAM_UTILS_POLYGON_UTILS_HPP
#endif
#include <string>
namespace rgbd_slam {
class CameraPolygon {
  public:
    std::vector<vector3> points;
    vector3 normal, center;
    CameraPolygon() = default;
    CameraPolygon(const std::vector<vector3>& pts, const vector3& n, const vector3& c) : points(pts), normal(n), center(c) {}
    bool is_valid(std::string&) const { return points.size() >= 3; }
    bool is_valid() const { return points.size() >= 3; }
    size_t boundary_length() const { return points.size(); }
};
}  // namespace rgbd_slam

//   src/features/primitives/shape_primitives.hpp (Plane / Cylinder value classes: covariances, polygons)  ->  records of what
//   find_primitives hands them
#ifndef RGBDSLAM_FEATURES_PRIMITIVES_PRIMITIVES_HPP
#define RGBDSLAM_FEATURES_PRIMITIVES_PRIMITIVES_HPP
#endif
namespace rgbd_slam::features::primitives {
class Plane {
  public:
    vector3 normal, centroid;
    double d = 0, mse = 0;
    unsigned point_count = 0;
    CameraPolygon polygon;
    template <class Segment>
    Plane(const Segment& s, const CameraPolygon& p)
        : normal(s.get_normal()), centroid(s.get_centroid()), d(s.get_plane_d()), mse(s.get_MSE()), point_count(s.get_point_count()), polygon(p)
    {
    }
    vector3 get_normal() const { return normal; }
    double get_d() const { return d; }
    const CameraPolygon& get_boundary_polygon() const { return polygon; }
};
class Cylinder {
  public:
    vector3 _normal;
    double _radius = 0;
    template <class Segment>
    explicit Cylinder(const Segment& s) : _normal(s.get_normal())
    {
        // Cylinder::Cylinder (shape_primitives.cpp:17-25): mean radius over the segments
        for (unsigned i = 0; i < s.get_segment_count(); ++i) _radius += s.get_radius(i);
        _radius /= s.get_segment_count();
    }
};
using plane_container = std::vector<Plane>;
using cylinder_container = std::vector<Cylinder>;
}  // namespace rgbd_slam::features::primitives


//   src/features/keypoints/keypoint_handler.hpp, src/features/lines/line_detection.hpp (OpenCV features2d / LSD)  ->  the type names
//   matches_containers.hpp mentions; nothing on the pose path touches them
#ifndef RGBDSLAM_FEATURES_KEYPOINTS_KEYPOINTS_HANDLER_HPP
#define RGBDSLAM_FEATURES_KEYPOINTS_KEYPOINTS_HANDLER_HPP
#endif
#ifndef RGBDSLAM_FEATURES_LINES_LINE_DETECTION_HPP
#define RGBDSLAM_FEATURES_LINES_LINE_DETECTION_HPP
#endif
#include "coordinates/point_coordinates.hpp"   // (the real header: keypoint_handler.hpp is how matches_containers.hpp gets it)
namespace rgbd_slam::features::keypoints {
struct Keypoint_Handler {};
struct KeypointsWithIdStruct {};
}  // namespace rgbd_slam::features::keypoints
namespace rgbd_slam::features::lines {
struct line_stub {};
using line_container = std::vector<line_stub>;
}  // namespace rgbd_slam::features::lines
