// TEST INFRASTRUCTURE - force-included in front of the reference's CAPE translation units (oracle/ref_shim/Makefile).
// Stands in for the reference headers that only exist to pull third-party code in, under their own include guards:
//   src/types.hpp (Eigen typedefs + Eigen-internal functor traits)  ->  the same type names on oracle/ref_shim/ref_eigen.hpp
#pragma once
#include <cmath>
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
struct quaternion {
    double w_ = 1, x_ = 0, y_ = 0, z_ = 0;
};
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
template <class M>
bool is_covariance_valid(const M& m)
{
    if (!m.allFinite()) return false;
    for (Eigen::Index i = 0; i < m.rows(); ++i)
        for (Eigen::Index j = 0; j < i; ++j)
            if (std::abs(m(i, j) - m(j, i)) > 1e-9 * (std::abs(m(i, j)) + std::abs(m(j, i)) + 1e-300)) return false;
    return true;
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
