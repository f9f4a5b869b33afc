// Reference-side binding of librgbdslam_b200.so: the translation unit a maintainer of BaptisteHudyma/RGB-D-SLAM adds
// to libprimitives / libposeOptimization so that the existing pipeline (src/rgbd_slam.cpp:110-112,198-199,295) runs the
// per-frame hot path on a B200 through the C-ABI of include/rgbdslam_b200.h.
//
// NOT compiled in this repository: it needs the reference's own headers (Eigen, OpenCV, boost), which are not
// available here. It is written against the reference at 183011f; every call names the member it replaces.
// Required reference-side patches (five small changes; the first three:)
//   1. Plane_Segment: public ctor `Plane_Segment(uint pointCount, const double sums[9])` that fills _pointCount and
//      _Sx.._Szx and calls fit_plane()   (members are private: plane_segment.hpp:118-140)
//   2. PointOptimizationFeature / PlaneOptimizationFeature / Point2dOptimizationFeature: `friend struct rs_adaptor::Flatten;`
//      (their _matchedPoint/_mapPoint/... members are protected: map_point.hpp:40-43, map_primitive.hpp:40-43)
//   3. CMake: link rgbdslam_b200 into `primitives` and `poseOptimization` (CMakeLists.txt:117-123,136-139)
#include <rgbdslam_b200.h>

#include <memory>
#include <stdexcept>
#include <limits>
#include <vector>

#include "features/primitives/primitive_detection.hpp"
#include "features/primitives/shape_primitives.hpp"
#include "map_management/map_features/map_point.hpp"
#include "map_management/map_features/map_point2d.hpp"
#include "map_management/map_features/map_primitive.hpp"
#include "outputs/logger.hpp"
#include "parameters.hpp"
#include "pose_optimization/pose_optimization.hpp"
#include "utils/random.hpp"

namespace rgbd_slam::rs_adaptor {

// ---- CAPE: Depth_Map_Transformation::get_organized_cloud_array + Primitive_Detection::find_primitives -------------

class CapeContext
{
  public:
    CapeContext(const uint width, const uint height)
    {
        // Parameters::get_camera_1_* are what Depth_Map_Transformation::init_matrices reads (depth_map_transformation.cpp:147-173)
        _ctx = rs_cape_create(int(width), int(height), int(parameters::detection::depthMapPatchSize_px),
                              Parameters::get_camera_1_focal_x(), Parameters::get_camera_1_focal_y(),
                              Parameters::get_camera_1_center_x(), Parameters::get_camera_1_center_y(),
                              /*max_batch*/ 1, /*device*/ 0);
        if (_ctx == nullptr) throw std::runtime_error(std::string("rs_cape_create: ") + rs_last_error());
        const size_t nc = size_t(rs_cape_cells_per_frame(_ctx));
        _planeLabels.resize(nc);
        _cylLabels.resize(nc);
        _planes.resize(RS_MAX_PLANES);
        _cyls.resize(RS_MAX_CYL_REGIONS);
        _boundary.resize(size_t(rs_cape_max_boundary(_ctx)) * 3);
    }
    ~CapeContext() { rs_cape_destroy(_ctx); }

    // RGBD_SLAM::rectify_depth (rgbd_slam.cpp:85-97 -> Depth_Map_Transformation::rectify_depth): instead of rectifying on
    // the host before track(), switch the device-side rectification on once; every find_primitives call below then
    // re-projects its input into camera 1's image before the plane fit (same last-writer-wins result as the serial scan).
    void enable_depth_rectification()
    {
        const matrix44 t = Parameters::get_camera_2_to_camera_1_transformation();
        double rowMajor[16];
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) rowMajor[4 * r + c] = t(r, c);   // Eigen stores column-major
        if (rs_cape_set_rectification(_ctx, rowMajor, 1) != RS_OK)
            throw std::runtime_error(std::string("rs_cape_set_rectification: ") + rs_last_error());
    }

    // Raw sensor image (CV_16U, what cv::imread(IMREAD_ANYDEPTH) returns in examples/main_TUM.cpp:221): the
    // convertTo(CV_32FC1, alpha) of :242 happens on the device, the PCIe copy is half the size.
    int run_u16(const cv::Mat_<ushort>& rawDepth, const double alpha, const rs_cape_outputs& out)
    {
        if (not rawDepth.isContinuous()) throw std::invalid_argument("depth image must be continuous");
        return rs_cape_run_u16(_ctx, rawDepth.ptr<ushort>(), alpha, 1, utils::Random::_seed, &out);
    }

    // Drop-in body of Primitive_Detection::find_primitives (primitive_detection.cpp:119-166). The organized cloud
    // argument of the reference is not needed: the back-projection is fused into the plane-fit kernel.
    void find_primitives(const cv::Mat_<float>& depthImage,
                         features::primitives::plane_container& planeContainer,
                         features::primitives::cylinder_container& cylinderContainer)
    {
        planeContainer.clear();
        cylinderContainer.clear();
        if (not depthImage.isContinuous()) throw std::invalid_argument("depth image must be continuous");
        rs_cape_outputs out {};
        out.plane_labels = _planeLabels.data();
        out.cyl_labels = _cylLabels.data();
        out.planes = _planes.data();
        out.cyls = _cyls.data();
        out.boundary_xyz = _boundary.data();
        out.info = &_info;
        // utils::Random::_seed: 0 under MAKE_DETERMINISTIC, time(0) otherwise (random.hpp:60-65); the reference
        // restarts the engine every frame because find_primitives runs on a fresh std::async thread.
        const int rc = rs_cape_run(_ctx, depthImage.ptr<float>(), 1, utils::Random::_seed, &out);
        if (rc != RS_OK)
        {
            outputs::log_error(std::string("rs_cape_run failed: ") + rs_last_error());
            return;
        }
        // add_planes_to_primitives (primitive_detection.cpp:562-648): polygon fit + validity stay on the host
        for (int k = 0; k < _info.n_planes; ++k)
        {
            const rs_plane_out& p = _planes[size_t(k)];
            if (not p.is_final) continue;
            if (p.n_boundary < 3)
            {
                outputs::log_warning("Could not find a correct boundary polygon, rejecting plane segment");
                continue;
            }
            std::vector<vector3> orderedBoundary;
            orderedBoundary.reserve(size_t(p.n_boundary));
            for (int i = 0; i < p.n_boundary; ++i)
            {
                const double* q = &_boundary[size_t(p.boundary_offset + i) * 3];
                orderedBoundary.emplace_back(q[0], q[1], q[2]);
            }
            const features::primitives::Plane_Segment planeSegment(uint(p.count), p.S);  // patch 1
            const CameraPolygon polygon(orderedBoundary, planeSegment.get_normal(), planeSegment.get_center());
            std::string debug;
            if (polygon.is_valid(debug) and polygon.boundary_length() >= 3)
                planeContainer.emplace_back(planeSegment, polygon);
            else
                outputs::log_error("Polyfit error: " + debug);
        }
        // add_cylinders_to_primitives (primitive_detection.cpp:705-734): only the axis is meaningful in the
        // reference's Cylinder (the radius is averaged over zero segments -> NaN)
        for (int r = 0; r < _info.n_cyl_regions; ++r)
        {
            const rs_cyl_out& c = _cyls[size_t(r)];
            for (int s = 0; s < c.n_segments; ++s)
                if (c.kept[s]) cylinderContainer.emplace_back(make_cylinder(c, s));
        }
    }

  private:
    // One reference Cylinder per kept (region, sub-segment) pair, as add_cylinders_to_primitives emplaces one per surviving
    // cylinder id (primitive_detection.cpp:705-734). The reference builds it from a COPY of the region's Cylinder_Segment whose
    // segment count the copy constructor resets to 0 (cylinder_segment.cpp:23-33), so Cylinder::Cylinder averages the radius
    // over zero segments: _radius = 0.0 / 0 = NaN, and only _normal (the region's axis) carries information
    // (shape_primitives.cpp:17-25). Drop-in means the same object: axis + NaN radius. Building with
    // -DRS_B200_CYLINDER_RADIUS hands the sub-segment's fitted radius over instead (what the reference presumably meant).
    // Needs patch 4 of INTEGRATION.md: the value constructor Cylinder(const vector3& normal, double radius).
    static features::primitives::Cylinder make_cylinder(const rs_cyl_out& c, const int segment)
    {
        const vector3 axis(c.axis[0], c.axis[1], c.axis[2]);
#ifdef RS_B200_CYLINDER_RADIUS
        return features::primitives::Cylinder(axis, c.radius[segment]);
#else
        (void)segment;
        return features::primitives::Cylinder(axis, std::numeric_limits<double>::quiet_NaN());
#endif
    }

    rs_cape_ctx* _ctx = nullptr;
    std::vector<int32_t> _planeLabels, _cylLabels;
    std::vector<rs_plane_out> _planes;
    std::vector<rs_cyl_out> _cyls;
    std::vector<double> _boundary;
    rs_cape_frame_info _info {};
};

// ---- pose: Pose_Optimization::compute_optimized_pose --------------------------------------------------------------

struct Flatten  // befriended by the two optimisation-feature classes (patch 2)
{
    static bool to_match(const matches_containers::feat_ptr& f, rs_match& m)
    {
        m = rs_match {};
        if (f->get_feature_type() == FeatureType::Point)
        {
            const auto& p = static_cast<const map_management::PointOptimizationFeature&>(*f);
            m.type = RS_FEAT_POINT;
            m.obs[0] = p._matchedPoint.x(), m.obs[1] = p._matchedPoint.y();
            for (int i = 0; i < 3; ++i) m.map[i] = p._mapPoint(i), m.sigma[i] = p._mapPointStandardDev(i);
            return true;
        }
        if (f->get_feature_type() == FeatureType::Plane)
        {
            const auto& p = static_cast<const map_management::PlaneOptimizationFeature&>(*f);
            m.type = RS_FEAT_PLANE;
            const vector4 o = p._matchedPlane.get_parametrization(), w = p._mapPlane.get_parametrization();
            for (int i = 0; i < 4; ++i) m.obs[i] = o(i), m.map[i] = w(i), m.sigma[i] = p._mapPlaneStandardDev(i);
            return true;
        }
        if (f->get_feature_type() == FeatureType::Point2d)
        {
            // inverse-depth map point (map_point2d.hpp; also befriended): packing documented at rs_match in the header
            const auto& p = static_cast<const map_management::Point2dOptimizationFeature&>(*f);
            m.type = RS_FEAT_POINT2D;
            m.obs[0] = p._matchedPoint.x(), m.obs[1] = p._matchedPoint.y();
            m.obs[2] = p._mapPoint.get_theta(), m.obs[3] = p._mapPoint.get_phi();
            const WorldCoordinate first = p._mapPoint.get_first_observation();
            for (int i = 0; i < 3; ++i) m.map[i] = first(i);
            m.map[3] = p._mapPoint.get_inverse_depth();
            // is_valid also wants the first observation's standard deviations finite and >= 0 (map_point2d.cpp:75-79)
            for (int i = 0; i < 3; ++i)
                if (not(p._mapPointStandardDev(i) >= 0)) return false;
            m.sigma[0] = p._mapPointStandardDev(InverseDepthWorldPoint::inverseDepthIndex);
            m.sigma[1] = p._mapPointStandardDev(InverseDepthWorldPoint::thetaIndex);
            m.sigma[2] = p._mapPointStandardDev(InverseDepthWorldPoint::phiIndex);
            return true;
        }
        return false;
    }
};

// Drop-in body of Pose_Optimization::compute_optimized_pose (pose_optimization.cpp:264-300).
inline bool compute_optimized_pose(rs_pose_ctx* ctx,
                                   const utils::PoseBase& currentPose,
                                   const matches_containers::match_container& matchedFeatures,
                                   utils::Pose& optimizedPose,
                                   matches_containers::match_sets& featureSets)
{
    std::vector<rs_match> flat;
    std::vector<matches_containers::feat_ptr> order;
    for (const auto& f: matchedFeatures)
    {
        rs_match m;
        if (not Flatten::to_match(f, m)) return false;
        flat.push_back(m);
        order.push_back(f);
    }
    const vector3 t = currentPose.get_position();
    const quaternion q = currentPose.get_orientation_quaternion();
    const double cur[7] = {t.x(), t.y(), t.z(), q.w(), q.x(), q.y(), q.z()};
    rs_pose_opts opts {};
    // The reference seeds its engines with time(0) unless it is built with MAKE_DETERMINISTIC (utils/random.hpp:57-64): only
    // the deterministic build promises a particular random stream, and only there is the reference's own std::mt19937
    // sequence reproduced (host draws on all host threads, one host round trip inside the solve: 6 k frames/s in batches).
    // The default build gets the counter-based on-device generator (no round trip: 38 k frames/s through the same call).
#ifdef MAKE_DETERMINISTIC
    opts.rng_mode = RS_RNG_REFERENCE;
#else
    opts.rng_mode = RS_RNG_DEVICE;
#endif
    opts.seed = utils::Random::_seed;
    opts.fx = Parameters::get_camera_1_focal_x(), opts.fy = Parameters::get_camera_1_focal_y();
    opts.cx = Parameters::get_camera_1_center_x(), opts.cy = Parameters::get_camera_1_center_y();
    rs_pose_out out {};
    std::vector<uint8_t> inlier(flat.size());
    if (rs_pose_solve(ctx, cur, flat.data(), int(flat.size()), &opts, &out, inlier.data()) != RS_OK)
    {
        outputs::log_error(std::string("rs_pose_solve failed: ") + rs_last_error());
        return false;
    }
    if (out.status != 1) return false;
    featureSets.clear();
    for (size_t i = 0; i < order.size(); ++i)
        (inlier[i] ? featureSets._inliers : featureSets._outliers).insert(
                (inlier[i] ? featureSets._inliers : featureSets._outliers).end(), order[i]);
    matrix66 cov;
    for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) cov(r, c) = out.cov[r * 6 + c];
    optimizedPose.set_parameters(vector3(out.pose[0], out.pose[1], out.pose[2]),
                                 quaternion(out.pose[3], out.pose[4], out.pose[5], out.pose[6]));
    optimizedPose.set_position_variance(cov);
    return true;
}

// ---- plane matching: the loop over the local map's planes that calls MapPlane::find_matches (map_primitive.cpp:91-161) -----
// One rs_plane_match call for the whole local map of the frame instead of one find_matches per map plane. Needs patch 5
// (INTEGRATION.md): `friend struct rs_adaptor::PlaneMatcher;` in utils::Polygon (the boost ring `_polygon` is protected,
// polygon.hpp:198-199) - the public get_unprojected_boundary() would round-trip the ring through 3-D and back.
struct PlaneMatcher
{
    // (n, d), the polygon's frame and its outer ring, as rs_polygon_plane + (x, y) pairs appended to `xy`
    static void flatten(const vector4& parametrization, const utils::Polygon& polygon, rs_polygon_plane& out, std::vector<double>& xy)
    {
        for (int i = 0; i < 3; ++i)
        {
            out.normal[i] = parametrization(i);
            out.center[i] = polygon.get_center()(i);
            out.x_axis[i] = polygon.get_x_axis()(i);
            out.y_axis[i] = polygon.get_y_axis()(i);
        }
        out.d = parametrization(3);
        out.first_vertex = int32_t(xy.size() / 2);
        const auto& ring = polygon._polygon.outer();   // closed ring (first == last), clockwise after boost::geometry::correct
        out.n_vertices = int32_t(ring.size());
        for (const auto& p: ring)
        {
            xy.push_back(p.x());
            xy.push_back(p.y());
        }
    }

    // selected[m] = index in `detectedPlanes` matched by mapPlanes[m], or -1: what the loop over the map's planes leaves in
    // each plane's matchIndexes (find_matches per plane, isDetectedFeatureMatched updated between two planes).
    static std::vector<int> find_matches(const map_management::DetectedPlaneObject& detectedPlanes,
                                         const std::vector<const map_management::MapPlane*>& mapPlanes,
                                         const WorldToCameraMatrix& worldToCamera,
                                         const vectorb& isDetectedFeatureMatched,
                                         const bool useAdvancedSearch)
    {
        std::vector<rs_polygon_plane> det(detectedPlanes.size()), map(mapPlanes.size());
        std::vector<double> detXY, mapXY;
        std::vector<uint8_t> matched(detectedPlanes.size());
        for (size_t k = 0; k < detectedPlanes.size(); ++k)
        {
            flatten(detectedPlanes[k].get_parametrization().get_parametrization(), detectedPlanes[k].get_boundary_polygon(), det[k], detXY);
            matched[k] = isDetectedFeatureMatched[long(k)] ? 1 : 0;
        }
        for (size_t m = 0; m < mapPlanes.size(); ++m)
            flatten(mapPlanes[m]->get_parametrization().get_parametrization(), mapPlanes[m]->get_boundary_polygon(), map[m], mapXY);
        double w2c[16];
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) w2c[4 * r + c] = worldToCamera(r, c);   // Eigen stores column-major
        const int32_t detFirst[2] = {0, int32_t(det.size())}, mapFirst[2] = {0, int32_t(map.size())};
        std::vector<int32_t> selected(map.size(), -1);
        std::vector<double> interArea(map.size(), 0.0);
        // sequential = 1: the map planes are served in the order given and a detection taken by one is marked matched for the
        // next, as the loop of Feature_Map::get_matches does (feature_map.hpp:652-669); `matched` returns the updated mask
        if (rs_plane_match(/*device*/ 0, 1, w2c, det.data(), detFirst, detXY.data(), map.data(), mapFirst, mapXY.data(),
                           matched.data(), useAdvancedSearch ? 1 : 0, /*sequential*/ 1, selected.data(), interArea.data(),
                           matched.data()) != RS_OK)
        {
            outputs::log_error(std::string("rs_plane_match failed: ") + rs_last_error());
            return std::vector<int>(map.size(), -1);
        }
        return std::vector<int>(selected.begin(), selected.end());
    }
};

}  // namespace rgbd_slam::rs_adaptor
