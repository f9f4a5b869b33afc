// TEST INFRASTRUCTURE — CPU oracle (see linalg.hpp header). Restates the reference's CAPE path:
//   src/features/primitives/{depth_map_transformation,plane_segment,histogram,primitive_detection,
//   cylinder_segment}.* and src/utils/covariances.cpp:12-19.
// PIN: the reference has no test, golden vector or dataset for this path; its own translation units do
// compile here against stand-in third-party headers (oracle/ref_shim -> oracle/_ref/libref_cape.so), and this restatement must
// equal that build bit for bit (tests/test_reference_build.py: cell fits, label grids, planes, boundary points, cylinders,
// rectify_depth on scene v0, edge cases, 96 random rooms). Unpinned: the arithmetic inside Eigen / OpenCV beyond the
// cross-checks of tests/test_oracle_cape.py (LAPACK eigh, cv2 morphology) - the stand-ins share this oracle's restatements there.
#pragma once
#include <cstdint>
#include <vector>

#include "../include/rgbdslam_b200.h"
#include "linalg.hpp"

namespace oracle {

struct CapeConfig {
    int width = 640, height = 480, cell = 20;
    double fx = 550, fy = 550, cx = 320, cy = 240;
};

// Plane_Segment (plane_segment.hpp): 9 sums, count, fit results.
struct PlaneSeg {
    int count = 0;
    bool planar = false;
    double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // Sx Sy Sz Sxs Sys Szs Sxy Syz Szx
    Vec3 centroid;
    Vec3 normal;
    double d = 0;
    double mse = DBL_MAX;
    double score = 0;
    void clear();                       // clear_plane_parameters, plane_segment.cpp:289-310
    void expand(const PlaneSeg& o);     // expand_segment, :170-190
    void fit_plane();                   // :232-284
    PlaneSeg copy() const;              // copy ctor (:18-36) incl. the PlaneCoordinates re-normalisation
    bool can_be_merged(const PlaneSeg& p, double maxMatchDistance) const;  // :322-326
};

double depth_quantization(double z);    // covariances.cpp:12-19

struct CapeFrame {
    std::vector<rs_cell_out> cells;
    std::vector<int32_t> plane_grid, plane_labels, cyl_labels, cyl_region_seg;
    std::vector<rs_plane_out> planes;   // n_planes entries
    std::vector<rs_cyl_out> cyls;       // n_cyl_regions entries
    std::vector<double> boundary_xyz;   // 3 * n_boundary
    rs_cape_frame_info info{};
};

// Back-projection factors kx[c], ky[r] (point_coordinates.cpp:79-83 via Eigen's 3x3 cofactor inverse).
void backprojection_factors(const CapeConfig& cfg, std::vector<double>& kx, std::vector<double>& ky);

// get_organized_cloud_array + init_planar_cell_fitting only (K1's scope). cloud (optional) receives the
// organized cloud [W*H x 3] column-major as the reference builds it.
void cape_cell_fit(const CapeConfig& cfg, const float* depth, std::vector<PlaneSeg>& grid, std::vector<float>& tols,
                   std::vector<float>* cloud = nullptr);

// Whole find_primitives (primitive_detection.cpp:119-166) for one frame. seed = utils::Random::_seed.
void cape_run(const CapeConfig& cfg, const float* depth, uint32_t seed, CapeFrame& out);

void cell_record(const PlaneSeg& s, float tol, int cellSize, rs_cell_out& o);

// Depth_Map_Transformation::rectify_depth (depth_map_transformation.cpp:23-87): depth image of camera 2 re-projected
// into the image of camera 1, serial row-major scan (the MAKE_DETERMINISTIC build), last writer wins.
// cam2_to_cam1 = Parameters::get_camera_2_to_camera_1_transformation(), row-major 4x4. out = H*W floats.
// Pinned by the compiled reference source (tests/test_reference_build.py::test_rectify_depth) up to the summation order inside
// Eigen's 4x4 * homogeneous product, taken as ((T0 x + T1 y) + T2 z) + T3 here and in the stand-in; it can only matter when a
// projected coordinate lies within an ulp of a pixel boundary.
void rectify_depth(const CapeConfig& cfg, const double cam2_to_cam1[16], const float* depth, float* out);

// The restatement of cv::erode / cv::dilate (3x3 square or cross kernel, anchor at the centre, one iteration) that the
// boundary and cylinder-opening steps use (primitive_detection.cpp:48-54,596-598,678-680,719-721). borderZero = the
// reference's explicit BORDER_CONSTANT / Scalar(0) erode; otherwise OpenCV's default morphology border. Pinned against
// the real OpenCV (cv2 wheel) in tests/test_oracle_cape.py.
std::vector<unsigned char> cape_morphology(const std::vector<unsigned char>& m, int rows, int cols, bool erode, bool cross,
                                           bool borderZero);

}  // namespace oracle
