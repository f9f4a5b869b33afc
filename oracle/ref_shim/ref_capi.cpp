// TEST INFRASTRUCTURE - C entry points of oracle/_ref/libref_cape.so: the reference's own CAPE translation units (compiled where
// they lie under /root/reference, unmodified, against the stand-in headers of this directory) run on a depth frame.
// Defined here: what the reference defines in translation units that cannot be built without third-party libraries and that
// the CAPE path only touches at its edges - the logger sinks, Parameters::load_defaut (the reference's default intrinsics,
// parameters.cpp:59-74) - and get_depth_quantization, whose body is taken verbatim from the reference's covariances.cpp at
// build time (oracle/ref_shim/Makefile writes it to a scratch file under /tmp that is deleted after the compile; nothing of it is committed or kept).
#define private public      // the label grids are private members of Primitive_Detection
#define protected public
#include "features/primitives/depth_map_transformation.hpp"
#include "features/primitives/primitive_detection.hpp"
#undef private
#undef protected
#include "outputs/logger.hpp"
#include "parameters.hpp"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>

#include "ref_common.inc"

using namespace rgbd_slam;

extern "C" {

// One frame (640 x 480 float depth, mm) through get_organized_cloud_array + find_primitives, on a fresh thread: the reference
// runs find_primitives through std::async, so the thread-local engine of utils/random.hpp restarts from its seed every frame
// (0 under MAKE_DETERMINISTIC). Outputs: the private label grids (plane segment index + 1, cylinder index + 1 per cell), per
// cell the fitted plane of init_planar_cell_fitting (planar flag, normal, d, MSE, point count), and the primitives returned.
// planes_out: [max_planes][6] = normal, d, mse, boundary point count; cyls_out: [max_cyls][4] = normal, mean radius.
// boundary_out (may be null): the ordered boundary points compute_plane_segment_boundary hands to the polygon fit, plane after
// plane, [max_boundary][3]; planes_out[k][5] says how many belong to plane k.
int ref_cape_run(const float* depth, int width, int height, int32_t* plane_grid, int32_t* cyl_grid, int32_t* cell_planar,
                 double* cell_normal_d_mse, int32_t* cell_count, double* planes_out, int max_planes, int32_t* n_planes, double* cyls_out,
                 int max_cyls, int32_t* n_cyls, double* boundary_out, int max_boundary)
{
    static bool loaded = false;
    if (!loaded) {
        Parameters::load_defaut();
        loaded = true;
    }
    int rc = 0;
    std::thread worker([&]() {
        const unsigned cell = parameters::detection::depthMapPatchSize_px;
        features::primitives::Depth_Map_Transformation depthOps(width, height, cell);
        features::primitives::Primitive_Detection detector(width, height);
        // The reference keeps ONE detector for the whole sequence, so after the first frames its segment vectors never
        // reallocate. That matters at the last bit: Plane_Segment has a copy constructor but no move constructor, a
        // reallocating push_back copies every stored segment, and every copy re-normalises its normal (PlaneCoordinates'
        // copy constructor) - a fresh detector per frame would add growth-dependent 1-ulp changes (seen on 4 of 1024 rooms).
        detector._planeSegments.reserve(1024);
        detector._cylinderSegments.reserve(1024);
        cv::Mat_<float> image(height, width);
        std::memcpy(image.data, depth, sizeof(float) * size_t(width) * height);
        matrixf cloud;
        if (!depthOps.get_organized_cloud_array(image, cloud)) {
            rc = 1;
            return;
        }
        features::primitives::plane_container planes;
        features::primitives::cylinder_container cylinders;
        detector.find_primitives(cloud, image, planes, cylinders);
        const int vc = detector._gridPlaneSegmentMap.rows, hc = detector._gridPlaneSegmentMap.cols;
        for (int r = 0; r < vc; ++r)
            for (int c = 0; c < hc; ++c) {
                const int i = r * hc + c;
                plane_grid[i] = detector._gridPlaneSegmentMap.at<int>(r, c);
                cyl_grid[i] = detector._gridCylinderSegMap.at<int>(r, c);
                const auto& cellSeg = detector._planeGrid[size_t(i)];
                cell_planar[i] = cellSeg.is_planar() ? 1 : 0;
                cell_count[i] = int32_t(cellSeg.get_point_count());
                const vector3 n = cellSeg.get_normal();
                double* o = cell_normal_d_mse + 5 * size_t(i);
                o[0] = n.x(), o[1] = n.y(), o[2] = n.z(), o[3] = cellSeg.get_plane_d(), o[4] = cellSeg.get_MSE();
            }
        *n_planes = int32_t(planes.size());
        int nb = 0;
        for (size_t k = 0; k < planes.size() && int(k) < max_planes; ++k) {
            double* o = planes_out + 6 * k;
            o[0] = planes[k].normal.x(), o[1] = planes[k].normal.y(), o[2] = planes[k].normal.z(), o[3] = planes[k].d;
            o[4] = planes[k].mse, o[5] = double(planes[k].polygon.points.size());
            for (const vector3& pt : planes[k].polygon.points)
                if (boundary_out && nb < max_boundary) {
                    boundary_out[3 * nb] = pt.x(), boundary_out[3 * nb + 1] = pt.y(), boundary_out[3 * nb + 2] = pt.z();
                    ++nb;
                }
        }
        *n_cyls = int32_t(cylinders.size());
        for (size_t k = 0; k < cylinders.size() && int(k) < max_cyls; ++k) {
            double* o = cyls_out + 4 * k;
            o[0] = cylinders[k]._normal.x(), o[1] = cylinders[k]._normal.y(), o[2] = cylinders[k]._normal.z(), o[3] = cylinders[k]._radius;
        }
    });
    worker.join();
    return rc;
}

// Depth_Map_Transformation::rectify_depth (depth_map_transformation.cpp:23-87) with the given camera-2 -> camera-1 transformation
// (row-major 4x4) and the default intrinsics on both cameras.
int ref_rectify_depth(const float* depth, int width, int height, const double* cam2_to_cam1, float* out)
{
    Parameters::load_defaut();
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) Parameters::_camera2toCamera1transformation(r, c) = cam2_to_cam1[4 * r + c];
    features::primitives::Depth_Map_Transformation depthOps(width, height, parameters::detection::depthMapPatchSize_px);
    cv::Mat_<float> image(height, width), rectified;
    std::memcpy(image.data, depth, sizeof(float) * size_t(width) * height);
    const bool ok = depthOps.rectify_depth(image, rectified);
    if (ok) std::memcpy(out, rectified.data, sizeof(float) * size_t(width) * height);
    Parameters::load_defaut();
    return ok ? 0 : 1;
}

int ref_cape_cell_px(void) { return int(parameters::detection::depthMapPatchSize_px); }

}  // extern "C"
