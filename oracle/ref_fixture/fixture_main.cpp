// TEST INFRASTRUCTURE - fixture generator built from the REFERENCE's own sources (see README.md; never compiled in the round's
// image). For every raw float32 depth file (640x480, row-major, mm) given on the command line it runs the reference's
// get_organized_cloud_array + find_primitives and writes <file>.labels:
//   int32 vc, hc | int32 plane grid [vc*hc] | int32 cylinder grid [vc*hc] | int32 n_planes | per plane: double normal[3], d |
//   int32 n_cylinders | per cylinder: double normal[3], radius
#define private public      // the label grids are private members of Primitive_Detection
#define protected public
#include "features/primitives/depth_map_transformation.hpp"
#include "features/primitives/primitive_detection.hpp"
#undef private
#undef protected
#include "parameters.hpp"

#include <cstdio>
#include <fstream>
#include <thread>
#include <vector>

using namespace rgbd_slam;

int main(int argc, char** argv)
{
    Parameters::load_defaut();
    constexpr unsigned W = 640, H = 480;
    const unsigned cell = parameters::detection::depthMapPatchSize_px;
    features::primitives::Depth_Map_Transformation depthOps(W, H, cell);
    for (int a = 1; a < argc; ++a) {
        cv::Mat_<float> depth(H, W);
        std::ifstream in(argv[a], std::ios::binary);
        in.read(reinterpret_cast<char*>(depth.data), sizeof(float) * W * H);
        if (!in) {
            std::fprintf(stderr, "cannot read %s\n", argv[a]);
            return 1;
        }
        // a fresh thread per frame: the thread-local engine of utils/random.hpp restarts from its seed (0 under MAKE_DETERMINISTIC),
        // as it does in RGBD_SLAM::track, where find_primitives runs through std::async
        std::thread worker([&]() {
            features::primitives::Primitive_Detection detector(W, H);
            matrixf cloud;
            if (!depthOps.get_organized_cloud_array(depth, cloud)) return;
            features::primitives::plane_container planes;
            features::primitives::cylinder_container cylinders;
            detector.find_primitives(cloud, depth, planes, cylinders);
            std::ofstream out(std::string(argv[a]) + ".labels", std::ios::binary);
            const int32_t vc = detector._gridPlaneSegmentMap.rows, hc = detector._gridPlaneSegmentMap.cols;
            out.write(reinterpret_cast<const char*>(&vc), 4);
            out.write(reinterpret_cast<const char*>(&hc), 4);
            for (int r = 0; r < vc; ++r)
                out.write(reinterpret_cast<const char*>(detector._gridPlaneSegmentMap.ptr<int>(r)), sizeof(int32_t) * hc);
            for (int r = 0; r < vc; ++r)
                out.write(reinterpret_cast<const char*>(detector._gridCylinderSegMap.ptr<int>(r)), sizeof(int32_t) * hc);
            const int32_t np = static_cast<int32_t>(planes.size());
            out.write(reinterpret_cast<const char*>(&np), 4);
            for (const auto& p: planes) {
                const double v[4] = {p.get_normal().x(), p.get_normal().y(), p.get_normal().z(), p.get_d()};
                out.write(reinterpret_cast<const char*>(v), sizeof(v));
            }
            const int32_t nc = static_cast<int32_t>(cylinders.size());
            out.write(reinterpret_cast<const char*>(&nc), 4);
            for (const auto& c: cylinders) {
                const double v[4] = {c._normal.x(), c._normal.y(), c._normal.z(), c._radius};   // mean radius over the segments
                out.write(reinterpret_cast<const char*>(v), sizeof(v));
            }
        });
        worker.join();
    }
    return 0;
}
