// TEST INFRASTRUCTURE — restates the synthetic scenario builders of the reference's own pose tests
// (tests/test_pose_optimization.cpp:23-213) so that the 40 reference scenarios can be replayed against the oracle
// and the CUDA library with the reference's tolerances (+-6 mm, +-0.1 deg; :230-242).
#include <cmath>
#include <cstdlib>
#include <random>
#include <vector>

#include "pose.hpp"

using namespace oracle;

namespace {
const unsigned NUMBER_OF_POINTS_IN_CUBE = unsigned(std::pow(4, 3));
const double CUBE_SIDE_SIZE = 20;
const double CUBE_START_X = 100, CUBE_START_Y = 100, CUBE_START_Z = 100;

struct Point_ {
    double x, y, z;
};

// :51-83
std::vector<Point_> get_cube_points(const unsigned numberOfPoints, const double error)
{
    std::mt19937 randomEngine(1000);
    std::uniform_real_distribution<double> errorDistribution(-error, error);
    const unsigned numberOfPointsByLine = static_cast<unsigned>(std::pow(static_cast<double>(numberOfPoints), 1.0 / 3.0));
    const double step = CUBE_SIDE_SIZE / static_cast<double>(numberOfPointsByLine - 1);
    std::vector<Point_> pts;
    for (unsigned a = 0; a <= numberOfPointsByLine; ++a)
        for (unsigned b = 0; b <= numberOfPointsByLine; ++b)
            for (unsigned c = 0; c <= numberOfPointsByLine; ++c) {
                Point_ p;
                p.x = CUBE_START_X + a * step + errorDistribution(randomEngine);
                p.y = CUBE_START_Y + b * step + errorDistribution(randomEngine);
                p.z = CUBE_START_Z + c * step + errorDistribution(randomEngine);
                pts.push_back(p);
            }
    return pts;
}

// Eigen's DenseBase::Random() coefficient for double: -1 + 2*rand()/RAND_MAX (Eigen/src/Core/MathFunctions.h)
double eigen_random() { return -1.0 + (1.0 - (-1.0)) * double(std::rand()) / double(RAND_MAX); }

bool project(const Intrinsics& K, const Mat4& w2c, const Point_& p, double uv[2])
{
    double h[4];
    for (int i = 0; i < 4; ++i) h[i] = ((w2c(i, 0) * p.x + w2c(i, 1) * p.y) + w2c(i, 2) * p.z) + w2c(i, 3);
    const double xc = h[0] / h[3], yc = h[1] / h[3], zc = h[2] / h[3];
    const double inv = 1.0 / zc;
    uv[0] = inv * ((K.fx * xc + 0.0 * yc) + K.cx * zc);
    uv[1] = inv * ((0.0 * xc + K.fy * yc) + K.cy * zc);
    return uv[0] == uv[0] and uv[1] == uv[1];
}
}  // namespace

extern "C" {

// Builds the match list of a reference scenario: planes first (get_matched_planes, :152-206) then points
// (get_matched_points, :85-150). outlier proportions < 0 disable that feature kind. Returns the match count.
int orc_ref_test_features(const double true_pose[7], double point_error, double point_outlier_prop, double plane_error,
                          double plane_outlier_prop, rs_match* out, int max_out)
{
    Intrinsics K;  // Parameters::load_defaut (parameters.cpp:59-74)
    const Mat4 w2c = world_to_camera(true_pose + 3, true_pose);
    std::vector<rs_match> all;

    if (plane_outlier_prop >= 0) {
        std::mt19937 randomEngine(1000);
        std::uniform_real_distribution<double> errorDistribution(-plane_error, plane_error);
        const Mat4 pm = plane_world_to_camera(w2c);
        const double planes0[4][4] = {{0.452271, -0.419436, -0.787099, 10},
                                      {-0.585607, -0.43009, 0.687085, 30},
                                      {-0.498271, 0.767552, -0.403223, -20},
                                      {0.706067, -0.0741267, -0.704255, 150}};
        double planes[4][4];
        for (int k = 0; k < 4; ++k) {
            const Vec3 n = normalized(Vec3{planes0[k][0], planes0[k][1], planes0[k][2]});
            planes[k][0] = n.x, planes[k][1] = n.y, planes[k][2] = n.z, planes[k][3] = planes0[k][3];
        }
        const double sd[4] = {std::sqrt(0.01 * 0.01), std::sqrt(0.01 * 0.01), std::sqrt(0.01 * 0.01), std::sqrt(1.0)};
        for (int k = 0; k < 4; ++k) {
            planes[k][3] += errorDistribution(randomEngine);
            double h[4];
            for (int i = 0; i < 4; ++i)
                h[i] = ((pm(i, 0) * planes[k][0] + pm(i, 1) * planes[k][1]) + pm(i, 2) * planes[k][2]) + pm(i, 3) * planes[k][3];
            const Vec3 nc = normalized(Vec3{h[0], h[1], h[2]});
            rs_match m{};
            m.type = RS_FEAT_PLANE;
            m.obs[0] = nc.x, m.obs[1] = nc.y, m.obs[2] = nc.z, m.obs[3] = h[3];
            for (int i = 0; i < 4; ++i) {
                m.map[i] = planes[k][i];
                m.sigma[i] = sd[i];
            }
            all.push_back(m);
        }
        std::uniform_real_distribution<float> indexDistribution(0, 1.0);
        std::uniform_real_distribution<double> DErrorDistribution(-100, 100);
        const size_t outlierToAdd = size_t(4 * plane_outlier_prop);
        for (size_t i = 0; i < outlierToAdd; ++i) {
            const size_t chosen = size_t(std::floor(4 * indexDistribution(randomEngine)));
            const double rx = eigen_random(), ry = eigen_random(), rz = eigen_random();
            const Vec3 n = normalized(normalized(Vec3{rx, ry, rz}));
            rs_match m{};
            m.type = RS_FEAT_PLANE;
            m.obs[0] = n.x, m.obs[1] = n.y, m.obs[2] = n.z, m.obs[3] = DErrorDistribution(randomEngine);
            for (int k = 0; k < 4; ++k) {
                m.map[k] = planes[chosen % 4][k];
                m.sigma[k] = sd[k];
            }
            all.push_back(m);
        }
    }

    if (point_outlier_prop >= 0) {
        std::mt19937 randomEngine(1000);
        const std::vector<Point_> cube = get_cube_points(NUMBER_OF_POINTS_IN_CUBE, point_error);
        for (const Point_& p : cube) {
            double uv[2];
            if (project(K, w2c, p, uv)) {
                rs_match m{};
                m.type = RS_FEAT_POINT;
                m.obs[0] = uv[0], m.obs[1] = uv[1];
                m.map[0] = p.x, m.map[1] = p.y, m.map[2] = p.z;
                m.sigma[0] = m.sigma[1] = m.sigma[2] = 1.0;
                all.push_back(m);
            }
        }
        const size_t dataSize = cube.size();
        const size_t outliersToAdd = size_t(dataSize * point_outlier_prop);
        std::uniform_real_distribution<float> indexDistribution(0, 1.0);
        for (size_t i = 0; i < outliersToAdd; ++i) {
            const size_t chosen = size_t(std::floor(dataSize * indexDistribution(randomEngine)));
            const Point_& feat = cube[chosen % dataSize];
            rs_match m{};
            m.type = RS_FEAT_POINT;
            m.obs[0] = eigen_random(), m.obs[1] = eigen_random();  // ScreenCoordinate2D = vector2::Random()
            m.map[0] = feat.x, m.map[1] = feat.y, m.map[2] = feat.z;
            m.sigma[0] = m.sigma[1] = m.sigma[2] = 1.0;
            all.push_back(m);
        }
    }
    const int n = int(all.size());
    for (int i = 0; i < n && i < max_out; ++i) out[i] = all[i];
    return n;
}

}  // extern "C"
