// TEST INFRASTRUCTURE — CPU oracle (see cape.hpp). Pinned bit for bit by the reference's own sources compiled against stand-in
// third-party headers (oracle/ref_shim, tests/test_reference_build.py); no reference test / golden vector exists for CAPE.
// Every function cites the reference lines it restates; quirks are kept on purpose.
#include "cape.hpp"

#include <cassert>
#include <random>

namespace oracle {

namespace {
// parameters.hpp:16-18,68-86
constexpr double kDepthSigmaError = 2.73, kDepthSigmaMultiplier = 0.74, kDepthSigmaMargin = -0.53;
constexpr double kMinimumPlaneSeedProportion = 0.8 / 100.0;
constexpr double kMinimumCellActivatedProportion = 0.65 / 100.0;
constexpr float kMinimumZeroDepthProportion = 0.7f;
constexpr float kMaximumPlaneAngleForMerge_d = 18.0f;
constexpr float kMaximumPlaneDistanceForMerge_mm = 50.0f;
constexpr float kCylinderRansacSqrtMaxDistance = 0.04f;
constexpr float kCylinderRansacMinimumScore = 75;
constexpr float kCylinderRansacInlierProportions = 0.33f;
constexpr float kCylinderRansacProbabilityOfSuccess = 0.8f;
}  // namespace

// covariances.cpp:12-19
double depth_quantization(const double depth)
{
    const double depthSigmaError = kDepthSigmaError * ((1.0 / 1000.0) * (1.0 / 1000.0));
    const double depthSigmaMultiplier = kDepthSigmaMultiplier / 1000.0;
    const double depthSigmaMargin = kDepthSigmaMargin;
    return std::max(depthSigmaMargin + depthSigmaMultiplier * depth + depthSigmaError * (depth * depth), 0.5);
}

// plane_segment.cpp:289-310
void PlaneSeg::clear()
{
    planar = false;
    count = 0;
    score = 0;
    mse = DBL_MAX;
    centroid = Vec3{};
    normal = Vec3{};
    d = 0;
    for (double& s : S) s = 0;
}

// plane_segment.cpp:170-190
void PlaneSeg::expand(const PlaneSeg& o)
{
    S[0] += o.S[0];
    S[1] += o.S[1];
    S[2] += o.S[2];
    S[3] += o.S[3];
    S[4] += o.S[4];
    S[5] += o.S[5];
    S[6] += o.S[6];
    S[7] += o.S[7];
    S[8] += o.S[8];
    count += o.count;
}

// plane_segment.cpp:18-36 — the member-wise copy goes through PlaneCoordinates' copy ctor, which
// re-normalises the normal (plane_coordinates.hpp:25-28).
PlaneSeg PlaneSeg::copy() const
{
    PlaneSeg r = *this;
    r.normal = normalized(normal);
    return r;
}

// plane_segment.cpp:205-284
void PlaneSeg::fit_plane()
{
    planar = false;
    const double oneOverCount = 1.0 / static_cast<double>(count);
    const double Sx = S[0], Sy = S[1], Sz = S[2], Sxs = S[3], Sys = S[4], Szs = S[5], Sxy = S[6], Syz = S[7],
                 Szx = S[8];
    centroid = Vec3{Sx * oneOverCount, Sy * oneOverCount, Sz * oneOverCount};

    // get_point_cloud_Huygen_covariance (:205-230)
    const double xx = std::max(0.0, Sxs - (Sx * Sx) * oneOverCount);
    const double yy = std::max(0.0, Sys - (Sy * Sy) * oneOverCount);
    const double zz = std::max(0.0, Szs - (Sz * Sz) * oneOverCount);
    const double xy = Sxy - Sx * Sy * oneOverCount;
    const double xz = Szx - Sx * Sz * oneOverCount;
    const double yz = Syz - Sy * Sz * oneOverCount;
    Mat3 cov;
    cov(0, 0) = xx, cov(0, 1) = xy, cov(0, 2) = xz;
    cov(1, 0) = xy, cov(1, 1) = yy, cov(1, 2) = yz;
    cov(2, 0) = xz, cov(2, 1) = yz, cov(2, 2) = zz;

    // degenerate covariance (:245): utils::double_equal(det, 0) with epsilon = DBL_EPSILON
    if (std::fabs(det3(cov) - 0.0) <= DBL_EPSILON) return;

    double ev[3];
    Mat3 evec;
    self_adjoint_eigen3(cov, ev, evec);
    const double l0 = std::fabs(ev[0]), l1 = std::fabs(ev[1]);
    const Vec3 eigenVector{evec(0, 0), evec(1, 0), evec(2, 0)};

    const Vec3 n = normalized(eigenVector);
    const double dd = -dot(n, centroid);
    // PlaneCoordinates(normal, d) normalises in its ctor, operator= normalises again (plane_coordinates.hpp:23,30-40)
    Vec3 pn;
    double pd;
    if (dd <= 0) {
        pn = -n;
        pd = -dd;
    }
    else {
        pn = n;
        pd = dd;
    }
    pn = normalized(pn);
    pn = normalized(pn);
    normal = pn;
    d = pd;

    mse = l0 * oneOverCount;
    score = l1 / std::max(l0, 1e-6);
    planar = true;
}

// plane_segment.cpp:322-326 ; abs() resolves to the FP overload (SURVEY.md App. A preamble).
bool PlaneSeg::can_be_merged(const PlaneSeg& p, const double maxMatchDistance) const
{
    static const double maximumMergeAngle = std::cos(kMaximumPlaneAngleForMerge_d * M_PI / 180.0);
    return dot(normal, p.normal) > maximumMergeAngle and std::fabs(dot(normal, p.centroid) + d) < maxMatchDistance;
}

void cell_record(const PlaneSeg& s, const float tol, const int cellSize, rs_cell_out& o)
{
    o.count = s.count;
    o.planar = s.planar ? 1 : 0;
    for (int i = 0; i < 9; ++i) o.S[i] = s.S[i];
    for (int i = 0; i < 3; ++i) {
        o.centroid[i] = s.centroid[i];
        o.normal[i] = s.normal[i];
    }
    o.d = s.d;
    o.mse = s.mse;
    o.score = s.score;
    o.tol = tol;
    // bin of init_histogram (primitive_detection.cpp:239-265, histogram.hpp:35-62), Histogram<cellSize>
    o.hist_bin = -1;
    if (s.planar) {
        const double theta = std::acos(-s.normal.z), phi = std::atan2(s.normal.x, s.normal.y);
        const int xQ = static_cast<int>(std::floor((cellSize - 1) * (theta - 0.0) / (M_PI - 0.0)));
        int yQ = 0;
        if (xQ > 0) yQ = static_cast<int>(std::floor((cellSize - 1) * (phi - (-M_PI)) / (M_PI - (-M_PI))));
        o.hist_bin = yQ * cellSize + xQ;
    }
}

// point_coordinates.cpp:79-83 (static inverse of the intrinsics) applied to (u, v, 1).
void backprojection_factors(const CapeConfig& cfg, std::vector<double>& kx, std::vector<double>& ky)
{
    Mat3 K;
    K(0, 0) = cfg.fx, K(0, 1) = 0, K(0, 2) = cfg.cx;
    K(1, 0) = 0, K(1, 1) = cfg.fy, K(1, 2) = cfg.cy;
    K(2, 0) = 0, K(2, 1) = 0, K(2, 2) = 1;
    const Mat3 Ki = inverse3(K);
    kx.resize(cfg.width);
    ky.resize(cfg.height);
    // Row 0 has a zero (possibly -0) v coefficient and row 1 a zero u coefficient, so the factors only depend on
    // the column / the row respectively; evaluated with Eigen's (a*u + b*v) + c association.
    for (int c = 0; c < cfg.width; ++c) kx[c] = (Ki(0, 0) * double(c) + Ki(0, 1) * 0.0) + Ki(0, 2) * 1.0;
    for (int r = 0; r < cfg.height; ++r) ky[r] = (Ki(1, 0) * 0.0 + Ki(1, 1) * double(r)) + Ki(1, 2) * 1.0;
}

namespace {

// plane_segment.cpp:44-60
bool is_continuous(const float pixelDepth, float& lastPixelDepth)
{
    if (pixelDepth > 0) {
        if (std::fabs(pixelDepth - lastPixelDepth) <= 4.0 * depth_quantization(pixelDepth)) {
            lastPixelDepth = pixelDepth;
            return true;
        }
        return false;
    }
    return true;
}

// plane_segment.cpp:62-80
bool is_cell_vertical_continuous(const float* z, const unsigned cellWidth, const unsigned ptsPerCell)
{
    const unsigned startValue = cellWidth / 2;
    const unsigned endValue = ptsPerCell - startValue;
    float last = std::max(z[startValue], z[startValue + cellWidth]);
    if (last <= 0) return false;
    for (unsigned i = startValue + cellWidth; i < endValue; i += cellWidth)
        if (not is_continuous(z[i], last)) return false;
    return true;
}

// plane_segment.cpp:82-100
bool is_cell_horizontal_continuous(const float* z, const unsigned cellWidth, const unsigned cellHeight)
{
    const unsigned startValue = static_cast<unsigned>(cellWidth * (cellHeight / 2.0));
    const unsigned endValue = startValue + cellWidth;
    float last = std::max(z[startValue], z[startValue + 1]);
    if (last <= 0) return false;
    for (unsigned i = startValue + 1; i < endValue; ++i)
        if (not is_continuous(z[i], last)) return false;
    return true;
}

// plane_segment.cpp:102-168
void init_plane_segment(PlaneSeg& seg, const float* x, const float* y, const float* z, const unsigned cellWidth,
                        const unsigned ptsPerCell, const unsigned minZeroPointCount)
{
    seg.clear();
    const unsigned cellHeight = ptsPerCell / cellWidth;
    if (not is_cell_horizontal_continuous(z, cellWidth, cellHeight) or
        not is_cell_vertical_continuous(z, cellWidth, ptsPerCell))
        return;
    unsigned positive = 0;
    for (unsigned i = 0; i < ptsPerCell; ++i) positive += (z[i] > 0) ? 1u : 0u;
    if (positive < ptsPerCell / 2) return;

    seg.count = 0;
    for (unsigned i = 0; i < ptsPerCell; ++i) {
        const float zi = z[i];
        if (zi > 0) {
            ++seg.count;
            const float xi = x[i];
            const float yi = y[i];
            // float products (SQR<float>, x*y), FP64 accumulation
            seg.S[0] += xi;
            seg.S[1] += yi;
            seg.S[2] += zi;
            seg.S[3] += xi * xi;
            seg.S[4] += yi * yi;
            seg.S[5] += zi * zi;
            seg.S[6] += xi * yi;
            seg.S[8] += xi * zi;
            seg.S[7] += yi * zi;
        }
    }
    if (static_cast<unsigned>(seg.count) < minZeroPointCount) return;
    seg.fit_plane();
    const double q = depth_quantization(seg.centroid.z);
    seg.planar = seg.mse <= q * q;
}

}  // namespace

// depth_map_transformation.cpp:89-173 + primitive_detection.cpp:187-237
void cape_cell_fit(const CapeConfig& cfg, const float* depth, std::vector<PlaneSeg>& grid, std::vector<float>& tols,
                   std::vector<float>* cloudOut)
{
    const int W = cfg.width, H = cfg.height, cs = cfg.cell;
    const int hc = W / cs, vc = H / cs, Nc = hc * vc;
    const unsigned P = unsigned(cs) * unsigned(cs);
    std::vector<double> kx, ky;
    backprojection_factors(cfg, kx, ky);

    // organized cloud, column-major W*H x 3, zero-filled (:96)
    const size_t n = size_t(W) * H;
    std::vector<float> cloud(3 * n, 0.0f);
    float* X = cloud.data();
    float* Y = X + n;
    float* Z = Y + n;
    for (int r = 0; r < H; ++r) {
        const unsigned cellR = unsigned(r) / cs, localR = unsigned(r) % cs;
        for (int c = 0; c < W; ++c) {
            const float z = depth[size_t(r) * W + c];
            if (z > 0) {
                const unsigned cellC = unsigned(c) / cs, localC = unsigned(c) % cs;
                if (cellR >= unsigned(vc) or cellC >= unsigned(hc)) continue;  // only when W,H are not multiples of cs
                const size_t id = size_t(cellR * hc + cellC) * P + localR * cs + localC;
                const double zd = z;
                X[id] = static_cast<float>(zd * kx[c]);
                Y[id] = static_cast<float>(zd * ky[r]);
                Z[id] = z;
            }
        }
    }

    // set_static_members (plane_segment.hpp:27-36)
    const unsigned minZeroPointCount =
            static_cast<unsigned>(std::floor(static_cast<float>(P) * kMinimumZeroDepthProportion));
    const float sinAngleForMerge = sinf(static_cast<float>(kMaximumPlaneAngleForMerge_d * M_PI / 180.0));
    const float planeMergeDistanceThreshold = kMaximumPlaneDistanceForMerge_mm;

    grid.assign(Nc, PlaneSeg{});
    tols.assign(Nc, 0.0f);
    for (int cell = 0; cell < Nc; ++cell) {
        const size_t off = size_t(cell) * P;
        PlaneSeg& seg = grid[cell];
        init_plane_segment(seg, X + off, Y + off, Z + off, cs, P, minZeroPointCount);
        if (seg.planar) {
            // cell diameter: float 3-vector norm of (last row) - (first row) of the cell block (:205-210)
            const float dx = X[off + P - 1] - X[off], dy = Y[off + P - 1] - Y[off], dz = Z[off + P - 1] - Z[off];
            const float cellDiameter = sqrtf((dx * dx + dy * dy) + dz * dz);
            tols[cell] = std::min(planeMergeDistanceThreshold,
                                  cellDiameter * sinAngleForMerge * sqrtf(static_cast<float>(seg.count)));
        }
        else {
            tols[cell] = 0;
        }
    }
    if (cloudOut) cloudOut->swap(cloud);
}

namespace {

// histogram.hpp — Histogram<Size> with Size = depthMapPatchSize_px (primitive_detection.hpp:199)
struct Histogram {
    int size = 0;
    std::vector<unsigned> hist;
    std::vector<int> bins;
    void reset(int sz)
    {
        size = sz;
        hist.assign(size_t(sz) * sz, 0u);
        bins.clear();
    }
    // :35-62
    void init(const std::vector<double>& theta, const std::vector<double>& phi, const std::vector<char>& mask)
    {
        const size_t n = mask.size();
        bins.assign(n, -1);
        const double minX = 0, minY = -M_PI, maxXminX = M_PI - minX, maxYminY = M_PI - minY;
        for (size_t i = 0; i < n; ++i) {
            if (mask[i]) {
                const int xQ = static_cast<int>(std::floor((size - 1) * (theta[i] - minX) / maxXminX));
                int yQ = 0;
                if (xQ > 0) yQ = static_cast<int>(std::floor((size - 1) * (phi[i] - minY) / maxYminY));
                const unsigned bin = unsigned(yQ * size + xQ);
                bins[i] = static_cast<int>(bin);
                if (bin < hist.size()) hist[bin] += 1;
            }
        }
    }
    // :69-98
    std::vector<unsigned> most_frequent() const
    {
        int mostFrequentBin = -1;
        unsigned maxOcc = 0;
        for (unsigned i = 0; i < hist.size(); ++i)
            if (hist[i] > maxOcc) {
                mostFrequentBin = int(i);
                maxOcc = hist[i];
            }
        std::vector<unsigned> ids;
        if (mostFrequentBin >= 0)
            for (unsigned i = 0; i < bins.size(); ++i)
                if (bins[i] == mostFrequentBin) ids.push_back(i);
        return ids;
    }
    // :103-113 — quirk: sets the bin to 1, not -1
    void remove_point(unsigned id)
    {
        if (bins[id] >= 0 and unsigned(bins[id]) < hist.size() and hist[bins[id]] != 0) hist[bins[id]] -= 1;
        bins[id] = 1;
    }
};

struct CylSeg {
    double radius;
    Vec3 center;
    double mse;
    std::vector<char> inliers;  // over local ids
};
struct CylinderSegment {
    int cellActivatedCount = 0;
    std::vector<unsigned> local2global;
    double pcaScore = 0;
    Vec3 axis;
    std::vector<CylSeg> segs;
};

struct Detector {
    const CapeConfig& cfg;
    int hc, vc, Nc;
    std::vector<PlaneSeg> grid;
    std::vector<float> tols;
    std::vector<char> unassigned;
    Histogram histogram;
    std::vector<PlaneSeg> planeSegments;
    std::vector<CylinderSegment> cylinderSegments;
    std::vector<std::pair<int, int>> cylinder2regionMap;
    std::vector<int32_t> gridPlane, gridCyl, gridCylRegionSeg;
    std::mt19937 rng;
    std::uniform_real_distribution<double> uni{0.0, 1.0};
    int nSeeds = 0;
    // per region/segment bookkeeping for the output record
    std::vector<std::vector<int>> cylAssigned;
    std::vector<std::vector<double>> cylPlaneMse;

    Detector(const CapeConfig& c, uint32_t seed) : cfg(c), rng(seed)
    {
        hc = c.width / c.cell;
        vc = c.height / c.cell;
        Nc = hc * vc;
    }

    // random.hpp:51-58
    unsigned random_uint(unsigned maxValue) { return 0 + static_cast<unsigned>(std::floor(uni(rng) * (maxValue - 0))); }

    // primitive_detection.cpp:778-818 (recursive 4-neighbour growth)
    void region_growing(unsigned x, unsigned y, const PlaneSeg& planeToExpand, std::vector<char>& activated)
    {
        const int index = int(x + unsigned(hc) * y);
        if (size_t(index) >= size_t(Nc)) return;
        if ((not unassigned[index]) or activated[index]) return;
        const PlaneSeg& patch = grid[index];
        if (planeToExpand.can_be_merged(patch, tols[index])) {
            activated[index] = 1;
            if (x > 0) region_growing(x - 1, y, patch, activated);
            if (x < unsigned(hc) - 1) region_growing(x + 1, y, patch, activated);
            if (y > 0) region_growing(x, y - 1, patch, activated);
            if (y < unsigned(vc) - 1) region_growing(x, y + 1, patch, activated);
        }
    }

    // cylinder_segment.cpp:227-322
    size_t run_ransac_loop(const CylinderSegment& cyl, unsigned maximumIterations, const std::vector<unsigned>& idsLeft,
                           const std::vector<Vec3>& planeNormals, const std::vector<Vec3>& projectedCentroids,
                           const std::vector<char>& idsLeftMask, std::vector<char>& isInlierFinal)
    {
        if (idsLeft.size() < 3) return 0;
        const unsigned planeIdsLeft = unsigned(idsLeft.size());
        const unsigned inliersAcceptedCount = unsigned(std::floor(0.9 * planeIdsLeft));
        const float maximumSqrtDistance = kCylinderRansacSqrtMaxDistance;
        double minHypothesisDist = maximumSqrtDistance * static_cast<float>(planeIdsLeft);
        std::vector<unsigned> finalInlierIndexes;

        for (unsigned iteration = 0; iteration < maximumIterations; ++iteration) {
            const unsigned id1 = idsLeft[random_uint(planeIdsLeft)];
            const unsigned id2 = idsLeft[random_uint(planeIdsLeft)];
            const unsigned id3 = idsLeft[random_uint(planeIdsLeft)];
            const Vec3& n1 = planeNormals[id1];
            const Vec3& n2 = planeNormals[id2];
            const Vec3& n3 = planeNormals[id3];
            const Vec3& c1 = projectedCentroids[id1];
            const Vec3& c2 = projectedCentroids[id2];
            const Vec3& c3 = projectedCentroids[id3];
            const Vec3 sumOfNormals = (n1 + n2) + n3;
            const Vec3 sumOfCenters = (c1 + c2) + c3;
            const double a = 1.0 - sqnorm(sumOfNormals) / 9.0;
            const double t0 = (n1.x * c1.x + n2.x * c2.x) + n3.x * c3.x;
            const double t1 = (n1.y * c1.y + n2.y * c2.y) + n3.y * c3.y;
            const double t2 = (n1.z * c1.z + n2.z * c2.z) + n3.z * c3.z;
            const double b = ((t0 + t1) + t2) / 3.0 - (dot(sumOfNormals, sumOfCenters) / 9.0);
            const double radius = b / a;
            const double oneOverRadiusSquared = 1.0 / (radius * radius);
            const Vec3 center = (sumOfCenters - radius * sumOfNormals) / 3.0;

            std::vector<unsigned> inlierIndexes;
            double dist = 0.0;
            for (unsigned i = 0; i < unsigned(cyl.cellActivatedCount); ++i) {
                if (not idsLeftMask[i]) continue;
                const double distance =
                        sqnorm((projectedCentroids[i] - radius * planeNormals[i]) - center) * oneOverRadiusSquared;
                if (distance < maximumSqrtDistance) {
                    dist += distance;
                    inlierIndexes.push_back(i);
                }
                else {
                    dist += maximumSqrtDistance;
                }
            }
            if (dist < minHypothesisDist) {
                minHypothesisDist = dist;
                finalInlierIndexes.swap(inlierIndexes);
                // quirk: tests the PREVIOUS best set (the vectors were just swapped), :310-312
                if (inlierIndexes.size() > inliersAcceptedCount) break;
            }
        }
        std::fill(isInlierFinal.begin(), isInlierFinal.end(), 0);
        for (unsigned i : finalInlierIndexes) isInlierFinal[i] = 1;
        return finalInlierIndexes.size();
    }

    // cylinder_segment.cpp:35-225
    CylinderSegment make_cylinder(const std::vector<char>& activated, unsigned cellActivatedCount)
    {
        CylinderSegment cyl;
        cyl.cellActivatedCount = int(cellActivatedCount);
        const size_t samplesCount = activated.size();
        const unsigned m = cellActivatedCount;
        std::vector<Vec3> planeNormals, planeCentroids;
        for (size_t i = 0; i < samplesCount; ++i)
            if (activated[i]) {
                planeNormals.push_back(grid[i].normal);
                planeCentroids.push_back(grid[i].centroid);
                cyl.local2global.push_back(unsigned(i));
            }
        // cov = [N -N][N -N]^T / (2m - 1)
        Mat3 cov;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                double s = 0;
                for (unsigned j = 0; j < m; ++j) s += planeNormals[j][r] * planeNormals[j][c];
                for (unsigned j = 0; j < m; ++j) s += (-planeNormals[j][r]) * (-planeNormals[j][c]);
                cov(r, c) = s / static_cast<double>(2 * m - 1);
            }
        double ev[3];
        Mat3 evec;
        self_adjoint_eigen3(cov, ev, evec);
        const double score = ev[2] / ev[0];
        cyl.pcaScore = score;
        if (score < kCylinderRansacMinimumScore) return cyl;

        const Vec3 axis{evec(0, 0), evec(1, 0), evec(2, 0)};
        cyl.axis = axis;

        std::vector<Vec3> projectedCentroids(m);
        for (unsigned j = 0; j < m; ++j) {
            const double cd = dot(axis, planeCentroids[j]);
            projectedCentroids[j] = Vec3{planeCentroids[j].x - cd * axis.x, planeCentroids[j].y - cd * axis.y,
                                         planeCentroids[j].z - cd * axis.z};
            const double nd = dot(axis, planeNormals[j]);
            Vec3 pn{planeNormals[j].x - nd * axis.x, planeNormals[j].y - nd * axis.y, planeNormals[j].z - nd * axis.z};
            const double nn = norm(pn);
            planeNormals[j] = Vec3{pn.x / nn, pn.y / nn, pn.z / nn};
        }

        const float pSuccess = kCylinderRansacProbabilityOfSuccess;
        const float w = kCylinderRansacInlierProportions;
        const unsigned maximumIterations = static_cast<unsigned>(logf(1.0f - pSuccess) / logf(1.0f - powf(w, 3.0f)));

        unsigned planeSegmentsLeft = m;
        std::vector<char> idsLeftMask(m, 1);
        std::vector<unsigned> idsLeft;
        for (unsigned i = 0; i < m; ++i) idsLeft.push_back(i);
        const size_t minimumCellActivated =
                static_cast<unsigned>(kMinimumCellActivatedProportion * static_cast<double>(samplesCount));
        while (planeSegmentsLeft > minimumCellActivated and planeSegmentsLeft > 0.1 * m) {
            std::vector<char> isInlierFinal(m, 1);
            const size_t maxInliersCount =
                    run_ransac_loop(cyl, maximumIterations, idsLeft, planeNormals, projectedCentroids, idsLeftMask,
                                    isInlierFinal);
            if (maxInliersCount < 6) break;

            double b = 0;
            Vec3 sumOfNormals, sumOfCenters;
            idsLeft.clear();
            for (unsigned i = 0; i < m; ++i) {
                if (isInlierFinal[i]) {
                    idsLeftMask[i] = 0;
                    planeSegmentsLeft--;
                    sumOfNormals = sumOfNormals + planeNormals[i];
                    sumOfCenters = sumOfCenters + projectedCentroids[i];
                    b += (planeNormals[i].x * projectedCentroids[i].x + planeNormals[i].y * projectedCentroids[i].y) +
                         planeNormals[i].z * projectedCentroids[i].z;
                }
                else if (idsLeftMask[i]) {
                    idsLeft.push_back(i);
                }
            }
            const double oneOverMaxInliersCountSquared = 1.0 / static_cast<double>(maxInliersCount * maxInliersCount);
            const double a = 1 - sqnorm(sumOfNormals) * oneOverMaxInliersCountSquared;
            b /= static_cast<double>(maxInliersCount);
            b -= dot(sumOfNormals, sumOfCenters) * oneOverMaxInliersCountSquared;
            double radius = b / a;
            const Vec3 center = (sumOfCenters - radius * sumOfNormals) / static_cast<double>(maxInliersCount);
            if (radius < 0) radius = -radius;

            const Vec3 P1d = center;
            const Vec3 P2d = center + axis;
            const double P1P2d = norm(P2d - P1d);
            double mse = 0;
            for (unsigned i = 0; i < m; ++i) {
                if (isInlierFinal[i]) {
                    const Vec3 P3 = planeCentroids[i];
                    const double t = norm(cross(P2d - P1d, P3 - P2d)) / P1P2d - radius;
                    mse += t * t;
                }
            }
            mse /= static_cast<double>(maxInliersCount);
            cyl.segs.push_back(CylSeg{radius, center, mse, isInlierFinal});
        }
        return cyl;
    }

    // primitive_detection.cpp:478-501 with :413-476
    void cylinder_fitting(unsigned cellActivatedCount, const std::vector<char>& activated)
    {
        cylinderSegments.push_back(make_cylinder(activated, cellActivatedCount));
        const CylinderSegment& cyl = cylinderSegments.back();
        const int region = int(cylinderSegments.size()) - 1;
        cylAssigned.emplace_back(cyl.segs.size(), 0);
        cylPlaneMse.emplace_back(cyl.segs.size(), DBL_MAX);
        for (unsigned segId = 0; segId < cyl.segs.size(); ++segId) {
            // mark the inlier mask for the parity output
            for (unsigned col = 0; col < cellActivatedCount; ++col)
                if (cyl.segs[segId].inliers[col] and segId < RS_MAX_CYL_SEGS and region < RS_MAX_CYL_REGIONS)
                    gridCylRegionSeg[cyl.local2global[col]] = 1 + region * RS_MAX_CYL_SEGS + int(segId);

            PlaneSeg newMergedPlane;
            newMergedPlane.clear();
            bool fitable = false;
            for (unsigned col = 0; col < cellActivatedCount; ++col)
                if (cyl.segs[segId].inliers[col]) {
                    const PlaneSeg& ps = grid[cyl.local2global[col]];
                    if (ps.planar) {
                        newMergedPlane.expand(ps);
                        fitable = true;
                    }
                }
            if (not fitable) continue;
            newMergedPlane.fit_plane();
            cylPlaneMse[region][segId] = newMergedPlane.mse;
            // add_cylinder_to_features (:437-476)
            if (newMergedPlane.mse < cyl.segs[segId].mse) {
                planeSegments.push_back(newMergedPlane.copy());
                const int currentPlaneCount = int(planeSegments.size());
                for (unsigned col = 0; col < cellActivatedCount; ++col)
                    if (cyl.segs[segId].inliers[col]) gridPlane[cyl.local2global[col]] = currentPlaneCount;
                cylAssigned[region][segId] = -currentPlaneCount;
            }
            else {
                cylinder2regionMap.emplace_back(region, int(segId));
                const int cylinderCount = int(cylinder2regionMap.size());
                for (unsigned col = 0; col < cellActivatedCount; ++col)
                    if (cyl.segs[segId].inliers[col]) gridCyl[cyl.local2global[col]] = cylinderCount;
                cylAssigned[region][segId] = cylinderCount;
            }
        }
    }

    // primitive_detection.cpp:312-389
    void grow_plane_segment_at_seed(unsigned seedId, unsigned& untriedPlanarCellsCount)
    {
        const PlaneSeg& planeToGrow = grid[seedId];
        if (not planeToGrow.planar) return;
        PlaneSeg newPlaneSegment = planeToGrow.copy();
        const unsigned y = seedId / unsigned(hc);
        const unsigned x = seedId % unsigned(hc);
        std::vector<char> activated(Nc, 0);
        region_growing(x, y, newPlaneSegment, activated);

        unsigned cellActivatedCount = 0;
        bool isPlaneFitable = false;
        for (int i = 0; i < Nc; ++i) {
            if (activated[i]) {
                const PlaneSeg& ps = grid[i];
                if (ps.planar) {
                    newPlaneSegment.expand(ps);
                    ++cellActivatedCount;
                    histogram.remove_point(unsigned(i));
                    unassigned[i] = 0;
                    --untriedPlanarCellsCount;
                    isPlaneFitable = true;
                }
            }
        }
        const unsigned minimumCellActivated = static_cast<unsigned>(kMinimumCellActivatedProportion * Nc);
        if (not isPlaneFitable or cellActivatedCount < minimumCellActivated) {
            histogram.remove_point(seedId);
            return;
        }
        newPlaneSegment.fit_plane();
        if (not newPlaneSegment.planar) return;
        if (newPlaneSegment.score > 100) {
            // add_plane_segment_to_features (:391-411)
            planeSegments.push_back(newPlaneSegment.copy());
            const int currentPlaneCount = int(planeSegments.size());
            for (int i = 0; i < Nc; ++i)
                if (activated[i]) gridPlane[i] = currentPlaneCount;
        }
        else if (cellActivatedCount > 5) {
            cylinder_fitting(cellActivatedCount, activated);
        }
    }

    // primitive_detection.cpp:267-310
    void grow_planes_and_cylinders(unsigned remainingPlanarCells)
    {
        unsigned untried = remainingPlanarCells;
        while (untried > 0) {
            const std::vector<unsigned> seedCandidates = histogram.most_frequent();
            const unsigned planeSeedCount = static_cast<unsigned>(kMinimumPlaneSeedProportion * Nc);
            if (seedCandidates.size() < planeSeedCount) break;
            unsigned seedId = 0;
            double minMSE = DBL_MAX;
            for (const unsigned cand : seedCandidates) {
                const double candidateMSE = grid[cand].mse;
                if (candidateMSE >= minMSE) continue;
                seedId = cand;
                minMSE = candidateMSE;
                if (minMSE <= 0) break;
            }
            if (minMSE >= DBL_MAX) break;
            ++nSeeds;
            grow_plane_segment_at_seed(seedId, untried);
        }
    }
};

// cv::erode / cv::dilate on 0/1 masks with 3x3 kernels. `cross` selects the cross-shaped kernel.
// borderZero: out-of-image neighbours count as 0 (BORDER_CONSTANT, Scalar(0)); otherwise they are ignored
// (OpenCV's default morphology border value: +inf for erode, -inf for dilate).
std::vector<unsigned char> morph(const std::vector<unsigned char>& m, int rows, int cols, bool erode, bool cross,
                                 bool borderZero)
{
    std::vector<unsigned char> r(m.size());
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            unsigned char v = erode ? 255 : 0;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    if (cross and dx != 0 and dy != 0) continue;
                    const int yy = y + dy, xx = x + dx;
                    unsigned char nv;
                    if (yy < 0 or yy >= rows or xx < 0 or xx >= cols) {
                        if (not borderZero) continue;
                        nv = 0;
                    }
                    else
                        nv = m[size_t(yy) * cols + xx];
                    v = erode ? std::min(v, nv) : std::max(v, nv);
                }
            r[size_t(y) * cols + x] = v;
        }
    return r;
}

}  // namespace

// test hook: the morphology restatement above, checked against the real cv2.erode / cv2.dilate in tests/test_oracle_cape.py
std::vector<unsigned char> cape_morphology(const std::vector<unsigned char>& m, int rows, int cols, bool erode, bool cross,
                                           bool borderZero)
{
    return morph(m, rows, cols, erode, cross, borderZero);
}

// primitive_detection.cpp:119-166
void cape_run(const CapeConfig& cfg, const float* depth, const uint32_t seed, CapeFrame& out)
{
    Detector det(cfg, seed);
    const int Nc = det.Nc, hc = det.hc, vc = det.vc, cs = cfg.cell;
    cape_cell_fit(cfg, depth, det.grid, det.tols);

    det.gridPlane.assign(Nc, 0);
    det.gridCyl.assign(Nc, 0);
    det.gridCylRegionSeg.assign(Nc, 0);
    det.unassigned.assign(Nc, 0);
    det.histogram.reset(cs);

    // init_histogram (:239-265)
    unsigned remainingPlanarCells = 0;
    std::vector<double> theta(Nc, 0.0), phi(Nc, 0.0);
    for (int i = 0; i < Nc; ++i) {
        const PlaneSeg& p = det.grid[i];
        if (p.planar) {
            theta[i] = std::acos(-p.normal.z);
            phi[i] = std::atan2(p.normal.x, p.normal.y);
            ++remainingPlanarCells;
            det.unassigned[i] = 1;
        }
    }
    det.histogram.init(theta, phi, det.unassigned);

    det.grow_planes_and_cylinders(remainingPlanarCells);

    // merge_planes (:503-560) with get_connected_components_matrix (:736-776)
    const int planeCount = int(det.planeSegments.size());
    std::vector<char> conn(size_t(planeCount) * planeCount, 0);
    if (planeCount > 0) {
        for (int row = 0; row < vc - 1; ++row)
            for (int col = 0; col < hc - 1; ++col) {
                const int planeId = det.gridPlane[row * hc + col];
                if (planeId <= 0) continue;
                const int nextPlaneId = det.gridPlane[row * hc + col + 1];
                const int belowPlaneId = det.gridPlane[(row + 1) * hc + col];
                if (nextPlaneId > 0 and planeId != nextPlaneId) {
                    conn[size_t(planeId - 1) * planeCount + nextPlaneId - 1] = 1;
                    conn[size_t(nextPlaneId - 1) * planeCount + planeId - 1] = 1;
                }
                if (belowPlaneId > 0 and planeId != belowPlaneId) {
                    conn[size_t(planeId - 1) * planeCount + belowPlaneId - 1] = 1;
                    conn[size_t(belowPlaneId - 1) * planeCount + planeId - 1] = 1;
                }
            }
    }
    std::vector<unsigned> planeMergeLabels(planeCount);
    for (int i = 0; i < planeCount; ++i) planeMergeLabels[i] = unsigned(i);
    for (int row = 0; row < planeCount; ++row) {
        bool wasPlaneExpanded = false;
        const unsigned planeId = planeMergeLabels[row];
        PlaneSeg& planeToExpand = det.planeSegments[planeId];
        if (not planeToExpand.planar) continue;
        for (int col = row + 1; col < planeCount; ++col) {
            if (not conn[size_t(row) * planeCount + col]) continue;
            const PlaneSeg& mergePlane = det.planeSegments[col];
            if (not mergePlane.planar) continue;
            if (planeToExpand.can_be_merged(mergePlane, kMaximumPlaneDistanceForMerge_mm)) {
                planeToExpand.expand(mergePlane);
                planeMergeLabels[col] = planeId;
                wasPlaneExpanded = true;
            }
            else {
                conn[size_t(row) * planeCount + col] = 0;
                conn[size_t(col) * planeCount + row] = 0;
            }
        }
        if (wasPlaneExpanded) planeToExpand.fit_plane();
    }

    // outputs
    out.cells.resize(Nc);
    for (int i = 0; i < Nc; ++i) cell_record(det.grid[i], det.tols[i], cfg.cell, out.cells[i]);
    out.plane_grid = det.gridPlane;
    out.cyl_labels = det.gridCyl;
    out.cyl_region_seg = det.gridCylRegionSeg;
    out.plane_labels.assign(Nc, 0);
    out.planes.assign(planeCount, rs_plane_out{});
    out.boundary_xyz.clear();
    out.info = rs_cape_frame_info{};
    out.info.status = RS_OK;
    if (planeCount > RS_MAX_PLANES or int(det.cylinderSegments.size()) > RS_MAX_CYL_REGIONS)
        out.info.status = RS_ERR_CAPACITY;

    std::vector<double> kx, ky;
    backprojection_factors(cfg, kx, ky);

    int nFinal = 0;
    for (int k = 0; k < planeCount; ++k) {
        const PlaneSeg& ps = det.planeSegments[k];
        rs_plane_out& po = out.planes[k];
        po.merge_label = int(planeMergeLabels[k]);
        po.planar = ps.planar ? 1 : 0;
        po.is_final = (planeMergeLabels[k] == unsigned(k) and ps.planar) ? 1 : 0;
        po.count = ps.count;
        for (int i = 0; i < 9; ++i) po.S[i] = ps.S[i];
        for (int i = 0; i < 3; ++i) {
            po.centroid[i] = ps.centroid[i];
            po.normal[i] = ps.normal[i];
        }
        po.d = ps.d;
        po.mse = ps.mse;
        po.score = ps.score;
        po.n_boundary = 0;
        po.boundary_offset = int(out.boundary_xyz.size() / 3);
        if (not po.is_final) continue;
        ++nFinal;

        // add_planes_to_primitives (:562-648): mask of all segments merged into k
        std::vector<unsigned char> mask(Nc, 0);
        for (int j = k; j < planeCount; ++j)
            if (planeMergeLabels[j] == planeMergeLabels[k])
                for (int c = 0; c < Nc; ++c)
                    if (det.gridPlane[c] == j + 1) mask[c] = 1;
        for (int c = 0; c < Nc; ++c)
            if (mask[c]) out.plane_labels[c] = k + 1;

        // compute_plane_segment_boundary (:650-703)
        const double maxBoundaryDistance = 3 * std::sqrt(ps.mse);
        const unsigned pixelPerCellSide = static_cast<unsigned>(sqrtf(static_cast<float>(cs * cs)));
        const std::vector<unsigned char> eroded = morph(mask, vc, hc, true, true, true);
        const std::vector<unsigned char> dilated = morph(mask, vc, hc, false, false, false);
        for (int row = 0; row < vc; ++row)
            for (int col = 0; col < hc; ++col) {
                const int idx = row * hc + col;
                const int v = int(dilated[idx]) - int(eroded[idx]);
                if (v <= 0) continue;
                const int centerX = int(col * pixelPerCellSide + pixelPerCellSide / 2);
                const int centerY = int(row * pixelPerCellSide + pixelPerCellSide / 2);
                const double depthValue = depth[size_t(centerY) * cfg.width + centerX];
                if (depthValue > 0) {
                    // ScreenCoordinate(x, y, depth).to_camera_coordinates() in FP64 (point_coordinates.cpp:150-167)
                    const Vec3 p{depthValue * kx[centerX], depthValue * ky[centerY], depthValue};
                    if (std::fabs(dot(ps.normal, p) + ps.d) < maxBoundaryDistance) {
                        out.boundary_xyz.push_back(p.x);
                        out.boundary_xyz.push_back(p.y);
                        out.boundary_xyz.push_back(p.z);
                        po.n_boundary++;
                    }
                }
            }
    }

    // cylinders: add_cylinders_to_primitives (:705-734) decides which ids survive
    std::vector<char> cylKept(det.cylinder2regionMap.size(), 0);
    for (size_t ci = 0; ci < det.cylinder2regionMap.size(); ++ci) {
        std::vector<unsigned char> mask(Nc, 0);
        for (int c = 0; c < Nc; ++c)
            if (det.gridCyl[c] == int(ci) + 1) mask[c] = 1;
        mask = morph(mask, vc, hc, false, true, false);
        mask = morph(mask, vc, hc, true, true, false);
        const std::vector<unsigned char> er = morph(mask, vc, hc, true, true, false);
        unsigned char mn = 255, mx = 0;
        for (unsigned char v : er) {
            mn = std::min(mn, v);
            mx = std::max(mx, v);
        }
        cylKept[ci] = not(mx <= 0 or mn >= mx);
    }
    out.cyls.assign(det.cylinderSegments.size(), rs_cyl_out{});
    for (size_t r = 0; r < det.cylinderSegments.size(); ++r) {
        const CylinderSegment& cs_ = det.cylinderSegments[r];
        rs_cyl_out& co = out.cyls[r];
        co.n_cells = cs_.cellActivatedCount;
        co.n_segments = int(cs_.segs.size());
        co.pca_score = cs_.pcaScore;
        for (int i = 0; i < 3; ++i) co.axis[i] = cs_.axis[i];
        if (co.n_segments > RS_MAX_CYL_SEGS) out.info.status = RS_ERR_CAPACITY;
        for (int s = 0; s < std::min<int>(co.n_segments, RS_MAX_CYL_SEGS); ++s) {
            co.radius[s] = cs_.segs[s].radius;
            for (int i = 0; i < 3; ++i) co.center[s][i] = cs_.segs[s].center[i];
            co.mse[s] = cs_.segs[s].mse;
            co.plane_mse[s] = det.cylPlaneMse[r][s];
            int ni = 0;
            for (char c : cs_.segs[s].inliers) ni += c ? 1 : 0;
            co.n_inliers[s] = ni;
            co.assigned[s] = det.cylAssigned[r][s];
            co.kept[s] = (co.assigned[s] > 0 and cylKept[co.assigned[s] - 1]) ? 1 : 0;
        }
    }

    int nPlanar = 0;
    for (int i = 0; i < Nc; ++i) nPlanar += det.grid[i].planar ? 1 : 0;
    out.info.n_planar_cells = nPlanar;
    out.info.n_seeds = det.nSeeds;
    out.info.n_planes = planeCount;
    out.info.n_final_planes = nFinal;
    out.info.n_cyl_regions = int(det.cylinderSegments.size());
    out.info.n_cylinders = int(det.cylinder2regionMap.size());
    out.info.n_boundary = int(out.boundary_xyz.size() / 3);
}

// ---- Depth_Map_Transformation::rectify_depth (depth_map_transformation.cpp:23-87) ------------------------------------
void rectify_depth(const CapeConfig& cfg, const double T[16], const float* depth, float* out)
{
    const int W = cfg.width, H = cfg.height;
    std::vector<double> kx, ky;
    backprojection_factors(cfg, kx, ky);   // ScreenCoordinate2D(col,row).to_camera_coordinates(), init_matrices :160-164
    std::vector<float> preX(W), preY(H);   // _Xpre / _Ypre are float images; they only depend on the column / the row
    for (int c = 0; c < W; ++c) preX[c] = static_cast<float>(kx[c]);
    for (int r = 0; r < H; ++r) preY[r] = static_cast<float>(ky[r]);
    std::fill(out, out + size_t(W) * H, 0.0f);
    for (int row = 0; row < H; ++row)
        for (int col = 0; col < W; ++col) {
            const float originalZ = depth[size_t(row) * W + col];
            if (originalZ <= 0) continue;
            // vector3 original(preX * z, preY * z, z): float products, widened by the vector3 constructor
            const double ox = static_cast<double>(preX[col] * originalZ);
            const double oy = static_cast<double>(preY[row] * originalZ);
            const double oz = static_cast<double>(originalZ);
            const double px = ((T[0] * ox + T[1] * oy) + T[2] * oz) + T[3];
            const double py = ((T[4] * ox + T[5] * oy) + T[6] * oz) + T[7];
            const double pz = ((T[8] * ox + T[9] * oy) + T[10] * oz) + T[11];
            // CameraCoordinate::to_screen_coordinates (point_coordinates.cpp:201-210): 1.0 / z * (K * p).head<2>()
            const double inv = 1.0 / pz;
            const double sx = inv * ((cfg.fx * px + 0.0 * py) + cfg.cx * pz);
            const double sy = inv * ((0.0 * px + cfg.fy * py) + cfg.cy * pz);
            if (sx != sx || sy != sy) continue;   // the reference exit(-1)s here; unreachable for finite inputs with z > 0
            // static_cast<uint>(floor(.)) of a negative double: x86-64 converts through a 64-bit integer and truncates,
            // giving a value >= 2^31 that fails the bounds test below
            const double fxs = std::floor(sx), fys = std::floor(sy);
            if (!(fxs > 0.0 && fys > 0.0 && fxs < double(W) && fys < double(H))) continue;
            out[size_t(fys) * W + size_t(fxs)] = static_cast<float>(pz);
        }
}

}  // namespace oracle
