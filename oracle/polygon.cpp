// TEST INFRASTRUCTURE - see polygon.hpp. Not part of the product.
#include "polygon.hpp"

#include <algorithm>
#include <cmath>
#include <vector>

namespace oracle {

namespace {
struct Edge {
    double xa, ya, xb, yb;   // xa < xb
    double at(double x) const { return ya + (x - xa) * ((yb - ya) / (xb - xa)); }
};

void ring_edges(const double* p, int n, std::vector<Edge>& out, std::vector<double>& xs)
{
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1) % n;
        double xa = p[2 * i], ya = p[2 * i + 1], xb = p[2 * j], yb = p[2 * j + 1];
        xs.push_back(xa);
        if (xa == xb) continue;   // vertical edges bound no slab interior
        if (xa > xb) std::swap(xa, xb), std::swap(ya, yb);
        out.push_back(Edge{xa, ya, xb, yb});
    }
}

// sorted (by y at the slab's midline) spanning edges of one ring -> inside intervals by the even-odd rule
void slab_intervals(const std::vector<Edge>& edges, double x0, double x1, std::vector<std::pair<const Edge*, const Edge*>>& iv)
{
    const double xm = 0.5 * (x0 + x1);
    std::vector<const Edge*> span;
    for (const Edge& e : edges)
        if (e.xa <= x0 && e.xb >= x1) span.push_back(&e);
    std::sort(span.begin(), span.end(), [xm](const Edge* a, const Edge* b) { return a->at(xm) < b->at(xm); });
    iv.clear();
    for (size_t k = 0; k + 1 < span.size(); k += 2) iv.emplace_back(span[k], span[k + 1]);
}
}  // namespace

double polygon_area(const double* a, int n)
{
    double s = 0;
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1) % n;
        s += a[2 * i] * a[2 * j + 1] - a[2 * j] * a[2 * i + 1];
    }
    return std::fabs(s) * 0.5;
}

double polygon_inter_area(const double* a, int na, const double* b, int nb)
{
    if (na < 3 || nb < 3) return 0.0;
    std::vector<Edge> ea, eb;
    std::vector<double> xs;
    ring_edges(a, na, ea, xs);
    ring_edges(b, nb, eb, xs);
    // proper crossings between the two boundaries are slab breakpoints too: inside a slab the vertical order is fixed
    for (const Edge& p : ea)
        for (const Edge& q : eb) {
            const double lo = std::max(p.xa, q.xa), hi = std::min(p.xb, q.xb);
            if (!(lo < hi)) continue;
            const double d0 = p.at(lo) - q.at(lo), d1 = p.at(hi) - q.at(hi);
            if ((d0 < 0 && d1 > 0) || (d0 > 0 && d1 < 0)) xs.push_back(lo + (hi - lo) * (d0 / (d0 - d1)));
        }
    std::sort(xs.begin(), xs.end());
    xs.erase(std::unique(xs.begin(), xs.end()), xs.end());
    double total = 0.0;
    std::vector<std::pair<const Edge*, const Edge*>> ia, ib;
    for (size_t s = 0; s + 1 < xs.size(); ++s) {
        const double x0 = xs[s], x1 = xs[s + 1];
        if (!(x1 > x0)) continue;
        slab_intervals(ea, x0, x1, ia);
        slab_intervals(eb, x0, x1, ib);
        const double xm = 0.5 * (x0 + x1);
        for (const auto& A : ia)
            for (const auto& B : ib) {
                const Edge* lo = A.first->at(xm) > B.first->at(xm) ? A.first : B.first;
                const Edge* hi = A.second->at(xm) < B.second->at(xm) ? A.second : B.second;
                if (!(hi->at(xm) > lo->at(xm))) continue;
                // the overlap length is linear over the slab (no crossing inside it): trapezoid rule is exact
                const double l0 = hi->at(x0) - lo->at(x0), l1 = hi->at(x1) - lo->at(x1);
                total += 0.5 * (std::max(l0, 0.0) + std::max(l1, 0.0)) * (x1 - x0);
            }
    }
    return total;
}

namespace {
struct V3 {
    double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 normalized(V3 a)
{
    const double n = std::sqrt(dot(a, a));
    return n > 0 ? (1.0 / n) * a : a;
}
inline V3 v3(const double* p) { return {p[0], p[1], p[2]}; }
inline V3 rot(const double* T, V3 p) { return {T[0] * p.x + T[1] * p.y + T[2] * p.z, T[4] * p.x + T[5] * p.y + T[6] * p.z, T[8] * p.x + T[9] * p.y + T[10] * p.z}; }
inline V3 xform(const double* T, V3 p) { return rot(T, p) + V3{T[3], T[7], T[11]}; }
}  // namespace

void plane_match_frame(const double* w2c, const rs_polygon_plane* det, int n_det, const double* det_xy, const rs_polygon_plane* map,
                       int n_map, const double* map_xy, const unsigned char* det_matched, int advanced_search, int sequential,
                       int* selected, double* inter, unsigned char* matched_out)
{
    // _isDetectedFeatureMatched as the caller's loop carries it (feature_map.hpp:652-669)
    std::vector<unsigned char> matched(size_t(n_det), 0);
    for (int k = 0; k < n_det; ++k) matched[k] = det_matched ? det_matched[k] : 0;
    const double minimumNormalDotDiff = std::fabs(std::cos(20.0 * M_PI / 180.0));   // maximumAngleForPlaneMatch_d
    const double maximumPlaneMatchDistance = 100.0;                                 // maximumDistanceForPlaneMatch_mm
    const double planeMinimalOverlap = static_cast<double>(0.4f);                   // minimumPlaneOverlapToConsiderMatch (float)
    const double threshold = advanced_search ? planeMinimalOverlap / 2 : planeMinimalOverlap;
    for (int m = 0; m < n_map; ++m) {
        selected[m] = -1, inter[m] = 0.0;
        const rs_polygon_plane& mp = map[m];
        // PlaneWorldCoordinates::to_camera_coordinates: n renormalised, d kept (plane_coordinates.cpp:20-24)
        const V3 nw = v3(mp.normal);
        V3 nc = rot(w2c, nw);
        const double dc = mp.d - dot(nc, V3{w2c[3], w2c[7], w2c[11]});
        nc = normalized(nc);
        // WorldPolygon::to_camera_space (polygon_coordinates.cpp:135-162)
        const V3 c = v3(mp.center), X = v3(mp.x_axis), Y = v3(mp.y_axis);
        const V3 nC = xform(w2c, c), nX = normalized(rot(w2c, X)), nY = normalized(rot(w2c, Y));
        std::vector<double> cam(2 * size_t(mp.n_vertices));
        for (int v = 0; v < mp.n_vertices; ++v) {
            const double px = map_xy[2 * (mp.first_vertex + v)], py = map_xy[2 * (mp.first_vertex + v) + 1];
            const V3 t = xform(w2c, c + px * X + py * Y);
            cam[2 * v] = dot(nX, t - nC), cam[2 * v + 1] = dot(nY, t - nC);
        }
        const double projectedArea = polygon_area(cam.data(), mp.n_vertices);
        if (projectedArea <= 0.0) continue;
        double greatest = 0.0;
        int sel = -1;
        for (int k = 0; k < n_det; ++k) {
            if (matched[k]) continue;
            const rs_polygon_plane& dp = det[k];
            if (!(std::fabs(dp.d - dc) < maximumPlaneMatchDistance)) continue;
            if (!(std::fabs(dot(v3(dp.normal), nc)) > minimumNormalDotDiff)) continue;
            // other.project(_xAxis, _yAxis, _center) (polygon.cpp:349-382): orthogonal projection onto the detected plane's frame
            const V3 dc3 = v3(dp.center), dX = v3(dp.x_axis), dY = v3(dp.y_axis);
            std::vector<double> prj(2 * size_t(mp.n_vertices));
            for (int v = 0; v < mp.n_vertices; ++v) {
                const V3 p3 = nC + cam[2 * v] * nX + cam[2 * v + 1] * nY;
                prj[2 * v] = dot(dX, p3 - dc3), prj[2 * v + 1] = dot(dY, p3 - dc3);
            }
            const double* dxy = det_xy + 2 * size_t(dp.first_vertex);
            const double newPlaneArea = polygon_area(dxy, dp.n_vertices);
            const double interArea = polygon_inter_area(dxy, dp.n_vertices, prj.data(), mp.n_vertices);
            if (interArea > greatest && interArea / newPlaneArea >= threshold) sel = k, greatest = interArea;
        }
        if (sel <= 0) continue;   // sic (map_primitive.cpp:146): detection 0 can never be matched
        selected[m] = sel, inter[m] = greatest;
        if (sequential) matched[sel] = 1;   // feature_map.hpp:666-667
    }
    if (matched_out)
        for (int k = 0; k < n_det; ++k) matched_out[k] = matched[k];
}

}  // namespace oracle
