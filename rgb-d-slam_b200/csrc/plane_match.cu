// Plane matching against the local map: MapPlane::find_matches (src/map_management/map_features/map_primitive.cpp:91-161)
// for every map plane of every frame, and the polygon intersection area behind it (Polygon::inter_area,
// src/utils/polygon.cpp:542-561 = summed area of boost::geometry::intersection of two valid polygons).
//
// Mapping (B200): ONE WARP PER MAP PLANE. The map polygon goes to camera space once (lanes over vertices, shared memory);
// for every candidate detection the warp projects it into the detection's frame and computes the intersection area WITHOUT
// constructing the intersection: a simple polygon is the signed sum of the fan triangles (O, p_i, p_i+1) - its indicator
// function is sum_i s_i 1[T_i] almost everywhere - so |A n B| = sum_i sum_j s_i t_j |T_i n U_j|, and each term is a
// triangle-triangle clip (Sutherland-Hodgman against three half-planes, at most seven vertices in registers). Lanes take
// the (i, j) pairs with a stride of 32; one shuffle reduction per candidate. No dense contraction, nothing for tensor cores;
// the data of a frame (a few polygons of tens of vertices) lives in L1 / shared memory, the kernel is FP64-latency bound.
#include <cuda_runtime.h>
#include <float.h>

#include <vector>

#include "common.cuh"

namespace rs {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kMaxVertices = 256;   // vertices of one polygon ring held in shared memory per warp (map side)
constexpr int WARPS = 4;

struct P2 {
    double x, y;
};

// area of the intersection of triangles (O, a0, a1) and (O, b0, b1), O = origin; both given counter-clockwise
__device__ double tri_tri_area(const P2 a0, const P2 a1, const P2 b0, const P2 b1)
{
    P2 poly[8], tmp[8];
    int n = 3;
    poly[0] = P2{0.0, 0.0}, poly[1] = a0, poly[2] = a1;
    const P2 clip[3] = {P2{0.0, 0.0}, b0, b1};
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        const P2 c0 = clip[e], c1 = clip[(e + 1) % 3];
        const double ex = c1.x - c0.x, ey = c1.y - c0.y;
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const P2 p = poly[i], q = poly[i + 1 == n ? 0 : i + 1];
            const double dp = ex * (p.y - c0.y) - ey * (p.x - c0.x);   // > 0: left of the clip edge = inside
            const double dq = ex * (q.y - c0.y) - ey * (q.x - c0.x);
            if (dp >= 0.0) tmp[m++] = p;
            if ((dp > 0.0 && dq < 0.0) || (dp < 0.0 && dq > 0.0)) {
                const double t = dp / (dp - dq);
                tmp[m++] = P2{p.x + t * (q.x - p.x), p.y + t * (q.y - p.y)};
            }
        }
        n = m;
        if (n < 3) return 0.0;
        for (int i = 0; i < n; ++i) poly[i] = tmp[i];
    }
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        const P2 p = poly[i], q = poly[i + 1 == n ? 0 : i + 1];
        s += p.x * q.y - q.x * p.y;
    }
    return 0.5 * fabs(s);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// signed shoelace area of a ring (all lanes call; result on every lane)
__device__ double ring_signed_area(const double* xy, const int n, const int lane)
{
    double s = 0.0;
    for (int i = lane; i < n; i += 32) {
        const int j = i + 1 == n ? 0 : i + 1;
        s += xy[2 * i] * xy[2 * j + 1] - xy[2 * j] * xy[2 * i + 1];
    }
    return 0.5 * warp_sum(s);
}

// |A n B| of two simple rings in a common frame (all lanes call; result on every lane). The fan origin is moved to A's
// first vertex: the decomposition holds for any origin, one near the polygons keeps the cancelling terms small.
__device__ double inter_area_warp(const double* a, const int na, const double* b, const int nb, const int lane)
{
    if (na < 3 || nb < 3) return 0.0;
    const double sa = ring_signed_area(a, na, lane), sb = ring_signed_area(b, nb, lane);
    if (sa == 0.0 || sb == 0.0) return 0.0;
    const double ox = a[0], oy = a[1];
    double acc = 0.0;
    const int pairs = na * nb;
    for (int t = lane; t < pairs; t += 32) {
        const int i = t / nb, j = t - i * nb;
        const int i1 = i + 1 == na ? 0 : i + 1, j1 = j + 1 == nb ? 0 : j + 1;
        P2 a0{a[2 * i] - ox, a[2 * i + 1] - oy}, a1{a[2 * i1] - ox, a[2 * i1 + 1] - oy};
        P2 b0{b[2 * j] - ox, b[2 * j + 1] - oy}, b1{b[2 * j1] - ox, b[2 * j1 + 1] - oy};
        double si = a0.x * a1.y - a1.x * a0.y, sj = b0.x * b1.y - b1.x * b0.y;   // twice the signed areas of the fan triangles
        if (si == 0.0 || sj == 0.0) continue;
        double sign = 1.0;
        if (si < 0.0) {   // make both triangles counter-clockwise, remember the sign
            const P2 w = a0;
            a0 = a1, a1 = w, sign = -sign;
        }
        if (sj < 0.0) {
            const P2 w = b0;
            b0 = b1, b1 = w, sign = -sign;
        }
        acc += sign * tri_tri_area(a0, a1, b0, b1);
    }
    acc = warp_sum(acc);
    // the rings' own orientations: a clockwise ring's fan sums to -1 inside
    if ((sa < 0.0) != (sb < 0.0)) acc = -acc;
    return acc > 0.0 ? acc : 0.0;
}

struct MatchArgs {
    int n_frames;
    const double* w2c;
    const rs_polygon_plane* det;
    const int32_t* det_first;
    const double* det_xy;
    const rs_polygon_plane* map;
    const int32_t* map_first;
    const double* map_xy;
    const uint8_t* det_matched;
    const int32_t* map_frame;   // frame of every map plane
    int n_map;
    int advanced;
    int sequential;
    double* cand;               // per map plane, one entry per detection of its frame: the intersection area if the pair
    const int32_t* cand_first;  // qualifies (distance, normal, overlap share), else 0; map plane m owns [cand_first[m], +n_det(frame))
    int32_t* selected;
    double* inter;
    uint8_t* matched_out;       // per detection: det_matched with this call's selections added (sequential mode), or null
};

__device__ __forceinline__ void rot3(const double* T, const double* p, double* o)
{
    o[0] = T[0] * p[0] + T[1] * p[1] + T[2] * p[2];
    o[1] = T[4] * p[0] + T[5] * p[1] + T[6] * p[2];
    o[2] = T[8] * p[0] + T[9] * p[1] + T[10] * p[2];
}
__device__ __forceinline__ void normalize3v(double* v)
{
    const double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (n > 0.0) v[0] /= n, v[1] /= n, v[2] /= n;
}

__global__ void __launch_bounds__(WARPS * 32) plane_match_kernel(const MatchArgs g)
{
    __shared__ double s_cam[WARPS][2 * kMaxVertices];   // the map polygon in camera space (its own frame there)
    __shared__ double s_prj[WARPS][2 * kMaxVertices];   // ... projected into the candidate detection's frame
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.x * WARPS + warp;
    if (m >= g.n_map) return;
    const rs_polygon_plane mp = g.map[m];
    const int f = g.map_frame[m];
    const double* T = g.w2c + size_t(f) * 16;
    const int k0 = g.det_first[f], k1 = g.det_first[f + 1];
    double* cand = g.cand + g.cand_first[m];
    for (int k = k0 + lane; k < k1; k += 32) cand[k - k0] = 0.0;
    __syncwarp();
    const int nv = mp.n_vertices;
    if (nv >= 3 && nv <= kMaxVertices) {
        const double minimumNormalDotDiff = fabs(cos(20.0 * kPi / 180.0));
        const double maximumPlaneMatchDistance = 100.0;
        const double planeMinimalOverlap = static_cast<double>(0.4f);
        const double threshold = g.advanced ? planeMinimalOverlap / 2 : planeMinimalOverlap;
        // plane to camera space: n renormalised, d kept
        double nc[3];
        rot3(T, mp.normal, nc);
        const double dc = mp.d - (nc[0] * T[3] + nc[1] * T[7] + nc[2] * T[11]);
        normalize3v(nc);
        // polygon frame to camera space
        double nC[3], nX[3], nY[3];
        rot3(T, mp.center, nC);
        nC[0] += T[3], nC[1] += T[7], nC[2] += T[11];
        rot3(T, mp.x_axis, nX);
        rot3(T, mp.y_axis, nY);
        normalize3v(nX);
        normalize3v(nY);
        double* cam = s_cam[warp];
        double* prj = s_prj[warp];
        for (int v = lane; v < nv; v += 32) {
            const double px = g.map_xy[2 * (size_t(mp.first_vertex) + v)], py = g.map_xy[2 * (size_t(mp.first_vertex) + v) + 1];
            const double w[3] = {mp.center[0] + px * mp.x_axis[0] + py * mp.y_axis[0], mp.center[1] + px * mp.x_axis[1] + py * mp.y_axis[1],
                                 mp.center[2] + px * mp.x_axis[2] + py * mp.y_axis[2]};
            double t[3];
            rot3(T, w, t);
            t[0] += T[3] - nC[0], t[1] += T[7] - nC[1], t[2] += T[11] - nC[2];
            cam[2 * v] = nX[0] * t[0] + nX[1] * t[1] + nX[2] * t[2];
            cam[2 * v + 1] = nY[0] * t[0] + nY[1] * t[1] + nY[2] * t[2];
        }
        __syncwarp();
        const double projectedArea = fabs(ring_signed_area(cam, nv, lane));
        if (projectedArea > 0.0) {
            for (int k = k0; k < k1; ++k) {
                if (!g.sequential && g.det_matched && g.det_matched[k]) continue;   // (sequential: the selection pass decides)
                const rs_polygon_plane& dp = g.det[k];
                if (!(fabs(dp.d - dc) < maximumPlaneMatchDistance)) continue;
                if (!(fabs(dp.normal[0] * nc[0] + dp.normal[1] * nc[1] + dp.normal[2] * nc[2]) > minimumNormalDotDiff)) continue;
                if (dp.n_vertices < 3) continue;
                for (int v = lane; v < nv; v += 32) {
                    const double p3[3] = {nC[0] + cam[2 * v] * nX[0] + cam[2 * v + 1] * nY[0] - dp.center[0],
                                          nC[1] + cam[2 * v] * nX[1] + cam[2 * v + 1] * nY[1] - dp.center[1],
                                          nC[2] + cam[2 * v] * nX[2] + cam[2 * v + 1] * nY[2] - dp.center[2]};
                    prj[2 * v] = dp.x_axis[0] * p3[0] + dp.x_axis[1] * p3[1] + dp.x_axis[2] * p3[2];
                    prj[2 * v + 1] = dp.y_axis[0] * p3[0] + dp.y_axis[1] * p3[1] + dp.y_axis[2] * p3[2];
                }
                __syncwarp();
                const double* dxy = g.det_xy + 2 * size_t(dp.first_vertex);
                const double newPlaneArea = fabs(ring_signed_area(dxy, dp.n_vertices, lane));
                const double interArea = inter_area_warp(dxy, dp.n_vertices, prj, nv, lane);
                if (lane == 0 && interArea / newPlaneArea >= threshold) cand[k - k0] = interArea;
                __syncwarp();
            }
        }
    }
}

// The selection of MapPlane::find_matches (map_primitive.cpp:114-147) and the caller's loop over the map's planes
// (feature_map.hpp:652-669): one thread per frame walks its map planes IN ORDER; a map plane takes the unmatched detection
// with the greatest qualifying intersection area (the first one on ties: `interArea > greatestSimilarity`), never detection 0
// (`if (selectedIndex <= 0) return`, sic), and in sequential mode that detection is marked matched for the map planes after it.
__global__ void plane_select_kernel(const MatchArgs g)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= g.n_frames) return;
    const int k0 = g.det_first[f], nd = g.det_first[f + 1] - k0;
    if (g.matched_out)
        for (int k = 0; k < nd; ++k) g.matched_out[k0 + k] = g.det_matched ? g.det_matched[k0 + k] : 0;
    for (int m = g.map_first[f]; m < g.map_first[f + 1]; ++m) {
        const double* cand = g.cand + g.cand_first[m];
        int sel = -1;
        double greatest = 0.0;
        for (int k = 0; k < nd; ++k) {
            const bool taken = g.sequential ? (g.matched_out[k0 + k] != 0) : (g.det_matched && g.det_matched[k0 + k]);
            if (taken) continue;
            if (cand[k] > greatest) sel = k, greatest = cand[k];
        }
        if (sel <= 0) sel = -1, greatest = 0.0;
        g.selected[m] = sel, g.inter[m] = greatest;
        if (sel > 0 && g.sequential) g.matched_out[k0 + sel] = 1;
    }
}

__global__ void __launch_bounds__(WARPS * 32) polygon_inter_area_kernel(const int n_pairs, const double* a_xy, const int32_t* a_first,
                                                                        const double* b_xy, const int32_t* b_first, double* area)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * WARPS + warp;
    if (p >= n_pairs) return;
    const double v = inter_area_warp(a_xy + 2 * size_t(a_first[p]), a_first[p + 1] - a_first[p], b_xy + 2 * size_t(b_first[p]),
                                     b_first[p + 1] - b_first[p], lane);
    if (lane == 0) area[p] = v;
}

// Device scratch of one call, from the device's stream-ordered memory pool (cudaMallocAsync): after the first call of a process
// the pool hands the same blocks back without touching the driver (a cudaMalloc / cudaFree pair per array cost 100 ms per
// call beside a process's other allocations - cudaFree synchronises the whole device).
struct CallStream {
    cudaStream_t s = nullptr;
    ~CallStream()
    {
        if (s) cudaStreamDestroy(s);
    }
    int open(const int device)
    {
        static std::mutex m;
        static bool pool_kept[64] = {};
        RS_CUDA_CHECK(cudaSetDevice(device));
        {
            std::lock_guard<std::mutex> lock(m);
            if (!pool_kept[device & 63]) {
                cudaMemPool_t pool;
                RS_CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device));
                uint64_t keep = UINT64_MAX;   // freed blocks stay in the pool for the next call
                RS_CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
                pool_kept[device & 63] = true;
            }
        }
        RS_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        return RS_OK;
    }
};

template <class T>
struct DevBuf {
    T* p = nullptr;
    cudaStream_t s = nullptr;
    ~DevBuf()
    {
        if (p) cudaFreeAsync(p, s);
    }
    int upload(const T* host, size_t n, cudaStream_t stream)
    {
        s = stream;
        RS_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&p), sizeof(T) * (n ? n : 1), s));
        if (n && host) RS_CUDA_CHECK(cudaMemcpyAsync(p, host, sizeof(T) * n, cudaMemcpyHostToDevice, s));
        return RS_OK;
    }
};

}  // namespace

int require_blackwell(int device);

}  // namespace rs

using namespace rs;

extern "C" {

int rs_plane_match(int device, int n_frames, const double* world_to_camera, const rs_polygon_plane* det, const int32_t* det_first,
                   const double* det_xy, const rs_polygon_plane* map, const int32_t* map_first, const double* map_xy,
                   const uint8_t* det_matched, int advanced_search, int sequential, int32_t* selected, double* inter_area,
                   uint8_t* det_matched_out)
{
    int rc = require_blackwell(device);
    if (rc != RS_OK) return rc;
    if (n_frames <= 0 || !world_to_camera || !det_first || !map_first || !selected || !inter_area) {
        set_last_error("rs_plane_match: invalid argument");
        return RS_ERR_INVALID_ARG;
    }
    for (int f = 0; f < n_frames; ++f)
        if (det_first[f] < 0 || map_first[f] < 0 || det_first[f + 1] < det_first[f] || map_first[f + 1] < map_first[f]) {
            set_last_error("rs_plane_match: det_first / map_first must be non-decreasing offsets starting at or above 0");
            return RS_ERR_INVALID_ARG;
        }
    const int n_det = det_first[n_frames], n_map = map_first[n_frames];
    if (n_map == 0) return RS_OK;
    if ((n_det && (!det || !det_xy)) || !map || !map_xy) {
        set_last_error("rs_plane_match: null plane / vertex arrays");
        return RS_ERR_INVALID_ARG;
    }
    size_t det_vertices = 0, map_vertices = 0;
    for (int k = 0; k < n_det; ++k) {
        if (det[k].first_vertex < 0 || det[k].n_vertices < 0) {
            set_last_error("rs_plane_match: negative vertex range");
            return RS_ERR_INVALID_ARG;
        }
        det_vertices = std::max(det_vertices, size_t(det[k].first_vertex) + det[k].n_vertices);
    }
    std::vector<int32_t> map_frame(n_map), cand_first(size_t(n_map) + 1, 0);
    for (int f = 0; f < n_frames; ++f)
        for (int m = map_first[f]; m < map_first[f + 1]; ++m) {
            map_frame[m] = f;
            cand_first[m + 1] = cand_first[m] + (det_first[f + 1] - det_first[f]);
        }
    for (int m = 0; m < n_map; ++m) {
        if (map[m].first_vertex < 0 || map[m].n_vertices < 0 || map[m].n_vertices > kMaxVertices) {
            set_last_error("rs_plane_match: a map polygon has more than 256 vertices (or a negative range)");
            return RS_ERR_INVALID_ARG;
        }
        map_vertices = std::max(map_vertices, size_t(map[m].first_vertex) + map[m].n_vertices);
    }
    CallStream cs;   // declared first: destroyed after the buffers below have queued their frees on it
    if ((rc = cs.open(device))) return rc;
    DevBuf<double> d_T, d_dxy, d_mxy, d_inter;
    DevBuf<rs_polygon_plane> d_det, d_map;
    DevBuf<int32_t> d_df, d_mf, d_sel;
    DevBuf<uint8_t> d_matched;
    if ((rc = d_T.upload(world_to_camera, size_t(n_frames) * 16, cs.s))) return rc;
    if ((rc = d_det.upload(det, n_det, cs.s))) return rc;
    if ((rc = d_map.upload(map, n_map, cs.s))) return rc;
    if ((rc = d_df.upload(det_first, size_t(n_frames) + 1, cs.s))) return rc;
    if ((rc = d_mf.upload(map_frame.data(), n_map, cs.s))) return rc;
    if ((rc = d_dxy.upload(det_xy, 2 * det_vertices, cs.s))) return rc;
    if ((rc = d_mxy.upload(map_xy, 2 * map_vertices, cs.s))) return rc;
    if (det_matched && (rc = d_matched.upload(det_matched, n_det, cs.s))) return rc;
    if ((rc = d_sel.upload(nullptr, n_map, cs.s))) return rc;
    if ((rc = d_inter.upload(nullptr, n_map, cs.s))) return rc;
    DevBuf<double> d_cand;
    DevBuf<int32_t> d_cf, d_mfirst;
    DevBuf<uint8_t> d_mout;
    if ((rc = d_cand.upload(nullptr, size_t(cand_first[n_map]), cs.s))) return rc;
    if ((rc = d_cf.upload(cand_first.data(), size_t(n_map) + 1, cs.s))) return rc;
    if ((rc = d_mfirst.upload(map_first, size_t(n_frames) + 1, cs.s))) return rc;
    const bool want_mask = sequential || det_matched_out;
    if (want_mask && (rc = d_mout.upload(nullptr, n_det, cs.s))) return rc;
    MatchArgs g;
    g.n_frames = n_frames, g.w2c = d_T.p, g.det = d_det.p, g.det_first = d_df.p, g.det_xy = d_dxy.p, g.map = d_map.p;
    g.map_first = d_mfirst.p, g.map_xy = d_mxy.p, g.det_matched = det_matched ? d_matched.p : nullptr, g.map_frame = d_mf.p;
    g.n_map = n_map, g.advanced = advanced_search, g.sequential = sequential ? 1 : 0, g.selected = d_sel.p, g.inter = d_inter.p;
    g.cand = d_cand.p, g.cand_first = d_cf.p, g.matched_out = want_mask ? d_mout.p : nullptr;
    plane_match_kernel<<<(n_map + WARPS - 1) / WARPS, WARPS * 32, 0, cs.s>>>(g);
    RS_LAUNCH_CHECK();
    plane_select_kernel<<<(n_frames + 127) / 128, 128, 0, cs.s>>>(g);
    RS_LAUNCH_CHECK();
    if (det_matched_out && n_det)
        RS_CUDA_CHECK(cudaMemcpyAsync(det_matched_out, d_mout.p, size_t(n_det), cudaMemcpyDeviceToHost, cs.s));
    RS_CUDA_CHECK(cudaMemcpyAsync(selected, d_sel.p, sizeof(int32_t) * n_map, cudaMemcpyDeviceToHost, cs.s));
    RS_CUDA_CHECK(cudaMemcpyAsync(inter_area, d_inter.p, sizeof(double) * n_map, cudaMemcpyDeviceToHost, cs.s));
    RS_CUDA_CHECK(cudaStreamSynchronize(cs.s));
    return RS_OK;
}

int rs_polygon_inter_area(int device, int n_pairs, const double* a_xy, const int32_t* a_first, const double* b_xy,
                          const int32_t* b_first, double* area)
{
    int rc = require_blackwell(device);
    if (rc != RS_OK) return rc;
    if (n_pairs < 0 || !a_first || !b_first || !area || !a_xy || !b_xy) {
        set_last_error("rs_polygon_inter_area: invalid argument");
        return RS_ERR_INVALID_ARG;
    }
    if (n_pairs == 0) return RS_OK;
    for (int p = 0; p < n_pairs; ++p)
        if (a_first[p] < 0 || b_first[p] < 0 || a_first[p + 1] < a_first[p] || b_first[p + 1] < b_first[p]) {
            set_last_error("rs_polygon_inter_area: a_first / b_first must be non-decreasing offsets");
            return RS_ERR_INVALID_ARG;
        }
    CallStream cs;
    if ((rc = cs.open(device))) return rc;
    DevBuf<double> d_a, d_b, d_area;
    DevBuf<int32_t> d_af, d_bf;
    if ((rc = d_a.upload(a_xy, 2 * size_t(a_first[n_pairs]), cs.s))) return rc;
    if ((rc = d_b.upload(b_xy, 2 * size_t(b_first[n_pairs]), cs.s))) return rc;
    if ((rc = d_af.upload(a_first, size_t(n_pairs) + 1, cs.s))) return rc;
    if ((rc = d_bf.upload(b_first, size_t(n_pairs) + 1, cs.s))) return rc;
    if ((rc = d_area.upload(nullptr, n_pairs, cs.s))) return rc;
    polygon_inter_area_kernel<<<(n_pairs + WARPS - 1) / WARPS, WARPS * 32, 0, cs.s>>>(n_pairs, d_a.p, d_af.p, d_b.p, d_bf.p, d_area.p);
    RS_LAUNCH_CHECK();
    RS_CUDA_CHECK(cudaMemcpyAsync(area, d_area.p, sizeof(double) * n_pairs, cudaMemcpyDeviceToHost, cs.s));
    RS_CUDA_CHECK(cudaStreamSynchronize(cs.s));
    return RS_OK;
}

}  // extern "C"
