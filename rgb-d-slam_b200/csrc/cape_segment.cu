// K2-K4 `cape_segment` — one CTA per frame over the cell graph produced by K1.
//
// Replaces  Primitive_Detection::{init_histogram, grow_planes_and_cylinders, grow_plane_segment_at_seed,
//           region_growing, cylinder_fitting, merge_planes, get_connected_components_matrix,
//           add_planes_to_primitives(mask + boundary points), add_cylinders_to_primitives(mask test)}
//           (primitive_detection.cpp:239-776), Histogram<N> (histogram.hpp:35-113) and
//           Cylinder_Segment's ctor + run_ransac_loop (cylinder_segment.cpp:35-322).
//
// The seed loop is inherently sequential (each seed depends on the histogram state left by the previous one), so
// the parallel axes are: frames (one CTA each), cells inside every scan, and the BFS frontier of the region growing
// (the reference's recursive DFS computes a reachability set in a directed graph; a frontier BFS gives the same set).
// Every FP64 sum whose order is fixed by the reference (merging cell sums in index order, LLS sums, MSAC cost) is
// evaluated in that same order by one thread so that results are bit-identical to the restated reference; all other
// work is spread over the CTA. Compiled with -fmad=false.
#include "cape_internal.cuh"
#include "plane_fit.cuh"

namespace rs {

namespace {

constexpr int T = 256;
constexpr int NW = T / 32;

struct Scalars {
    int best_bin, cand_count, seed, changed;
    int untried, n_planes, n_cyl_regions, n_cylinders, n_seeds, n_boundary, status;
    int cnt, flag, uniform_cursor;
    double dscr[NW];
    int iscr[NW];
    int iscr2[NW];
    PlaneModel work;       // plane being grown / refit
    PlaneModel work2;
    // cylinder scalars
    double axis[3], radius, inv_r2, center[3];
    double dval[8];
    int nleft, nids, best_count, cand_cnt_tmp;
};

struct Smem {
    double* cn;    // [Nc][3] cell normals
    double* cd;    // [Nc]
    double* cc;    // [Nc][3] centroids
    double* cmse;  // [Nc]
    double* pn;    // [Nc][3] projected normals (cylinder, local ids)
    double* pc;    // [Nc][3] projected centroids
    double* val;   // [Nc] scratch values
    float* tol;    // [Nc]
    int* bins;     // [Nc]
    int* hist;     // [cs*cs]
    int* list;     // [Nc] ordered activated cells (local -> global)
    int* ids;      // [Nc] remaining local ids (cylinder RANSAC)
    int* plab;     // [Nc] final merged labels
    short* gplane; // [Nc] _gridPlaneSegmentMap
    short* gcyl;   // [Nc] _gridCylinderSegMap
    unsigned char *unassigned, *activated, *planar, *mleft, *inlA, *inlB, *m0, *m1;
    // planes (compact copy for predicates; sums live in the global record)
    double* pln;   // [RS_MAX_PLANES][3]
    double* plc;   // [RS_MAX_PLANES][3]
    double* pld;   // [RS_MAX_PLANES]
    int* plabel;   // [RS_MAX_PLANES] merge labels
    int* pplanar;  // [RS_MAX_PLANES]
    unsigned* conn; // [RS_MAX_PLANES][RS_MAX_PLANES/32]
    Scalars* sc;
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ __device__ inline size_t carve(Smem* s, unsigned char* base, int Nc, int nbins)
{
    size_t o = 0;
    auto take = [&](size_t bytes) {
        unsigned char* p = base ? base + o : nullptr;
        o = align_up(o + bytes, 16);
        return p;
    };
    double* cn = reinterpret_cast<double*>(take(sizeof(double) * 3 * Nc));
    double* cd = reinterpret_cast<double*>(take(sizeof(double) * Nc));
    double* cc = reinterpret_cast<double*>(take(sizeof(double) * 3 * Nc));
    double* cmse = reinterpret_cast<double*>(take(sizeof(double) * Nc));
    double* pn = reinterpret_cast<double*>(take(sizeof(double) * 3 * Nc));
    double* pc = reinterpret_cast<double*>(take(sizeof(double) * 3 * Nc));
    double* val = reinterpret_cast<double*>(take(sizeof(double) * Nc));
    double* pln = reinterpret_cast<double*>(take(sizeof(double) * 3 * RS_MAX_PLANES));
    double* plc = reinterpret_cast<double*>(take(sizeof(double) * 3 * RS_MAX_PLANES));
    double* pld = reinterpret_cast<double*>(take(sizeof(double) * RS_MAX_PLANES));
    Scalars* sc = reinterpret_cast<Scalars*>(take(sizeof(Scalars)));
    float* tol = reinterpret_cast<float*>(take(sizeof(float) * Nc));
    int* bins = reinterpret_cast<int*>(take(sizeof(int) * Nc));
    int* hist = reinterpret_cast<int*>(take(sizeof(int) * nbins));
    int* list = reinterpret_cast<int*>(take(sizeof(int) * Nc));
    int* ids = reinterpret_cast<int*>(take(sizeof(int) * Nc));
    int* plab = reinterpret_cast<int*>(take(sizeof(int) * Nc));
    int* plabel = reinterpret_cast<int*>(take(sizeof(int) * RS_MAX_PLANES));
    int* pplanar = reinterpret_cast<int*>(take(sizeof(int) * RS_MAX_PLANES));
    unsigned* conn = reinterpret_cast<unsigned*>(take(sizeof(unsigned) * RS_MAX_PLANES * (RS_MAX_PLANES / 32)));
    short* gplane = reinterpret_cast<short*>(take(sizeof(short) * Nc));
    short* gcyl = reinterpret_cast<short*>(take(sizeof(short) * Nc));
    unsigned char* u8[8];
    for (int i = 0; i < 8; ++i) u8[i] = take(size_t(Nc));
    if (s) {
        s->cn = cn, s->cd = cd, s->cc = cc, s->cmse = cmse, s->pn = pn, s->pc = pc, s->val = val;
        s->pln = pln, s->plc = plc, s->pld = pld, s->sc = sc, s->tol = tol, s->bins = bins, s->hist = hist;
        s->list = list, s->ids = ids, s->plab = plab, s->plabel = plabel, s->pplanar = pplanar, s->conn = conn;
        s->gplane = gplane, s->gcyl = gcyl;
        s->unassigned = u8[0], s->activated = u8[1], s->planar = u8[2], s->mleft = u8[3];
        s->inlA = u8[4], s->inlB = u8[5], s->m0 = u8[6], s->m1 = u8[7];
    }
    return o;
}

// ---- block-wide helpers (all threads must call) ----------------------------------------------------
__device__ __forceinline__ int block_sum_int(int v, int* scr)
{
    v = __reduce_add_sync(0xffffffffu, v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) scr[w] = v;
    __syncthreads();
    int r = 0;
#pragma unroll
    for (int i = 0; i < NW; ++i) r += scr[i];
    return r;
}

// exclusive prefix (in index order) of a 0/1 flag over the CTA; returns the CTA total through *total
__device__ __forceinline__ int block_scan_flag(bool flag, int* scr, int* total)
{
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) scr[w] = __popc(m);
    __syncthreads();
    int before = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        if (i < w) before += scr[i];
        tot += scr[i];
    }
    *total = tot;
    return before + __popc(m & ((1u << l) - 1u));
}

// lexicographic minimum of (key, idx) over the CTA
__device__ __forceinline__ void block_min_key(double& key, int& idx, double* dscr, int* iscr)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ok = __shfl_xor_sync(0xffffffffu, key, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ok < key || (ok == key && oi < idx)) key = ok, idx = oi;
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) dscr[w] = key, iscr[w] = idx;
    __syncthreads();
    key = dscr[0], idx = iscr[0];
#pragma unroll
    for (int i = 1; i < NW; ++i)
        if (dscr[i] < key || (dscr[i] == key && iscr[i] < idx)) key = dscr[i], idx = iscr[i];
}

// Histogram::remove_point (histogram.hpp:103-113) — quirk: the bin becomes 1, not -1.
__device__ __forceinline__ void hist_remove_atomic(const Smem& s, int cell, int nbins)
{
    const int b = s.bins[cell];
    if (b >= 0 && b < nbins) {
        const int old = atomicSub(&s.hist[b], 1);
        if (old <= 0) atomicAdd(&s.hist[b], 1);
    }
    s.bins[cell] = 1;
}

__device__ __forceinline__ void store_plane_compact(const Smem& s, int k, const PlaneModel& p)
{
    for (int i = 0; i < 3; ++i) s.pln[3 * k + i] = p.n[i], s.plc[3 * k + i] = p.c[i];
    s.pld[k] = p.d;
    s.pplanar[k] = p.planar;
}

__device__ __forceinline__ void store_plane_record(rs_plane_out& o, const PlaneModel& p)
{
    o.count = p.count;
    o.planar = p.planar;
    for (int i = 0; i < 9; ++i) o.S[i] = p.S[i];
    for (int i = 0; i < 3; ++i) o.centroid[i] = p.c[i], o.normal[i] = p.n[i];
    o.d = p.d;
    o.mse = p.mse;
    o.score = p.score;
}

// new plane segment := copy-construct (re-normalises the normal, plane_segment.cpp:18-36) and push_back.
// Must be called by thread 0 only. Returns the 1-based plane id or 0 when the capacity is exhausted.
__device__ int push_plane(const Smem& s, rs_plane_out* planes, const PlaneModel& src)
{
    Scalars& sc = *s.sc;
    if (sc.n_planes >= RS_MAX_PLANES) {
        sc.status = RS_ERR_CAPACITY;
        return 0;
    }
    PlaneModel p = src;
    normalize3(p.n);
    const int k = sc.n_planes++;
    store_plane_compact(s, k, p);
    store_plane_record(planes[k], p);
    s.plabel[k] = k;
    return k + 1;
}

// S/count of `dst` += sums of the listed cells, in list order (threads 0..9; one running sum each).
__device__ __forceinline__ void ordered_expand(PlaneModel& dst, const rs_cell_out* cells, const int* list, int n,
                                               const unsigned char* filter)
{
    const int t = threadIdx.x;
    if (t < 9) {
        double acc = dst.S[t];
        for (int j = 0; j < n; ++j)
            if (!filter || filter[j]) acc += cells[list[j]].S[t];
        dst.S[t] = acc;
    }
    else if (t == 9) {
        int acc = dst.count;
        for (int j = 0; j < n; ++j)
            if (!filter || filter[j]) acc += cells[list[j]].count;
        dst.count = acc;
    }
}

// cv::erode / cv::dilate on the cell grid, 3x3 kernels, one output cell.
__device__ __forceinline__ unsigned char morph_at(const unsigned char* m, int row, int col, int vc, int hc, bool erode,
                                                  bool cross, bool borderZero)
{
    unsigned char v = erode ? 255 : 0;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            if (cross && dx != 0 && dy != 0) continue;
            const int yy = row + dy, xx = col + dx;
            unsigned char nv;
            if (yy < 0 || yy >= vc || xx < 0 || xx >= hc) {
                if (!borderZero) continue;
                nv = 0;
            }
            else
                nv = m[yy * hc + xx];
            v = erode ? (nv < v ? nv : v) : (nv > v ? nv : v);
        }
    return v;
}

// ---- cylinder branch (cylinder_segment.cpp:35-322 + primitive_detection.cpp:413-501) ------------
__device__ void cylinder_fitting(const Smem& s, const SegmentParams& prm, const rs_cell_out* cells, rs_plane_out* planes,
                                 rs_cyl_out* cyls, int32_t* cyl_region_seg, const int m, const int Nc)
{
    Scalars& sc = *s.sc;
    const int t = threadIdx.x;
    if (sc.n_cyl_regions >= RS_MAX_CYL_REGIONS) {
        if (t == 0) sc.status = RS_ERR_CAPACITY;
        return;
    }
    const int region = sc.n_cyl_regions;
    rs_cyl_out& co = cyls[region];
    __syncthreads();

    // covariance of [N -N] (3 x 2m), summed column by column: six unique entries, one thread each
    if (t < 6) {
        const int r = (t < 3) ? 0 : (t < 5 ? 1 : 2);
        const int c = (t < 3) ? t : (t < 5 ? t - 2 : 2);
        double acc = 0;
        for (int j = 0; j < m; ++j) acc += s.cn[3 * s.list[j] + r] * s.cn[3 * s.list[j] + c];
        for (int j = 0; j < m; ++j) acc += (-s.cn[3 * s.list[j] + r]) * (-s.cn[3 * s.list[j] + c]);
        sc.dval[t] = acc / static_cast<double>(2 * m - 1);
    }
    __syncthreads();
    if (t == 0) {
        // entries: 0:(0,0) 1:(0,1) 2:(0,2) 3:(1,1) 4:(1,2) 5:(2,2); lower triangle a10=(0,1), a20=(0,2), a21=(1,2)
        double ev[3], q[3][3];
        self_adjoint_eigen3(sc.dval[0], sc.dval[1], sc.dval[3], sc.dval[2], sc.dval[4], sc.dval[5], ev, q);
        const double score = ev[2] / ev[0];
        co.n_cells = m;
        co.n_segments = 0;
        co.pca_score = score;
        sc.axis[0] = q[0][0], sc.axis[1] = q[1][0], sc.axis[2] = q[2][0];
        sc.flag = !(score < 75.0);  // cylinderRansacMinimumScore (float 75)
        if (sc.flag)
            for (int i = 0; i < 3; ++i) co.axis[i] = sc.axis[i];
        sc.n_cyl_regions = region + 1;
    }
    __syncthreads();
    if (!sc.flag) return;

    const double ax = sc.axis[0], ay = sc.axis[1], az = sc.axis[2];
    for (int j = t; j < m; j += T) {
        const int gi = s.list[j];
        const double c0 = s.cc[3 * gi], c1 = s.cc[3 * gi + 1], c2 = s.cc[3 * gi + 2];
        const double cdot = (ax * c0 + ay * c1) + az * c2;
        s.pc[3 * j] = c0 - cdot * ax;
        s.pc[3 * j + 1] = c1 - cdot * ay;
        s.pc[3 * j + 2] = c2 - cdot * az;
        const double n0 = s.cn[3 * gi], n1 = s.cn[3 * gi + 1], n2 = s.cn[3 * gi + 2];
        const double ndot = (ax * n0 + ay * n1) + az * n2;
        const double p0 = n0 - ndot * ax, p1 = n1 - ndot * ay, p2 = n2 - ndot * az;
        const double nn = sqrt((p0 * p0 + p1 * p1) + p2 * p2);
        s.pn[3 * j] = p0 / nn;
        s.pn[3 * j + 1] = p1 / nn;
        s.pn[3 * j + 2] = p2 / nn;
        s.mleft[j] = 1;
        s.ids[j] = j;
    }
    if (t == 0) {
        sc.nleft = m;
        sc.nids = m;
    }
    __syncthreads();

    const unsigned minimumCellActivated = static_cast<unsigned>(0.65 / 100.0 * static_cast<double>(Nc));
    const float maxSqrtDist = 0.04f;
    const double maxSqrtDistD = static_cast<double>(maxSqrtDist);
    int segId = 0;
    while (static_cast<unsigned>(sc.nleft) > minimumCellActivated && sc.nleft > 0.1 * m) {
        // ---- run_ransac_loop ----
        const int nIds = sc.nids;
        unsigned char* best = s.inlA;
        unsigned char* cand = s.inlB;
        int bestCount = 0;
        if (nIds >= 3) {
            const unsigned accepted = static_cast<unsigned>(floor(0.9 * nIds));
            double minHypothesisDist = static_cast<double>(maxSqrtDist * static_cast<float>(nIds));
            for (int j = t; j < m; j += T) best[j] = 0;
            for (int it = 0; it < RS_CYL_RANSAC_ITERS; ++it) {
                if (t == 0) {
                    int id[3];
                    for (int k = 0; k < 3; ++k) {
                        double u = 0.0;
                        if (sc.uniform_cursor < prm.n_uniforms)
                            u = prm.uniforms[sc.uniform_cursor++];
                        else
                            sc.status = RS_ERR_CAPACITY;
                        id[k] = s.ids[static_cast<unsigned>(floor(u * static_cast<double>(static_cast<unsigned>(nIds))))];
                    }
                    const double* n1 = s.pn + 3 * id[0];
                    const double* n2 = s.pn + 3 * id[1];
                    const double* n3 = s.pn + 3 * id[2];
                    const double* c1 = s.pc + 3 * id[0];
                    const double* c2 = s.pc + 3 * id[1];
                    const double* c3 = s.pc + 3 * id[2];
                    double sn[3], sm[3], tt[3];
                    for (int k = 0; k < 3; ++k) {
                        sn[k] = (n1[k] + n2[k]) + n3[k];
                        sm[k] = (c1[k] + c2[k]) + c3[k];
                        tt[k] = (n1[k] * c1[k] + n2[k] * c2[k]) + n3[k] * c3[k];
                    }
                    const double a = 1.0 - dot3(sn, sn) / 9.0;
                    const double b = ((tt[0] + tt[1]) + tt[2]) / 3.0 - (dot3(sn, sm) / 9.0);
                    const double radius = b / a;
                    sc.radius = radius;
                    sc.inv_r2 = 1.0 / (radius * radius);
                    for (int k = 0; k < 3; ++k) sc.center[k] = (sm[k] - radius * sn[k]) / 3.0;
                }
                __syncthreads();
                const double radius = sc.radius, invr2 = sc.inv_r2, e0 = sc.center[0], e1 = sc.center[1], e2 = sc.center[2];
                for (int j = t; j < m; j += T) {
                    double v = -1.0;  // not part of this RANSAC round
                    unsigned char in = 0;
                    if (s.mleft[j]) {
                        const double d0 = (s.pc[3 * j] - radius * s.pn[3 * j]) - e0;
                        const double d1 = (s.pc[3 * j + 1] - radius * s.pn[3 * j + 1]) - e1;
                        const double d2 = (s.pc[3 * j + 2] - radius * s.pn[3 * j + 2]) - e2;
                        const double distance = ((d0 * d0 + d1 * d1) + d2 * d2) * invr2;
                        if (distance < maxSqrtDistD) {
                            v = distance;
                            in = 1;
                        }
                        else
                            v = maxSqrtDistD;
                    }
                    s.val[j] = v;
                    cand[j] = in;
                }
                __syncthreads();
                if (t == 0) {
                    // MSAC cost in cell order, exactly as the reference accumulates it
                    double dist = 0.0;
                    int cnt = 0;
                    for (int j = 0; j < m; ++j) {
                        const double v = s.val[j];
                        if (v >= 0.0) dist += v;
                        cnt += cand[j];
                    }
                    sc.flag = 0;
                    sc.cnt = 0;
                    if (dist < minHypothesisDist) {
                        sc.dval[0] = dist;
                        sc.flag = 1;
                        sc.cand_cnt_tmp = cnt;
                    }
                }
                __syncthreads();
                bool stop = false;
                if (sc.flag) {
                    minHypothesisDist = sc.dval[0];
                    unsigned char* tmp = best;
                    best = cand;
                    cand = tmp;
                    const int prevCount = bestCount;
                    bestCount = sc.cand_cnt_tmp;
                    // quirk: the early stop looks at the PREVIOUS best set (vectors swapped before the test)
                    if (static_cast<unsigned>(prevCount) > accepted) stop = true;
                }
                __syncthreads();
                if (stop) break;
            }
        }
        if (bestCount < 6) break;
        if (segId >= RS_MAX_CYL_SEGS) {
            if (t == 0) sc.status = RS_ERR_CAPACITY;
            break;
        }

        // ---- LLS over the inliers, ordered sums (7 running sums) ----
        if (t < 7) {
            double acc = 0.0;
            for (int j = 0; j < m; ++j)
                if (best[j]) {
                    if (t < 3)
                        acc += s.pn[3 * j + t];
                    else if (t < 6)
                        acc += s.pc[3 * j + (t - 3)];
                    else
                        acc += (s.pn[3 * j] * s.pc[3 * j] + s.pn[3 * j + 1] * s.pc[3 * j + 1]) + s.pn[3 * j + 2] * s.pc[3 * j + 2];
                }
            sc.dval[t] = acc;
        }
        __syncthreads();
        if (t == 0) {
            // rebuild the remaining id list in index order, drop the inliers from the mask
            int nl = sc.nleft, k = 0;
            for (int j = 0; j < m; ++j) {
                if (best[j]) {
                    s.mleft[j] = 0;
                    nl--;
                }
                else if (s.mleft[j])
                    s.ids[k++] = j;
            }
            sc.nleft = nl;
            sc.nids = k;
            const double cntd = static_cast<double>(static_cast<size_t>(bestCount));
            const double inv2 = 1.0 / static_cast<double>(static_cast<size_t>(bestCount) * static_cast<size_t>(bestCount));
            const double* sn = sc.dval;
            const double* sm = sc.dval + 3;
            const double a = 1 - dot3(sn, sn) * inv2;
            double b = sc.dval[6];
            b /= cntd;
            b -= dot3(sn, sm) * inv2;
            double radius = b / a;
            for (int k2 = 0; k2 < 3; ++k2) sc.center[k2] = (sm[k2] - radius * sn[k2]) / cntd;
            if (radius < 0) radius = -radius;
            sc.radius = radius;
        }
        __syncthreads();
        {
            // per-inlier squared point-to-axis distance error, using the UNPROJECTED centroids (:199-218)
            const double P1[3] = {sc.center[0], sc.center[1], sc.center[2]};
            const double P2[3] = {P1[0] + sc.axis[0], P1[1] + sc.axis[1], P1[2] + sc.axis[2]};
            const double D[3] = {P2[0] - P1[0], P2[1] - P1[1], P2[2] - P1[2]};
            const double P1P2 = sqrt(dot3(D, D));
            const double radius = sc.radius;
            for (int j = t; j < m; j += T) {
                double v = 0.0;
                if (best[j]) {
                    const int gi = s.list[j];
                    const double E[3] = {s.cc[3 * gi] - P2[0], s.cc[3 * gi + 1] - P2[1], s.cc[3 * gi + 2] - P2[2]};
                    const double X[3] = {D[1] * E[2] - D[2] * E[1], D[2] * E[0] - D[0] * E[2], D[0] * E[1] - D[1] * E[0]};
                    const double tdist = sqrt(dot3(X, X)) / P1P2 - radius;
                    v = tdist * tdist;
                    cyl_region_seg[gi] = 1 + region * RS_MAX_CYL_SEGS + segId;
                }
                s.val[j] = v;
            }
        }
        // plane refit over the segment's inlier cells (find_plane_segment_in_cylinder)
        if (t == 0) plane_clear(sc.work2);
        __syncthreads();
        ordered_expand(sc.work2, cells, s.list, m, best);
        __syncthreads();
        if (t == 0) {
            double mse = 0.0;
            for (int j = 0; j < m; ++j)
                if (best[j]) mse += s.val[j];
            mse /= static_cast<double>(static_cast<size_t>(bestCount));
            co.radius[segId] = sc.radius;
            for (int k = 0; k < 3; ++k) co.center[segId][k] = sc.center[k];
            co.mse[segId] = mse;
            co.n_inliers[segId] = bestCount;
            co.n_segments = segId + 1;

            plane_fit(sc.work2);
            co.plane_mse[segId] = sc.work2.mse;
            int assigned = 0;
            if (sc.work2.mse < mse) {
                const int id = push_plane(s, planes, sc.work2);
                assigned = -id;
                sc.flag = id;   // > 0: label inliers as plane `id`
            }
            else {
                const int id = ++sc.n_cylinders;
                assigned = id;
                sc.flag = -id;  // < 0: label inliers as cylinder `id`
            }
            co.assigned[segId] = assigned;
            co.kept[segId] = 0;
        }
        __syncthreads();
        {
            const int f = sc.flag;
            for (int j = t; j < m; j += T)
                if (best[j]) {
                    if (f > 0)
                        s.gplane[s.list[j]] = static_cast<short>(f);
                    else if (f < 0)
                        s.gcyl[s.list[j]] = static_cast<short>(-f);
                }
        }
        __syncthreads();
        ++segId;
    }
}

__global__ void __launch_bounds__(T, 1) cape_segment_kernel(const SegmentParams prm, const SegmentBuffers buf)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Nc = prm.hc * prm.vc, hc = prm.hc, vc = prm.vc, cs = prm.cell;
    const int nbins = cs * cs;
    Smem s;
    carve(&s, smem_raw, Nc, nbins);
    Scalars& sc = *s.sc;
    const int t = threadIdx.x;
    const int frame = blockIdx.x;
    const rs_cell_out* cells = buf.cells + size_t(frame) * Nc;
    rs_plane_out* planes = buf.planes + size_t(frame) * RS_MAX_PLANES;
    rs_cyl_out* cyls = buf.cyls + size_t(frame) * RS_MAX_CYL_REGIONS;
    int32_t* out_region_seg = buf.cyl_region_seg + size_t(frame) * Nc;
    const float* depth = buf.depth + size_t(frame) * prm.W * prm.H;
    double* boundary = buf.boundary_xyz + size_t(frame) * prm.max_boundary * 3;

    // ---- load the cell graph, init_histogram (primitive_detection.cpp:239-265, histogram.hpp:35-62) ----
    for (int i = t; i < nbins; i += T) s.hist[i] = 0;
    if (t == 0) {
        sc.n_planes = 0, sc.n_cyl_regions = 0, sc.n_cylinders = 0, sc.n_seeds = 0, sc.n_boundary = 0;
        sc.status = RS_OK, sc.uniform_cursor = 0;
    }
    __syncthreads();
    int myPlanar = 0;
    for (int i = t; i < Nc; i += T) {
        const rs_cell_out& c = cells[i];
        s.cn[3 * i] = c.normal[0], s.cn[3 * i + 1] = c.normal[1], s.cn[3 * i + 2] = c.normal[2];
        s.cc[3 * i] = c.centroid[0], s.cc[3 * i + 1] = c.centroid[1], s.cc[3 * i + 2] = c.centroid[2];
        s.cd[i] = c.d;
        s.cmse[i] = c.mse;
        s.tol[i] = c.tol;
        s.planar[i] = c.planar ? 1 : 0;
        s.unassigned[i] = c.planar ? 1 : 0;
        s.gplane[i] = 0;
        s.gcyl[i] = 0;
        out_region_seg[i] = 0;
        int bin = -1;
        if (c.planar) {
            ++myPlanar;
            const double theta = acos(-c.normal[2]);
            const double phi = atan2(c.normal[0], c.normal[1]);
            const int xQ = static_cast<int>(floor((cs - 1) * (theta - 0.0) / (kPi - 0.0)));
            int yQ = 0;
            if (xQ > 0) yQ = static_cast<int>(floor((cs - 1) * (phi - (-kPi)) / (kPi - (-kPi))));
            bin = yQ * cs + xQ;
            if (bin >= 0 && bin < nbins) atomicAdd(&s.hist[bin], 1);
        }
        s.bins[i] = bin;
    }
    const int nPlanar = block_sum_int(myPlanar, sc.iscr);
    if (t == 0) sc.untried = nPlanar;
    __syncthreads();

    const unsigned planeSeedCount = static_cast<unsigned>(0.8 / 100.0 * Nc);
    const unsigned minimumCellActivated = static_cast<unsigned>(0.65 / 100.0 * Nc);

    // ---- grow_planes_and_cylinders (primitive_detection.cpp:267-310) ----
    int guard = 0;
    while (sc.untried > 0) {
        // most frequent bin: strictly greatest count, lowest index on ties (histogram.hpp:69-84)
        {
            double key = 1.0;  // -count as key so that the minimum is the fullest bin
            int idx = 0x7fffffff;
            for (int i = t; i < nbins; i += T) {
                const int h = s.hist[i];
                if (h > 0) {
                    const double k = -static_cast<double>(h);
                    if (k < key || (k == key && i < idx)) key = k, idx = i;
                }
            }
            block_min_key(key, idx, sc.dscr, sc.iscr);
            if (t == 0) sc.best_bin = (key < 0.0) ? idx : -1;
        }
        __syncthreads();
        const int bestBin = sc.best_bin;
        // candidates of that bin, min-MSE seed (first strictly smallest; stop at MSE <= 0) (:286-298)
        {
            double key = DBL_MAX;
            int idx = 0x7fffffff;
            int cnt = 0;
            if (bestBin >= 0) {
                for (int i = t; i < Nc; i += T)
                    if (s.bins[i] == bestBin) {
                        ++cnt;
                        const double mse = s.cmse[i];
                        const double k = (mse <= 0.0) ? 0.0 : mse;
                        if (k < key || (k == key && i < idx)) key = k, idx = i;
                    }
            }
            cnt = block_sum_int(cnt, sc.iscr2);
            block_min_key(key, idx, sc.dscr, sc.iscr);
            if (t == 0) {
                sc.cand_count = cnt;
                sc.seed = (key < DBL_MAX) ? idx : -1;
            }
        }
        __syncthreads();
        if (static_cast<unsigned>(sc.cand_count) < planeSeedCount) break;
        if (sc.seed < 0) break;
        const int seed = sc.seed;
        if (t == 0) sc.n_seeds++;

        // ---- grow_plane_segment_at_seed (:312-389) ----
        // A non-planar seed changes no state in the reference (:318-322), which would then spin forever; it is
        // unreachable (only planar cells carry a bin id). Leave the loop instead of hanging the GPU.
        if (!s.planar[seed] || ++guard > 4 * Nc + nbins) {
            if (t == 0) sc.status = RS_ERR_CAPACITY;
            break;
        }
        for (int i = t; i < Nc; i += T) s.activated[i] = 0;
        if (t == 0) {
            // newPlaneSegment = copy of the seed (normal re-normalised by the copy ctor)
            PlaneModel& w = sc.work;
            const rs_cell_out& c = cells[seed];
            w.count = c.count;
            w.planar = c.planar;
            for (int k = 0; k < 9; ++k) w.S[k] = c.S[k];
            for (int k = 0; k < 3; ++k) w.c[k] = c.centroid[k], w.n[k] = c.normal[k];
            w.d = c.d, w.mse = c.mse, w.score = c.score;
            normalize3(w.n);
            // region_growing's first test: new segment -> seed cell (:795-803)
            sc.flag = 0;
            if (s.unassigned[seed] &&
                plane_can_merge(w.n, w.d, s.cn + 3 * seed, s.cc + 3 * seed, static_cast<double>(s.tol[seed]), prm.cos_merge))
                sc.flag = 1;
        }
        __syncthreads();
        if (sc.flag) {
            if (t == 0) s.activated[seed] = 1;
            __syncthreads();
            // frontier BFS: a cell joins when some activated 4-neighbour u satisfies can_be_merged(u, cell, tol[cell])
            while (true) {
                int changed = 0;
                for (int i = t; i < Nc; i += T) {
                    if (!s.unassigned[i] || s.activated[i]) continue;
                    const int y = i / hc, x = i - y * hc;
                    const double tolI = static_cast<double>(s.tol[i]);
                    bool join = false;
                    if (x > 0 && s.activated[i - 1] &&
                        plane_can_merge(s.cn + 3 * (i - 1), s.cd[i - 1], s.cn + 3 * i, s.cc + 3 * i, tolI, prm.cos_merge))
                        join = true;
                    if (!join && x < hc - 1 && s.activated[i + 1] &&
                        plane_can_merge(s.cn + 3 * (i + 1), s.cd[i + 1], s.cn + 3 * i, s.cc + 3 * i, tolI, prm.cos_merge))
                        join = true;
                    if (!join && y > 0 && s.activated[i - hc] &&
                        plane_can_merge(s.cn + 3 * (i - hc), s.cd[i - hc], s.cn + 3 * i, s.cc + 3 * i, tolI, prm.cos_merge))
                        join = true;
                    if (!join && y < vc - 1 && s.activated[i + hc] &&
                        plane_can_merge(s.cn + 3 * (i + hc), s.cd[i + hc], s.cn + 3 * i, s.cc + 3 * i, tolI, prm.cos_merge))
                        join = true;
                    if (join) {
                        s.m0[i] = 1;
                        changed = 1;
                    }
                    else
                        s.m0[i] = 0;
                }
                changed = __syncthreads_or(changed);
                if (!changed) break;
                for (int i = t; i < Nc; i += T)
                    if (s.unassigned[i] && !s.activated[i] && s.m0[i]) s.activated[i] = 1;
                __syncthreads();
            }
        }
        // ordered list of the activated cells (index order)
        {
            int base = 0;
            for (int c0 = 0; c0 < Nc; c0 += T) {
                const int i = c0 + t;
                const bool f = (i < Nc) && s.activated[i];
                int tot;
                const int pos = block_scan_flag(f, sc.iscr, &tot);
                if (f) s.list[base + pos] = i;
                base += tot;
            }
            if (t == 0) sc.cnt = base;
        }
        __syncthreads();
        const int cnt = sc.cnt;
        // merge activated cells & remove them from the histogram (:343-360)
        ordered_expand(sc.work, cells, s.list, cnt, nullptr);
        for (int j = t; j < cnt; j += T) {
            const int i = s.list[j];
            hist_remove_atomic(s, i, nbins);
            s.unassigned[i] = 0;
        }
        __syncthreads();
        if (t == 0) sc.untried -= cnt;
        if (cnt == 0 || static_cast<unsigned>(cnt) < minimumCellActivated) {
            if (t == 0) {
                // _histogram.remove_point(seedId)
                const int b = s.bins[seed];
                if (b >= 0 && b < nbins && s.hist[b] != 0) s.hist[b] -= 1;
                s.bins[seed] = 1;
            }
            __syncthreads();
            continue;
        }
        if (t == 0) {
            plane_fit(sc.work);
            sc.flag = 0;
            if (sc.work.planar) {
                if (sc.work.score > 100)
                    sc.flag = push_plane(s, planes, sc.work);  // add_plane_segment_to_features (:391-411)
                else if (cnt > 5)
                    sc.flag = -1;
            }
        }
        __syncthreads();
        const int f = sc.flag;
        if (f > 0) {
            for (int j = t; j < cnt; j += T) s.gplane[s.list[j]] = static_cast<short>(f);
        }
        else if (f < 0) {
            cylinder_fitting(s, prm, cells, planes, cyls, out_region_seg, cnt, Nc);
        }
        __syncthreads();
    }
    __syncthreads();

    // ---- merge_planes (:503-560) with get_connected_components_matrix (:736-776) ----
    const int P = sc.n_planes;
    constexpr int CW = RS_MAX_PLANES / 32;
    for (int i = t; i < RS_MAX_PLANES * CW; i += T) s.conn[i] = 0;
    __syncthreads();
    for (int i = t; i < Nc; i += T) {
        const int row = i / hc, col = i - row * hc;
        if (row >= vc - 1 || col >= hc - 1) continue;  // last row / column are never scan origins
        const int id = s.gplane[i];
        if (id <= 0) continue;
        const int nx = s.gplane[i + 1], bl = s.gplane[i + hc];
        if (nx > 0 && id != nx) {
            atomicOr(&s.conn[(id - 1) * CW + ((nx - 1) >> 5)], 1u << ((nx - 1) & 31));
            atomicOr(&s.conn[(nx - 1) * CW + ((id - 1) >> 5)], 1u << ((id - 1) & 31));
        }
        if (bl > 0 && id != bl) {
            atomicOr(&s.conn[(id - 1) * CW + ((bl - 1) >> 5)], 1u << ((bl - 1) & 31));
            atomicOr(&s.conn[(bl - 1) * CW + ((id - 1) >> 5)], 1u << ((id - 1) & 31));
        }
    }
    __syncthreads();
    if (t == 0) {
        for (int row = 0; row < P; ++row) {
            bool expanded = false;
            const int planeId = s.plabel[row];
            if (!s.pplanar[planeId]) continue;
            rs_plane_out& target = planes[planeId];
            for (int col = row + 1; col < P; ++col) {
                if (!((s.conn[row * CW + (col >> 5)] >> (col & 31)) & 1u)) continue;
                if (!s.pplanar[col]) continue;
                if (plane_can_merge(s.pln + 3 * planeId, s.pld[planeId], s.pln + 3 * col, s.plc + 3 * col, 50.0, prm.cos_merge)) {
                    const rs_plane_out& src = planes[col];
                    for (int k = 0; k < 9; ++k) target.S[k] += src.S[k];
                    target.count += src.count;
                    s.plabel[col] = planeId;
                    expanded = true;
                }
                else {
                    s.conn[row * CW + (col >> 5)] &= ~(1u << (col & 31));
                    s.conn[col * CW + (row >> 5)] &= ~(1u << (row & 31));
                }
            }
            if (expanded) {
                PlaneModel& w = sc.work;
                w.count = target.count;
                for (int k = 0; k < 9; ++k) w.S[k] = target.S[k];
                for (int k = 0; k < 3; ++k) w.c[k] = target.centroid[k], w.n[k] = target.normal[k];
                w.d = target.d, w.mse = target.mse, w.score = target.score;
                plane_fit(w);
                store_plane_compact(s, planeId, w);
                store_plane_record(target, w);
            }
        }
    }
    __syncthreads();

    // ---- final labels + per-plane record tail ----
    for (int i = t; i < Nc; i += T) {
        const int id = s.gplane[i];
        int lab = 0;
        if (id > 0) {
            const int root = s.plabel[id - 1];
            if (s.plabel[root] == root && s.pplanar[root]) lab = root + 1;
        }
        s.plab[i] = lab;
    }
    __syncthreads();

    // ---- boundary points per final plane (compute_plane_segment_boundary, :650-703) ----
    const unsigned pixelPerCellSide = static_cast<unsigned>(sqrtf(static_cast<float>(cs * cs)));
    int nFinal = 0;
    for (int k = 0; k < P; ++k) {
        const bool isFinal = (s.plabel[k] == k) && s.pplanar[k];
        if (t == 0) {
            planes[k].merge_label = s.plabel[k];
            planes[k].is_final = isFinal ? 1 : 0;
            planes[k].n_boundary = 0;
            planes[k].boundary_offset = sc.n_boundary;
        }
        if (!isFinal) continue;
        ++nFinal;
        for (int i = t; i < Nc; i += T) s.m0[i] = (s.plab[i] == k + 1) ? 1 : 0;
        __syncthreads();
        const double maxBoundaryDistance = 3 * sqrt(planes[k].mse);
        const double n0 = s.pln[3 * k], n1 = s.pln[3 * k + 1], n2 = s.pln[3 * k + 2], dd = s.pld[k];
        int base = sc.n_boundary;
        const int start = base;
        for (int c0 = 0; c0 < Nc; c0 += T) {
            const int i = c0 + t;
            bool keep = false;
            double px = 0, py = 0, pz = 0;
            if (i < Nc) {
                const int row = i / hc, col = i - row * hc;
                const int er = morph_at(s.m0, row, col, vc, hc, true, true, true);
                const int di = morph_at(s.m0, row, col, vc, hc, false, false, false);
                if (di - er > 0) {
                    const int centerX = static_cast<int>(col * pixelPerCellSide + pixelPerCellSide / 2);
                    const int centerY = static_cast<int>(row * pixelPerCellSide + pixelPerCellSide / 2);
                    const double z = static_cast<double>(depth[size_t(centerY) * prm.W + centerX]);
                    if (z > 0) {
                        px = z * prm.kx[centerX];
                        py = z * prm.ky[centerY];
                        pz = z;
                        if (fabs(((n0 * px + n1 * py) + n2 * pz) + dd) < maxBoundaryDistance) keep = true;
                    }
                }
            }
            int tot;
            const int pos = block_scan_flag(keep, sc.iscr, &tot);
            if (keep) {
                const int o = base + pos;
                if (o < prm.max_boundary) {
                    boundary[3 * o] = px, boundary[3 * o + 1] = py, boundary[3 * o + 2] = pz;
                }
            }
            base += tot;
        }
        __syncthreads();
        if (t == 0) {
            if (base > prm.max_boundary) {
                sc.status = RS_ERR_CAPACITY;
                base = prm.max_boundary;
            }
            planes[k].n_boundary = base - start;
            sc.n_boundary = base;
        }
        __syncthreads();
    }

    // ---- cylinders: opening test of add_cylinders_to_primitives (:705-734) ----
    for (int ci = 1; ci <= sc.n_cylinders; ++ci) {
        for (int i = t; i < Nc; i += T) s.m0[i] = (s.gcyl[i] == ci) ? 1 : 0;
        __syncthreads();
        for (int i = t; i < Nc; i += T) s.m1[i] = morph_at(s.m0, i / hc, i % hc, vc, hc, false, true, false);
        __syncthreads();
        for (int i = t; i < Nc; i += T) s.m0[i] = morph_at(s.m1, i / hc, i % hc, vc, hc, true, true, false);
        __syncthreads();
        int mn = 255, mx = 0;
        for (int i = t; i < Nc; i += T) {
            const int v = morph_at(s.m0, i / hc, i % hc, vc, hc, true, true, false);
            mn = min(mn, v);
            mx = max(mx, v);
        }
        mn = __reduce_min_sync(0xffffffffu, mn);
        mx = __reduce_max_sync(0xffffffffu, mx);
        __syncthreads();
        if ((t & 31) == 0) sc.iscr[t >> 5] = mn, sc.iscr2[t >> 5] = mx;
        __syncthreads();
        if (t == 0) {
            for (int w = 1; w < NW; ++w) mn = min(mn, sc.iscr[w]), mx = max(mx, sc.iscr2[w]);
            const int kept = !(mx <= 0 || mn >= mx);
            for (int r = 0; r < sc.n_cyl_regions; ++r)
                for (int sg = 0; sg < cyls[r].n_segments; ++sg)
                    if (cyls[r].assigned[sg] == ci) cyls[r].kept[sg] = kept;
        }
        __syncthreads();
    }

    // ---- write the label grids and the frame info ----
    int32_t* o_grid = buf.plane_grid + size_t(frame) * Nc;
    int32_t* o_lab = buf.plane_labels + size_t(frame) * Nc;
    int32_t* o_cyl = buf.cyl_labels + size_t(frame) * Nc;
    for (int i = t; i < Nc; i += T) {
        o_grid[i] = s.gplane[i];
        o_lab[i] = s.plab[i];
        o_cyl[i] = s.gcyl[i];
    }
    if (t == 0) {
        rs_cape_frame_info& info = buf.info[frame];
        info.status = sc.status;
        info.n_planar_cells = nPlanar;
        info.n_seeds = sc.n_seeds;
        info.n_planes = P;
        info.n_final_planes = nFinal;
        info.n_cyl_regions = sc.n_cyl_regions;
        info.n_cylinders = sc.n_cylinders;
        info.n_boundary = sc.n_boundary;
    }
}

}  // namespace

int launch_cape_segment(const SegmentParams& prm, const SegmentBuffers& buf, cudaStream_t stream)
{
    const int Nc = prm.hc * prm.vc;
    const size_t smem = carve(nullptr, nullptr, Nc, prm.cell * prm.cell);
    if (smem > 227 * 1024) {
        set_last_error("cape_segment: the cell grid does not fit in shared memory (" + std::to_string(smem) + " B > 227 KB)");
        return RS_ERR_INVALID_ARG;
    }
    if (Nc > 32767) {
        set_last_error("cape_segment: too many cells");
        return RS_ERR_INVALID_ARG;
    }
    static size_t configured = 0;
    if (smem > configured) {
        RS_CUDA_CHECK(cudaFuncSetAttribute(cape_segment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        configured = smem;
    }
    // the plane/cylinder record arrays are zeroed so that unused entries read as empty
    RS_CUDA_CHECK(cudaMemsetAsync(buf.planes, 0, sizeof(rs_plane_out) * size_t(prm.batch) * RS_MAX_PLANES, stream));
    RS_CUDA_CHECK(cudaMemsetAsync(buf.cyls, 0, sizeof(rs_cyl_out) * size_t(prm.batch) * RS_MAX_CYL_REGIONS, stream));
    cape_segment_kernel<<<prm.batch, T, smem, stream>>>(prm, buf);
    RS_LAUNCH_CHECK();
    return RS_OK;
}

}  // namespace rs
