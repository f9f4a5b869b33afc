// K2-K4 `cape_segment` — ONE WARP PER FRAME over the cell graph produced by K1.
//
// Replaces  Primitive_Detection::{init_histogram, grow_planes_and_cylinders, grow_plane_segment_at_seed,
//           region_growing, cylinder_fitting, merge_planes, get_connected_components_matrix,
//           add_planes_to_primitives(mask + boundary points), add_cylinders_to_primitives(mask test)}
//           (primitive_detection.cpp:239-776), Histogram<N> (histogram.hpp:35-113) and
//           Cylinder_Segment's ctor + run_ransac_loop (cylinder_segment.cpp:35-322).
//
// The seed loop is inherently sequential (each seed depends on the histogram state left by the previous one) and the
// chains inside it (ordered FP64 sums, the 3x3 eigen-solve of every refit, the MSAC cost) are serial by construction,
// so a frame is latency bound. The mapping therefore spends as little of the SM as possible per frame - one warp, no
// block-wide barrier anywhere - and lets the batch supply the parallelism (two frames per SM, every frame of a
// 256-frame batch resident at once, and room left on each SM for the pose kernels that run beside this one):
//   * the cell graph (normals, centroids, d, MSE, tolerances) is staged once in shared memory;
//   * the merge predicate can_be_merged(u -> v) of every grid edge depends only on the two cells, so it is evaluated
//     ONCE per frame into four bit planes (one 64-bit word per cell row and direction) instead of once per visit;
//   * region growing (the reference's recursive DFS = reachability in that directed graph) is a bit-parallel flood
//     fill: one lane per cell row, a shift/AND/OR row sweep until the row is stable, neighbour rows read from shared
//     memory, repeated until no row changes - tens of instructions per pass instead of a CTA-wide scan with barriers;
//   * every FP64 sum whose order is fixed by the reference (merging cell sums in index order, LLS sums, MSAC cost) is
//     evaluated in that same order by one lane, from operands gathered by the whole warp, so that results are
//     bit-identical to the restated reference. Compiled with -fmad=false.
#include "cape_internal.cuh"
#include "plane_fit.cuh"

namespace rs {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int STAGE_CELLS = 128;  // cells gathered per round of an ordered sum
constexpr int MAX_ROWS = 64;      // cell rows / columns the bit planes can hold
constexpr int PC_CAP = 200;       // cylinder regions up to this many cells keep their projected normals / centroids in smem
constexpr int UBUF = 3 * RS_CYL_RANSAC_ITERS + 3;  // uniforms one RANSAC run can consume
typedef unsigned long long u64;

struct Scalars {
    int n_planes, n_cyl_regions, n_cylinders, n_seeds, n_boundary, status, uniform_cursor;
    PlaneModel work;       // plane being grown / refit
    PlaneModel work2;
    double axis[3], radius, inv_r2, center[3];
    double dval[8];
};

struct Smem {
    double* cn;    // [Nc][3] cell normals
    double* cc;    // [Nc][3] centroids
    double* cd;    // [Nc]
    double* cmse;  // [Nc]
    double* val;   // [Nc] scratch values (cylinder branch)
    double* stage; // [STAGE_CELLS][10] operands of the ordered sums
    double* pnpc;  // [2][PC_CAP][3] projected normals / centroids of the cylinder branch (HBM scratch beyond PC_CAP)
    double* ubuf;  // [UBUF] uniforms of the current cylinder RANSAC run
    double* ckx;   // [MAX_ROWS] back-projection factor of every cell column's centre pixel
    double* cky;   // [MAX_ROWS] ... of every cell row's centre pixel
    double* pln;   // [RS_MAX_PLANES][3] planes (compact copy for predicates; sums live in the global record)
    double* plc;   // [RS_MAX_PLANES][3]
    double* pld;   // [RS_MAX_PLANES]
    u64 *U, *ACT, *EL, *ER, *EU, *ED;  // [MAX_ROWS] bit planes: unassigned, activated, merge edges from the 4 neighbours
    Scalars* sc;
    float* tol;    // [Nc]
    float* cz;     // [Nc] depth of the cell's centre pixel (boundary points)
    int* hist;     // [cs*cs]
    int* plabel;   // [RS_MAX_PLANES] merge labels
    int* pplanar;  // [RS_MAX_PLANES]
    unsigned* conn; // [RS_MAX_PLANES][RS_MAX_PLANES/32]
    short* bins;   // [Nc]
    short* list;   // [Nc] ordered activated cells (local -> global)
    short* ids;    // [Nc] remaining local ids (cylinder RANSAC)
    short* plab;   // [Nc] final merged labels
    short* gplane; // [Nc] _gridPlaneSegmentMap
    short* gcyl;   // [Nc] _gridCylinderSegMap
    unsigned char *planar, *mleft, *inlA, *inlB;
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
    double* cc = reinterpret_cast<double*>(take(sizeof(double) * 3 * Nc));
    double* cd = reinterpret_cast<double*>(take(sizeof(double) * Nc));
    double* cmse = reinterpret_cast<double*>(take(sizeof(double) * Nc));
    double* val = reinterpret_cast<double*>(take(sizeof(double) * Nc));
    double* stage = reinterpret_cast<double*>(take(sizeof(double) * STAGE_CELLS * 10));
    double* pnpc = reinterpret_cast<double*>(take(sizeof(double) * 6 * PC_CAP));
    double* ubuf = reinterpret_cast<double*>(take(sizeof(double) * UBUF));
    double* ckx = reinterpret_cast<double*>(take(sizeof(double) * MAX_ROWS));
    double* cky = reinterpret_cast<double*>(take(sizeof(double) * MAX_ROWS));
    double* pln = reinterpret_cast<double*>(take(sizeof(double) * 3 * RS_MAX_PLANES));
    double* plc = reinterpret_cast<double*>(take(sizeof(double) * 3 * RS_MAX_PLANES));
    double* pld = reinterpret_cast<double*>(take(sizeof(double) * RS_MAX_PLANES));
    u64* planes64 = reinterpret_cast<u64*>(take(sizeof(u64) * 6 * MAX_ROWS));
    Scalars* sc = reinterpret_cast<Scalars*>(take(sizeof(Scalars)));
    float* tol = reinterpret_cast<float*>(take(sizeof(float) * Nc));
    float* cz = reinterpret_cast<float*>(take(sizeof(float) * Nc));
    int* hist = reinterpret_cast<int*>(take(sizeof(int) * nbins));
    int* plabel = reinterpret_cast<int*>(take(sizeof(int) * RS_MAX_PLANES));
    int* pplanar = reinterpret_cast<int*>(take(sizeof(int) * RS_MAX_PLANES));
    unsigned* conn = reinterpret_cast<unsigned*>(take(sizeof(unsigned) * RS_MAX_PLANES * (RS_MAX_PLANES / 32)));
    short* sh[6];
    for (int i = 0; i < 6; ++i) sh[i] = reinterpret_cast<short*>(take(sizeof(short) * Nc));
    unsigned char* u8[4];
    for (int i = 0; i < 4; ++i) u8[i] = take(size_t(Nc));
    if (s) {
        s->cn = cn, s->cc = cc, s->cd = cd, s->cmse = cmse, s->val = val, s->stage = stage, s->pnpc = pnpc, s->ubuf = ubuf, s->ckx = ckx, s->cky = cky;
        s->pln = pln, s->plc = plc, s->pld = pld, s->sc = sc, s->tol = tol, s->cz = cz, s->hist = hist;
        s->U = planes64, s->ACT = planes64 + MAX_ROWS, s->EL = planes64 + 2 * MAX_ROWS, s->ER = planes64 + 3 * MAX_ROWS;
        s->EU = planes64 + 4 * MAX_ROWS, s->ED = planes64 + 5 * MAX_ROWS;
        s->plabel = plabel, s->pplanar = pplanar, s->conn = conn;
        s->bins = sh[0], s->list = sh[1], s->ids = sh[2], s->plab = sh[3], s->gplane = sh[4], s->gcyl = sh[5];
        s->planar = u8[0], s->mleft = u8[1], s->inlA = u8[2], s->inlB = u8[3];
    }
    return o;
}

// ---- warp-wide helpers (all 32 lanes must call) ---------------------------------------------------
// lexicographic minimum of (key, idx) over the warp
__device__ __forceinline__ void warp_min_key(double& key, int& idx)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ok = __shfl_xor_sync(FULL, key, o);
        const int oi = __shfl_xor_sync(FULL, idx, o);
        if (ok < key || (ok == key && oi < idx)) key = ok, idx = oi;
    }
}

// exclusive prefix of `v` over the lanes; *total = warp sum
__device__ __forceinline__ int warp_excl_scan(const int v, const int lane, int* total)
{
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    *total = __shfl_sync(FULL, inc, 31);
    return inc - v;
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
// Must be called by lane 0 only. Returns the 1-based plane id or 0 when the capacity is exhausted.
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

// S/count of `dst` += sums of the listed cells, in list order. The operands are gathered STAGE_CELLS cells at a time
// by the whole warp (coalesced 80-byte reads of the records K1 wrote); lanes 0..9 then run one ordered sum each.
__device__ void ordered_expand(const Smem& s, PlaneModel& dst, const rs_cell_out* __restrict__ cells, const short* list,
                               const int n, const unsigned char* filter, const int lane)
{
    double acc = 0.0;
    if (lane < 9) acc = dst.S[lane];
    if (lane == 9) acc = static_cast<double>(dst.count);  // point counts stay far below 2^53: the double sum is exact
    for (int base = 0; base < n; base += STAGE_CELLS) {
        const int m = min(STAGE_CELLS, n - base);
        for (int idx = lane; idx < m * 10; idx += 32) {
            const int j = idx / 10, t = idx - 10 * j;
            double v = 0.0;
            if (!filter || filter[base + j]) {
                const rs_cell_out* c = cells + list[base + j];
                v = t < 9 ? c->S[t] : static_cast<double>(c->count);
            }
            s.stage[idx] = v;
        }
        __syncwarp();
        if (lane < 10) {
            for (int j = 0; j < m; ++j)
                if (!filter || filter[base + j]) acc += s.stage[j * 10 + lane];
        }
        __syncwarp();
    }
    if (lane < 9) dst.S[lane] = acc;
    if (lane == 9) dst.count = static_cast<int>(acc);
    __syncwarp();
}

// ---- cylinder branch (cylinder_segment.cpp:35-322 + primitive_detection.cpp:413-501) ------------
// pn / pc: projected normals / centroids of the region's cells, [m][3] each, in this frame's global scratch.
__device__ void cylinder_fitting(const Smem& s, const SegmentParams& prm, const rs_cell_out* cells, rs_plane_out* planes,
                                 rs_cyl_out* cyls, int32_t* cyl_region_seg, double* pn, double* pc, const int m,
                                 const int Nc, const int lane)
{
    Scalars& sc = *s.sc;
    if (m <= PC_CAP) {   // the usual case: keep the region's projected normals / centroids in shared memory
        pn = s.pnpc;
        pc = s.pnpc + 3 * PC_CAP;
    }
    if (sc.n_cyl_regions >= RS_MAX_CYL_REGIONS) {
        if (lane == 0) sc.status = RS_ERR_CAPACITY;
        __syncwarp();
        return;
    }
    const int region = sc.n_cyl_regions;
    rs_cyl_out& co = cyls[region];
    __syncwarp();

    // covariance of [N -N] (3 x 2m), summed column by column: six unique entries, one lane each
    if (lane < 6) {
        const int r = (lane < 3) ? 0 : (lane < 5 ? 1 : 2);
        const int c = (lane < 3) ? lane : (lane < 5 ? lane - 2 : 2);
        double acc = 0;
        for (int j = 0; j < m; ++j) acc += s.cn[3 * s.list[j] + r] * s.cn[3 * s.list[j] + c];
        for (int j = 0; j < m; ++j) acc += (-s.cn[3 * s.list[j] + r]) * (-s.cn[3 * s.list[j] + c]);
        sc.dval[lane] = acc / static_cast<double>(2 * m - 1);
    }
    __syncwarp();
    int flag = 0;
    if (lane == 0) {
        // entries: 0:(0,0) 1:(0,1) 2:(0,2) 3:(1,1) 4:(1,2) 5:(2,2); lower triangle a10=(0,1), a20=(0,2), a21=(1,2)
        double ev[3], q[3][3];
        self_adjoint_eigen3(sc.dval[0], sc.dval[1], sc.dval[3], sc.dval[2], sc.dval[4], sc.dval[5], ev, q);
        const double score = ev[2] / ev[0];
        co.n_cells = m;
        co.n_segments = 0;
        co.pca_score = score;
        sc.axis[0] = q[0][0], sc.axis[1] = q[1][0], sc.axis[2] = q[2][0];
        flag = !(score < 75.0);  // cylinderRansacMinimumScore (float 75)
        if (flag)
            for (int i = 0; i < 3; ++i) co.axis[i] = sc.axis[i];
        sc.n_cyl_regions = region + 1;
    }
    flag = __shfl_sync(FULL, flag, 0);
    __syncwarp();
    if (!flag) return;

    const double ax = sc.axis[0], ay = sc.axis[1], az = sc.axis[2];
    for (int j = lane; j < m; j += 32) {
        const int gi = s.list[j];
        const double c0 = s.cc[3 * gi], c1 = s.cc[3 * gi + 1], c2 = s.cc[3 * gi + 2];
        const double cdot = (ax * c0 + ay * c1) + az * c2;
        pc[3 * j] = c0 - cdot * ax;
        pc[3 * j + 1] = c1 - cdot * ay;
        pc[3 * j + 2] = c2 - cdot * az;
        const double n0 = s.cn[3 * gi], n1 = s.cn[3 * gi + 1], n2 = s.cn[3 * gi + 2];
        const double ndot = (ax * n0 + ay * n1) + az * n2;
        const double p0 = n0 - ndot * ax, p1 = n1 - ndot * ay, p2 = n2 - ndot * az;
        const double nn = sqrt((p0 * p0 + p1 * p1) + p2 * p2);
        pn[3 * j] = p0 / nn;
        pn[3 * j + 1] = p1 / nn;
        pn[3 * j + 2] = p2 / nn;
        s.mleft[j] = 1;
        s.ids[j] = static_cast<short>(j);
    }
    __syncwarp();
    int nleft = m, nIds = m;   // warp-uniform

    const unsigned minimumCellActivated = static_cast<unsigned>(0.65 / 100.0 * static_cast<double>(Nc));
    const float maxSqrtDist = 0.04f;
    const double maxSqrtDistD = static_cast<double>(maxSqrtDist);
    int segId = 0;
    while (static_cast<unsigned>(nleft) > minimumCellActivated && nleft > 0.1 * m) {
        // ---- run_ransac_loop ----
        unsigned char* best = s.inlA;
        unsigned char* cand = s.inlB;
        int bestCount = 0;
        if (nIds >= 3) {
            const unsigned accepted = static_cast<unsigned>(floor(0.9 * nIds));
            double minHypothesisDist = static_cast<double>(maxSqrtDist * static_cast<float>(nIds));
            for (int j = lane; j < m; j += 32) best[j] = 0;
            // this run's uniforms (3 per iteration, consumed in order from the frame's stream), one coalesced fetch
            const int ubase = sc.uniform_cursor;
            for (int k = lane; k < UBUF - 3; k += 32) s.ubuf[k] = (ubase + k < prm.n_uniforms) ? prm.uniforms[ubase + k] : -1.0;
            __syncwarp();
            for (int it = 0; it < RS_CYL_RANSAC_ITERS; ++it) {
                if (lane == 0) {
                    int id[3];
                    for (int k = 0; k < 3; ++k) {
                        double u = s.ubuf[3 * it + k];
                        if (u < 0.0) {
                            u = 0.0;
                            sc.status = RS_ERR_CAPACITY;
                        }
                        id[k] = s.ids[static_cast<unsigned>(floor(u * static_cast<double>(static_cast<unsigned>(nIds))))];
                    }
                    sc.uniform_cursor = ubase + 3 * (it + 1);
                    const double* n1 = pn + 3 * id[0];
                    const double* n2 = pn + 3 * id[1];
                    const double* n3 = pn + 3 * id[2];
                    const double* c1 = pc + 3 * id[0];
                    const double* c2 = pc + 3 * id[1];
                    const double* c3 = pc + 3 * id[2];
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
                __syncwarp();
                const double radius = sc.radius, invr2 = sc.inv_r2, e0 = sc.center[0], e1 = sc.center[1], e2 = sc.center[2];
                double partial = 0.0;   // this lane's share of the MSAC cost (any order: only used to rule hypotheses out)
                int candCount = 0;
                for (int j = lane; j < m; j += 32) {
                    double v = -1.0;  // not part of this RANSAC round
                    unsigned char in = 0;
                    if (s.mleft[j]) {
                        const double d0 = (pc[3 * j] - radius * pn[3 * j]) - e0;
                        const double d1 = (pc[3 * j + 1] - radius * pn[3 * j + 1]) - e1;
                        const double d2 = (pc[3 * j + 2] - radius * pn[3 * j + 2]) - e2;
                        const double distance = ((d0 * d0 + d1 * d1) + d2 * d2) * invr2;
                        if (distance < maxSqrtDistD) {
                            v = distance;
                            in = 1;
                        }
                        else
                            v = maxSqrtDistD;
                        partial += v;
                    }
                    s.val[j] = v;
                    cand[j] = in;
                    candCount += in;
                }
                candCount = __reduce_add_sync(FULL, candCount);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) partial += __shfl_xor_sync(FULL, partial, o);
                __syncwarp();
                // The reference accumulates the cost in cell order and compares it with the best so far. A sum of at most
                // Nc non-negative doubles in any order is within Nc * 2^-53 of it, so a hypothesis whose unordered sum is
                // clearly above the threshold is rejected without the serial pass; otherwise the ordered sum decides.
                int better = 0;
                double dist = 0.0;
                if (!(partial > minHypothesisDist * (1.0 + 1e-9))) {
                    if (lane == 0) {
                        for (int j = 0; j < m; ++j) {
                            const double v = s.val[j];
                            if (v >= 0.0) dist += v;
                        }
                        better = dist < minHypothesisDist ? 1 : 0;
                    }
                    better = __shfl_sync(FULL, better, 0);
                }
                bool stop = false;
                if (better) {
                    minHypothesisDist = __shfl_sync(FULL, dist, 0);
                    unsigned char* tmp = best;
                    best = cand;
                    cand = tmp;
                    const int prevCount = bestCount;
                    bestCount = candCount;
                    // quirk: the early stop looks at the PREVIOUS best set (vectors swapped before the test)
                    if (static_cast<unsigned>(prevCount) > accepted) stop = true;
                }
                __syncwarp();
                if (stop) break;
            }
        }
        if (bestCount < 6) break;
        if (segId >= RS_MAX_CYL_SEGS) {
            if (lane == 0) sc.status = RS_ERR_CAPACITY;
            break;
        }

        // ---- LLS over the inliers, ordered sums (7 running sums) ----
        if (lane < 7) {
            double acc = 0.0;
            for (int j = 0; j < m; ++j)
                if (best[j]) {
                    if (lane < 3)
                        acc += pn[3 * j + lane];
                    else if (lane < 6)
                        acc += pc[3 * j + (lane - 3)];
                    else
                        acc += (pn[3 * j] * pc[3 * j] + pn[3 * j + 1] * pc[3 * j + 1]) + pn[3 * j + 2] * pc[3 * j + 2];
                }
            sc.dval[lane] = acc;
        }
        __syncwarp();
        if (lane == 0) {
            // rebuild the remaining id list in index order, drop the inliers from the mask
            int nl = nleft, k = 0;
            for (int j = 0; j < m; ++j) {
                if (best[j]) {
                    s.mleft[j] = 0;
                    nl--;
                }
                else if (s.mleft[j])
                    s.ids[k++] = static_cast<short>(j);
            }
            nleft = nl;
            nIds = k;
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
        nleft = __shfl_sync(FULL, nleft, 0);
        nIds = __shfl_sync(FULL, nIds, 0);
        __syncwarp();
        {
            // per-inlier squared point-to-axis distance error, using the UNPROJECTED centroids (:199-218)
            const double P1[3] = {sc.center[0], sc.center[1], sc.center[2]};
            const double P2[3] = {P1[0] + sc.axis[0], P1[1] + sc.axis[1], P1[2] + sc.axis[2]};
            const double D[3] = {P2[0] - P1[0], P2[1] - P1[1], P2[2] - P1[2]};
            const double P1P2 = sqrt(dot3(D, D));
            const double radius = sc.radius;
            for (int j = lane; j < m; j += 32) {
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
        if (lane == 0) plane_clear(sc.work2);
        __syncwarp();
        ordered_expand(s, sc.work2, cells, s.list, m, best, lane);
        int f = 0;
        if (lane == 0) {
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
                f = id;    // > 0: label inliers as plane `id`
            }
            else {
                const int id = ++sc.n_cylinders;
                assigned = id;
                f = -id;   // < 0: label inliers as cylinder `id`
            }
            co.assigned[segId] = assigned;
            co.kept[segId] = 0;
        }
        f = __shfl_sync(FULL, f, 0);
        __syncwarp();
        for (int j = lane; j < m; j += 32)
            if (best[j]) {
                if (f > 0)
                    s.gplane[s.list[j]] = static_cast<short>(f);
                else if (f < 0)
                    s.gcyl[s.list[j]] = static_cast<short>(-f);
            }
        __syncwarp();
        ++segId;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(32) cape_segment_kernel(const SegmentParams prm, const SegmentBuffers buf)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Nc = prm.hc * prm.vc, hc = prm.hc, vc = prm.vc, cs = prm.cell;
    const int nbins = cs * cs;
    Smem s;
    carve(&s, smem_raw, Nc, nbins);
    Scalars& sc = *s.sc;
    const int lane = threadIdx.x;
    const int frame = blockIdx.x;
    const rs_cell_out* cells = buf.cells + size_t(frame) * Nc;
    rs_plane_out* planes = buf.planes + size_t(frame) * RS_MAX_PLANES;
    rs_cyl_out* cyls = buf.cyls + size_t(frame) * RS_MAX_CYL_REGIONS;
    int32_t* out_region_seg = buf.cyl_region_seg + size_t(frame) * Nc;
    const float* depth = buf.depth + size_t(frame) * prm.W * prm.H;
    double* boundary = buf.boundary_xyz + size_t(frame) * prm.max_boundary * 3;
    double* pn = buf.scratch + size_t(frame) * Nc * 6;
    double* pc = pn + size_t(Nc) * 3;

    // ---- load the cell graph, init_histogram (primitive_detection.cpp:239-265, histogram.hpp:35-62) ----
    for (int i = lane; i < nbins; i += 32) s.hist[i] = 0;
    for (int i = lane; i < 6 * MAX_ROWS; i += 32) s.U[i] = 0ull;   // the six bit planes are contiguous
    if (lane == 0) {
        sc.n_planes = 0, sc.n_cyl_regions = 0, sc.n_cylinders = 0, sc.n_seeds = 0, sc.n_boundary = 0;
        sc.status = RS_OK, sc.uniform_cursor = 0;
    }
    __syncwarp();
    const unsigned pixelPerCellSide = static_cast<unsigned>(sqrtf(static_cast<float>(cs * cs)));
    // centre pixel of every cell (the only depth the boundary step reads): fetched here, all loads in flight at once
    for (int i = lane; i < Nc; i += 32) {
        const int row = i / hc, col = i - row * hc;
        const int centerX = static_cast<int>(col * pixelPerCellSide + pixelPerCellSide / 2);
        const int centerY = static_cast<int>(row * pixelPerCellSide + pixelPerCellSide / 2);
        s.cz[i] = __ldg(depth + size_t(centerY) * prm.W + centerX);
    }
    for (int x = lane; x < hc; x += 32) s.ckx[x] = prm.kx[static_cast<int>(x * pixelPerCellSide + pixelPerCellSide / 2)];
    for (int y = lane; y < vc; y += 32) s.cky[y] = prm.ky[static_cast<int>(y * pixelPerCellSide + pixelPerCellSide / 2)];
    int nPlanar = 0;
    for (int i = lane; i < Nc; i += 32) {
        const rs_cell_out& c = cells[i];
        s.cn[3 * i] = c.normal[0], s.cn[3 * i + 1] = c.normal[1], s.cn[3 * i + 2] = c.normal[2];
        s.cc[3 * i] = c.centroid[0], s.cc[3 * i + 1] = c.centroid[1], s.cc[3 * i + 2] = c.centroid[2];
        s.cd[i] = c.d;
        s.cmse[i] = c.mse;
        s.tol[i] = c.tol;
        s.planar[i] = c.planar ? 1 : 0;
        s.gplane[i] = 0;
        s.gcyl[i] = 0;
        out_region_seg[i] = 0;
        int bin = -1;
        if (c.planar) {
            ++nPlanar;
            bin = c.hist_bin;   // init_histogram's bin, computed by the per-cell fit kernel
            if (bin >= 0 && bin < nbins) atomicAdd(&s.hist[bin], 1);
        }
        s.bins[i] = static_cast<short>(bin);
    }
    nPlanar = __reduce_add_sync(FULL, nPlanar);
    __syncwarp();

    // ---- merge edges of the cell graph, once per frame: bit x of E?[y] = can_be_merged(neighbour -> (y, x)) ----
    // one row at a time, lane = column (two halves when the grid is wider than 32 cells): the words come out of ballots
    for (int y = 0; y < vc; ++y) {
        u64 wU = 0, wL = 0, wR = 0, wUp = 0, wDn = 0;
        for (int x0 = 0; x0 < hc; x0 += 32) {
            const int x = x0 + lane;
            bool pl = false, eL = false, eR = false, eU = false, eD = false;
            if (x < hc) {
                const int i = y * hc + x;
                pl = s.planar[i] != 0;
                if (pl) {
                    const double tolI = static_cast<double>(s.tol[i]);
                    eL = x > 0 && s.planar[i - 1] &&
                         plane_can_merge(s.cn + 3 * (i - 1), s.cd[i - 1], s.cn + 3 * i, s.cc + 3 * i, tolI, prm.cos_merge);
                    eR = x < hc - 1 && s.planar[i + 1] &&
                         plane_can_merge(s.cn + 3 * (i + 1), s.cd[i + 1], s.cn + 3 * i, s.cc + 3 * i, tolI, prm.cos_merge);
                    eU = y > 0 && s.planar[i - hc] &&
                         plane_can_merge(s.cn + 3 * (i - hc), s.cd[i - hc], s.cn + 3 * i, s.cc + 3 * i, tolI, prm.cos_merge);
                    eD = y < vc - 1 && s.planar[i + hc] &&
                         plane_can_merge(s.cn + 3 * (i + hc), s.cd[i + hc], s.cn + 3 * i, s.cc + 3 * i, tolI, prm.cos_merge);
                }
            }
            wU |= static_cast<u64>(__ballot_sync(FULL, pl)) << x0;
            wL |= static_cast<u64>(__ballot_sync(FULL, eL)) << x0;
            wR |= static_cast<u64>(__ballot_sync(FULL, eR)) << x0;
            wUp |= static_cast<u64>(__ballot_sync(FULL, eU)) << x0;
            wDn |= static_cast<u64>(__ballot_sync(FULL, eD)) << x0;
        }
        if (lane == 0) s.U[y] = wU, s.EL[y] = wL, s.ER[y] = wR, s.EU[y] = wUp, s.ED[y] = wDn;
    }
    __syncwarp();

    const unsigned planeSeedCount = static_cast<unsigned>(0.8 / 100.0 * Nc);
    const unsigned minimumCellActivated = static_cast<unsigned>(0.65 / 100.0 * Nc);

    // ---- grow_planes_and_cylinders (primitive_detection.cpp:267-310) ----
    int untried = nPlanar;
    int guard = 0;
    while (untried > 0) {
        // most frequent bin: strictly greatest count, lowest index on ties (histogram.hpp:69-84)
        int bestBin;
        {
            double key = 1.0;  // -count as key so that the minimum is the fullest bin
            int idx = 0x7fffffff;
            for (int i = lane; i < nbins; i += 32) {
                const int h = s.hist[i];
                if (h > 0) {
                    const double k = -static_cast<double>(h);
                    if (k < key || (k == key && i < idx)) key = k, idx = i;
                }
            }
            warp_min_key(key, idx);
            bestBin = (key < 0.0) ? idx : -1;
        }
        // candidates of that bin, min-MSE seed (first strictly smallest; stop at MSE <= 0) (:286-298)
        int seed, candCount = 0;
        {
            double key = DBL_MAX;
            int idx = 0x7fffffff;
            if (bestBin >= 0) {
                for (int i = lane; i < Nc; i += 32)
                    if (s.bins[i] == bestBin) {
                        ++candCount;
                        const double mse = s.cmse[i];
                        const double k = (mse <= 0.0) ? 0.0 : mse;
                        if (k < key || (k == key && i < idx)) key = k, idx = i;
                    }
            }
            candCount = __reduce_add_sync(FULL, candCount);
            warp_min_key(key, idx);
            seed = (key < DBL_MAX) ? idx : -1;
        }
        if (static_cast<unsigned>(candCount) < planeSeedCount) break;
        if (seed < 0) break;
        if (lane == 0) sc.n_seeds++;

        // ---- grow_plane_segment_at_seed (:312-389) ----
        // A non-planar seed changes no state in the reference (:318-322), which would then spin forever; it is
        // unreachable (only planar cells carry a bin id). Leave the loop instead of hanging the GPU.
        if (!s.planar[seed] || ++guard > 4 * Nc + nbins) {
            if (lane == 0) sc.status = RS_ERR_CAPACITY;
            break;
        }
        for (int r = lane; r < MAX_ROWS; r += 32) s.ACT[r] = 0ull;
        const int seedY = seed / hc, seedX = seed - seedY * hc;
        int flag = 0;
        if (lane == 0) {
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
            if (((s.U[seedY] >> seedX) & 1ull) &&
                plane_can_merge(w.n, w.d, s.cn + 3 * seed, s.cc + 3 * seed, static_cast<double>(s.tol[seed]), prm.cos_merge))
                flag = 1;
        }
        flag = __shfl_sync(FULL, flag, 0);
        __syncwarp();
        if (flag) {
            if (lane == 0) s.ACT[seedY] = 1ull << seedX;
            __syncwarp();
            // bit-parallel flood fill: a cell joins when some activated 4-neighbour u has can_be_merged(u -> cell)
            while (true) {
                bool changed = false;
                for (int r = lane; r < vc; r += 32) {
                    const u64 a = s.ACT[r];
                    const u64 open = s.U[r];
                    const u64 up = r > 0 ? s.ACT[r - 1] : 0ull;
                    const u64 dn = r < vc - 1 ? s.ACT[r + 1] : 0ull;
                    const u64 el = s.EL[r], er = s.ER[r];
                    u64 na = a | (open & ((up & s.EU[r]) | (dn & s.ED[r])));
                    while (true) {
                        const u64 nb = na | (open & (((na << 1) & el) | ((na >> 1) & er)));
                        if (nb == na) break;
                        na = nb;
                    }
                    if (na != a) {
                        s.ACT[r] = na;
                        changed = true;
                    }
                }
                __syncwarp();
                if (!__any_sync(FULL, changed)) break;
            }
        }
        // ordered list of the activated cells (index order)
        int cnt;
        {
            int base = 0;
            for (int r0 = 0; r0 < vc; r0 += 32) {
                const int r = r0 + lane;
                u64 bits = r < vc ? s.ACT[r] : 0ull;
                int tot;
                int pos = base + warp_excl_scan(__popcll(bits), lane, &tot);
                while (bits) {
                    const int x = __ffsll(static_cast<long long>(bits)) - 1;
                    bits &= bits - 1;
                    s.list[pos++] = static_cast<short>(r * hc + x);
                }
                base += tot;
            }
            cnt = base;
        }
        __syncwarp();
        // merge activated cells & remove them from the histogram (:343-360)
        ordered_expand(s, sc.work, cells, s.list, cnt, nullptr, lane);
        for (int j = lane; j < cnt; j += 32) hist_remove_atomic(s, s.list[j], nbins);
        for (int r = lane; r < vc; r += 32) s.U[r] &= ~s.ACT[r];
        __syncwarp();
        untried -= cnt;
        if (cnt == 0 || static_cast<unsigned>(cnt) < minimumCellActivated) {
            if (lane == 0) {
                // _histogram.remove_point(seedId)
                const int b = s.bins[seed];
                if (b >= 0 && b < nbins && s.hist[b] != 0) s.hist[b] -= 1;
                s.bins[seed] = 1;
            }
            __syncwarp();
            continue;
        }
        int f = 0;
        if (lane == 0) {
            plane_fit(sc.work);
            if (sc.work.planar) {
                if (sc.work.score > 100)
                    f = push_plane(s, planes, sc.work);  // add_plane_segment_to_features (:391-411)
                else if (cnt > 5)
                    f = -1;
            }
        }
        f = __shfl_sync(FULL, f, 0);
        __syncwarp();
        if (f > 0) {
            for (int j = lane; j < cnt; j += 32) s.gplane[s.list[j]] = static_cast<short>(f);
        }
        else if (f < 0) {
            cylinder_fitting(s, prm, cells, planes, cyls, out_region_seg, pn, pc, cnt, Nc, lane);
        }
        __syncwarp();
    }
    __syncwarp();

    // ---- merge_planes (:503-560) with get_connected_components_matrix (:736-776) ----
    const int P = sc.n_planes;
    constexpr int CW = RS_MAX_PLANES / 32;
    for (int i = lane; i < RS_MAX_PLANES * CW; i += 32) s.conn[i] = 0;
    __syncwarp();
    for (int i = lane; i < Nc; i += 32) {
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
    __syncwarp();
    if (lane == 0) {
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
    __syncwarp();

    // ---- final labels + per-plane record tail ----
    for (int i = lane; i < Nc; i += 32) {
        const int id = s.gplane[i];
        int lab = 0;
        if (id > 0) {
            const int root = s.plabel[id - 1];
            if (s.plabel[root] == root && s.pplanar[root]) lab = root + 1;
        }
        s.plab[i] = static_cast<short>(lab);
    }
    __syncwarp();

    // ---- boundary points per final plane (compute_plane_segment_boundary, :650-703) ----
    // mask rows as 64-bit words (ACT is free now): boundary = dilate3x3(mask) & ~erode_cross(mask), erosion with a zero
    // border, dilation ignoring it - a handful of shifts per row instead of 14 taps per cell.
    const u64 rowMask = hc >= 64 ? ~0ull : ((1ull << hc) - 1ull);
    u64* MK = s.ACT;   // [MAX_ROWS] plane mask
    int nFinal = 0;
    int nBoundary = 0;   // warp-uniform running count
    int status = RS_OK;
    for (int k = 0; k < P; ++k) {
        const bool isFinal = (s.plabel[k] == k) && s.pplanar[k];
        if (lane == 0) {
            planes[k].merge_label = s.plabel[k];
            planes[k].is_final = isFinal ? 1 : 0;
            planes[k].n_boundary = 0;
            planes[k].boundary_offset = nBoundary;
        }
        if (!isFinal) continue;
        ++nFinal;
        for (int y = 0; y < vc; ++y) {
            u64 w = 0;
            for (int x0 = 0; x0 < hc; x0 += 32) {
                const int x = x0 + lane;
                w |= static_cast<u64>(__ballot_sync(FULL, x < hc && s.plab[y * hc + x] == k + 1)) << x0;
            }
            if (lane == 0) MK[y] = w;
        }
        __syncwarp();
        const double maxBoundaryDistance = 3 * sqrt(planes[k].mse);
        const double n0 = s.pln[3 * k], n1 = s.pln[3 * k + 1], n2 = s.pln[3 * k + 2], dd = s.pld[k];
        const int start = nBoundary;
        int base = nBoundary;
        for (int r0 = 0; r0 < vc; r0 += 32) {
            const int r = r0 + lane;
            u64 keepBits = 0;
            if (r < vc) {
                const u64 m = MK[r];
                const u64 up = r > 0 ? MK[r - 1] : 0ull, dn = r < vc - 1 ? MK[r + 1] : 0ull;
                const u64 er = m & (m << 1) & (m >> 1) & up & dn;
                const u64 di = ((m | (m << 1) | (m >> 1)) | (up | (up << 1) | (up >> 1)) | (dn | (dn << 1) | (dn >> 1))) & rowMask;
                u64 bits = di & ~er;
                while (bits) {
                    const int x = __ffsll(static_cast<long long>(bits)) - 1;
                    bits &= bits - 1;
                    const double z = static_cast<double>(s.cz[r * hc + x]);
                    if (z > 0) {
                        const double px = z * s.ckx[x], py = z * s.cky[r];
                        if (fabs(((n0 * px + n1 * py) + n2 * z) + dd) < maxBoundaryDistance) keepBits |= 1ull << x;
                    }
                }
            }
            int tot;
            int o = base + warp_excl_scan(__popcll(keepBits), lane, &tot);
            while (keepBits) {
                const int x = __ffsll(static_cast<long long>(keepBits)) - 1;
                keepBits &= keepBits - 1;
                if (o < prm.max_boundary) {
                    const double z = static_cast<double>(s.cz[r * hc + x]);
                    boundary[3 * o] = z * s.ckx[x], boundary[3 * o + 1] = z * s.cky[r], boundary[3 * o + 2] = z;
                }
                ++o;
            }
            base += tot;
        }
        if (base > prm.max_boundary) {
            status = RS_ERR_CAPACITY;
            base = prm.max_boundary;
        }
        if (lane == 0) planes[k].n_boundary = base - start;
        nBoundary = base;
        __syncwarp();
    }

    // ---- cylinders: opening test of add_cylinders_to_primitives (:705-734): dilate, erode, erode with the cross
    // kernel, borders ignored (taps outside the grid do not contribute) ----
    const int nCyl = sc.n_cylinders;
    for (int ci = 1; ci <= nCyl; ++ci) {
        u64* A = s.ACT;
        u64* Bm = s.EL;
        for (int y = 0; y < vc; ++y) {
            u64 w = 0;
            for (int x0 = 0; x0 < hc; x0 += 32) {
                const int x = x0 + lane;
                w |= static_cast<u64>(__ballot_sync(FULL, x < hc && s.gcyl[y * hc + x] == ci)) << x0;
            }
            if (lane == 0) A[y] = w;
        }
        __syncwarp();
        const u64 first = 1ull, last = 1ull << (hc - 1);
        for (int r = lane; r < vc; r += 32) {   // dilate
            const u64 m = A[r];
            Bm[r] = (m | (m << 1) | (m >> 1) | (r > 0 ? A[r - 1] : 0ull) | (r < vc - 1 ? A[r + 1] : 0ull)) & rowMask;
        }
        __syncwarp();
        for (int pass = 0; pass < 2; ++pass) {  // erode twice: Bm -> A -> Bm
            const u64* src = pass == 0 ? Bm : A;
            u64* dst = pass == 0 ? A : Bm;
            for (int r = lane; r < vc; r += 32) {
                const u64 m = src[r];
                dst[r] = m & ((m << 1) | first) & ((m >> 1) | last) & (r > 0 ? src[r - 1] : ~0ull) &
                         (r < vc - 1 ? src[r + 1] : ~0ull) & rowMask;
            }
            __syncwarp();
        }
        bool anySet = false, anyClear = false;
        for (int r = lane; r < vc; r += 32) {
            anySet = anySet || Bm[r] != 0ull;
            anyClear = anyClear || Bm[r] != rowMask;
        }
        anySet = __any_sync(FULL, anySet);
        anyClear = __any_sync(FULL, anyClear);
        if (lane == 0) {
            const int kept = (anySet && anyClear) ? 1 : 0;   // !(max <= 0 || min >= max) on a 0/1 mask
            for (int r = 0; r < sc.n_cyl_regions; ++r)
                for (int sg = 0; sg < cyls[r].n_segments; ++sg)
                    if (cyls[r].assigned[sg] == ci) cyls[r].kept[sg] = kept;
        }
        __syncwarp();
    }

    // ---- write the label grids and the frame info ----
    int32_t* o_grid = buf.plane_grid + size_t(frame) * Nc;
    int32_t* o_lab = buf.plane_labels + size_t(frame) * Nc;
    int32_t* o_cyl = buf.cyl_labels + size_t(frame) * Nc;
    for (int i = lane; i < Nc; i += 32) {
        o_grid[i] = s.gplane[i];
        o_lab[i] = s.plab[i];
        o_cyl[i] = s.gcyl[i];
    }
    if (lane == 0) {
        rs_cape_frame_info& info = buf.info[frame];
        info.status = sc.status != RS_OK ? sc.status : status;
        info.n_planar_cells = nPlanar;
        info.n_seeds = sc.n_seeds;
        info.n_planes = P;
        info.n_final_planes = nFinal;
        info.n_cyl_regions = sc.n_cyl_regions;
        info.n_cylinders = sc.n_cylinders;
        info.n_boundary = nBoundary;
    }
}

}  // namespace

size_t cape_segment_scratch_doubles_per_frame(int n_cells) { return size_t(n_cells) * 6; }

// Geometries the segmentation kernel can take (checked by rs_cape_create, so that an unsupported grid fails at creation
// instead of on the first run).
int cape_segment_validate(const int hc, const int vc, const int cell)
{
    const int Nc = hc * vc;
    if (hc > MAX_ROWS || vc > MAX_ROWS || Nc > 32767 || cell * cell > 32767) {
        set_last_error("cape_segment: at most 64 x 64 cells (one 64-bit word per cell row and merge direction)");
        return RS_ERR_INVALID_ARG;
    }
    const size_t smem = carve(nullptr, nullptr, Nc, cell * cell);
    if (smem > size_t(kMaxDynamicSmem)) {
        set_last_error("cape_segment: the cell grid does not fit in shared memory (" + std::to_string(smem) + " B > 227 KB)");
        return RS_ERR_INVALID_ARG;
    }
    return RS_OK;
}

int launch_cape_segment(const SegmentParams& prm, const SegmentBuffers& buf, cudaStream_t stream)
{
    const int Nc = prm.hc * prm.vc;
    const int rc = cape_segment_validate(prm.hc, prm.vc, prm.cell);
    if (rc != RS_OK) return rc;
    const size_t smem = carve(nullptr, nullptr, Nc, prm.cell * prm.cell);
    static SmemOptIn optin;
    RS_CUDA_CHECK(optin.ensure(cape_segment_kernel, smem));
    // the plane/cylinder record arrays are zeroed so that unused entries read as empty
    RS_CUDA_CHECK(cudaMemsetAsync(buf.planes, 0, sizeof(rs_plane_out) * size_t(prm.batch) * RS_MAX_PLANES, stream));
    RS_CUDA_CHECK(cudaMemsetAsync(buf.cyls, 0, sizeof(rs_cyl_out) * size_t(prm.batch) * RS_MAX_CYL_REGIONS, stream));
    cape_segment_kernel<<<prm.batch, 32, smem, stream>>>(prm, buf);
    RS_LAUNCH_CHECK();
    return RS_OK;
}

}  // namespace rs
