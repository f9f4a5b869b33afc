// TEST INFRASTRUCTURE — C entry points of the CPU oracle for ctypes (tests/, __graft_entry__.smoke(), bench.py's
// cpu_baseline / --impl reference legs only). Not part of the product; the product library never links this.
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

#include "cape.hpp"
#include "kalman.hpp"
#include "polygon.hpp"
#include "pose.hpp"

using namespace oracle;

namespace {
void store_frame(const CapeFrame& f, int Nc, int maxBoundary, int b, const rs_cape_outputs* o)
{
    if (o->cells) std::memcpy(o->cells + size_t(b) * Nc, f.cells.data(), sizeof(rs_cell_out) * Nc);
    if (o->plane_grid) std::memcpy(o->plane_grid + size_t(b) * Nc, f.plane_grid.data(), 4 * size_t(Nc));
    if (o->plane_labels) std::memcpy(o->plane_labels + size_t(b) * Nc, f.plane_labels.data(), 4 * size_t(Nc));
    if (o->cyl_labels) std::memcpy(o->cyl_labels + size_t(b) * Nc, f.cyl_labels.data(), 4 * size_t(Nc));
    if (o->cyl_region_seg) std::memcpy(o->cyl_region_seg + size_t(b) * Nc, f.cyl_region_seg.data(), 4 * size_t(Nc));
    rs_cape_frame_info info = f.info;
    if (o->planes) {
        rs_plane_out* dst = o->planes + size_t(b) * RS_MAX_PLANES;
        std::memset(dst, 0, sizeof(rs_plane_out) * RS_MAX_PLANES);
        const size_t n = std::min<size_t>(f.planes.size(), RS_MAX_PLANES);
        if (n) std::memcpy(dst, f.planes.data(), sizeof(rs_plane_out) * n);
    }
    if (o->cyls) {
        rs_cyl_out* dst = o->cyls + size_t(b) * RS_MAX_CYL_REGIONS;
        std::memset(dst, 0, sizeof(rs_cyl_out) * RS_MAX_CYL_REGIONS);
        const size_t n = std::min<size_t>(f.cyls.size(), RS_MAX_CYL_REGIONS);
        if (n) std::memcpy(dst, f.cyls.data(), sizeof(rs_cyl_out) * n);
    }
    if (o->boundary_xyz) {
        double* dst = o->boundary_xyz + size_t(b) * maxBoundary * 3;
        std::memset(dst, 0, sizeof(double) * 3 * size_t(maxBoundary));
        size_t n = f.boundary_xyz.size() / 3;
        if (n > size_t(maxBoundary)) {
            n = size_t(maxBoundary);
            info.status = RS_ERR_CAPACITY;
        }
        if (n) std::memcpy(dst, f.boundary_xyz.data(), sizeof(double) * 3 * n);
    }
    if (o->info) o->info[b] = info;
}
}  // namespace

extern "C" {

int orc_kalman_new_state(int N, int M, const double* F, const double* H, const double* Q, const double* x, const double* P,
                         const double* z, const double* R, double* x_out, double* P_out)
{
    if (N < 1 || N > KF_MAX || M < 1 || M > KF_MAX) return -100;
    return kalman_new_state(N, M, F, H, Q, x, P, z, R, x_out, P_out);
}

int orc_kalman_track_points(int n, const double* x, const double* P, const double* z, const double* R, double q, double* x_out,
                            double* P_out, double* score, uint8_t* moving, int32_t* status)
{
    for (int i = 0; i < n; ++i)
        kalman_track_point(x + 3 * i, P + 9 * i, z + 3 * i, R + 9 * i, q, x_out + 3 * i, P_out + 9 * i, score + i, moving + i,
                           status + i);
    return 0;
}

int orc_kalman_track_planes(int n, const double* x, const double* P, const double* z, const double* R, double q, double* x_out,
                            double* P_out, double* score, int32_t* status)
{
    for (int i = 0; i < n; ++i)
        kalman_track_plane(x + 4 * i, P + 16 * i, z + 4 * i, R + 16 * i, q, x_out + 4 * i, P_out + 16 * i, score + i, status + i);
    return 0;
}

double orc_polygon_inter_area(const double* a, int na, const double* b, int nb) { return polygon_inter_area(a, na, b, nb); }
double orc_polygon_area(const double* a, int n) { return polygon_area(a, n); }

// batched like rs_plane_match: frame f owns detections [det_first[f], det_first[f+1]) and map planes [map_first[f], map_first[f+1])
int orc_plane_match(int n_frames, const double* w2c, const rs_polygon_plane* det, const int32_t* det_first, const double* det_xy,
                    const rs_polygon_plane* map, const int32_t* map_first, const double* map_xy, const uint8_t* det_matched,
                    int advanced_search, int sequential, int32_t* selected, double* inter, uint8_t* matched_out)
{
    for (int f = 0; f < n_frames; ++f) {
        const int d0 = det_first[f], m0 = map_first[f];
        plane_match_frame(w2c + 16 * size_t(f), det + d0, det_first[f + 1] - d0, det_xy, map + m0, map_first[f + 1] - m0, map_xy,
                          det_matched ? det_matched + d0 : nullptr, advanced_search, sequential, selected + m0, inter + m0,
                          matched_out ? matched_out + d0 : nullptr);
    }
    return 0;
}

int orc_rectify_depth(int W, int H, double fx, double fy, double cx, double cy, const double* cam2_to_cam1, const float* depth,
                      int batch, float* out)
{
    CapeConfig cfg;
    cfg.width = W, cfg.height = H, cfg.fx = fx, cfg.fy = fy, cfg.cx = cx, cfg.cy = cy;
    for (int b = 0; b < batch; ++b) rectify_depth(cfg, cam2_to_cam1, depth + size_t(b) * W * H, out + size_t(b) * W * H);
    return 0;
}

int orc_cape_run(int W, int H, int cell, double fx, double fy, double cx, double cy, const float* depth, int batch,
                 uint32_t seed, int max_boundary, const rs_cape_outputs* out)
{
    CapeConfig cfg;
    cfg.width = W, cfg.height = H, cfg.cell = cell, cfg.fx = fx, cfg.fy = fy, cfg.cx = cx, cfg.cy = cy;
    const int Nc = (W / cell) * (H / cell);
    for (int b = 0; b < batch; ++b) {
        CapeFrame f;
        cape_run(cfg, depth + size_t(b) * W * H, seed, f);
        store_frame(f, Nc, max_boundary, b, out);
    }
    return 0;
}

int orc_cape_cell_fit(int W, int H, int cell, double fx, double fy, double cx, double cy, const float* depth, int batch,
                      rs_cell_out* cells, float* cloud)
{
    CapeConfig cfg;
    cfg.width = W, cfg.height = H, cfg.cell = cell, cfg.fx = fx, cfg.fy = fy, cfg.cx = cx, cfg.cy = cy;
    const int Nc = (W / cell) * (H / cell);
    for (int b = 0; b < batch; ++b) {
        std::vector<PlaneSeg> grid;
        std::vector<float> tols, cl;
        cape_cell_fit(cfg, depth + size_t(b) * W * H, grid, tols, cloud ? &cl : nullptr);
        for (int i = 0; i < Nc; ++i) cell_record(grid[i], tols[i], cell, cells[size_t(b) * Nc + i]);
        if (cloud) std::memcpy(cloud + size_t(b) * 3 * W * H, cl.data(), sizeof(float) * cl.size());
    }
    return 0;
}

void orc_morphology(const unsigned char* mask, int rows, int cols, int erode, int cross, int border_zero, unsigned char* out)
{
    const std::vector<unsigned char> m(mask, mask + size_t(rows) * cols);
    const std::vector<unsigned char> r = cape_morphology(m, rows, cols, erode != 0, cross != 0, border_zero != 0);
    std::copy(r.begin(), r.end(), out);
}

void orc_eigen3(const double a[9], double evals[3], double evecs[9])
{
    Mat3 m, v;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) m(i, j) = a[i * 3 + j];
    self_adjoint_eigen3(m, evals, v);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) evecs[i * 3 + j] = v(i, j);
}

void orc_world_to_camera(const double pose[7], double w2c[16], double plane_w2c[16])
{
    const Mat4 m = world_to_camera(pose + 3, pose);
    const Mat4 p = plane_world_to_camera(m);
    std::memcpy(w2c, &m.m[0][0], sizeof(double) * 16);
    std::memcpy(plane_w2c, &p.m[0][0], sizeof(double) * 16);
}

void orc_pose_coefficients(const double pose[7], double x[6])
{
    Pose7 p;
    for (int i = 0; i < 3; ++i) p.t[i] = pose[i];
    for (int i = 0; i < 4; ++i) p.q[i] = pose[3 + i];
    coefficients_from_pose(p, x);
}
void orc_pose_from_coefficients(const double x[6], double pose[7], double v6[6])
{
    const Pose7 p = pose_from_coefficients(x);
    for (int i = 0; i < 3; ++i) pose[i] = p.t[i];
    for (int i = 0; i < 4; ++i) pose[3 + i] = p.q[i];
    pose_vector6(p, v6);
}
void orc_quaternion_from_euler(double yaw, double pitch, double roll, double q[4]) { quaternion_from_euler(yaw, pitch, roll, q); }

int orc_residual_count(const rs_match* m, int n)
{
    std::vector<rs_match> f(m, m + n);
    return residual_count(f);
}
void orc_pose_residuals(const double K[4], const rs_match* m, int n, const double x[6], double* fvec)
{
    Intrinsics I{K[0], K[1], K[2], K[3]};
    std::vector<rs_match> f(m, m + n);
    pose_residuals(I, f, x, fvec);
}
// LM on the pose residuals from the coefficient vector x (in/out). Returns the Eigen status; nfev in *nfev.
int orc_pose_lm(const double K[4], const rs_match* m, int n, double x[6], int maxfev, int* nfev)
{
    Intrinsics I{K[0], K[1], K[2], K[3]};
    std::vector<rs_match> f(m, m + n);
    const int cnt = residual_count(f);
    const ResidualFn fn = [&](const double* xx, double* fvec) { pose_residuals(I, f, xx, fvec); };
    const LMResult r = lm_minimize(fn, cnt, x, maxfev);
    if (nfev) *nfev = r.nfev;
    return r.status;
}

// IOptimizationFeature::is_inlier of every match under a pose (features as given: no normalisation)
void orc_pose_inliers(const double K[4], const double pose[7], const rs_match* m, int n, uint8_t* mask)
{
    Intrinsics I{K[0], K[1], K[2], K[3]};
    const Mat4 w2c = world_to_camera(pose + 3, pose);
    const Mat4 pw2c = plane_world_to_camera(w2c);
    for (int i = 0; i < n; ++i) mask[i] = feature_is_inlier(I, m[i], w2c, pw2c) ? 1 : 0;
}

int orc_ransac_default_iterations() { return ransac_default_iterations(); }

// One compute_optimized_pose. subsets/normals: optional explicit random inputs (layout of rs_pose_export_random,
// for ONE frame). cand_* (optional): per-iteration taps [max_iterations].
int orc_pose_solve(const double K[4], const double cur_pose[7], const rs_match* m, int n, int max_iterations,
                   int n_variance, uint32_t seed, const int32_t* subsets, const double* normals, int max_matches,
                   int lm_max_fev, rs_pose_out* out, uint8_t* inlier_mask, double* cand_poses, int32_t* cand_ok,
                   double* cand_scores, int32_t* subsets_out)
{
    Intrinsics I{K[0], K[1], K[2], K[3]};
    Pose7 cur;
    for (int i = 0; i < 3; ++i) cur.t[i] = cur_pose[i];
    for (int i = 0; i < 4; ++i) cur.q[i] = cur_pose[3 + i];
    std::vector<rs_match> f(m, m + n);
    PoseRandom rnd(seed);
    rnd.subsets = subsets;
    rnd.normals = normals;
    rnd.max_matches = max_matches;
    const PoseSolveResult r = pose_solve(I, cur, f, max_iterations, n_variance, rnd, lm_max_fev > 0 ? lm_max_fev : 400);
    *out = r.out;
    if (inlier_mask)
        for (int i = 0; i < n; ++i) inlier_mask[i] = r.inlier_mask[i];
    for (size_t it = 0; it < r.candidate_ok.size(); ++it) {
        if (cand_ok) cand_ok[it] = r.candidate_ok[it];
        if (cand_scores) cand_scores[it] = r.candidate_scores[it];
        if (cand_poses) {
            for (int i = 0; i < 3; ++i) cand_poses[it * 7 + i] = r.candidate_poses[it].t[i];
            for (int i = 0; i < 4; ++i) cand_poses[it * 7 + 3 + i] = r.candidate_poses[it].q[i];
        }
        if (subsets_out) {
            for (int k = 0; k < RS_MAX_SUBSET; ++k)
                subsets_out[it * RS_MAX_SUBSET + k] = k < int(r.subsets[it].size()) ? r.subsets[it][k] : -1;
        }
    }
    return 0;
}

// CPU-baseline driver: full frames (CAPE + pose solve) over `n_threads` host threads, frames split statically.
// Returns elapsed seconds. do_cape / do_pose select the stages. Outputs are discarded except poses (7/frame).
double orc_process_frames(int W, int H, int cell, const double K[4], const float* depth, const double* cur_pose,
                          const rs_match* matches, const int32_t* n_matches, int max_matches, int batch, int do_cape,
                          int do_pose, int max_iterations, int n_variance, uint32_t seed, int n_threads, double* poses_out)
{
    CapeConfig cfg;
    cfg.width = W, cfg.height = H, cfg.cell = cell, cfg.fx = K[0], cfg.fy = K[1], cfg.cx = K[2], cfg.cy = K[3];
    Intrinsics I{K[0], K[1], K[2], K[3]};
    if (n_threads < 1) n_threads = 1;
    auto work = [&](int t) {
        for (int b = t; b < batch; b += n_threads) {
            if (do_cape) {
                CapeFrame f;
                cape_run(cfg, depth + size_t(b) * W * H, seed, f);
            }
            if (do_pose) {
                Pose7 cur;
                for (int i = 0; i < 3; ++i) cur.t[i] = cur_pose[b * 7 + i];
                for (int i = 0; i < 4; ++i) cur.q[i] = cur_pose[b * 7 + 3 + i];
                std::vector<rs_match> f(matches + size_t(b) * max_matches, matches + size_t(b) * max_matches + n_matches[b]);
                PoseRandom rnd(seed + uint32_t(b));
                const PoseSolveResult r = pose_solve(I, cur, f, max_iterations, n_variance, rnd);
                if (poses_out)
                    for (int i = 0; i < 7; ++i) poses_out[b * 7 + i] = r.out.pose[i];
            }
        }
    };
    const auto t0 = std::chrono::steady_clock::now();
    if (n_threads == 1) {
        work(0);
    }
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
