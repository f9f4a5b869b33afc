// C-ABI glue for the pose path (include/rgbdslam_b200.h): context, device buffers, host-side reference RNG streams.
// No CPU fallback: every entry point needs the sm_100 device the context was created on.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <random>
#include <string>
#include <vector>

#include "pose_internal.cuh"

using namespace rs;

struct rs_pose_ctx {
    int max_batch, M, max_iterations, max_variance, device;
    rs_match* d_matches = nullptr;
    double* d_cur = nullptr;
    int32_t* d_n = nullptr;
    int32_t* d_subsets_in = nullptr;
    double* d_normals = nullptr;  // lazily allocated: B x max_variance x M x 4
    PoseBuffers buf{};
    cudaStream_t stream = nullptr;
    // host mirrors kept for the reference RNG mode and for export
    std::vector<int32_t> h_n;
    std::vector<int32_t> h_type;  // B x M
    bool has_point2d = false;     // the uploaded batch carries an inverse-depth (RS_FEAT_POINT2D) feature
    PoseLaunch last{};
    int last_batch = 0;
    std::vector<cudaEvent_t> events;  // 5 per timing slot
    cudaEvent_t ransac_done = nullptr;   // recorded after the RANSAC + final LM kernel (rs_pose_stream_wait_ransac)
    // rs_pose_opts::sub_batches > 1: side streams of the frame groups 1.. (group 0 runs on the caller's stream)
    static constexpr int kMaxGroups = 8;
    cudaStream_t group_stream[kMaxGroups - 1] = {};
    cudaEvent_t group_fork = nullptr, group_ransac[kMaxGroups - 1] = {}, group_done[kMaxGroups - 1] = {};
    int groups_last = 1;                 // groups of the most recent solve (how many group_ransac events are live)
    int timing_slots = 0;
    uint64_t run_counter = 0;
};

namespace {

template <class T>
int dev_alloc(T** p, size_t n)
{
    RS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(p), sizeof(T) * std::max<size_t>(n, 1)));
    RS_CUDA_CHECK(cudaMemset(*p, 0, sizeof(T) * std::max<size_t>(n, 1)));
    return RS_OK;
}

int create_impl(rs_pose_ctx* c)
{
    int rc = require_blackwell(c->device);
    if (rc != RS_OK) return rc;
    const size_t B = size_t(c->max_batch), M = size_t(c->M);
    PoseBuffers& b = c->buf;
    b.max_matches = c->M, b.max_iterations = c->max_iterations, b.max_variance = c->max_variance;
    if ((rc = dev_alloc(&c->d_matches, B * M))) return rc;
    if ((rc = dev_alloc(&c->d_cur, B * 7))) return rc;
    if ((rc = dev_alloc(&c->d_n, B))) return rc;
    if ((rc = dev_alloc(&c->d_subsets_in, B * size_t(c->max_iterations) * RS_MAX_SUBSET))) return rc;
    if ((rc = dev_alloc(&b.type, B * M))) return rc;
    if ((rc = dev_alloc(&b.obs, B * 4 * M))) return rc;
    if ((rc = dev_alloc(&b.map, B * 4 * M))) return rc;
    if ((rc = dev_alloc(&b.sigma, B * 4 * M))) return rc;
    if ((rc = dev_alloc(&b.aux, B * 4 * M))) return rc;
    if ((rc = dev_alloc(&b.state, B))) return rc;
    if ((rc = dev_alloc(&b.out, B))) return rc;
    if ((rc = dev_alloc(&b.mask, B * M))) return rc;
    if ((rc = dev_alloc(&b.inlier_idx, B * M))) return rc;
    if ((rc = dev_alloc(&b.poses, B * 7))) return rc;
    if ((rc = dev_alloc(&b.subsets_used, B * size_t(c->max_iterations) * RS_MAX_SUBSET))) return rc;
    if ((rc = dev_alloc(&b.v6, B * size_t(c->max_variance) * 6))) return rc;
    if ((rc = dev_alloc(&b.v_ok, B * size_t(c->max_variance)))) return rc;
    b.matches_aos = c->d_matches, b.cur_pose = c->d_cur, b.n_matches = c->d_n;
    b.subsets_in = nullptr, b.normals_in = nullptr;
    RS_CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    RS_CUDA_CHECK(cudaEventCreateWithFlags(&c->ransac_done, cudaEventDisableTiming));
    RS_CUDA_CHECK(cudaEventCreateWithFlags(&c->group_fork, cudaEventDisableTiming));
    for (int g = 0; g < rs_pose_ctx::kMaxGroups - 1; ++g) {
        RS_CUDA_CHECK(cudaStreamCreateWithFlags(&c->group_stream[g], cudaStreamNonBlocking));
        RS_CUDA_CHECK(cudaEventCreateWithFlags(&c->group_ransac[g], cudaEventDisableTiming));
        RS_CUDA_CHECK(cudaEventCreateWithFlags(&c->group_done[g], cudaEventDisableTiming));
    }
    c->h_n.assign(B, 0);
    c->h_type.assign(B * M, 0);
    return RS_OK;
}

int ensure_normals(rs_pose_ctx* c)
{
    if (c->d_normals) return RS_OK;
    return dev_alloc(&c->d_normals, size_t(c->max_batch) * c->max_variance * c->M * 4);
}

int resolve_launch(const rs_pose_ctx* c, const rs_pose_opts* opts, int batch, PoseLaunch& prm)
{
    rs_pose_opts o{};
    if (opts) o = *opts;
    prm.batch = batch;
    // maximumIterations = ceil(logf(1 - 0.8f) / logf(1 - powf(0.65f, 10.0f))) = 119 (pose_optimization.cpp:129-132)
    prm.max_iterations = o.max_iterations > 0 ? o.max_iterations : 119;
    prm.n_variance = o.n_variance < 0 ? 100 : o.n_variance;
    prm.lm_max_fev = o.lm_max_fev > 0 ? o.lm_max_fev : 400;
    prm.rng_mode = o.rng_mode;
    prm.seed = o.seed;
    prm.has_point2d = c->has_point2d ? 1 : 0;
    prm.sub_batches = o.sub_batches < 1 ? 1 : (o.sub_batches > rs_pose_ctx::kMaxGroups ? rs_pose_ctx::kMaxGroups : o.sub_batches);
    if (o.fx == 0 && o.fy == 0 && o.cx == 0 && o.cy == 0)
        prm.K = PoseIntrinsics{550.0, 550.0, 320.0, 240.0};  // Parameters::load_defaut (parameters.cpp:59-74)
    else
        prm.K = PoseIntrinsics{o.fx, o.fy, o.cx, o.cy};
    if (prm.max_iterations > c->max_iterations || prm.n_variance > c->max_variance) {
        set_last_error("rs_pose: max_iterations / n_variance exceed the capacities given to rs_pose_create");
        return RS_ERR_INVALID_ARG;
    }
    if (prm.rng_mode != RS_RNG_REFERENCE && prm.rng_mode != RS_RNG_DEVICE) {
        set_last_error("rs_pose: unknown rng_mode");
        return RS_ERR_INVALID_ARG;
    }
    return RS_OK;
}

int upload_impl(rs_pose_ctx* c, const double* cur_pose, const rs_match* matches, const int32_t* n_matches, int batch,
                cudaStream_t s, bool host_types = true)
{
    if (!c || !cur_pose || !matches || !n_matches || batch <= 0 || batch > c->max_batch) {
        set_last_error("rs_pose: invalid argument (null pointer or batch out of range)");
        return RS_ERR_INVALID_ARG;
    }
    for (int b = 0; b < batch; ++b)
        if (n_matches[b] < 0 || n_matches[b] > c->M) {
            set_last_error("rs_pose: n_matches out of range");
            return RS_ERR_INVALID_ARG;
        }
    // one pass over the match lists: feature types must be ones both sides of the solve know (the device scoring and the
    // host-drawn reference subsets would otherwise disagree silently), and the batch says which kernel instantiation runs
    bool p2d = false;
    for (int b = 0; b < batch; ++b)
        for (int i = 0; i < n_matches[b]; ++i) {
            const int32_t ty = matches[size_t(b) * c->M + i].type;
            if (ty != RS_FEAT_POINT && ty != RS_FEAT_PLANE && ty != RS_FEAT_POINT2D) {
                set_last_error("rs_pose: rs_match.type " + std::to_string(ty) + " of frame " + std::to_string(b) + ", match " +
                               std::to_string(i) + " is not RS_FEAT_POINT / RS_FEAT_PLANE / RS_FEAT_POINT2D");
                return RS_ERR_INVALID_ARG;
            }
            p2d = p2d || ty == RS_FEAT_POINT2D;
        }
    RS_CUDA_CHECK(cudaSetDevice(c->device));
    RS_CUDA_CHECK(cudaMemcpyAsync(c->d_cur, cur_pose, sizeof(double) * 7 * batch, cudaMemcpyHostToDevice, s));
    RS_CUDA_CHECK(cudaMemcpyAsync(c->d_matches, matches, sizeof(rs_match) * size_t(batch) * c->M, cudaMemcpyHostToDevice, s));
    RS_CUDA_CHECK(cudaMemcpyAsync(c->d_n, n_matches, sizeof(int32_t) * batch, cudaMemcpyHostToDevice, s));
    c->has_point2d = p2d;
    for (int b = 0; b < batch; ++b) {
        c->h_n[b] = n_matches[b];
        if (host_types)  // only the host-side reference RNG (std::shuffle / normal draws per inlier) reads the types
            for (int i = 0; i < n_matches[b]; ++i) c->h_type[size_t(b) * c->M + i] = matches[size_t(b) * c->M + i].type;
    }
    return RS_OK;
}

inline double feature_score(int type) { return type == RS_FEAT_PLANE ? 1.0 / 3.0 : 1.0 / 5.0; }  // point and point2d: 1/5

// ransac::get_random_subset_with_score (ransac.hpp:77-103) with the reference's engine: std::shuffle over the
// whole list, shuffled prefix until the score reaches 1, every pick prepended.
void reference_subsets(const rs_pose_ctx* c, int b, uint32_t seed, int iterations, int32_t* out /* iterations x 16 */)
{
    const int n = c->h_n[b];
    const int32_t* type = c->h_type.data() + size_t(b) * c->M;
    std::mt19937 engine(seed);
    std::vector<int> order(n);
    for (int it = 0; it < iterations; ++it) {
        int32_t* dst = out + size_t(it) * RS_MAX_SUBSET;
        std::fill(dst, dst + RS_MAX_SUBSET, -1);
        std::iota(order.begin(), order.end(), 0);
        std::shuffle(order.begin(), order.end(), engine);
        double cumulated = 0.0;
        std::vector<int> picked;
        for (int idx : order) {
            cumulated += feature_score(type[idx]);
            picked.insert(picked.begin(), idx);
            if (cumulated >= 1.0 || int(picked.size()) == RS_MAX_SUBSET) break;
        }
        if (cumulated >= 1.0)
            for (size_t k = 0; k < picked.size(); ++k) dst[k] = picked[k];
    }
}

// Gaussian draws of compute_random_variation_of_pose for one frame, continuing the frame's engine after the
// `started` shuffles the RANSAC loop consumed: for each sample, for each inlier in list order, 3 (point) or
// 4 (plane) std::normal_distribution draws (map_point.cpp:49-58, map_primitive.cpp:66-77).
void reference_normals(const rs_pose_ctx* c, int b, uint32_t seed, int started, int n_variance, const uint8_t* mask,
                       double* out /* n_variance x M x 4 */)
{
    const int n = c->h_n[b];
    const int32_t* type = c->h_type.data() + size_t(b) * c->M;
    std::mt19937 engine(seed);
    std::vector<int> order(n);
    for (int it = 0; it < started; ++it) {
        std::iota(order.begin(), order.end(), 0);
        std::shuffle(order.begin(), order.end(), engine);
    }
    std::normal_distribution<double> normal(0.0, 1.0);
    for (int s = 0; s < n_variance; ++s)
        for (int i = 0; i < n; ++i) {
            if (!mask[i]) continue;
            const int nd = type[i] == RS_FEAT_POINT ? 3 : (type[i] == RS_FEAT_POINT2D ? 2 : 4);
            double* dst = out + (size_t(s) * c->M + i) * 4;
            for (int k = 0; k < nd; ++k) dst[k] = normal(engine);
        }
}

int solve_impl(rs_pose_ctx* c, int batch, const PoseLaunch& prm, cudaStream_t s, bool reference_rng)
{
    int rc;
    PoseBuffers buf = c->buf;
    std::vector<int32_t> h_subsets;
    if (reference_rng) {
        h_subsets.resize(size_t(batch) * c->max_iterations * RS_MAX_SUBSET, -1);
        for (int b = 0; b < batch; ++b)
            reference_subsets(c, b, prm.seed + uint32_t(b), prm.max_iterations,
                              h_subsets.data() + size_t(b) * c->max_iterations * RS_MAX_SUBSET);
        RS_CUDA_CHECK(cudaMemcpyAsync(c->d_subsets_in, h_subsets.data(), sizeof(int32_t) * h_subsets.size(),
                                      cudaMemcpyHostToDevice, s));
        buf.subsets_in = c->d_subsets_in;
    }
    cudaEvent_t* ev = c->timing_slots > 0 ? &c->events[size_t(c->run_counter % uint64_t(c->timing_slots)) * 5] : nullptr;
    ++c->run_counter;
    const int groups = (!reference_rng && prm.n_variance > 0) ? std::min(prm.sub_batches, batch) : 1;
    c->groups_last = groups;
    if (groups > 1) {
        // Frame groups: the RANSAC kernel of the whole batch lasts as long as its slowest frame (a latency chain), and the
        // throughput-bound Monte-Carlo kernel cannot start before it ends. Split into groups of frames whose
        // prepare -> RANSAC -> Monte-Carlo -> covariance chains run on their own streams, a group's Monte-Carlo solves start as
        // soon as ITS slowest frame is through and fill the SMs while the other groups' RANSAC chains are still running.
        // Frames are independent and the random draws are keyed by the frame index: the results do not change.
        // Timing slots in this mode: [0,1] prepare of group 0, [1,2] RANSAC of group 0, [2,3] first group's Monte-Carlo
        // kernel ... the whole solve is [0,4] (ev[4] is recorded after the join).
        RS_CUDA_CHECK(cudaEventRecord(c->group_fork, s));
        const int per = (batch + groups - 1) / groups;
        for (int g = 0; g < groups; ++g) {
            cudaStream_t gs = g == 0 ? s : c->group_stream[g - 1];
            PoseLaunch gp = prm;
            gp.frame0 = g * per;
            gp.batch = std::min(per, batch - gp.frame0);
            if (gp.batch <= 0) {
                c->groups_last = g;
                break;
            }
            if (g > 0) RS_CUDA_CHECK(cudaStreamWaitEvent(gs, c->group_fork, 0));
            if (g == 0 && ev) RS_CUDA_CHECK(cudaEventRecord(ev[0], gs));
            if ((rc = launch_pose_prepare(buf, gp, gs)) != RS_OK) return rc;
            if (g == 0 && ev) RS_CUDA_CHECK(cudaEventRecord(ev[1], gs));
            if ((rc = launch_pose_ransac(buf, gp, gs)) != RS_OK) return rc;
            RS_CUDA_CHECK(cudaEventRecord(g == 0 ? c->ransac_done : c->group_ransac[g - 1], gs));
            if (g == 0 && ev) RS_CUDA_CHECK(cudaEventRecord(ev[2], gs));
            if ((rc = launch_pose_variance(buf, gp, gs)) != RS_OK) return rc;
            if (g == 0 && ev) RS_CUDA_CHECK(cudaEventRecord(ev[3], gs));
            if ((rc = launch_pose_covariance(buf, gp, gs)) != RS_OK) return rc;
            if (g > 0) RS_CUDA_CHECK(cudaEventRecord(c->group_done[g - 1], gs));
        }
        for (int g = 1; g < c->groups_last; ++g) RS_CUDA_CHECK(cudaStreamWaitEvent(s, c->group_done[g - 1], 0));
        if (ev) RS_CUDA_CHECK(cudaEventRecord(ev[4], s));
        c->last = prm;
        c->last_batch = batch;
        return RS_OK;
    }
    if (ev) RS_CUDA_CHECK(cudaEventRecord(ev[0], s));
    if ((rc = launch_pose_prepare(buf, prm, s)) != RS_OK) return rc;
    if (ev) RS_CUDA_CHECK(cudaEventRecord(ev[1], s));
    if ((rc = launch_pose_ransac(buf, prm, s)) != RS_OK) return rc;
    RS_CUDA_CHECK(cudaEventRecord(c->ransac_done, s));
    if (ev) {
        RS_CUDA_CHECK(cudaEventRecord(ev[2], s));
        if (prm.n_variance <= 0) {
            RS_CUDA_CHECK(cudaEventRecord(ev[3], s));
            RS_CUDA_CHECK(cudaEventRecord(ev[4], s));
        }
    }
    if (prm.n_variance > 0) {
        std::vector<double> h_normals;
        if (reference_rng) {
            // the Gaussian stream continues where the RANSAC loop stopped: needs iterations_run and the inlier mask
            std::vector<rs_pose_out> h_out(batch);
            std::vector<uint8_t> h_mask(size_t(batch) * c->M);
            RS_CUDA_CHECK(cudaMemcpyAsync(h_out.data(), buf.out, sizeof(rs_pose_out) * batch, cudaMemcpyDeviceToHost, s));
            RS_CUDA_CHECK(cudaMemcpyAsync(h_mask.data(), buf.mask, h_mask.size(), cudaMemcpyDeviceToHost, s));
            RS_CUDA_CHECK(cudaStreamSynchronize(s));
            if ((rc = ensure_normals(c)) != RS_OK) return rc;
            h_normals.assign(size_t(batch) * prm.n_variance * c->M * 4, 0.0);
            for (int b = 0; b < batch; ++b)
                if (h_out[b].status == -2)  // final pose available, covariance pending
                    reference_normals(c, b, prm.seed + uint32_t(b), h_out[b].iterations_run, prm.n_variance,
                                      h_mask.data() + size_t(b) * c->M,
                                      h_normals.data() + size_t(b) * prm.n_variance * c->M * 4);
            RS_CUDA_CHECK(cudaMemcpyAsync(c->d_normals, h_normals.data(), sizeof(double) * h_normals.size(),
                                          cudaMemcpyHostToDevice, s));
            buf.normals_in = c->d_normals;
        }
        if (ev && reference_rng) RS_CUDA_CHECK(cudaEventRecord(ev[2], s));  // exclude the host RNG round trip
        if ((rc = launch_pose_variance(buf, prm, s)) != RS_OK) return rc;
        if (ev) RS_CUDA_CHECK(cudaEventRecord(ev[3], s));
        if ((rc = launch_pose_covariance(buf, prm, s)) != RS_OK) return rc;
        if (ev) RS_CUDA_CHECK(cudaEventRecord(ev[4], s));
        if (reference_rng) RS_CUDA_CHECK(cudaStreamSynchronize(s));  // h_normals must outlive the copy
    }
    c->last = prm;
    c->last_batch = batch;
    return RS_OK;
}

int download_impl(rs_pose_ctx* c, int batch, rs_pose_out* out, uint8_t* inlier_mask, cudaStream_t s)
{
    if (!c || !out || batch <= 0 || batch > c->max_batch) {
        set_last_error("rs_pose_download: invalid argument");
        return RS_ERR_INVALID_ARG;
    }
    RS_CUDA_CHECK(cudaMemcpyAsync(out, c->buf.out, sizeof(rs_pose_out) * batch, cudaMemcpyDeviceToHost, s));
    if (inlier_mask)
        RS_CUDA_CHECK(cudaMemcpyAsync(inlier_mask, c->buf.mask, size_t(batch) * c->M, cudaMemcpyDeviceToHost, s));
    RS_CUDA_CHECK(cudaStreamSynchronize(s));
    return RS_OK;
}

}  // namespace

extern "C" {

rs_pose_ctx* rs_pose_create(int max_batch, int max_matches, int max_iterations, int max_variance, int device)
{
    if (max_batch <= 0 || max_matches <= 0 || max_matches > pose_max_matches_supported()) {
        set_last_error("rs_pose_create: invalid capacities (max_matches must be in 1.." + std::to_string(pose_max_matches_supported()) +
                       ": one frame's match list is staged in shared memory)");
        return nullptr;
    }
    rs_pose_ctx* c = new rs_pose_ctx();
    c->max_batch = max_batch, c->M = max_matches, c->device = device;
    c->max_iterations = max_iterations > 0 ? max_iterations : 119;
    c->max_variance = max_variance >= 0 ? std::max(max_variance, 1) : 100;
    if (create_impl(c) != RS_OK) {
        rs_pose_destroy(c);
        return nullptr;
    }
    return c;
}

void rs_pose_destroy(rs_pose_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaFree(c->d_matches);
    cudaFree(c->d_cur);
    cudaFree(c->d_n);
    cudaFree(c->d_subsets_in);
    cudaFree(c->d_normals);
    PoseBuffers& b = c->buf;
    cudaFree(b.type);
    cudaFree(b.obs);
    cudaFree(b.map);
    cudaFree(b.sigma);
    cudaFree(b.aux);
    cudaFree(b.state);
    cudaFree(b.out);
    cudaFree(b.mask);
    cudaFree(b.inlier_idx);
    cudaFree(b.poses);
    cudaFree(b.subsets_used);
    cudaFree(b.v6);
    cudaFree(b.v_ok);
    for (cudaEvent_t e : c->events) cudaEventDestroy(e);
    if (c->ransac_done) cudaEventDestroy(c->ransac_done);
    if (c->group_fork) cudaEventDestroy(c->group_fork);
    for (int g = 0; g < rs_pose_ctx::kMaxGroups - 1; ++g) {
        if (c->group_ransac[g]) cudaEventDestroy(c->group_ransac[g]);
        if (c->group_done[g]) cudaEventDestroy(c->group_done[g]);
        if (c->group_stream[g]) cudaStreamDestroy(c->group_stream[g]);
    }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int rs_pose_set_timing(rs_pose_ctx* c, int n_slots)
{
    if (!c || n_slots < 0 || n_slots > 4096) {
        set_last_error("rs_pose_set_timing: invalid argument");
        return RS_ERR_INVALID_ARG;
    }
    RS_CUDA_CHECK(cudaSetDevice(c->device));
    for (cudaEvent_t e : c->events) cudaEventDestroy(e);
    c->events.assign(size_t(n_slots) * 5, nullptr);
    for (cudaEvent_t& e : c->events) RS_CUDA_CHECK(cudaEventCreate(&e));
    c->timing_slots = n_slots;
    c->run_counter = 0;
    return RS_OK;
}

int rs_pose_kernel_ms(rs_pose_ctx* c, int slot, float ms[4])
{
    if (!c || !ms || slot < 0 || slot >= c->timing_slots || uint64_t(slot) >= c->run_counter) {
        set_last_error("rs_pose_kernel_ms: timing is off or that slot has not been recorded");
        return RS_ERR_INVALID_ARG;
    }
    cudaEvent_t* ev = &c->events[size_t(slot) * 5];
    RS_CUDA_CHECK(cudaEventSynchronize(ev[4]));
    for (int k = 0; k < 4; ++k) RS_CUDA_CHECK(cudaEventElapsedTime(&ms[k], ev[k], ev[k + 1]));
    return RS_OK;
}

int rs_pose_solve_batched_begin(rs_pose_ctx* c, const double* cur_pose, const rs_match* matches, const int32_t* n_matches,
                                int batch, const rs_pose_opts* opts, rs_pose_out* out, uint8_t* inlier_mask)
{
    if (!c || !out) {
        set_last_error("rs_pose_solve_batched: null context or output");
        return RS_ERR_INVALID_ARG;
    }
    PoseLaunch prm;
    int rc = resolve_launch(c, opts, batch, prm);
    if (rc != RS_OK) return rc;
    const bool reference_rng = prm.rng_mode == RS_RNG_REFERENCE;
    if ((rc = upload_impl(c, cur_pose, matches, n_matches, batch, c->stream, reference_rng)) != RS_OK) return rc;
    prm.has_point2d = c->has_point2d ? 1 : 0;   // known only once this batch's feature types have been seen
    if ((rc = solve_impl(c, batch, prm, c->stream, reference_rng)) != RS_OK) return rc;
    RS_CUDA_CHECK(cudaMemcpyAsync(out, c->buf.out, sizeof(rs_pose_out) * batch, cudaMemcpyDeviceToHost, c->stream));
    if (inlier_mask)
        RS_CUDA_CHECK(cudaMemcpyAsync(inlier_mask, c->buf.mask, size_t(batch) * c->M, cudaMemcpyDeviceToHost, c->stream));
    return RS_OK;
}

int rs_pose_solve_batched_end(rs_pose_ctx* c)
{
    if (!c) {
        set_last_error("rs_pose_solve_batched_end: null context");
        return RS_ERR_INVALID_ARG;
    }
    RS_CUDA_CHECK(cudaSetDevice(c->device));
    RS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return RS_OK;
}

int rs_pose_solve_batched(rs_pose_ctx* c, const double* cur_pose, const rs_match* matches, const int32_t* n_matches,
                          int batch, const rs_pose_opts* opts, rs_pose_out* out, uint8_t* inlier_mask)
{
    const int rc = rs_pose_solve_batched_begin(c, cur_pose, matches, n_matches, batch, opts, out, inlier_mask);
    if (rc != RS_OK) return rc;
    return rs_pose_solve_batched_end(c);
}

int rs_pose_solve(rs_pose_ctx* c, const double cur_pose[7], const rs_match* matches, int n_matches,
                  const rs_pose_opts* opts, rs_pose_out* out, uint8_t* inlier_mask)
{
    if (!c || !matches || n_matches < 0 || n_matches > c->M) {
        set_last_error("rs_pose_solve: invalid argument");
        return RS_ERR_INVALID_ARG;
    }
    // the batched entry point expects a max_matches stride: stage the frame's matches in a padded row
    std::vector<rs_match> padded(size_t(c->M));
    std::memcpy(padded.data(), matches, sizeof(rs_match) * size_t(n_matches));
    const int32_t n = n_matches;
    std::vector<uint8_t> mask(size_t(c->M));
    const int rc = rs_pose_solve_batched(c, cur_pose, padded.data(), &n, 1, opts, out, mask.data());
    if (rc == RS_OK && inlier_mask) std::memcpy(inlier_mask, mask.data(), size_t(n_matches));
    return rc;
}

int rs_pose_upload(rs_pose_ctx* c, const double* cur_pose, const rs_match* matches, const int32_t* n_matches, int batch)
{
    if (!c) {
        set_last_error("rs_pose_upload: null context");
        return RS_ERR_INVALID_ARG;
    }
    const int rc = upload_impl(c, cur_pose, matches, n_matches, batch, c->stream);
    if (rc != RS_OK) return rc;
    RS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return RS_OK;
}

int rs_pose_solve_device(rs_pose_ctx* c, int batch, const rs_pose_opts* opts, void* stream)
{
    if (!c || batch <= 0 || batch > c->max_batch) {
        set_last_error("rs_pose_solve_device: invalid argument");
        return RS_ERR_INVALID_ARG;
    }
    PoseLaunch prm;
    int rc = resolve_launch(c, opts, batch, prm);
    if (rc != RS_OK) return rc;
    if (prm.rng_mode != RS_RNG_DEVICE) {
        set_last_error("rs_pose_solve_device: only RS_RNG_DEVICE runs without host round trips");
        return RS_ERR_INVALID_ARG;
    }
    RS_CUDA_CHECK(cudaSetDevice(c->device));
    return solve_impl(c, batch, prm, static_cast<cudaStream_t>(stream), false);
}

int rs_pose_stream_wait_ransac(rs_pose_ctx* c, void* stream)
{
    if (!c) {
        set_last_error("rs_pose_stream_wait_ransac: null context");
        return RS_ERR_INVALID_ARG;
    }
    RS_CUDA_CHECK(cudaSetDevice(c->device));
    RS_CUDA_CHECK(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), c->ransac_done, 0));
    for (int g = 1; g < c->groups_last; ++g)
        RS_CUDA_CHECK(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), c->group_ransac[g - 1], 0));
    return RS_OK;
}

int rs_pose_download(rs_pose_ctx* c, int batch, rs_pose_out* out, uint8_t* inlier_mask)
{
    if (!c) {
        set_last_error("rs_pose_download: null context");
        return RS_ERR_INVALID_ARG;
    }
    RS_CUDA_CHECK(cudaSetDevice(c->device));
    RS_CUDA_CHECK(cudaDeviceSynchronize());
    return download_impl(c, batch, out, inlier_mask, c->stream);
}

double* rs_pose_device_poses(rs_pose_ctx* c) { return c ? c->buf.poses : nullptr; }

int rs_pose_export_random(rs_pose_ctx* c, int batch, int32_t* subsets, double* normals)
{
    if (!c || batch <= 0 || batch > c->max_batch || batch > c->last_batch) {
        set_last_error("rs_pose_export_random: no solve of that batch size has run");
        return RS_ERR_INVALID_ARG;
    }
    RS_CUDA_CHECK(cudaSetDevice(c->device));
    RS_CUDA_CHECK(cudaDeviceSynchronize());
    if (subsets) {
        // caller layout: B x last.max_iterations x RS_MAX_SUBSET ; device layout uses the context capacity
        for (int b = 0; b < batch; ++b)
            RS_CUDA_CHECK(cudaMemcpy(subsets + size_t(b) * c->last.max_iterations * RS_MAX_SUBSET,
                                     c->buf.subsets_used + size_t(b) * c->max_iterations * RS_MAX_SUBSET,
                                     sizeof(int32_t) * size_t(c->last.max_iterations) * RS_MAX_SUBSET, cudaMemcpyDeviceToHost));
    }
    if (normals && c->last.n_variance > 0) {
        int rc = ensure_normals(c);
        if (rc != RS_OK) return rc;
        PoseLaunch prm = c->last;
        prm.batch = batch;
        if ((rc = launch_pose_export_normals(c->buf, prm, c->d_normals, c->stream)) != RS_OK) return rc;
        RS_CUDA_CHECK(cudaMemcpyAsync(normals, c->d_normals, sizeof(double) * size_t(batch) * prm.n_variance * c->M * 4,
                                      cudaMemcpyDeviceToHost, c->stream));
        RS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    return RS_OK;
}

}  // extern "C"
