// C-ABI glue for the pose path (include/rgbdslam_b200.h): context, device buffers, host-side reference RNG streams.
// No CPU fallback: every entry point needs the sm_100 device the context was created on.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "pose_internal.cuh"

using namespace rs;

struct rs_pose_ctx {
    int max_batch, M, max_iterations, max_variance, device;
    int sm_count = 148;
    rs_match* d_matches = nullptr;
    double* d_cur = nullptr;
    int32_t* d_n = nullptr;
    int32_t* d_subsets_in = nullptr;
    double* d_normals = nullptr;  // lazily allocated: B x max_variance x M x 4
    double* h_normals = nullptr;  // pinned staging of the same size (RS_RNG_REFERENCE)
    std::vector<int32_t> h_subsets;
    std::vector<rs_pose_out> h_out;
    std::vector<uint8_t> h_mask;
    PoseBuffers buf{};
    cudaStream_t stream = nullptr;
    // host mirrors kept for the reference RNG mode and for export
    std::vector<int32_t> h_n;
    std::vector<int32_t> h_type;  // B x M
    bool has_point2d = false;     // the uploaded batch carries an inverse-depth (RS_FEAT_POINT2D) feature
    PoseLaunch last{};
    int last_batch = 0;
    bool last_fused = false;             // the most recent solve ran the fused kernel (rs_pose_add_workers / rs_pose_phase_ms apply)
    bool last_mc_only = false;           // the most recent solve-kernel launch was the second half of an RS_RNG_REFERENCE solve
    std::vector<cudaEvent_t> events;  // 5 per timing slot
    cudaEvent_t ransac_done = nullptr;   // recorded after the RANSAC + final LM kernel (rs_pose_stream_wait_ransac)
    // rs_pose_opts::sub_batches > 1: side streams of the frame groups 1.. (group 0 runs on the caller's stream)
    static constexpr int kMaxGroups = 8;
    cudaStream_t group_stream[kMaxGroups - 1] = {};
    cudaEvent_t group_fork = nullptr, group_ransac[kMaxGroups - 1] = {}, group_done[kMaxGroups - 1] = {};
    int groups_last = 1;                 // groups of the most recent solve (how many group_ransac events are live)
    int timing_slots = 0;
    uint64_t run_counter = 0;
    int prepared_batch = 0;              // rs_pose_prepare_device ran for this batch size and no solve has consumed it yet
    cudaEvent_t prepared = nullptr;      // recorded after that preparation kernel
};

namespace {

template <class T>
int dev_alloc(T** p, size_t n)
{
    RS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(p), sizeof(T) * std::max<size_t>(n, 1)));
    RS_CUDA_CHECK(cudaMemset(*p, 0, sizeof(T) * std::max<size_t>(n, 1)));
    return RS_OK;
}

constexpr int kWideMinIterations = 256;   // beyond this many hypotheses per frame the one-hypothesis-per-lane kernel is the default

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
    if ((rc = dev_alloc(&b.work, 1))) return rc;
    if ((rc = dev_alloc(&b.rframe, B))) return rc;
    if ((rc = dev_alloc(&b.ring_mask, B * size_t(kRansacRing + 1) * ((M + 31) / 32)))) return rc;
    if ((rc = dev_alloc(&b.ready, B))) return rc;
    if ((rc = dev_alloc(&b.open_list, B))) return rc;
    if ((rc = dev_alloc(&b.mc_done, B))) return rc;
    if ((rc = dev_alloc(&b.frame_times, B * 4))) return rc;
    b.hyp = nullptr, b.hyp_mask = nullptr, b.fold = nullptr, b.fold_mask = nullptr;
    if (c->max_iterations > kWideMinIterations && pose_wide_supports(c->M) && pose_chain_supports(c->M)) {
        // one hypothesis per lane (pose_wide.cu): a record and an inlier mask per hypothesis
        if ((rc = dev_alloc(&b.hyp, B * size_t(c->max_iterations)))) return rc;
        if ((rc = dev_alloc(&b.hyp_mask, B * size_t(c->max_iterations) * ((M + 31) / 32)))) return rc;
        if ((rc = dev_alloc(&b.fold, B))) return rc;
        if ((rc = dev_alloc(&b.fold_mask, B * ((M + 31) / 32)))) return rc;
    }
    RS_CUDA_CHECK(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device));
    b.matches_aos = c->d_matches, b.cur_pose = c->d_cur, b.n_matches = c->d_n;
    b.subsets_in = nullptr, b.normals_in = nullptr;
    RS_CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    RS_CUDA_CHECK(cudaEventCreateWithFlags(&c->ransac_done, cudaEventDisableTiming));
    RS_CUDA_CHECK(cudaEventCreateWithFlags(&c->group_fork, cudaEventDisableTiming));
    RS_CUDA_CHECK(cudaEventCreateWithFlags(&c->prepared, cudaEventDisableTiming));
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
    const size_t n = size_t(c->max_batch) * c->max_variance * c->M * 4;
    if (!c->h_normals) RS_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&c->h_normals), sizeof(double) * n, cudaHostAllocDefault));
    if (c->d_normals) return RS_OK;
    return dev_alloc(&c->d_normals, n);
}

// Runs fn(frame) for every frame of the batch on the host's hardware threads (the reference's random draws of different
// frames are independent: each frame owns an engine).
template <class F>
void parallel_frames(const int batch, F fn)
{
    const int nt = int(std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), unsigned(batch)));
    if (nt <= 1) {
        for (int b = 0; b < batch; ++b) fn(b);
        return;
    }
    std::atomic<int> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&] {
            for (int b = next.fetch_add(1); b < batch; b = next.fetch_add(1)) fn(b);
        });
    for (std::thread& t : th) t.join();
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
    prm.ctas_per_sm = o.worker_ctas_per_sm;
    prm.solver = o.solver;
    if (prm.solver < 0 || prm.solver > 3) {
        set_last_error("rs_pose: unknown solver (0 = by shape, 1 = chain, 2 = fused, 3 = one hypothesis per lane)");
        return RS_ERR_INVALID_ARG;
    }
    if (const char* e = std::getenv("RS_POSE_SPLIT")) prm.split = std::atoi(e);
    if (const char* e = std::getenv("RS_POSE_MC_CAP")) prm.mc_cap = std::atoi(e);       // experiment knobs (tools/exp_pose_timeline.py)
    if (const char* e = std::getenv("RS_POSE_HELP_MIN")) prm.help_min = std::atoi(e);
    if (const char* e = std::getenv("RS_POSE_MC_BATCHED")) prm.mc_batched = std::atoi(e);
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

// RS_RNG_REFERENCE, between the two halves of a solve: the Gaussian stream of a frame continues where its RANSAC loop left
// the engine, so the host fetches iterations_run and the inlier masks, draws what the reference would draw (host threads, a
// frame each) and uploads the draws; buf.normals_in then points at them.
int reference_round_trip(rs_pose_ctx* c, int batch, const PoseLaunch& lp, cudaStream_t s, PoseBuffers& buf)
{
    int rc;
    std::vector<rs_pose_out>& h_out = c->h_out;
    std::vector<uint8_t>& h_mask = c->h_mask;
    h_out.resize(batch);
    h_mask.resize(size_t(batch) * c->M);
    RS_CUDA_CHECK(cudaMemcpyAsync(h_out.data(), buf.out, sizeof(rs_pose_out) * batch, cudaMemcpyDeviceToHost, s));
    RS_CUDA_CHECK(cudaMemcpyAsync(h_mask.data(), buf.mask, h_mask.size(), cudaMemcpyDeviceToHost, s));
    RS_CUDA_CHECK(cudaStreamSynchronize(s));
    if ((rc = ensure_normals(c)) != RS_OK) return rc;
    const size_t per_frame = size_t(lp.n_variance) * c->M * 4;
    double* h_normals = c->h_normals;
    parallel_frames(batch, [&](int b) {
        double* dst = h_normals + size_t(b) * per_frame;
        std::fill(dst, dst + per_frame, 0.0);
        if (h_out[b].status == -2)  // final pose available, covariance pending
            reference_normals(c, b, lp.seed + uint32_t(b), h_out[b].iterations_run, lp.n_variance,
                              h_mask.data() + size_t(b) * c->M, dst);
    });
    RS_CUDA_CHECK(cudaMemcpyAsync(c->d_normals, h_normals, sizeof(double) * per_frame * batch, cudaMemcpyHostToDevice, s));
    buf.normals_in = c->d_normals;
    return RS_OK;
}

int solve_impl(rs_pose_ctx* c, int batch, const PoseLaunch& prm, cudaStream_t s, bool reference_rng)
{
    int rc;
    PoseBuffers buf = c->buf;
    if (reference_rng) {
        // the reference's draws, frame by frame (every frame has its own engine, mt19937(seed + frame)): host threads
        std::vector<int32_t>& h_subsets = c->h_subsets;
        h_subsets.resize(size_t(batch) * c->max_iterations * RS_MAX_SUBSET);
        parallel_frames(batch, [&](int b) {
            reference_subsets(c, b, prm.seed + uint32_t(b), prm.max_iterations,
                              h_subsets.data() + size_t(b) * c->max_iterations * RS_MAX_SUBSET);
        });
        RS_CUDA_CHECK(cudaMemcpyAsync(c->d_subsets_in, h_subsets.data(), sizeof(int32_t) * h_subsets.size(),
                                      cudaMemcpyHostToDevice, s));
        RS_CUDA_CHECK(cudaStreamSynchronize(s));   // pageable source: the copy has been staged when this returns
        buf.subsets_in = c->d_subsets_in;
    }
    cudaEvent_t* ev = c->timing_slots > 0 ? &c->events[size_t(c->run_counter % uint64_t(c->timing_slots)) * 5] : nullptr;
    ++c->run_counter;
    c->groups_last = 1;   // rs_pose_opts.sub_batches is accepted and has no effect: the fused kernel hands over per frame
    PoseLaunch lp = prm;
    lp.frame0 = 0;
    if (ev) RS_CUDA_CHECK(cudaEventRecord(ev[0], s));
    if (c->prepared_batch == batch && !reference_rng) {
        // rs_pose_prepare_device already ran the preparation kernel for this batch (possibly on another stream)
        RS_CUDA_CHECK(cudaStreamWaitEvent(s, c->prepared, 0));
    }
    else if ((rc = launch_pose_prepare(buf, lp, s)) != RS_OK)
        return rc;
    c->prepared_batch = 0;
    if (ev) RS_CUDA_CHECK(cudaEventRecord(ev[1], s));
    // one hypothesis per lane (pose_wide.cu) when a frame has hundreds of hypotheses; it knows points and planes only
    const bool wide_ok = buf.hyp != nullptr && !lp.has_point2d;
    if (lp.solver == 3 && !wide_ok) {
        set_last_error("rs_pose: solver 3 needs a context created with max_iterations > 256 and a batch without RS_FEAT_POINT2D features");
        return RS_ERR_INVALID_ARG;
    }
    const bool wide = lp.solver == 3 || (lp.solver == 0 && wide_ok && lp.max_iterations > kWideMinIterations);
    const bool fused = !wide && (lp.solver == 2 || (lp.solver == 0 && lp.max_iterations > 256) || !pose_chain_supports(c->M));
    c->last_fused = fused;
    if (!fused) {
        // the three-launch chain (pose_chain.cu), or the hypotheses one per lane and the chain's kernels for the rest
        if (wide) {
            if ((rc = launch_pose_wide_hypotheses(buf, lp, s, c->sm_count)) != RS_OK) return rc;
            lp.final_only = 1;
        }
        if ((rc = launch_pose_chain_ransac(buf, lp, s)) != RS_OK) return rc;
        RS_CUDA_CHECK(cudaEventRecord(c->ransac_done, s));
        if (ev) RS_CUDA_CHECK(cudaEventRecord(ev[2], s));
        if (lp.n_variance > 0) {
            if (reference_rng) {
                if ((rc = reference_round_trip(c, batch, lp, s, buf)) != RS_OK) return rc;
                if (ev) RS_CUDA_CHECK(cudaEventRecord(ev[2], s));  // exclude the host RNG round trip
            }
            if ((rc = launch_pose_chain_variance(buf, lp, s)) != RS_OK) return rc;
        }
        if (ev) {
            RS_CUDA_CHECK(cudaEventRecord(ev[3], s));
            RS_CUDA_CHECK(cudaEventRecord(ev[4], s));
        }
        if (reference_rng && lp.n_variance > 0) RS_CUDA_CHECK(cudaStreamSynchronize(s));  // the pinned Gaussian staging buffer is reused
    }
    else if (!reference_rng || lp.n_variance <= 0) {
        // Hypotheses, final LM, Monte-Carlo solves and covariance of the batch feed on one set of work queues. Two launches of
        // the same kernel work them, side by side: one with the frame role only (a CTA per frame, 28 KB of shared memory: it
        // fits beside whatever else the caller runs, and its CTAs leave as their frames finish), one with the Monte-Carlo role
        // (58 KB per CTA: a sample per warp holds a perturbed copy of the map side) whose CTAs move in as the first ones leave
        // and start on a frame's samples the moment its final LM is done.
        const bool frames_only = lp.n_variance > 0 && lp.ctas_per_sm < 0;   // the caller launches the Monte-Carlo role (rs_pose_add_workers)
        const bool split = lp.n_variance > 0 && lp.split != 0 && !frames_only;
        cudaStream_t aux = c->group_stream[0];
        if (split) {
            RS_CUDA_CHECK(cudaEventRecord(c->group_fork, s));
            RS_CUDA_CHECK(cudaStreamWaitEvent(aux, c->group_fork, 0));
            PoseLaunch fr = lp;
            fr.run_ransac = 1, fr.run_mc = 0, fr.linger = lp.max_iterations >= 256 ? 1 : 0, fr.ctas_per_sm = 0;
            if ((rc = launch_pose_fused(buf, fr, s)) != RS_OK) return rc;
            RS_CUDA_CHECK(cudaEventRecord(c->ransac_done, s));
            PoseLaunch mc = lp;
            // both roles: whichever launch the hardware dispatches first, every resident CTA can take whatever work there is
            // (CTAs that could only wait for frames would starve the frames' own CTAs of registers if they got there first)
            mc.run_ransac = 1, mc.run_mc = 1;
            if ((rc = launch_pose_fused(buf, mc, aux)) != RS_OK) return rc;
            RS_CUDA_CHECK(cudaEventRecord(c->group_done[0], aux));
            RS_CUDA_CHECK(cudaStreamWaitEvent(s, c->group_done[0], 0));
        }
        else if (frames_only) {
            PoseLaunch fr = lp;
            fr.run_ransac = 1, fr.run_mc = 0, fr.linger = lp.max_iterations >= 256 ? 1 : 0, fr.ctas_per_sm = 0;
            if ((rc = launch_pose_fused(buf, fr, s)) != RS_OK) return rc;
            RS_CUDA_CHECK(cudaEventRecord(c->ransac_done, s));
        }
        else {
            lp.run_ransac = 1, lp.run_mc = 1;
            if ((rc = launch_pose_fused(buf, lp, s)) != RS_OK) return rc;
            RS_CUDA_CHECK(cudaEventRecord(c->ransac_done, s));
        }
        if (ev)
            for (int k = 2; k <= 4; ++k) RS_CUDA_CHECK(cudaEventRecord(ev[k], s));
    }
    else {
        // RS_RNG_REFERENCE: the Gaussian stream of a frame continues where its RANSAC loop left the engine, so the host
        // needs iterations_run and the inlier mask between the two halves of the kernel
        lp.run_ransac = 1, lp.run_mc = 0, lp.linger = 1;
        if ((rc = launch_pose_fused(buf, lp, s)) != RS_OK) return rc;
        RS_CUDA_CHECK(cudaEventRecord(c->ransac_done, s));
        if (ev) RS_CUDA_CHECK(cudaEventRecord(ev[2], s));
        if ((rc = reference_round_trip(c, batch, lp, s, buf)) != RS_OK) return rc;
        if (ev) RS_CUDA_CHECK(cudaEventRecord(ev[2], s));  // exclude the host RNG round trip
        lp.run_ransac = 0, lp.run_mc = 1, lp.publish_all = 1;
        if ((rc = launch_pose_fused(buf, lp, s)) != RS_OK) return rc;
        if (ev) {
            RS_CUDA_CHECK(cudaEventRecord(ev[3], s));
            RS_CUDA_CHECK(cudaEventRecord(ev[4], s));
        }
        RS_CUDA_CHECK(cudaStreamSynchronize(s));  // the pinned Gaussian staging buffer is reused by the next solve
    }
    c->last = prm;
    c->last_batch = batch;
    c->last_mc_only = lp.run_ransac == 0;
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
    if (c->h_normals) cudaFreeHost(c->h_normals);
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
    cudaFree(b.work);
    cudaFree(b.rframe);
    cudaFree(b.ring_mask);
    cudaFree(b.ready);
    cudaFree(b.open_list);
    cudaFree(b.mc_done);
    cudaFree(b.frame_times);
    cudaFree(b.hyp);
    cudaFree(b.hyp_mask);
    cudaFree(b.fold);
    cudaFree(b.fold_mask);
    for (cudaEvent_t e : c->events) cudaEventDestroy(e);
    if (c->ransac_done) cudaEventDestroy(c->ransac_done);
    if (c->group_fork) cudaEventDestroy(c->group_fork);
    if (c->prepared) cudaEventDestroy(c->prepared);
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

int rs_pose_prepare_device(rs_pose_ctx* c, int batch, const rs_pose_opts* opts, void* stream)
{
    if (!c || batch <= 0 || batch > c->max_batch) {
        set_last_error("rs_pose_prepare_device: invalid argument");
        return RS_ERR_INVALID_ARG;
    }
    PoseLaunch prm;
    int rc = resolve_launch(c, opts, batch, prm);
    if (rc != RS_OK) return rc;
    RS_CUDA_CHECK(cudaSetDevice(c->device));
    prm.frame0 = 0;
    if ((rc = launch_pose_prepare(c->buf, prm, static_cast<cudaStream_t>(stream))) != RS_OK) return rc;
    RS_CUDA_CHECK(cudaEventRecord(c->prepared, static_cast<cudaStream_t>(stream)));
    c->prepared_batch = batch;
    return RS_OK;
}

int rs_pose_add_workers(rs_pose_ctx* c, int ctas_per_sm, void* stream)
{
    if (!c || c->last_batch <= 0) {
        set_last_error("rs_pose_add_workers: no solve has been launched through this context");
        return RS_ERR_INVALID_ARG;
    }
    if (!c->last_fused) return RS_OK;   // the chain has no work queues to feed on
    RS_CUDA_CHECK(cudaSetDevice(c->device));
    PoseLaunch prm = c->last;
    prm.batch = c->last_batch;
    prm.frame0 = 0;
    prm.ctas_per_sm = ctas_per_sm;
    prm.run_ransac = c->last_mc_only ? 0 : 1, prm.run_mc = 1, prm.publish_all = 0;
    PoseBuffers buf = c->buf;
    if (c->last_mc_only) buf.normals_in = c->d_normals;
    return launch_pose_fused(buf, prm, static_cast<cudaStream_t>(stream));
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

int rs_pose_debug_counters(rs_pose_ctx* c, uint64_t out[8])
{
    if (!c || !out) {
        set_last_error("rs_pose_debug_counters: invalid argument");
        return RS_ERR_INVALID_ARG;
    }
    RS_CUDA_CHECK(cudaSetDevice(c->device));
    RS_CUDA_CHECK(cudaDeviceSynchronize());
    PoseWork w;
    RS_CUDA_CHECK(cudaMemcpy(&w, c->buf.work, sizeof(PoseWork), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 8; ++k) out[k] = w.dbg[k];
    return RS_OK;
}

int rs_pose_debug_frame_times(rs_pose_ctx* c, int batch, double* ms /* batch x 4 */)
{
    if (!c || !ms || batch <= 0 || batch > c->max_batch) {
        set_last_error("rs_pose_debug_frame_times: invalid argument");
        return RS_ERR_INVALID_ARG;
    }
    RS_CUDA_CHECK(cudaSetDevice(c->device));
    RS_CUDA_CHECK(cudaDeviceSynchronize());
    PoseWork w;
    RS_CUDA_CHECK(cudaMemcpy(&w, c->buf.work, sizeof(PoseWork), cudaMemcpyDeviceToHost));
    std::vector<unsigned long long> t(size_t(batch) * 4);
    RS_CUDA_CHECK(cudaMemcpy(t.data(), c->buf.frame_times, sizeof(unsigned long long) * t.size(), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < t.size(); ++i) ms[i] = t[i] >= w.t_first && w.t_first != ~0ull ? double(t[i] - w.t_first) * 1e-6 : -1.0;
    return RS_OK;
}

int rs_pose_phase_ms(rs_pose_ctx* c, float ms[2])
{
    if (!c || !ms) {
        set_last_error("rs_pose_phase_ms: invalid argument");
        return RS_ERR_INVALID_ARG;
    }
    RS_CUDA_CHECK(cudaSetDevice(c->device));
    RS_CUDA_CHECK(cudaDeviceSynchronize());
    return pose_work_times(c->buf, &ms[0], &ms[1], c->stream);
}

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
