// K5 / K6 — per-frame RANSAC + Levenberg-Marquardt pose solve and its Monte-Carlo covariance.
//
// Replaces  Pose_Optimization::{compute_optimized_pose, compute_pose_with_ransac, compute_optimized_global_pose,
//           get_features_inliers_outliers, compute_pose_variance, compute_random_variation_of_pose}
//           (src/pose_optimization/pose_optimization.cpp:33-72,107-437,482-501),
//           Global_Pose_Estimator::operator() + the coefficient <-> pose maps (levenberg_marquardt_functors.cpp:14-38,
//           74-98,128-169), ransac::get_random_subset_with_score (ransac.hpp:77-103), the Point/Plane
//           IOptimizationFeature implementations (map_point.cpp:16-65, map_primitive.cpp:15-85) and the transforms
//           they call (camera_transformation.cpp:11-71, point_coordinates.cpp:245-278, plane_coordinates.cpp:20-56),
//           plus Eigen::LevenbergMarquardt<NumericalDiff<F, Forward>>::minimize (MINPACK lmdif: forward-difference
//           Jacobian, trust region with lmpar / qrsolv).
//
// Mapping (B200): ONE WARP PER LM PROBLEM. Features are lane-strided; every lane evaluates its features' residuals
// at x and at the six forward-difference points, forms its rows of J, and accumulates the 6x6 J^T J, the 6x1 J^T r
// and |r|^2 in FP64 registers; a warp-shuffle butterfly reduces the 28 sums. The m x 6 Jacobian is never stored.
// Lane 0 then does the MINPACK step on shared-memory 6x6 state: the triangular factor R of J P = Q R is obtained as
// the pivoted Cholesky factor of the column-scaled J^T J (same R up to row signs, to which lmpar is invariant), and
// Q^T r = R^-T P^T J^T r. The trust-region logic (lmpar, qrsolv, ratio tests, stop codes 1-8, nfev accounting incl.
// the redundant f(x) of NumericalDiff) is MINPACK's, so iterates follow the reference's path up to rounding.
//   pose_prepare_kernel  : AoS -> SoA of the match lists, validity, total score, reset of the per-frame work state.
//   pose_fused_kernel    : persistent CTAs; a frame's hypotheses are claimed dynamically by the warps working on it
//                          (hypotheses are independent: each LM starts from the current pose) and the reference's serial
//                          best-so-far / early-stop rule is applied in iteration order as results complete; the warp that
//                          applies the terminal hypothesis runs the final LM on the winning inlier set and publishes the
//                          frame's Monte-Carlo solves (one warp per sample, perturbed copy of the inlier set in shared
//                          memory), which any CTA picks up; the CTA finishing a frame's last samples reduces its 6x6
//                          covariance (one entry per lane, summed in sample order) and validates it.
// FP64 throughout (forward differences with h = 1.49e-8 |x| on mm-scale coordinates need it).
#include "pose_lm.cuh"

namespace rs {

namespace {

#ifndef RS_POSE_CTAS_PER_SM
#define RS_POSE_CTAS_PER_SM 3        // 3 x 4 warps at 168 registers; 4 would cap the LM body at 128 registers (spills on the serial chains)
#endif
constexpr int WARPS = 4;             // warps per CTA of the fused kernel (one hypothesis / one Monte-Carlo sample each)
constexpr int THREADS = WARPS * 32;

// ---- kernels ----------------------------------------------------------------------------------------------------------

// AoS -> SoA, PlaneCoordinates normalisation of both plane sides, validity (compute_optimized_pose :269-282),
// total score in list order. One CTA per frame.
__global__ void __launch_bounds__(THREADS) pose_prepare_kernel(const PoseBuffers buf, const PoseLaunch prm)
{
    const int b = prm.frame0 + blockIdx.x;
    const int M = buf.max_matches;
    int n = buf.n_matches[b];
    n = n < 0 ? 0 : (n > M ? M : n);
    const rs_match* src = buf.matches_aos + size_t(b) * M;
    int32_t* type = buf.type + size_t(b) * M;
    double* obs = buf.obs + size_t(b) * 4 * M;
    double* map = buf.map + size_t(b) * 4 * M;
    double* sig = buf.sigma + size_t(b) * 4 * M;
    __shared__ int s_invalid;
    if (threadIdx.x == 0) s_invalid = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        if (i < n) {
            const rs_match f = src[i];
            const int ty = (f.type == RS_FEAT_POINT) ? RS_FEAT_POINT : (f.type == RS_FEAT_POINT2D ? RS_FEAT_POINT2D : RS_FEAT_PLANE);
            double o[4] = {f.obs[0], f.obs[1], f.obs[2], f.obs[3]};
            double m[4] = {f.map[0], f.map[1], f.map[2], f.map[3]};
            double sg[4] = {0.0, 0.0, 0.0, 0.0}, ax[4] = {0.0, 0.0, 0.0, 0.0};
            // is_valid: map_point.cpp:60-64, map_primitive.cpp:79-83, map_point2d.cpp:75-79
            const int k = (ty == RS_FEAT_POINT) ? 3 : 4;
            const int ko = (ty == RS_FEAT_POINT) ? 2 : 4;
            const int ks = (ty == RS_FEAT_POINT2D) ? 3 : k;
            bool bad = false;
            for (int c = 0; c < ko; ++c) bad = bad || (o[c] != o[c]);
            for (int c = 0; c < k; ++c) bad = bad || (m[c] != m[c]);
            for (int c = 0; c < ks; ++c) bad = bad || !(f.sigma[c] >= 0.0);
            if (bad) atomicOr(&s_invalid, 1);
            if (ty == RS_FEAT_PLANE) {
                normalize3(o);
                normalize3(m);
                for (int c = 0; c < 4; ++c) sg[c] = f.sigma[c];
            }
            else if (ty == RS_FEAT_POINT) {
                o[2] = 0.0, o[3] = 0.0, m[3] = 0.0;
                for (int c = 0; c < 3; ++c) sg[c] = f.sigma[c];
            }
            else {
                // inverse-depth point: device layout o = (u, v, dFar, dNear), m = (theta, phi), ax = first observation
                // (+ inverse depth), sg = (sigma theta, sigma phi). get_furthest/closest_estimation
                // (inverse_depth_coordinates.cpp:142-154) with the square root taken of the STANDARD DEVIATION, as
                // Point2dOptimizationFeature::get_distance passes it in the covariance's place (map_point2d.cpp:42-43, :159).
                const double depthVariation = sqrt(f.sigma[0]) * 3;
                ax[0] = f.map[0], ax[1] = f.map[1], ax[2] = f.map[2], ax[3] = f.map[3];
                m[0] = f.obs[2], m[1] = f.obs[3], m[2] = 0.0, m[3] = 0.0;
                o[2] = fmin(f.map[3] - depthVariation, 1e-9);
                o[3] = fmin(f.map[3] + depthVariation, 1e-9);
                sg[0] = f.sigma[1], sg[1] = f.sigma[2];
            }
            type[i] = ty;
            for (int c = 0; c < 4; ++c) {
                obs[c * M + i] = o[c];
                map[c * M + i] = m[c];
                sig[c * M + i] = sg[c];
                buf.aux[(size_t(b) * 4 + c) * M + i] = ax[c];
            }
            buf.mask[size_t(b) * M + i] = 0;
        }
        else {
            type[i] = RS_FEAT_POINT;
            for (int c = 0; c < 4; ++c)
                obs[c * M + i] = 0.0, map[c * M + i] = 0.0, sig[c * M + i] = 0.0, buf.aux[(size_t(b) * 4 + c) * M + i] = 0.0;
            buf.mask[size_t(b) * M + i] = 0;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        PoseFrameState st;
        st.n = n;
        st.valid = s_invalid ? 0 : 1;
        double score = 0.0;
        int res = 0;
        for (int i = 0; i < n; ++i) {
            const int ty = type[i];
            score += score_of(ty);
            res += parts_of(ty);
        }
        st.total_score = score;
        st.residuals = res;
        st.stage = 0;
        for (int j = 0; j < 6; ++j) st.final_x[j] = 0.0;
        buf.state[b] = st;
        rs_pose_out o;
        o.status = 0, o.n_inliers = 0, o.iterations_run = 0, o.best_iteration = -1, o.n_variance_ok = 0, o.reserved = 0;
        o.score = 0.0;
        for (int j = 0; j < 7; ++j) o.pose[j] = buf.cur_pose[b * 7 + j];
        for (int j = 0; j < 36; ++j) o.cov[j] = 0.0;
        buf.out[b] = o;
        for (int j = 0; j < 7; ++j) buf.poses[b * 7 + j] = o.pose[j];
    }
    // ---- work state of the fused kernel for this frame ----
    {
        RansacFrame* g = buf.rframe + b;
        int* gi = reinterpret_cast<int*>(g);
        for (int i = threadIdx.x; i < int(sizeof(RansacFrame) / sizeof(int)); i += blockDim.x) gi[i] = 0;
        const int words = (M + 31) / 32;
        unsigned* best = buf.ring_mask + (size_t(b) * (kRansacRing + 1) + kRansacRing) * words;
        for (int i = threadIdx.x; i < words; i += blockDim.x) best[i] = 0u;
        __syncthreads();
        if (threadIdx.x == 0) {
            double x0[6];
            coefficients_from_pose(buf.cur_pose + b * 7, x0);
            for (int j = 0; j < 6; ++j) g->best_x[j] = x0[j];
            g->max_score = 1.0;
            g->best_iteration = -1;
            buf.mc_done[b] = 0;
            for (int k = 0; k < 4; ++k) buf.frame_times[b * 4 + k] = 0ull;
            buf.ready[blockIdx.x] = 0;
            buf.open_list[blockIdx.x] = 0;
            if (blockIdx.x == 0) {
                PoseWork w;
                w.join_ticket = 0, w.frames_done = 0, w.n_ready = 0, w.mc_head = 0, w.n_open = 0;
                w.mc_done_tasks = 0, w.pad[0] = w.pad[1] = 0;
                w.t_first = ~0ull, w.t_ransac_end = 0ull, w.t_last = 0ull;
                for (int k = 0; k < 8; ++k) w.dbg[k] = 0ull;
                *buf.work = w;
            }
        }
    }
}


// ---- the fused solve kernel ------------------------------------------------------------------------------------------------
// compute_pose_with_ransac (pose_optimization.cpp:107-262), compute_pose_variance (:361-437) and
// compute_random_variation_of_pose (:482-501) of a whole batch in ONE persistent kernel. The three stages of a frame used
// to be three launches, and a stage of the batch lasted as long as its slowest frame: the hypothesis stage is a latency
// chain that leaves four fifths of the issue slots idle, and the Monte-Carlo solves (throughput bound) could not start
// before the last frame's final LM. Here a frame's Monte-Carlo solves are published the moment ITS final LM is done and
// are picked up by whichever CTA has nothing more urgent to do; the covariance is reduced by the CTA that finishes a
// frame's last sample group. CTAs (four warps) are persistent and take, in this order:
//   1. a frame nobody has started: stage its matches in shared memory and run its hypotheses (the warps claim iteration
//      indices; iterations 0..3 are always needed - the reference cannot stop before the fourth), then the final LM on the
//      warp that applied the terminal hypothesis;
//   2. a Monte-Carlo task (frame, group of four samples), one sample per warp;
//   3. a frame on the help list (its loop went past the minimum of four and has iterations left): the frame's ring of
//      hypothesis results lives in global memory, so any number of CTAs can claim its iterations - a 1024-hypothesis
//      frame spreads over the whole GPU;
//   4. nothing: sleep and look again, or leave when every frame is through and no task is left.
// Every wait is for work held by a CTA that is running (tickets are only ever taken by resident CTAs), so the kernel makes
// progress whatever part of the grid is resident - CTAs that only become resident when another kernel's CTAs retire (the
// cell-graph segmentation runs beside the hypothesis stage) simply join in.
// One LM call site per kernel, as before: the three kinds of problem (hypothesis, final, Monte-Carlo sample) are set up in
// front of it and taken apart behind it.
// Cross-CTA state is read with volatile / .cg loads (L2) and published fence-then-flag; gpu-scope fences are kept off the
// per-iteration paths (on sm_100 each one also invalidates the SM's L1, where register spills live).

__device__ __forceinline__ int ld_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ void st_volatile(int* p, const int v) { *reinterpret_cast<volatile int*>(p) = v; }
__device__ __forceinline__ double ld_volatile(const double* p) { return *reinterpret_cast<const volatile double*>(p); }
__device__ __forceinline__ void st_volatile(double* p, const double v) { *reinterpret_cast<volatile double*>(p) = v; }
__device__ __forceinline__ unsigned long long global_timer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
struct FrameFlags {
    int can_quit, closed, opened, applied;
};
__device__ __forceinline__ FrameFlags load_flags(const RansacFrame* g)   // the first 16 bytes of the frame's state, one load
{
    FrameFlags f;
    asm volatile("ld.volatile.global.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(f.can_quit), "=r"(f.closed), "=r"(f.opened), "=r"(f.applied) : "l"(g));
    return f;
}
__device__ __forceinline__ void count(PoseWork* W, const int k, const unsigned long long v = 1ull) { atomicAdd(&W->dbg[k], v); }

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) / 16 * 16; }

// shared-memory carve-up. Monte-Carlo task: the frame's observations, and per warp the perturbed map side of its sample.
// Hypothesis task: the observations and, in the first warp's region, the map side.
struct FusedSmem {
    double* obs;        // [4][M] the frame's observations
    double* pmap;       // [warps][4][M]
    WarpLM* lm;         // [warps]
    int32_t* type;      // [M]
    short* idx;         // [M] inlier indices (final LM / Monte-Carlo)
    unsigned* wmask;    // [warps][words] inlier mask of the hypothesis a warp is scoring
    short* subset;      // [warps][RS_MAX_SUBSET]
    int* ctl;           // [8] task hand-over
};
__host__ __device__ inline size_t fused_carve(FusedSmem* s, unsigned char* base, const int M, const int warps, const bool mc_role = true)
{
    const int words = (M + 31) / 32;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        unsigned char* p = base ? base + o : nullptr;
        o = align16(o + bytes);
        return p;
    };
    unsigned char* obs = take(sizeof(double) * 4 * M);
    unsigned char* pm = take(sizeof(double) * size_t(mc_role ? warps : 1) * 4 * M);   // without the Monte-Carlo role: the map side only
    unsigned char* lm = take(sizeof(WarpLM) * warps);
    unsigned char* type = take(sizeof(int32_t) * M);
    unsigned char* idx = take(sizeof(short) * M);
    unsigned char* wm = take(sizeof(unsigned) * size_t(warps) * words);
    unsigned char* sub = take(sizeof(short) * warps * RS_MAX_SUBSET);
    unsigned char* ctl = take(sizeof(int) * 8);
    if (s) {
        s->obs = reinterpret_cast<double*>(obs), s->pmap = reinterpret_cast<double*>(pm);
        s->lm = reinterpret_cast<WarpLM*>(lm), s->type = reinterpret_cast<int32_t*>(type);
        s->idx = reinterpret_cast<short*>(idx), s->wmask = reinterpret_cast<unsigned*>(wm);
        s->subset = reinterpret_cast<short*>(sub), s->ctl = reinterpret_cast<int*>(ctl);
    }
    return o;
}

enum { TASK_EXIT = 0, TASK_RANSAC = 1, TASK_MC = 2 };
enum { DBG_HYP_FIRST = 0, DBG_HYP_HELPER, DBG_APPLY_BATCHES, DBG_APPLIED, DBG_HELPER_JOINS, DBG_MC_TASKS, DBG_ABORTED, DBG_IDLE_LOOPS };

// The reference's serial bookkeeping (pose_optimization.cpp:151-227), applied strictly in iteration order by the warp that
// holds the frame's lock (all 32 lanes call): every finished hypothesis whose predecessors have all been applied is folded
// into the best-so-far state; the early stop freezes the state exactly where the serial loop would have left it.
// Up to 32 consecutive finished hypotheses are taken per round: their results are loaded one per lane, the serial rule runs
// over them in registers, the winner's pose and inlier mask are copied, and the frame's flags are published once.
// Returns true when THIS call closed the hypothesis stage (early stop, or the last iteration applied): the calling warp
// then owns the frame's final LM.
__device__ __noinline__ bool apply_ready_warp(RansacFrame* g, unsigned* ring, const int words, const int maxIterations,
                                             const unsigned inliersToStop, PoseWork* W, int32_t* open_list, const int frame,
                                             const int lane)
{
    for (;;) {
        const FrameFlags f = load_flags(g);
        if (f.can_quit || f.closed || f.applied >= maxIterations) return false;
        const int ap = f.applied;
        const int mine = ap + lane;
        const int slot = mine % kRansacRing;
        const bool ready = mine < maxIterations && ld_volatile(&g->done[slot]) == mine + 1;
        const unsigned rb = __ballot_sync(FULL, ready);
        const int cnt = rb == FULL ? 32 : __ffs(~rb) - 1;   // leading run of finished hypotheses
        if (cnt <= 0) return false;
        double hs = 0.0;
        int hin = 0, hok = 0;
        if (lane < cnt) {
            hs = ld_volatile(&g->slot[slot].score);
            hin = ld_volatile(&g->slot[slot].inliers);
            hok = ld_volatile(&g->slot[slot].ok);
        }
        double ms = ld_volatile(&g->max_score);
        int bi = ld_volatile(&g->best_inliers), bit = ld_volatile(&g->best_iteration), started = ld_volatile(&g->started);
        int best_lane = -1, used = 0;
        bool quit = false;
        for (int j = 0; j < cnt; ++j) {
            const double s_j = __shfl_sync(FULL, hs, j);
            const int in_j = __shfl_sync(FULL, hin, j), ok_j = __shfl_sync(FULL, hok, j);
            const int i = ap + j;
            ++started, ++used;
            if (ok_j && s_j >= 1.0) {
                const bool canOverload = (s_j > ms) || (fabs(s_j - ms) <= 0.1 && bi < in_j);
                if (canOverload) ms = s_j, bi = in_j, bit = i, best_lane = j;
                if (i >= 3 && unsigned(bi) > inliersToStop) {
                    quit = true;
                    break;
                }
            }
        }
        if (best_lane >= 0) {   // before the slots can be handed out again
            const int bs = (ap + best_lane) % kRansacRing;
            if (lane < 6) st_volatile(&g->best_x[lane], ld_volatile(&g->slot[bs].x[lane]));
            volatile unsigned* src = ring + size_t(bs) * words;
            volatile unsigned* dst = ring + size_t(kRansacRing) * words;
            for (int k = lane; k < words; k += 32) dst[k] = src[k];
        }
        __syncwarp();
        const int nap = ap + used;
        const bool closed_now = quit || nap == maxIterations;
        if (lane == 0) {
            st_volatile(&g->max_score, ms);
            st_volatile(&g->best_inliers, bi);
            st_volatile(&g->best_iteration, bit);
            st_volatile(&g->started, started);
            __threadfence();   // the state before the flags
            if (quit) st_volatile(&g->can_quit, 1);
            st_volatile(&g->applied, nap);
            if (closed_now) st_volatile(&g->closed, 1);
            else if (nap >= 4 && !f.opened) {
                // the minimum of four hypotheses did not stop the loop: more warps pay off from here on
                st_volatile(&g->opened, 1);
                const int s = atomicAdd(&W->n_open, 1);
                st_volatile(&open_list[s], frame + 1);
            }
            count(W, DBG_APPLY_BATCHES);
            count(W, DBG_APPLIED, unsigned(used));
        }
        __syncwarp();
        if (closed_now) return true;
    }
}

// Takes the frame's bookkeeping lock if something is ready to be applied and applies it (all lanes call).
// Returns 1 when this call closed the hypothesis stage, 0 otherwise; *busy = the lock was held by another warp.
__device__ __forceinline__ int try_apply(RansacFrame* g, unsigned* ring, const int words, const int maxIterations,
                                         const unsigned inliersToStop, PoseWork* W, int32_t* open_list, const int frame,
                                         const int lane, const int applied)
{
    int got = 0;
    if (lane == 0)
        got = applied < maxIterations && ld_volatile(&g->done[applied % kRansacRing]) == applied + 1 && atomicCAS(&g->lock, 0, 1) == 0;
    got = __shfl_sync(FULL, got, 0);
    if (!got) return -1;
    const bool closed_now = apply_ready_warp(g, ring, words, maxIterations, inliersToStop, W, open_list, frame, lane);
    if (lane == 0) atomicExch(&g->lock, 0);
    return closed_now ? 1 : 0;
}

// Hypotheses + final LM of frame b (pose_optimization.cpp:107-262), run by all warps of the CTA; `arg` != 0 for the frame's
// first CTA. Called inline by the kernel instantiation that has the frame role only, out of line by the one that also
// carries the Monte-Carlo role (there it is the fallback that keeps every resident CTA able to take whatever work there is).
template <bool P2D>
__device__ __forceinline__ void frame_task(const PoseBuffers& buf, const PoseLaunch& prm, const FusedSmem& sm, PoseWork* W, const int b,
                                           const int arg, const int n, const int warp, const int lane)
{
    const int M = buf.max_matches;
    const int words = (M + 31) / 32;
    const int maxIterations = prm.max_iterations;
    const PoseFrameState* stp = buf.state + b;
    WarpLM& S = sm.lm[warp];
    short* subset = sm.subset + warp * RS_MAX_SUBSET;
    unsigned* wmask = sm.wmask + size_t(warp) * words;
    // ======================= hypotheses + final LM of frame b (pose_optimization.cpp:107-262) =======================
    const bool first = arg != 0;
    const bool usable = __ldcg(&stp->valid) && __ldcg(&stp->total_score) >= 1.0;
    if (!usable) {   // out[b] already says status 0, pose = current pose (helpers never come here: the frame never opens)
        if (threadIdx.x == 0) atomicAdd(&W->frames_done, 1);
        __syncthreads();
        return;
    }
    double* s_obs = sm.obs;
    double* s_map = sm.pmap;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        sm.type[i] = buf.type[size_t(b) * M + i];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            s_obs[c * M + i] = buf.obs[(size_t(b) * 4 + c) * M + i];
            s_map[c * M + i] = buf.map[(size_t(b) * 4 + c) * M + i];
        }
    }
    double x0[6];
    coefficients_from_pose(buf.cur_pose + b * 7, x0);
    RansacFrame* g = buf.rframe + b;
    unsigned* ring = buf.ring_mask + size_t(b) * (kRansacRing + 1) * words;
    const unsigned inliersToStop = unsigned(ceil(double(n) * kEarlyStopProportion));
    Problem P;
    P.type = sm.type, P.obs = s_obs, P.map = s_map, P.M = M;
    P.aux = buf.aux + size_t(b) * 4 * M;
    rs_pose_out* out = buf.out + b;
    if (first && threadIdx.x == 0) buf.frame_times[b * 4 + 0] = global_timer();
    __syncthreads();

    bool finalPass = false;
    for (;;) {
        int it = -1;
        if (!finalPass) {
            // ---- claim the next iteration index (all lanes walk this loop together) ----
            int fin = 0;
            unsigned backoff = 200;
            for (;;) {
                const FrameFlags f = load_flags(g);
                if (f.can_quit || f.closed) {
                    it = -1;
                    break;
                }
                if (it < 0) {
                    if (!first && !f.opened) {   // (a helper only joins opened frames; kept for safety)
                        __nanosleep(backoff);
                        continue;
                    }
                    int t = -1;
                    if (lane == 0 && ld_volatile(&g->next_iter) < maxIterations) t = atomicAdd(&g->next_iter, 1);
                    t = __shfl_sync(FULL, t, 0);
                    if (t < 0 || t >= maxIterations) {   // all handed out: whoever applies the last one takes the final LM
                        it = -1;
                        break;
                    }
                    it = t;
                }
                if (it < f.applied + kRansacRing) break;   // the ring slot of this iteration is free
                // ring full: help the bookkeeping catch up, or wait for it
                const int r = try_apply(g, ring, words, maxIterations, inliersToStop, W, buf.open_list, b, lane, f.applied);
                if (r == 1) {
                    fin = 1;   // the early stop fired here: the ticket is dropped, this warp runs the final LM
                    break;
                }
                if (r < 0) {
                    __nanosleep(backoff);
                    if (backoff < 1000) backoff += 200;
                }
            }
            if (fin) finalPass = true;
            else if (it < 0) break;   // nothing left for this warp
            if (fin && lane == 0) buf.frame_times[b * 4 + 1] = global_timer();
        }
        int cnt = 0, m = 0;
        double cumulated = 0.0;
        double xs[6];
        if (!finalPass) {
            // ---- random subset: ransac::get_random_subset_with_score (ransac.hpp:77-103) ----
            if (lane == 0) {
                count(W, first ? DBG_HYP_FIRST : DBG_HYP_HELPER);
                int32_t* used = buf.subsets_used + (size_t(b) * buf.max_iterations + it) * RS_MAX_SUBSET;
                if (buf.subsets_in) {
                    // host-drawn (std::mt19937 + std::shuffle), already in the reference's prepended order
                    const int32_t* in = buf.subsets_in + (size_t(b) * buf.max_iterations + it) * RS_MAX_SUBSET;
                    for (int k = 0; k < RS_MAX_SUBSET; ++k) {
                        const int idx = in[k];
                        if (idx >= 0 && idx < n) subset[cnt++] = short(idx);
                    }
                    for (int k = cnt - 1; k >= 0; --k) cumulated += score_of(sm.type[subset[k]]);
                }
                else {
                    // distinct uniform picks until the cumulated score reaches 1, each pick PREPENDED
                    const uint64_t key = rng_key(prm.seed, 1u, uint32_t(b)) ^ mix64(uint64_t(uint32_t(it)) << 20);
                    short picks[RS_MAX_SUBSET];
                    uint64_t ctr = 0;
                    while (cnt < RS_MAX_SUBSET && cnt < n && cumulated < 1.0) {
                        const int idx = int(mix64(key + ctr++) % uint64_t(n));
                        bool dup = false;
                        for (int k = 0; k < cnt; ++k) dup = dup || (picks[k] == idx);
                        if (dup) continue;
                        picks[cnt++] = short(idx);
                        cumulated += score_of(sm.type[idx]);
                    }
                    for (int k = 0; k < cnt; ++k) subset[k] = picks[cnt - 1 - k];
                }
                for (int k = 0; k < RS_MAX_SUBSET; ++k) used[k] = k < cnt ? int(subset[k]) : -1;
                for (int k = 0; k < cnt; ++k) m += parts_of(sm.type[subset[k]]);
            }
            P.idx = subset;
#pragma unroll
            for (int j = 0; j < 6; ++j) xs[j] = x0[j];
        }
        else {
            // ---- final optimisation on the winning inlier set, from the winning pose (:229-262) ----
            if (lane == 0) {
                volatile unsigned* best = ring + size_t(kRansacRing) * words;
                for (int w = 0; w < words; ++w) {
                    unsigned bits = best[w];
                    while (bits) {
                        const int i = w * 32 + (__ffs(bits) - 1);
                        bits &= bits - 1;
                        sm.idx[cnt++] = short(i);
                        cumulated += score_of(sm.type[i]);
                        m += parts_of(sm.type[i]);
                    }
                }
                out->n_inliers = ld_volatile(&g->best_inliers);
                out->iterations_run = ld_volatile(&g->started);
                out->best_iteration = ld_volatile(&g->best_iteration);
                out->score = ld_volatile(&g->max_score);
            }
            P.idx = sm.idx;
#pragma unroll
            for (int j = 0; j < 6; ++j) xs[j] = ld_volatile(&g->best_x[j]);
        }
        cnt = __shfl_sync(FULL, cnt, 0);
        m = __shfl_sync(FULL, m, 0);
        cumulated = __shfl_sync(FULL, cumulated, 0);
        __syncwarp();
        bool ok = cumulated >= 1.0;  // a hypothesis without enough score is skipped; final: status stays 0
        if (ok) {
            P.n = cnt;
            ok = optimize_pose_warp<P2D>(S, P, prm.K, xs, m, cumulated, prm.lm_max_fev, lane, finalPass ? nullptr : &g->can_quit);
        }
        if (finalPass) {
            int stage = 0;
            if (cumulated >= 1.0 && !ok) {
                if (lane == 0) out->status = -1;
            }
            else if (ok) {
                volatile unsigned* best = ring + size_t(kRansacRing) * words;
                for (int i = lane; i < n; i += 32) buf.mask[size_t(b) * M + i] = (best[i >> 5] >> (i & 31)) & 1u;
                if (lane == 0) {
                    double q[4];
                    quaternion_from_coefficients(S.x, q);
                    for (int j = 0; j < 3; ++j) out->pose[j] = S.x[j];
                    for (int j = 0; j < 4; ++j) out->pose[3 + j] = q[j];
                    for (int j = 0; j < 7; ++j) buf.poses[b * 7 + j] = out->pose[j];
                    out->status = prm.n_variance == 0 ? 1 : -2;  // -2 until the covariance validates it
                    PoseFrameState* sp = buf.state + b;
                    sp->stage = 1;
                    for (int j = 0; j < 6; ++j) sp->final_x[j] = S.x[j];
                    sp->n_inliers = cnt;
                    sp->inlier_residuals = m;
                    sp->inlier_score = cumulated;
                }
                for (int k = lane; k < cnt; k += 32) buf.inlier_idx[size_t(b) * M + k] = sm.idx[k];
                stage = 1;
            }
            // hand the frame over: its Monte-Carlo solves may start on any SM from here on
            __syncwarp();
            if (lane == 0) {
                __threadfence();
                if (stage == 1 && prm.n_variance > 0) {
                    const int s = atomicAdd(&W->n_ready, 1);
                    st_volatile(&buf.ready[s], b + 1);
                }
                buf.frame_times[b * 4 + 2] = global_timer();
                atomicMax(&W->t_ransac_end, buf.frame_times[b * 4 + 2]);
                __threadfence();
                atomicAdd(&W->frames_done, 1);
            }
            break;
        }
        const int slot = it % kRansacRing;
        int nIn = 0;
        double score = 0.0;
        if (ok) {
            // ---- get_features_inliers_outliers (pose_optimization.cpp:33-72) over all features ----
            if (lane == 0) make_xform(S.x, S.T);
            __syncwarp();
            for (int w = 0; w < words; ++w) {
                const int i = w * 32 + lane;
                bool in = false;
                if (i < n) {
                    double o[4], mm[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) o[c] = s_obs[c * M + i], mm[c] = s_map[c * M + i];
                    in = feature_is_inlier<P2D>(sm.type[i], o, mm, S.T, prm.K, P.aux, M, i);
                }
                const unsigned bits = __ballot_sync(FULL, in);
                if (lane == 0) wmask[w] = bits;
                nIn += __popc(bits);
            }
            __syncwarp();
            if (lane == 0) {
                // the score is accumulated in list order, like the reference's running double
                for (int w = 0; w < words; ++w) {
                    unsigned bits = wmask[w];
                    while (bits) {
                        const int i = w * 32 + (__ffs(bits) - 1);
                        bits &= bits - 1;
                        score += score_of(sm.type[i]);
                    }
                }
            }
            // the mask travels to the frame's ring slot (free: iteration `it` only started once it - ring was applied)
            volatile unsigned* dst = ring + size_t(slot) * words;
            for (int w = lane; w < words; w += 32) dst[w] = wmask[w];
        }
        else if (lane == 0 && ld_volatile(&g->can_quit))
            count(W, DBG_ABORTED);
        // ---- publish the result, then fold in whatever is ready ----
        if (lane < 6) st_volatile(&g->slot[slot].x[lane], S.x[lane]);
        if (lane == 0) {
            st_volatile(&g->slot[slot].score, score);
            st_volatile(&g->slot[slot].inliers, nIn);
            st_volatile(&g->slot[slot].ok, ok ? 1 : 0);
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            st_volatile(&g->done[slot], it + 1);
        }
        __syncwarp();
        for (;;) {
            const FrameFlags f = load_flags(g);
            if (f.can_quit || f.closed) break;
            // if another warp holds the lock it re-checks after releasing it
            const int r = try_apply(g, ring, words, maxIterations, inliersToStop, W, buf.open_list, b, lane, f.applied);
            if (r == 1) {
                finalPass = true;
                if (lane == 0) buf.frame_times[b * 4 + 1] = global_timer();
            }
            if (r != 0) break;
        }
    }
    __syncthreads();
    if (!first && threadIdx.x == 0) atomicSub(&g->joiners, 1);
}

template <bool P2D>
__device__ __noinline__ void frame_task_outlined(const PoseBuffers& buf, const PoseLaunch& prm, const FusedSmem& sm, PoseWork* W,
                                                const int b, const int arg, const int n, const int warp, const int lane)
{
    frame_task<P2D>(buf, prm, sm, W, b, arg, n, warp, lane);
}

// MC_ROLE = false: frame role only (28 KB of shared memory per CTA). MC_ROLE = true: Monte-Carlo role, and the frame role out
// of line. Both at 168 registers, three CTAs per SM: 58 KB per Monte-Carlo CTA allows no fourth, and capping the LM body at
// 128 registers costs the sample solves a third of their speed (measured).
template <bool P2D, bool MC_ROLE>
__global__ void __launch_bounds__(THREADS, RS_POSE_CTAS_PER_SM) pose_fused_kernel(const PoseBuffers buf, const PoseLaunch prm)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int M = buf.max_matches;
    const int words = (M + 31) / 32;
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool do_ransac = prm.run_ransac != 0;
    const bool do_mc = MC_ROLE && prm.n_variance > 0 && prm.run_mc != 0;
    FusedSmem sm;
    fused_carve(&sm, smem_raw, M, nwarps, do_mc);
    PoseWork* W = buf.work;
    const int B = prm.batch;
    const int groups = do_mc ? (prm.n_variance + nwarps - 1) / nwarps : 1;   // Monte-Carlo tasks per frame
    const int maxIterations = prm.max_iterations;

    if (threadIdx.x == 0) atomicMin(&W->t_first, global_timer());

    for (;;) {
        // ================================ pick a task (warp 0) ================================
        if (warp == 0) {
            int kind = -1, frame = -1, arg = 0;
            unsigned backoff = 200;
            while (kind < 0) {
                // 1. a frame nobody has started
                if (do_ransac && ld_volatile(&W->join_ticket) < B) {
                    int t = 0;
                    if (lane == 0) t = atomicAdd(&W->join_ticket, 1);
                    t = __shfl_sync(FULL, t, 0);
                    if (t < B) {
                        kind = TASK_RANSAC, frame = prm.frame0 + t, arg = 1;   // arg 1: first CTA of the frame
                        break;
                    }
                }
                // 2. a Monte-Carlo task. Tickets, not compare-and-swap: with hundreds of CTAs looking for work at once a CAS
                // loop hands out one task per L2 round trip. A ticket is only drawn when a task is there to be had; the race
                // may still push it past the last published frame, then the CTA waits for that frame (its hypotheses are
                // running on some CTA: every frame has been started by the time this branch is reached).
                if (do_mc) {
                    int t = -1;
                    if (lane == 0) {
                        const int h = ld_volatile(&W->mc_head);
                        const int s = h / groups;
                        // while frames are still in their RANSAC stage the Monte-Carlo solves (FP64 throughput work) are held
                        // to prm.mc_cap tasks in flight: every FP64 instruction of theirs queues in front of the one-lane
                        // algebra of the hypothesis chains, which is what the batch is waiting for
                        const bool throttled = ld_volatile(&W->frames_done) < B && h - ld_volatile(&W->mc_done_tasks) >= prm.mc_cap;
                        if (!throttled && s < B && ld_volatile(&buf.ready[s]) != 0) t = atomicAdd(&W->mc_head, 1);
                    }
                    t = __shfl_sync(FULL, t, 0);
                    if (t >= 0) {
                        const int s = t / groups;
                        int fr = 0;
                        unsigned wait = 200;
                        while (s < B) {
                            fr = ld_volatile(&buf.ready[s]);
                            if (fr != 0) break;
                            if (ld_volatile(&W->frames_done) >= B && s >= ld_volatile(&W->n_ready)) break;   // no such frame will come
                            __nanosleep(wait);
                            if (wait < 2000) wait += 200;
                        }
                        if (fr != 0) {
                            kind = TASK_MC, frame = fr - 1, arg = t % groups;
                            break;
                        }
                        continue;   // the ticket was beyond the last frame
                    }
                }
                // 3. a frame that takes helpers: the one with the most iterations left per CTA already on it
                if (do_ransac) {
                    const int nopen = ld_volatile(&W->n_open);
                    int best_f = -1, best_gain = 0;
                    for (int k = lane; k < nopen; k += 32) {
                        const int f = ld_volatile(&buf.open_list[k]) - 1;
                        if (f < 0) continue;
                        const RansacFrame* g = buf.rframe + f;
                        const FrameFlags ff = load_flags(g);
                        if (ff.can_quit || ff.closed) continue;
                        const int left = maxIterations - ld_volatile(&g->next_iter);
                        const int gain = left / (ld_volatile(&g->joiners) + 2);
                        if (gain > best_gain) best_gain = gain, best_f = f;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const int og = __shfl_xor_sync(FULL, best_gain, o), of = __shfl_xor_sync(FULL, best_f, o);
                        if (og > best_gain || (og == best_gain && of > best_f)) best_gain = og, best_f = of;
                    }
                    if (best_f >= 0 && best_gain >= prm.help_min) {
                        if (lane == 0) {
                            atomicAdd(&buf.rframe[best_f].joiners, 1);
                            count(W, DBG_HELPER_JOINS);
                        }
                        kind = TASK_RANSAC, frame = best_f, arg = 0;
                        break;
                    }
                }
                // 4. everything handed out?
                if (!do_mc && !prm.linger && ld_volatile(&W->join_ticket) >= B) {   // every frame has its CTA: nothing left for this one
                    kind = TASK_EXIT;
                    break;
                }
                if (ld_volatile(&W->frames_done) >= B) {
                    bool more = false;
                    if (do_mc) {
                        const int s = ld_volatile(&W->mc_head) / groups;
                        more = s < B && ld_volatile(&buf.ready[s]) != 0;
                    }
                    if (!more) {
                        kind = TASK_EXIT;
                        break;
                    }
                    continue;
                }
                __nanosleep(backoff);
                if (backoff < 2000) backoff += 200;
            }
            if (lane == 0) sm.ctl[0] = kind, sm.ctl[1] = frame, sm.ctl[2] = arg;
        }
        __syncthreads();
        const int kind = sm.ctl[0], b = sm.ctl[1], arg = sm.ctl[2];
        if (kind == TASK_EXIT) break;

        // per-frame state: written by the prepare kernel (an earlier launch) and, for a Monte-Carlo task, completed by the
        // frame's final LM on another SM of THIS launch - read past L1
        const int n = __ldcg(&buf.state[b].n);

        if (kind == TASK_RANSAC) {
            if (MC_ROLE)
                frame_task_outlined<P2D>(buf, prm, sm, W, b, arg, n, warp, lane);
            else
                frame_task<P2D>(buf, prm, sm, W, b, arg, n, warp, lane);
            continue;
        }

        // ======================= Monte-Carlo task: `nwarps` samples of frame b (pose_optimization.cpp:379-412) =======================
        if (MC_ROLE) {
            const PoseFrameState* stp = buf.state + b;
            WarpLM& S = sm.lm[warp];
            const int cnt = __ldcg(&stp->n_inliers);
            for (int i = threadIdx.x; i < M; i += blockDim.x) {
                sm.type[i] = buf.type[size_t(b) * M + i];
                if (i < cnt) sm.idx[i] = __ldcg(buf.inlier_idx + size_t(b) * M + i);
#pragma unroll
                for (int c = 0; c < 4; ++c) sm.obs[c * M + i] = buf.obs[(size_t(b) * 4 + c) * M + i];
            }
            if (threadIdx.x == 0) count(W, DBG_MC_TASKS);
            __syncthreads();
            const int sample = arg * nwarps + warp;
            if (sample < prm.n_variance) {
                double* pmap = sm.pmap + size_t(warp) * 4 * M;
                const double* gmap = buf.map + size_t(b) * 4 * M;
                const double* gsig = buf.sigma + size_t(b) * 4 * M;
                for (int k = lane; k < cnt; k += 32) {
                    const int i = sm.idx[k];
                    double gn[4];
                    if (buf.normals_in) {
                        const double* src = buf.normals_in + ((size_t(b) * prm.n_variance + sample) * M + i) * 4;
                        gn[0] = src[0], gn[1] = src[1], gn[2] = src[2], gn[3] = src[3];
                    }
                    else {
                        device_normals(prm.seed, b, sample, i, gn);
                    }
                    if (sm.type[i] == RS_FEAT_POINT) {
                        // map_point.cpp:49-58
#pragma unroll
                        for (int c = 0; c < 3; ++c) pmap[c * M + i] = gmap[c * M + i] + gn[c] * gsig[c * M + i];
                        pmap[3 * M + i] = 0.0;
                    }
                    else if (P2D && sm.type[i] == RS_FEAT_POINT2D) {
                        // map_point2d.cpp:49-73: theta then phi, clamped to [0, pi] / [-pi, pi]; nothing else varies
                        const double th = gmap[i] + gn[0] * gsig[i], ph = gmap[M + i] + gn[1] * gsig[M + i];
                        pmap[i] = th < 0.0 ? 0.0 : (kPi < th ? kPi : th);
                        pmap[M + i] = ph < -kPi ? -kPi : (kPi < ph ? kPi : ph);
                        pmap[2 * M + i] = 0.0;
                        pmap[3 * M + i] = 0.0;
                    }
                    else {
                        // map_primitive.cpp:66-77: perturbed normal renormalised (twice: vector + PlaneCoordinates ctor)
                        double nn[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c) nn[c] = gmap[c * M + i] + gn[c] * gsig[c * M + i];
                        normalize3(nn);
                        normalize3(nn);
#pragma unroll
                        for (int c = 0; c < 3; ++c) pmap[c * M + i] = nn[c];
                        pmap[3 * M + i] = gmap[3 * M + i] + gn[3] * gsig[3 * M + i];
                    }
                }
                __syncwarp();
                Problem P;
                P.n = cnt, P.idx = sm.idx, P.type = sm.type, P.obs = sm.obs, P.map = pmap, P.M = M;
                P.aux = buf.aux + size_t(b) * 4 * M;
                double xs[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) xs[j] = __ldcg(&stp->final_x[j]);
                const bool ok = optimize_pose_warp<P2D>(S, P, prm.K, xs, __ldcg(&stp->inlier_residuals), __ldcg(&stp->inlier_score),
                                                        prm.lm_max_fev, lane, nullptr);
                if (lane == 0) {
                    double v[6] = {0, 0, 0, 0, 0, 0};
                    if (ok) pose_vector6(S.x, v);
                    double* dst = buf.v6 + (size_t(b) * buf.max_variance + sample) * 6;
                    for (int j = 0; j < 6; ++j) dst[j] = v[j];
                    buf.v_ok[size_t(b) * buf.max_variance + sample] = ok ? 1 : 0;
                }
            }
            __syncthreads();
            // the CTA that finishes a frame's last sample group reduces its covariance (pose_optimization.cpp:414-437)
            if (threadIdx.x == 0) {
                __threadfence();
                sm.ctl[3] = atomicAdd(&buf.mc_done[b], 1) == groups - 1 ? 1 : 0;
                atomicAdd(&W->mc_done_tasks, 1);
            }
            __syncthreads();
            if (sm.ctl[3] && warp == 0) {
                frame_covariance_warp(buf, prm, b, lane);
                if (lane == 0) buf.frame_times[b * 4 + 3] = global_timer();
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) atomicMax(&W->t_last, global_timer());
}

// RS_RNG_REFERENCE runs the fused kernel in two halves around the host's Gaussian draws: this fills the hand-over list of
// the second half with every frame whose final pose is available.
__global__ void pose_publish_all_kernel(const PoseBuffers buf, const PoseLaunch prm)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    PoseWork* W = buf.work;
    int nr = 0;
    for (int t = 0; t < prm.batch; ++t) {
        const int b = prm.frame0 + t;
        buf.mc_done[b] = 0;
        buf.ready[t] = 0;
        if (buf.state[b].stage == 1) buf.ready[nr++] = b + 1;
    }
    W->n_ready = nr;
    W->mc_head = 0;
    W->frames_done = prm.batch;
    W->join_ticket = prm.batch;
}

__global__ void pose_export_normals_kernel(const PoseLaunch prm, const int M, double* normals)
{
    const size_t total = size_t(prm.batch) * prm.n_variance * M;
    for (size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += size_t(gridDim.x) * blockDim.x) {
        const int i = int(t % M);
        const int sample = int((t / M) % prm.n_variance);
        const int b = int(t / (size_t(M) * prm.n_variance));
        double g[4];
        device_normals(prm.seed, b, sample, i, g);
        for (int c = 0; c < 4; ++c) normals[t * 4 + c] = g[c];
    }
}


constexpr size_t kSmemPerCta = 232448;   // 227 KB opt-in limit of sm_100
constexpr size_t kSmemPerSm = 233472;    // 228 KB, 1 KB reserved per resident CTA

// Warps per CTA of the fused kernel for a match capacity M: the choice that keeps the most warps resident per SM (four warps,
// four CTAs per SM, up to M = 400: every warp of a Monte-Carlo task holds a perturbed copy of the frame's map side; fewer
// warps per CTA beyond). 0 = even one warp does not fit.
int fused_warps_for(const int M)
{
    int best = 0, best_resident = 0;
    for (int w = WARPS; w >= 1; w >>= 1) {
        const size_t smem = fused_carve(nullptr, nullptr, M, w);
        if (smem > kSmemPerCta) continue;
        const int ctas = int(std::min<size_t>(kSmemPerSm / (smem + 1024), size_t(64 / w)));
        if (w * ctas > best_resident) best = w, best_resident = w * ctas;
    }
    return best;
}

}  // namespace

int pose_max_matches_supported()
{
    static int limit = 0;
    if (limit == 0) {
        int lo = 1, hi = 32767;   // the fused kernel stages one frame's match list in shared memory
        while (lo < hi) {
            const int mid = (lo + hi + 1) / 2;
            if (fused_warps_for(mid) > 0) lo = mid; else hi = mid - 1;
        }
        limit = lo;
    }
    return limit;
}

int launch_pose_prepare(const PoseBuffers& buf, const PoseLaunch& prm, cudaStream_t stream)
{
    pose_prepare_kernel<<<prm.batch, THREADS, 0, stream>>>(buf, prm);
    RS_LAUNCH_CHECK();
    return RS_OK;
}

// Two instantiations of the fused kernel: the usual one knows only points and planes; the other also carries the
// inverse-depth ("line") residual, whose sincos / two-projection / seven-evaluation code would otherwise cost the hot loops
// registers (measured on the former Monte-Carlo kernel: +50 % when merely compiled in). prm.has_point2d selects.
int launch_pose_fused(const PoseBuffers& buf, const PoseLaunch& prm, cudaStream_t stream)
{
    const int warps = fused_warps_for(buf.max_matches);
    if (warps == 0) return RS_ERR_INVALID_ARG;   // rs_pose_create refuses such capacities (pose_max_matches_supported)
    const bool mc_role = prm.run_mc && prm.n_variance > 0;
    if (!prm.run_ransac && !mc_role) return RS_OK;
    const size_t smem = fused_carve(nullptr, nullptr, buf.max_matches, warps, mc_role);
    static SmemOptIn optin[4];
    if (mc_role) {
        RS_CUDA_CHECK(optin[2].ensure(pose_fused_kernel<false, true>, smem));
        RS_CUDA_CHECK(optin[3].ensure(pose_fused_kernel<true, true>, smem));
    }
    else {
        RS_CUDA_CHECK(optin[0].ensure(pose_fused_kernel<false, false>, smem));
        RS_CUDA_CHECK(optin[1].ensure(pose_fused_kernel<true, false>, smem));
    }
    if (prm.publish_all) {
        pose_publish_all_kernel<<<1, 32, 0, stream>>>(buf, prm);
        RS_LAUNCH_CHECK();
    }
    // persistent grid: what is resident at once, or fewer when the batch cannot feed that many CTAs
    int dev = 0, sms = 0, per_sm = 0;
    RS_CUDA_CHECK(cudaGetDevice(&dev));
    RS_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    void (*kernel)(const PoseBuffers, const PoseLaunch) =
            mc_role ? (prm.has_point2d ? pose_fused_kernel<true, true> : pose_fused_kernel<false, true>)
                    : (prm.has_point2d ? pose_fused_kernel<true, false> : pose_fused_kernel<false, false>);
    RS_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, warps * 32, smem));
    if (per_sm < 1) per_sm = 1;
    if (prm.ctas_per_sm > 0 && prm.ctas_per_sm < per_sm) per_sm = prm.ctas_per_sm;
    const int groups = mc_role ? (prm.n_variance + warps - 1) / warps : 0;
    long long useful = 0;
    if (prm.run_ransac) useful += (long long)prm.batch * ((prm.linger || mc_role) ? std::max(1, (prm.max_iterations + 8 * warps - 1) / (8 * warps)) : 1);
    if (mc_role) useful += (long long)prm.batch * groups;
    const int grid = int(std::max<long long>(1, std::min<long long>((long long)sms * per_sm, useful)));
    kernel<<<grid, warps * 32, smem, stream>>>(buf, prm);
    RS_LAUNCH_CHECK();
    return RS_OK;
}

int pose_work_times(const PoseBuffers& buf, float* ransac_phase_ms, float* total_ms, cudaStream_t stream)
{
    PoseWork w;
    RS_CUDA_CHECK(cudaMemcpyAsync(&w, buf.work, sizeof(PoseWork), cudaMemcpyDeviceToHost, stream));
    RS_CUDA_CHECK(cudaStreamSynchronize(stream));
    const bool ok = w.t_first != ~0ull && w.t_last >= w.t_first;
    if (ransac_phase_ms) *ransac_phase_ms = ok && w.t_ransac_end >= w.t_first ? float(double(w.t_ransac_end - w.t_first) * 1e-6) : 0.f;
    if (total_ms) *total_ms = ok ? float(double(w.t_last - w.t_first) * 1e-6) : 0.f;
    return RS_OK;
}

int launch_pose_export_normals(const PoseBuffers& buf, const PoseLaunch& prm, double* normals, cudaStream_t stream)
{
    pose_export_normals_kernel<<<sm_count() * 4, 256, 0, stream>>>(prm, buf.max_matches, normals);
    RS_LAUNCH_CHECK();
    return RS_OK;
}

}  // namespace rs
