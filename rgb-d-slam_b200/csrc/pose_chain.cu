// The three-launch chain of the pose solve: preparation (pose_solve.cu), then the two kernels here.
//   pose_ransac_kernel   : one CTA of four warps per frame, the frame's ring of hypothesis results in SHARED memory; warps claim
//                          hypothesis indices dynamically, the serial best-so-far / early-stop rule is applied in iteration order
//                          as results complete, then the final LM on the winning inlier set (pose_optimization.cpp:107-262).
//   pose_variance_kernel : one warp per Monte-Carlo sample, eight samples per CTA (:361-412, 482-501).
//   pose_covariance_kernel: one warp per frame, 6x6 covariance (one entry per lane, summed in sample order) + validity (:414-437).
//                          (Folding it into the Monte-Carlo kernel - last CTA of a frame - was measured: the fences and barriers
//                          cost that kernel 0.04 ms per 256 frames, the separate launch 0.016.)
// The fused kernel of pose_solve.cu (one persistent launch, the ring in global memory, any CTA can work on any frame) is what
// a solve ALONE or with hundreds of hypotheses per frame is fastest with. This chain is kept for the shape the benchmark step
// has - 119 hypotheses of which four or five run, beside the cell-graph segmentation kernel: its CTAs are small (21 KB) or
// short-lived, so the segmentation's 108 KB CTAs find room between them, which persistent 58 KB CTAs do not leave
// (measured: 1.69 ms per 256-frame step with the chain, 1.78-2.0 ms with the fused kernel; 1.50 against 1.20 ms for the solve
// alone). api_pose.cu picks by hypothesis count and by what the caller says runs beside the solve.
#define RS_ABORT_AT_TOP   // the early-stop flag of a frame lives in shared memory here: test it where the loop starts
#include "pose_lm.cuh"

namespace rs {

namespace {

constexpr int WARPS = 8;             // warps per CTA of the Monte-Carlo kernel (one sample each)
constexpr int THREADS = WARPS * 32;
// RANSAC: warps per frame, each running one hypothesis at a time. The reference never stops before iteration 3, so
// hypotheses 0..3 are always needed; beyond them the warps run ahead of the serial early-stop rule speculatively.
constexpr int RWARPS = 4;
constexpr int RTHREADS = RWARPS * 32;

constexpr int RRING = 8;   // hypotheses that may be in flight or finished-but-unapplied beyond the serial rule's position

struct RansacShared {
    double best_x[6];
    double max_score;
    int best_inliers, best_iteration, can_quit, started;
    int next_iter;   // next hypothesis index to hand out
    int applied;     // hypotheses whose bookkeeping has been applied, in iteration order
    int lock;        // guards the in-order bookkeeping
    int done[RRING]; // iteration + 1 once the slot's result is complete
    double hyp_x[RRING][6];
    double hyp_score[RRING];
    int hyp_ok[RRING];
    int hyp_inliers[RRING];
};

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) / 16 * 16; }

// shared-memory carve-up of the RANSAC kernel
struct RansacSmem {
    int32_t* type;
    double* obs;
    double* map;
    WarpLM* lm;
    unsigned* hyp_mask;   // [RRING][words]
    unsigned* best_mask;  // [words]
    short* subset;        // [RWARPS][RS_MAX_SUBSET]
    short* inlier_idx;    // [M]
    RansacShared* sh;
};
__host__ __device__ inline size_t ransac_carve(RansacSmem* s, unsigned char* base, const int M)
{
    const int words = (M + 31) / 32;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        unsigned char* p = base ? base + o : nullptr;
        o = align16(o + bytes);
        return p;
    };
    unsigned char* obs = take(sizeof(double) * 4 * M);
    unsigned char* map = take(sizeof(double) * 4 * M);
    unsigned char* lm = take(sizeof(WarpLM) * RWARPS);
    unsigned char* sh = take(sizeof(RansacShared));
    unsigned char* type = take(sizeof(int32_t) * M);
    unsigned char* hm = take(sizeof(unsigned) * RRING * words);
    unsigned char* bm = take(sizeof(unsigned) * words);
    unsigned char* sub = take(sizeof(short) * RWARPS * RS_MAX_SUBSET);
    unsigned char* ii = take(sizeof(short) * M);
    if (s) {
        s->obs = reinterpret_cast<double*>(obs), s->map = reinterpret_cast<double*>(map);
        s->lm = reinterpret_cast<WarpLM*>(lm), s->sh = reinterpret_cast<RansacShared*>(sh);
        s->type = reinterpret_cast<int32_t*>(type), s->hyp_mask = reinterpret_cast<unsigned*>(hm);
        s->best_mask = reinterpret_cast<unsigned*>(bm), s->subset = reinterpret_cast<short*>(sub);
        s->inlier_idx = reinterpret_cast<short*>(ii);
    }
    return o;
}

// compute_pose_with_ransac (pose_optimization.cpp:107-262): one CTA per frame.
template <bool P2D>
__global__ void __launch_bounds__(RTHREADS, 3) pose_ransac_kernel(const PoseBuffers buf, const PoseLaunch prm)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = prm.frame0 + blockIdx.x;
    const int M = buf.max_matches;
    const int words = (M + 31) / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    RansacSmem sm;
    ransac_carve(&sm, smem_raw, M);
    const PoseFrameState st = buf.state[b];
    const int n = st.n;
    if (!st.valid || st.total_score < 1.0) return;  // out[b] already says status 0, pose = current pose

    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        sm.type[i] = buf.type[size_t(b) * M + i];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            sm.obs[c * M + i] = buf.obs[(size_t(b) * 4 + c) * M + i];
            sm.map[c * M + i] = buf.map[(size_t(b) * 4 + c) * M + i];
        }
    }
    double x0[6];
    coefficients_from_pose(buf.cur_pose + b * 7, x0);
    RansacShared& sh = *sm.sh;
    if (threadIdx.x == 0) {
        sh.max_score = 1.0;
        sh.best_inliers = 0, sh.best_iteration = -1, sh.can_quit = 0, sh.started = 0;
        sh.next_iter = 0, sh.applied = 0, sh.lock = 0;
        for (int k = 0; k < RRING; ++k) sh.done[k] = 0;
        for (int j = 0; j < 6; ++j) sh.best_x[j] = x0[j];
    }
    for (int i = threadIdx.x; i < words; i += blockDim.x) sm.best_mask[i] = 0u;
    __syncthreads();
    if (prm.final_only) {
        // the hypotheses were evaluated and folded elsewhere (pose_wide.cu): take the serial loop's end state from there
        const HypFold& f = buf.fold[b];
        if (threadIdx.x == 0) {
            sh.max_score = f.max_score;
            sh.best_inliers = f.best_inliers, sh.best_iteration = f.best_iteration, sh.started = f.started;
            for (int j = 0; j < 6; ++j) sh.best_x[j] = f.best_x[j];
        }
        for (int i = threadIdx.x; i < words; i += blockDim.x) sm.best_mask[i] = buf.fold_mask[size_t(b) * words + i];
        __syncthreads();
    }

    const int maxIterations = prm.final_only ? 0 : prm.max_iterations;
    const unsigned inliersToStop = unsigned(ceil(double(n) * kEarlyStopProportion));
    WarpLM& S = sm.lm[warp];
    short* subset = sm.subset + warp * RS_MAX_SUBSET;
    Problem P;
    P.type = sm.type, P.obs = sm.obs, P.map = sm.map, P.M = M;
    P.aux = buf.aux + size_t(b) * 4 * M;
    volatile int* v_can_quit = &sh.can_quit;
    volatile int* v_next = &sh.next_iter;
    volatile int* v_applied = &sh.applied;
    volatile int* v_done = sh.done;

    // The reference's serial bookkeeping (:151-227), applied strictly in iteration order by whoever holds the lock:
    // every finished hypothesis whose predecessors have all been applied is folded into the best-so-far state; the early
    // stop freezes the state, exactly where the serial loop would have left it.
    auto apply_ready = [&]() {
        while (!*v_can_quit) {
            const int i = *v_applied;
            if (i >= maxIterations || v_done[i % RRING] != i + 1) break;
            __threadfence_block();
            const int slot = i % RRING;
            ++sh.started;
            if (sh.hyp_ok[slot]) {
                const double hs = sh.hyp_score[slot];
                if (hs >= 1.0) {
                    const bool canOverload =
                            (hs > sh.max_score) || (fabs(hs - sh.max_score) <= 0.1 && sh.best_inliers < sh.hyp_inliers[slot]);
                    if (canOverload) {
                        sh.max_score = hs;
                        for (int j = 0; j < 6; ++j) sh.best_x[j] = sh.hyp_x[slot][j];
                        for (int k = 0; k < words; ++k) sm.best_mask[k] = sm.hyp_mask[slot * words + k];
                        sh.best_inliers = sh.hyp_inliers[slot];
                        sh.best_iteration = i;
                    }
                    if (i >= 3 && unsigned(sh.best_inliers) > inliersToStop) *v_can_quit = 1;
                }
            }
            __threadfence_block();
            *v_applied = i + 1;
        }
    };

    rs_pose_out* out = buf.out + b;
    // One loop, one LM call site (the LM body is inlined once). Hypothesis passes: every warp keeps claiming the next
    // iteration index (hypotheses are independent: each LM starts from the current pose) up to RRING ahead of the serial
    // rule; a hypothesis that finishes after the early stop is simply never applied, and one still running is dropped at
    // its next LM iteration. When nothing is left to claim the warps meet once, warp 0 runs the final optimisation on
    // the winning inlier set and the others leave.
    bool finalPass = false;
    for (;;) {
        int it = -1;
        if (!finalPass) {
            if (lane == 0) {
                while (!*v_can_quit) {
                    const int cur = *v_next;
                    if (cur >= maxIterations) break;
                    if (cur >= *v_applied + RRING) {   // ring full: help the bookkeeping catch up, or wait for it
                        if (v_done[*v_applied % RRING] == *v_applied + 1 && atomicCAS(&sh.lock, 0, 1) == 0) {
                            apply_ready();
                            __threadfence_block();
                            atomicExch(&sh.lock, 0);
                        }
                        else
                            __nanosleep(200);
                        continue;
                    }
                    if (atomicCAS(&sh.next_iter, cur, cur + 1) == cur) {
                        it = cur;
                        break;
                    }
                }
            }
            it = __shfl_sync(FULL, it, 0);
            if (it < 0) {
                __syncthreads();   // every warp arrives here exactly once; all claimed hypotheses are finished
                if (threadIdx.x == 0) apply_ready();
                __syncthreads();
                if (warp != 0) return;
                finalPass = true;
            }
        }
        int cnt = 0, m = 0;
        double cumulated = 0.0;
        double xs[6];
        if (!finalPass) {
            // ---- random subset: ransac::get_random_subset_with_score (ransac.hpp:77-103) ----
            if (lane == 0) {
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
                for (int w = 0; w < words; ++w) {
                    unsigned bits = sm.best_mask[w];
                    while (bits) {
                        const int i = w * 32 + (__ffs(bits) - 1);
                        bits &= bits - 1;
                        sm.inlier_idx[cnt++] = short(i);
                        cumulated += score_of(sm.type[i]);
                        m += parts_of(sm.type[i]);
                    }
                }
                out->n_inliers = sh.best_inliers;
                out->iterations_run = sh.started;
                out->best_iteration = sh.best_iteration;
                out->score = sh.max_score;
            }
            P.idx = sm.inlier_idx;
#pragma unroll
            for (int j = 0; j < 6; ++j) xs[j] = sh.best_x[j];
        }
        cnt = __shfl_sync(FULL, cnt, 0);
        m = __shfl_sync(FULL, m, 0);
        cumulated = __shfl_sync(FULL, cumulated, 0);
        __syncwarp();
        bool ok = cumulated >= 1.0;  // a hypothesis without enough score is skipped; final: status stays 0
        if (finalPass && !ok) return;
        if (ok) {
            P.n = cnt;
            ok = optimize_pose_warp<P2D>(S, P, prm.K, xs, m, cumulated, prm.lm_max_fev, lane, finalPass ? nullptr : v_can_quit);
        }
        if (finalPass) {
            if (!ok) {
                if (lane == 0) out->status = -1;
                return;
            }
            for (int i = lane; i < n; i += 32) buf.mask[size_t(b) * M + i] = (sm.best_mask[i >> 5] >> (i & 31)) & 1u;
            if (lane == 0) {
                double q[4];
                quaternion_from_coefficients(S.x, q);
                for (int j = 0; j < 3; ++j) out->pose[j] = S.x[j];
                for (int j = 0; j < 4; ++j) out->pose[3 + j] = q[j];
                for (int j = 0; j < 7; ++j) buf.poses[b * 7 + j] = out->pose[j];
                out->status = prm.n_variance == 0 ? 1 : -2;  // -2 until the covariance kernel validates it
                PoseFrameState* stp = buf.state + b;
                stp->stage = 1;
                for (int j = 0; j < 6; ++j) stp->final_x[j] = S.x[j];
                stp->n_inliers = cnt;
                stp->inlier_residuals = m;
                stp->inlier_score = cumulated;
            }
            for (int k = lane; k < cnt; k += 32) buf.inlier_idx[size_t(b) * M + k] = sm.inlier_idx[k];
            return;
        }
        const int slot = it % RRING;
        unsigned* hmask = sm.hyp_mask + slot * words;
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
                    for (int c = 0; c < 4; ++c) o[c] = sm.obs[c * M + i], mm[c] = sm.map[c * M + i];
                    in = feature_is_inlier<P2D>(sm.type[i], o, mm, S.T, prm.K, P.aux, M, i);
                }
                const unsigned bits = __ballot_sync(FULL, in);
                if (lane == 0) hmask[w] = bits;
                nIn += __popc(bits);
            }
            __syncwarp();
            if (lane == 0) {
                // the score is accumulated in list order, like the reference's running double
                for (int w = 0; w < words; ++w) {
                    unsigned bits = hmask[w];
                    while (bits) {
                        const int i = w * 32 + (__ffs(bits) - 1);
                        bits &= bits - 1;
                        score += score_of(sm.type[i]);
                    }
                }
            }
        }
        if (lane == 0) {
            sh.hyp_ok[slot] = ok ? 1 : 0;
            sh.hyp_score[slot] = score;
            sh.hyp_inliers[slot] = nIn;
            for (int j = 0; j < 6; ++j) sh.hyp_x[slot][j] = S.x[j];
            __threadfence_block();
            v_done[slot] = it + 1;
            // fold in whatever is ready; if another warp holds the lock it re-checks after releasing it
            while (!*v_can_quit && v_done[*v_applied % RRING] == *v_applied + 1 && *v_applied < maxIterations) {
                if (atomicCAS(&sh.lock, 0, 1) != 0) break;
                apply_ready();
                __threadfence_block();
                atomicExch(&sh.lock, 0);
            }
        }
        __syncwarp();
    }
}

// compute_pose_variance's loop body (pose_optimization.cpp:379-412) + compute_random_variation_of_pose (:482-501):
// grid (ceil(n_variance / warps), B), one warp per Monte-Carlo sample; warps per CTA = blockDim.x / 32 (8 unless the
// perturbed copies of a very long match list would not fit in shared memory, see launch_pose_variance).
// BATCHED: the serial 6x6 parts of the CTA's samples run together on the lanes of warp 0 (lm_minimize_cta).
template <bool P2D, bool BATCHED>
__global__ void __launch_bounds__(THREADS, 2) pose_variance_kernel(const PoseBuffers buf, const PoseLaunch prm)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = prm.frame0 + blockIdx.y;
    const int M = buf.max_matches;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PoseFrameState st = buf.state[b];
    if (st.stage != 1) return;
    const int n = st.n;
    // carve: obs[4][M] | pmap[warps][4][M] | WarpLM[warps] | type[M] | idx[M] | count
    const int nwarps = blockDim.x >> 5;
    double* s_obs = reinterpret_cast<double*>(smem_raw);
    double* s_pmap = s_obs + 4 * M;
    WarpLM* s_lm = reinterpret_cast<WarpLM*>(s_pmap + size_t(nwarps) * 4 * M);
    int32_t* s_type = reinterpret_cast<int32_t*>(s_lm + nwarps);
    short* s_idx = reinterpret_cast<short*>(s_type + M);
    const int cnt = st.n_inliers;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        s_type[i] = buf.type[size_t(b) * M + i];
        if (i < cnt) s_idx[i] = buf.inlier_idx[size_t(b) * M + i];
#pragma unroll
        for (int c = 0; c < 4; ++c) s_obs[c * M + i] = buf.obs[(size_t(b) * 4 + c) * M + i];
    }
    __syncthreads();
    const int sample = blockIdx.x * nwarps + warp;
    const bool mine = sample < prm.n_variance;
    if (!BATCHED && !mine) return;
    double* pmap = s_pmap + size_t(warp) * 4 * M;
    const double* gmap = buf.map + size_t(b) * 4 * M;
    const double* gsig = buf.sigma + size_t(b) * 4 * M;
    for (int k = lane; mine && k < cnt; k += 32) {
        const int i = s_idx[k];
        double g[4];
        if (buf.normals_in) {
            const double* src = buf.normals_in + ((size_t(b) * prm.n_variance + sample) * M + i) * 4;
            g[0] = src[0], g[1] = src[1], g[2] = src[2], g[3] = src[3];
        }
        else {
            device_normals(prm.seed, b, sample, i, g);
        }
        if (s_type[i] == RS_FEAT_POINT) {
            // map_point.cpp:49-58
#pragma unroll
            for (int c = 0; c < 3; ++c) pmap[c * M + i] = gmap[c * M + i] + g[c] * gsig[c * M + i];
            pmap[3 * M + i] = 0.0;
        }
        else if (P2D && s_type[i] == RS_FEAT_POINT2D) {
            // map_point2d.cpp:49-73: theta then phi, clamped to [0, pi] / [-pi, pi]; nothing else varies
            const double th = gmap[i] + g[0] * gsig[i], ph = gmap[M + i] + g[1] * gsig[M + i];
            pmap[i] = th < 0.0 ? 0.0 : (kPi < th ? kPi : th);
            pmap[M + i] = ph < -kPi ? -kPi : (kPi < ph ? kPi : ph);
            pmap[2 * M + i] = 0.0;
            pmap[3 * M + i] = 0.0;
        }
        else {
            // map_primitive.cpp:66-77: perturbed normal renormalised (twice: vector + PlaneCoordinates ctor)
            double nn[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) nn[c] = gmap[c * M + i] + g[c] * gsig[c * M + i];
            normalize3(nn);
            normalize3(nn);
#pragma unroll
            for (int c = 0; c < 3; ++c) pmap[c * M + i] = nn[c];
            pmap[3 * M + i] = gmap[3 * M + i] + g[3] * gsig[3 * M + i];
        }
    }
    __syncwarp();
    Problem P;
    P.n = cnt, P.idx = s_idx, P.type = s_type, P.obs = s_obs, P.map = pmap, P.M = M;
    P.aux = buf.aux + size_t(b) * 4 * M;
    WarpLM& S = s_lm[warp];
    double x0[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) x0[j] = st.final_x[j];
    bool ok;
    if (BATCHED)
        ok = optimize_pose_cta<P2D>(s_lm, nwarps, P, prm.K, x0, st.inlier_residuals, st.inlier_score, prm.lm_max_fev, warp, lane, mine);
    else
        ok = optimize_pose_warp<P2D>(S, P, prm.K, x0, st.inlier_residuals, st.inlier_score, prm.lm_max_fev, lane);
    if (mine && lane == 0) {
        double v[6] = {0, 0, 0, 0, 0, 0};
        if (ok) pose_vector6(S.x, v);
        double* dst = buf.v6 + (size_t(b) * buf.max_variance + sample) * 6;
        for (int j = 0; j < 6; ++j) dst[j] = v[j];
        buf.v_ok[size_t(b) * buf.max_variance + sample] = ok ? 1 : 0;
    }
}

// compute_pose_variance's reduction (pose_optimization.cpp:414-437): one warp per frame. Lanes split the samples for
// the mean, then lane e < 21 owns one entry of the upper triangle and sums it over the samples in sample order.
__global__ void __launch_bounds__(128) pose_covariance_kernel(const PoseBuffers buf, const PoseLaunch prm)
{
    const int b = prm.frame0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= prm.frame0 + prm.batch) return;
    if (buf.state[b].stage != 1 || prm.n_variance <= 0) return;
    const double* v6 = buf.v6 + size_t(b) * buf.max_variance * 6;
    const int32_t* vok = buf.v_ok + size_t(b) * buf.max_variance;
    rs_pose_out* out = buf.out + b;
    double medium[6] = {0, 0, 0, 0, 0, 0};
    int cnt = 0;
    for (int s = lane; s < prm.n_variance; s += 32)
        if (vok[s]) {
#pragma unroll
            for (int j = 0; j < 6; ++j) medium[j] += v6[s * 6 + j];
            ++cnt;
        }
    cnt = __reduce_add_sync(FULL, cnt);
#pragma unroll
    for (int j = 0; j < 6; ++j) medium[j] = warp_sum(medium[j]);
    if (lane == 0) out->n_variance_ok = cnt;
    if (unsigned(cnt) < unsigned(prm.n_variance) / 2u) {
        if (lane == 0) out->status = -2;
        return;
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) medium[j] /= double(cnt);
    int ei = 0, ej = 0;
    {
        int rem = lane < 21 ? lane : 0;
        while (rem >= 6 - ei) rem -= 6 - ei, ++ei;
        ej = ei + rem;
    }
    double mi = 0.0, mj = 0.0;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        if (j == ei) mi = medium[j];
        if (j == ej) mj = medium[j];
    }
    double acc = 0.0;
    for (int s = 0; s < prm.n_variance; ++s)
        if (vok[s]) acc += (v6[s * 6 + ei] - mi) * (v6[s * 6 + ej] - mj);
    acc /= double(cnt - 1);
    if (ei == ej) acc += 0.001;
    if (lane < 21) {
        out->cov[ei * 6 + ej] = acc;
        out->cov[ej * 6 + ei] = acc;
    }
    __syncwarp();
    if (lane == 0) {
        double cov[36];
        for (int i = 0; i < 36; ++i) cov[i] = out->cov[i];
        out->status = covariance_valid(cov) ? 1 : -2;
    }
}


size_t variance_smem_bytes(const int M, const int warps)
{
    return sizeof(double) * 4 * M + sizeof(double) * size_t(warps) * 4 * M + sizeof(WarpLM) * warps + sizeof(int32_t) * M +
           sizeof(short) * M + 16;
}

constexpr size_t kSmemPerCta = 232448;   // 227 KB opt-in limit of sm_100
constexpr size_t kSmemPerSm = 233472;    // 228 KB, 1 KB reserved per resident CTA

// Warps per CTA of the Monte-Carlo kernel for a match capacity M: the choice that keeps the most warps resident per SM
// (8 warps, two CTAs per SM, up to M = 370; fewer, fatter samples per CTA beyond). 0 = even one warp does not fit.
int variance_warps_for(const int M)
{
    int best = 0, best_resident = 0;
    for (int w = WARPS; w >= 1; w >>= 1) {
        const size_t smem = variance_smem_bytes(M, w);
        if (smem > kSmemPerCta) continue;
        const int ctas = int(std::min<size_t>(kSmemPerSm / (smem + 1024), size_t(64 / w)));
        if (w * ctas > best_resident) best = w, best_resident = w * ctas;
    }
    return best;
}

}  // namespace

// longest match list the chain's kernels can stage in the shared memory of one SM
bool pose_chain_supports(const int max_matches)
{
    return ransac_carve(nullptr, nullptr, max_matches) <= kSmemPerCta && variance_warps_for(max_matches) > 0;
}

int launch_pose_chain_ransac(const PoseBuffers& buf, const PoseLaunch& prm, cudaStream_t stream)
{
    const size_t smem = ransac_carve(nullptr, nullptr, buf.max_matches);
    if (smem > kSmemPerCta) return RS_ERR_INVALID_ARG;
    static SmemOptIn optin[2];
    RS_CUDA_CHECK(optin[0].ensure(pose_ransac_kernel<false>, smem));
    RS_CUDA_CHECK(optin[1].ensure(pose_ransac_kernel<true>, smem));
    if (prm.has_point2d)
        pose_ransac_kernel<true><<<prm.batch, RTHREADS, smem, stream>>>(buf, prm);
    else
        pose_ransac_kernel<false><<<prm.batch, RTHREADS, smem, stream>>>(buf, prm);
    RS_LAUNCH_CHECK();
    return RS_OK;
}

int launch_pose_chain_variance(const PoseBuffers& buf, const PoseLaunch& prm, cudaStream_t stream)
{
    if (prm.n_variance <= 0) return RS_OK;
    const int warps = variance_warps_for(buf.max_matches);
    if (warps == 0) return RS_ERR_INVALID_ARG;
    const size_t smem = variance_smem_bytes(buf.max_matches, warps);
    static SmemOptIn optin[4];
    RS_CUDA_CHECK(optin[0].ensure(pose_variance_kernel<false, false>, smem));
    RS_CUDA_CHECK(optin[1].ensure(pose_variance_kernel<true, false>, smem));
    RS_CUDA_CHECK(optin[2].ensure(pose_variance_kernel<false, true>, smem));
    RS_CUDA_CHECK(optin[3].ensure(pose_variance_kernel<true, true>, smem));
    const dim3 grid((prm.n_variance + warps - 1) / warps, prm.batch);
    const bool batched = prm.mc_batched && warps > 1;
    if (prm.has_point2d) {
        if (batched)
            pose_variance_kernel<true, true><<<grid, warps * 32, smem, stream>>>(buf, prm);
        else
            pose_variance_kernel<true, false><<<grid, warps * 32, smem, stream>>>(buf, prm);
    }
    else {
        if (batched)
            pose_variance_kernel<false, true><<<grid, warps * 32, smem, stream>>>(buf, prm);
        else
            pose_variance_kernel<false, false><<<grid, warps * 32, smem, stream>>>(buf, prm);
    }
    RS_LAUNCH_CHECK();
    pose_covariance_kernel<<<(prm.batch + 3) / 4, 128, 0, stream>>>(buf, prm);
    RS_LAUNCH_CHECK();
    return RS_OK;
}

}  // namespace rs
