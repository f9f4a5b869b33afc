// RANSAC with hundreds of hypotheses per frame (BASELINE configs[4]: 1024): ONE HYPOTHESIS PER LANE.
//
// A hypothesis is a Levenberg-Marquardt solve on a minimal subset - five points, three planes, or a mix (at most RS_MAX_SUBSET
// features) - followed by an inlier test over the frame's whole match list (pose_optimization.cpp:151-227). With a warp per
// hypothesis (pose_chain.cu / pose_solve.cu) five lanes of 32 carry features and one lane runs the 6x6 trust-region algebra:
// 42.6 k warp-instructions per hypothesis at ~10 active lanes, 13.3 ms for 64 x 1024 hypotheses. Here every lane runs a whole
// LM of its own - the Jacobian pass is a serial loop over the subset's features, the normal equations accumulate in the
// lane's registers (no shuffles), the trust-region algebra is the same code on the lane's own 6x6 state - so a warp carries
// 32 hypotheses through the same instruction stream. Iteration counts differ wildly between hypotheses (a subset with an
// outlier wanders until maxfev): the lanes of a warp therefore advance in lock step ONE LM ITERATION at a time, and a lane whose
// solve has ended takes the next hypothesis of its CTA's share at once (a counter in shared memory), so the warp stays full.
// Finished poses queue up per warp; when 32 are waiting (or nothing is left) the warp tests them against the match list
// cooperatively (lanes over features, one ballot per 32 features = one word of the hypothesis's inlier mask), then every
// lane sums the inlier score of one queued hypothesis in list order, as the reference's running double does.
//
// The reference's serial best-so-far / early-stop rule is applied afterwards, in iteration order, by pose_fold_kernel (one
// warp per frame over the records this kernel leaves in global memory); hypotheses are evaluated in chunks of consecutive
// iterations so that frames whose loop has stopped early cost nothing in the following chunks. The final optimisation on
// the winning inlier set and the Monte-Carlo solves are the chain's (pose_chain.cu: pose_ransac_kernel in its final-only
// mode, pose_variance_kernel).
#include "pose_lm_lane.cuh"

namespace rs {

namespace {

constexpr int HTHREADS = 128;   // lanes = concurrent hypotheses per CTA
constexpr int HWARPS = HTHREADS / 32;
constexpr int HQUEUE = 64;      // finished poses a warp may hold before it must test them (>= 32 + the 32 that may end at once)

// The features of a lane's hypothesis: its minimal subset (indices in shared memory) of the frame's match list
struct SubsetSource {
    const short* idx;
    int n;
    const int32_t* type;
    const double* obs;
    const double* map;
    int M;
    __device__ __forceinline__ int count() const { return n; }
    __device__ __forceinline__ int load(const int k, int& ty, double o[4], double m[4]) const
    {
        const int i = idx[k];
        ty = type[i];
#pragma unroll
        for (int c = 0; c < 4; ++c) o[c] = obs[c * M + i], m[c] = map[c * M + i];
        return i;
    }
};

struct WideSmem {
    int32_t* type;
    double* obs;
    double* map;
    short* subset;        // [HTHREADS][RS_MAX_SUBSET]
    double* q_x;          // [HWARPS][HQUEUE][6]
    int* q_it;            // [HWARPS][HQUEUE]
    unsigned* q_mask;     // [HWARPS][32][words]  masks of the batch being tested
    unsigned* plane_bits; // [words] : feature is a plane (score 1/3; points and inverse-depth points score 1/5)
    int* q_n;             // [HWARPS]
    int* next;            // 1 : next hypothesis of this CTA's share
};

__host__ __device__ inline size_t wide_align16(size_t v) { return (v + 15) / 16 * 16; }
__host__ __device__ inline size_t wide_carve(WideSmem* s, unsigned char* base, const int M)
{
    const int words = (M + 31) / 32;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        unsigned char* p = base ? base + o : nullptr;
        o = wide_align16(o + bytes);
        return p;
    };
    unsigned char* obs = take(sizeof(double) * 4 * M);
    unsigned char* map = take(sizeof(double) * 4 * M);
    unsigned char* qx = take(sizeof(double) * HWARPS * HQUEUE * 6);
    unsigned char* type = take(sizeof(int32_t) * M);
    unsigned char* sub = take(sizeof(short) * HTHREADS * RS_MAX_SUBSET);
    unsigned char* qit = take(sizeof(int) * HWARPS * HQUEUE);
    unsigned char* qm = take(sizeof(unsigned) * HWARPS * 32 * words);
    unsigned char* pb = take(sizeof(unsigned) * words);
    unsigned char* qn = take(sizeof(int) * HWARPS);
    unsigned char* nx = take(sizeof(int));
    if (s) {
        s->obs = reinterpret_cast<double*>(obs), s->map = reinterpret_cast<double*>(map);
        s->q_x = reinterpret_cast<double*>(qx), s->type = reinterpret_cast<int32_t*>(type);
        s->subset = reinterpret_cast<short*>(sub), s->q_it = reinterpret_cast<int*>(qit);
        s->q_mask = reinterpret_cast<unsigned*>(qm), s->plane_bits = reinterpret_cast<unsigned*>(pb);
        s->q_n = reinterpret_cast<int*>(qn), s->next = reinterpret_cast<int*>(nx);
    }
    return o;
}

// grid (CTAs per frame, frames); the CTAs of a frame split the chunk's iterations [prm.iter0, prm.iter0 + prm.iter_count)
__global__ void __launch_bounds__(HTHREADS, 2) pose_hypotheses_kernel(const PoseBuffers buf, const PoseLaunch prm)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = prm.frame0 + blockIdx.y;
    const int M = buf.max_matches;
    const int words = (M + 31) / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PoseFrameState st = buf.state[b];
    if (!st.valid || st.total_score < 1.0) return;
    if (buf.fold[b].can_quit) return;   // the serial loop of this frame stopped in an earlier chunk
    const int n = st.n;
    // this CTA's share of the chunk
    const int per_cta = (prm.iter_count + gridDim.x - 1) / gridDim.x;
    const int first = prm.iter0 + blockIdx.x * per_cta;
    const int last = min(min(first + per_cta, prm.iter0 + prm.iter_count), prm.max_iterations);
    if (first >= last) return;

    WideSmem sm;
    wide_carve(&sm, smem_raw, M);
    for (int i = threadIdx.x; i < M; i += HTHREADS) {
        sm.type[i] = buf.type[size_t(b) * M + i];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            sm.obs[c * M + i] = buf.obs[(size_t(b) * 4 + c) * M + i];
            sm.map[c * M + i] = buf.map[(size_t(b) * 4 + c) * M + i];
        }
    }
    for (int w = threadIdx.x; w < words; w += HTHREADS) {
        unsigned bits = 0u;
        for (int k = 0; k < 32; ++k) {
            const int i = w * 32 + k;
            if (i < n && buf.type[size_t(b) * M + i] == RS_FEAT_PLANE) bits |= 1u << k;
        }
        sm.plane_bits[w] = bits;
    }
    if (threadIdx.x < HWARPS) sm.q_n[threadIdx.x] = 0;
    if (threadIdx.x == 0) *sm.next = first;
    __syncthreads();

    double x0[6];
    coefficients_from_pose(buf.cur_pose + b * 7, x0);
    short* my_subset = sm.subset + threadIdx.x * RS_MAX_SUBSET;
    SubsetSource sub;
    sub.idx = my_subset, sub.n = 0, sub.type = sm.type, sub.obs = sm.obs, sub.map = sm.map, sub.M = M;
    DRLocal dR;
    double* q_x = sm.q_x + size_t(warp) * HQUEUE * 6;
    int* q_it = sm.q_it + warp * HQUEUE;
    unsigned* q_mask = sm.q_mask + size_t(warp) * 32 * words;
    int* q_n = sm.q_n + warp;
    HypRecord* rec = buf.hyp + size_t(b) * buf.max_iterations;
    unsigned* rec_mask = buf.hyp_mask + size_t(b) * buf.max_iterations * words;

    // inlier tests of the first `cnt` queued poses (cnt <= 32), then the queue is shifted down
    auto test_queued = [&](const int cnt) {
        for (int e = 0; e < cnt; ++e) {
            double xe[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) xe[j] = q_x[e * 6 + j];
            Xform T;
            make_xform(xe, T);   // every lane builds the same transform: cheaper than a round trip through shared memory
            for (int w = 0; w < words; ++w) {
                const int i = w * 32 + lane;
                bool in = false;
                if (i < n) {
                    double o[4], mm[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) o[c] = sm.obs[c * M + i], mm[c] = sm.map[c * M + i];
                    in = feature_is_inlier<false>(sm.type[i], o, mm, T, prm.K, nullptr, M, i);
                }
                const unsigned bits = __ballot_sync(FULL, in);
                if (lane == 0) q_mask[e * words + w] = bits;
            }
        }
        __syncwarp();
        if (lane < cnt) {
            // get_features_inliers_outliers' running score (pose_optimization.cpp:33-72), in list order
            const int it = q_it[lane];
            double score = 0.0;
            int nIn = 0;
            for (int w = 0; w < words; ++w) {
                unsigned bits = q_mask[lane * words + w];
                const unsigned planes = sm.plane_bits[w];
                rec_mask[size_t(it) * words + w] = bits;
                nIn += __popc(bits);
                while (bits) {
                    const unsigned low = bits & (0u - bits);
                    bits ^= low;
                    score += (planes & low) ? kPlaneScore : kPointScore;
                }
            }
            HypRecord r;
#pragma unroll
            for (int j = 0; j < 6; ++j) r.x[j] = q_x[lane * 6 + j];
            r.score = score, r.inliers = nIn, r.ok = 1;
            rec[it] = r;
        }
        __syncwarp();
        // shift the rest of the queue down
        const int total = *q_n;
        for (int e = cnt + lane; e < total; e += 32) {
            double xe[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) xe[j] = q_x[e * 6 + j];
            const int it = q_it[e];   // (at most 31 entries remain and they move below index 32: sources and targets are disjoint)
#pragma unroll
            for (int j = 0; j < 6; ++j) q_x[(e - cnt) * 6 + j] = xe[j];
            q_it[e - cnt] = it;
        }
        __syncwarp();
        if (lane == 0) *q_n = total - cnt;
        __syncwarp();
    };

    LaneLM S;
    S.status = 0;
    bool active = false, exhausted = false;
    int it = -1;
    for (;;) {
        // ---- lanes without a running solve take the next hypothesis of the share ----
        if (!active && !exhausted) {
            it = atomicAdd(sm.next, 1);
            if (it >= last)
                exhausted = true;
            else {
                // random subset: ransac::get_random_subset_with_score (ransac.hpp:77-103)
                int cnt = 0, m = 0;
                double cumulated = 0.0;
                int32_t* used = buf.subsets_used + (size_t(b) * buf.max_iterations + it) * RS_MAX_SUBSET;
                if (buf.subsets_in) {
                    const int32_t* in = buf.subsets_in + (size_t(b) * buf.max_iterations + it) * RS_MAX_SUBSET;
                    for (int k = 0; k < RS_MAX_SUBSET; ++k) {
                        const int idx = in[k];
                        if (idx >= 0 && idx < n) my_subset[cnt++] = short(idx);
                    }
                    for (int k = cnt - 1; k >= 0; --k) cumulated += score_of(sm.type[my_subset[k]]);
                }
                else {
                    // distinct uniform picks until the cumulated score reaches 1, each pick PREPENDED (same draws as the
                    // warp-per-hypothesis kernels: keyed by frame and iteration)
                    const uint64_t key = rng_key(prm.seed, 1u, uint32_t(b)) ^ mix64(uint64_t(uint32_t(it)) << 20);
                    uint64_t ctr = 0;
                    while (cnt < RS_MAX_SUBSET && cnt < n && cumulated < 1.0) {
                        const int idx = int(mix64(key + ctr++) % uint64_t(n));
                        bool dup = false;
                        for (int k = 0; k < cnt; ++k) dup = dup || (my_subset[k] == idx);
                        if (dup) continue;
                        my_subset[cnt++] = short(idx);
                        cumulated += score_of(sm.type[idx]);
                    }
                    for (int k = 0; k < cnt / 2; ++k) {   // prepended order = reverse pick order
                        const short t = my_subset[k];
                        my_subset[k] = my_subset[cnt - 1 - k], my_subset[cnt - 1 - k] = t;
                    }
                }
                for (int k = 0; k < RS_MAX_SUBSET; ++k) used[k] = k < cnt ? int(my_subset[k]) : -1;
                for (int k = 0; k < cnt; ++k) m += parts_of(sm.type[my_subset[k]]);
                sub.n = cnt;
                bool finite = true;
#pragma unroll
                for (int j = 0; j < 6; ++j) finite = finite && isfinite(x0[j]);
                // compute_optimized_global_pose's guards (pose_optimization.cpp:302-321)
                if (cumulated >= 1.0 && finite && m > 1) {
                    lane_lm_begin(S, sub, prm.K, x0, m, prm.lm_max_fev);
                    active = true;   // a solve that stopped inside begin is retired below
                }
                else {
                    HypRecord r = {};
                    rec[it] = r;   // ok = 0: skipped by the serial rule
                }
            }
        }
        if (!__any_sync(FULL, active)) {
            if (__all_sync(FULL, exhausted)) break;
            continue;
        }
        // ---- one LM iteration for every lane that has a solve running ----
        if (active && S.status == kRunning) lane_lm_step(S, sub, dR, prm.K, prm.lm_max_fev);
        // ---- retire finished solves ----
        if (active && S.status != kRunning) {
            active = false;
            bool ok = S.status > 0;
#pragma unroll
            for (int j = 0; j < 6; ++j) ok = ok && isfinite(S.x[j]);
            ok = ok && isfinite(S.x[3] * S.x[3] + S.x[4] * S.x[4] + S.x[5] * S.x[5]);
            if (ok) {
                const int slot = atomicAdd(q_n, 1);
#pragma unroll
                for (int j = 0; j < 6; ++j) q_x[slot * 6 + j] = S.x[j];
                q_it[slot] = it;
            }
            else {
                HypRecord r = {};
                rec[it] = r;
            }
        }
        __syncwarp();
        if (*q_n >= 32) test_queued(32);
    }
    __syncwarp();
    while (*q_n > 0) test_queued(min(*q_n, 32));
}

// The reference's serial bookkeeping (pose_optimization.cpp:151-227) over the records of one chunk, in iteration order:
// one warp per frame, lanes fetch 32 records at a time, lane 0 applies the rule.
__global__ void __launch_bounds__(128) pose_fold_kernel(const PoseBuffers buf, const PoseLaunch prm)
{
    const int b = prm.frame0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= prm.frame0 + prm.batch) return;
    const PoseFrameState st = buf.state[b];
    if (!st.valid || st.total_score < 1.0) return;
    HypFold* f = buf.fold + b;
    if (f->can_quit) return;
    const int M = buf.max_matches, words = (M + 31) / 32;
    const unsigned inliersToStop = unsigned(ceil(double(st.n) * kEarlyStopProportion));
    const HypRecord* rec = buf.hyp + size_t(b) * buf.max_iterations;
    double max_score = f->max_score;
    const int prev_best = f->best_iteration;
    int best_inliers = f->best_inliers, best_iteration = prev_best, started = f->started, can_quit = 0;
    const int end = min(prm.iter0 + prm.iter_count, prm.max_iterations);
    for (int base = prm.iter0; base < end && !can_quit; base += 32) {
        const int i = base + lane;
        double score = 0.0;
        int inl = 0, ok = 0;
        if (i < end) {
            const HypRecord r = rec[i];
            score = r.score, inl = r.inliers, ok = r.ok;
        }
        const int cnt = min(32, end - base);
        for (int k = 0; k < cnt; ++k) {
            const double hs = __shfl_sync(FULL, score, k);
            const int hi = __shfl_sync(FULL, inl, k);
            const int hok = __shfl_sync(FULL, ok, k);
            ++started;
            if (hok && hs >= 1.0) {
                const bool canOverload = (hs > max_score) || (fabs(hs - max_score) <= 0.1 && best_inliers < hi);
                if (canOverload) max_score = hs, best_inliers = hi, best_iteration = base + k;
                if (base + k >= 3 && unsigned(best_inliers) > inliersToStop) {
                    can_quit = 1;
                    break;
                }
            }
        }
    }
    __syncwarp();
    if (best_iteration != prev_best) {
        const unsigned* src = buf.hyp_mask + (size_t(b) * buf.max_iterations + best_iteration) * words;
        unsigned* dst = buf.fold_mask + size_t(b) * words;
        for (int w = lane; w < words; w += 32) dst[w] = src[w];
        if (lane < 6) f->best_x[lane] = rec[best_iteration].x[lane];
    }
    __syncwarp();
    if (lane == 0) {
        f->max_score = max_score, f->best_inliers = best_inliers, f->best_iteration = best_iteration;
        f->started = started, f->can_quit = can_quit;
    }
}

__global__ void pose_fold_reset_kernel(const PoseBuffers buf, const PoseLaunch prm)
{
    const int b = prm.frame0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= prm.frame0 + prm.batch) return;
    HypFold* f = buf.fold + b;
    f->max_score = 1.0;
    f->best_inliers = 0, f->best_iteration = -1, f->started = 0, f->can_quit = 0;
    double x0[6];
    coefficients_from_pose(buf.cur_pose + b * 7, x0);
    for (int j = 0; j < 6; ++j) f->best_x[j] = x0[j];
    const int words = (buf.max_matches + 31) / 32;
    for (int w = 0; w < words; ++w) buf.fold_mask[size_t(b) * words + w] = 0u;
}

constexpr size_t kWideSmemPerCta = 232448;

}  // namespace

bool pose_wide_supports(const int max_matches) { return 2 * (wide_carve(nullptr, nullptr, max_matches) + 1024) <= 233472; }

// Hypotheses [0, max_iterations) of every frame in chunks, each followed by the serial fold; leaves HypFold / fold_mask for the
// final-only mode of the chain's RANSAC kernel.
int launch_pose_wide_hypotheses(const PoseBuffers& buf, const PoseLaunch& prm, cudaStream_t stream, const int sm_count)
{
    if (!buf.hyp || !buf.fold) return RS_ERR_INVALID_ARG;
    const size_t smem = wide_carve(nullptr, nullptr, buf.max_matches);
    if (smem > kWideSmemPerCta) return RS_ERR_INVALID_ARG;
    static SmemOptIn optin;
    RS_CUDA_CHECK(optin.ensure(pose_hypotheses_kernel, smem));
    pose_fold_reset_kernel<<<(prm.batch + 127) / 128, 128, 0, stream>>>(buf, prm);
    RS_LAUNCH_CHECK();
    // chunk: enough hypotheses for every lane the GPU holds to take four in a row (refills are what keeps the warps full),
    // a multiple of 32 per frame; a frame's chunk is split over CTAs of 128 lanes so that each lane expects ~4 hypotheses
    const long lanes = long(sm_count) * 2 * HTHREADS;
    int chunk = int((4 * lanes + prm.batch - 1) / prm.batch);
    chunk = (chunk + 31) / 32 * 32;
    chunk = std::max(64, std::min(chunk, prm.max_iterations));
    for (int it0 = 0; it0 < prm.max_iterations; it0 += chunk) {
        PoseLaunch lp = prm;
        lp.iter0 = it0, lp.iter_count = std::min(chunk, prm.max_iterations - it0);
        // CTAs per frame: fill every resident slot of the GPU, but leave each lane at least two hypotheses to take in turn
        const int want = (2 * sm_count + prm.batch - 1) / prm.batch;
        const int cap = (lp.iter_count + 2 * HTHREADS - 1) / (2 * HTHREADS);
        // (measured at 64 frames x 1024 hypotheses: 4 CTAs per frame 1.94 ms, 2 / 3 / 8 CTAs 2.13 / 2.08 / 2.19 ms; three CTAs per SM
        // at 168 registers 2.64 ms - the spills cost more than the extra warps hide. Also tried: parking a solve that is still
        // running after 8 .. 28 LM iterations (its persistent MINPACK state to global memory) and finishing the parked ones in a
        // second pass, one per lane - 2.01 .. 2.25 ms against 1.91 .. 1.97: what a lock-step tick costs is the divergence INSIDE
        // it (lmpar's Newton loop runs at 6 of 32 lanes, plane features at 2), not the tail of long solves)
        const int ctas = std::max(1, std::min(want, cap));
        pose_hypotheses_kernel<<<dim3(ctas, prm.batch), HTHREADS, smem, stream>>>(buf, lp);
        RS_LAUNCH_CHECK();
        pose_fold_kernel<<<(prm.batch + 3) / 4, 128, 0, stream>>>(buf, lp);
        RS_LAUNCH_CHECK();
    }
    return RS_OK;
}

}  // namespace rs
