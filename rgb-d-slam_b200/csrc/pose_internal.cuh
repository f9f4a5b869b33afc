// Internal declarations shared by the pose kernels and their C-ABI glue (not part of the public header).
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace rs {

struct PoseIntrinsics {
    double fx, fy, cx, cy;
};

// Per-frame RANSAC / LM state kept on the device between the kernels of one solve.
struct PoseFrameState {
    int n;                // features of this frame
    int valid;            // every feature finite, sigma >= 0 (compute_optimized_pose :269-282)
    int residuals;        // 2 per point + 3 per plane
    int stage;            // 0 = failed before/inside RANSAC, 1 = final pose available (variance may run)
    double total_score;   // sum of get_score() in list order
    double final_x[6];    // LM coefficients of the final pose (start point of the Monte-Carlo solves)
    int n_inliers;        // winning inlier set: size, residual count and score in list order (inputs of every Monte-Carlo solve)
    int inlier_residuals;
    double inlier_score;
};

// ---- work state of the fused solve kernel (global memory, reset by the prepare kernel) -------------------------------------
// Hypotheses of one frame may be run by the warps of several CTAs, so the reference's serial best-so-far bookkeeping
// lives in global memory: a ring of finished-but-unapplied hypotheses, folded in strictly in iteration order under a lock.
constexpr int kRansacRing = 64;   // hypotheses that may be in flight or finished-but-unapplied beyond the serial rule's position
struct RansacSlot {               // result of one hypothesis (64 bytes)
    double x[6];
    double score;
    int inliers;
    int ok;
};
struct RansacFrame {
    // read together by every warp that looks for work on the frame (one 16-byte load)
    int can_quit;     // the early stop has fired
    int closed;       // the hypothesis stage is over (early stop, or every iteration applied): the final LM has an owner
    int opened;       // the loop went past the minimum of four hypotheses: more warps pay off, the frame is on the help list
    int applied;      // hypotheses whose bookkeeping has been applied, in iteration order
    int next_iter;    // next hypothesis index to hand out
    int lock;         // guards the in-order bookkeeping
    int joiners;      // other CTAs working on this frame's hypotheses
    int started;      // iterations the serial loop would have started (rs_pose_out.iterations_run)
    double max_score;
    int best_inliers, best_iteration;
    double best_x[6];
    int done[kRansacRing];          // iteration + 1 once the slot's result is complete
    RansacSlot slot[kRansacRing];
};
// ---- one hypothesis per lane (pose_wide.cu): every hypothesis of a chunk leaves a record, the serial rule is folded over them
using HypRecord = RansacSlot;
struct HypFold {       // best-so-far state of a frame's serial RANSAC loop between chunks (pose_optimization.cpp:151-227)
    double max_score;
    double best_x[6];
    int best_inliers, best_iteration, started, can_quit;
};

struct PoseWork {
    int join_ticket;   // frames handed to a first CTA so far
    int frames_done;   // frames whose RANSAC + final LM stage is over (whatever the outcome)
    int n_ready;       // frames published for their Monte-Carlo solves (completion order)
    int mc_head;       // next Monte-Carlo task: frame slot = mc_head / groups, sample group = mc_head % groups
    int n_open;        // frames on the help list
    int mc_done_tasks; // Monte-Carlo tasks finished
    int pad[2];
    unsigned long long dbg[8];   // counters: hypotheses run by a frame's first CTA / by helpers, bookkeeping batches, hypotheses applied, ...
    unsigned long long t_first, t_ransac_end, t_last;   // %globaltimer stamps: first CTA in, last final LM out, last CTA out
};

// Device buffers of one pose context (SoA feature layout: component-major, stride = max_matches).
struct PoseBuffers {
    int max_matches, max_iterations, max_variance;
    const rs_match* matches_aos;   // B x M (staging copy of the caller's array)
    const double* cur_pose;        // B x 7
    const int32_t* n_matches;      // B
    int32_t* type;                 // B x M
    double* obs;                   // B x 4 x M
    double* map;                   // B x 4 x M
    double* sigma;                 // B x 4 x M
    double* aux;                   // B x 4 x M : first observation + inverse depth of the inverse-depth (point2d) features
    PoseFrameState* state;         // B
    rs_pose_out* out;              // B
    uint8_t* mask;                 // B x M
    int16_t* inlier_idx;           // B x M : indices of the winning inlier set, ascending (written by the RANSAC kernel)
    double* poses;                 // B x 7 (the all-gather payload)
    const int32_t* subsets_in;     // B x max_iterations x RS_MAX_SUBSET (RS_RNG_REFERENCE), or null
    int32_t* subsets_used;         // B x max_iterations x RS_MAX_SUBSET
    const double* normals_in;      // B x n_variance x M x 4 (RS_RNG_REFERENCE), or null
    double* v6;                    // B x max_variance x 6 : [pos, eulerAngles(0,1,2)] per Monte-Carlo solve
    int32_t* v_ok;                 // B x max_variance
    // fused solve kernel
    PoseWork* work;                // 1
    RansacFrame* rframe;           // B
    unsigned* ring_mask;           // B x (kRansacRing + 1) x words : inlier masks of the ring slots, then of the best hypothesis
    int32_t* ready;                // B : frame + 1, in the order the frames finished their RANSAC stage
    int32_t* open_list;            // B : frame + 1, frames whose hypothesis loop went past the minimum and takes helpers
    int32_t* mc_done;              // B : Monte-Carlo sample groups finished per frame (the last one reduces the covariance)
    // one hypothesis per lane (allocated when the context's max_iterations makes that path eligible)
    HypRecord* hyp;                // B x max_iterations
    unsigned* hyp_mask;            // B x max_iterations x words
    HypFold* fold;                 // B
    unsigned* fold_mask;           // B x words : inlier mask of the best hypothesis so far
    unsigned long long* frame_times;   // B x 4 %globaltimer stamps: hypotheses started, hypothesis stage closed, final LM done, covariance done
};

struct PoseLaunch {
    int batch;            // frames this launch covers: [frame0, frame0 + batch)
    int frame0 = 0;       // first frame (the host splits a batch into groups whose kernel chains run on separate streams)
    int sub_batches = 1;  // host side only: number of such groups (rs_pose_opts::sub_batches)
    int max_iterations;   // RANSAC hypotheses (119 default)
    int n_variance;       // Monte-Carlo solves
    int lm_max_fev;       // 400
    int rng_mode;
    int has_point2d;      // some frame of the batch carries an RS_FEAT_POINT2D feature: run the kernels that know the type
    int mc_cap = 1 << 30; // Monte-Carlo tasks that may be in flight while frames are still in their RANSAC stage
    int help_min = 64;    // a CTA joins an opened frame when that leaves at least this many iterations per CTA on it (at 119
                          // iterations nobody joins: measured, the hypotheses helpers run there are mostly dropped by the early stop)
    int solver = 0;       // host side: 0 = by shape (the three-launch chain up to 256 hypotheses per frame, the fused kernel beyond),
                          // 1 = chain, 2 = fused kernel (rs_pose_opts.solver)
    int iter0 = 0, iter_count = 0;   // pose_wide.cu: the chunk of RANSAC iterations a launch covers
    int mc_batched = 1;   // chain Monte-Carlo kernel: serial 6x6 parts of a CTA's samples batched on one warp (lm_minimize_cta)
    int final_only = 0;   // chain RANSAC kernel: skip the hypotheses, run the final optimisation from the HypFold state
    int split = 1;        // host side: frame role and Monte-Carlo role in two launches side by side (0: one launch with both)
    int ctas_per_sm = 0;  // resident CTAs per SM the fused kernel is launched with (<= 0: what fits)
    // roles of one launch of the fused kernel (any number of launches may feed on the same work queues)
    int run_ransac = 1;   // its CTAs take frames (hypotheses + final LM)
    int run_mc = 1;       // its CTAs take Monte-Carlo tasks (and reduce covariances)
    int linger = 1;       // CTAs of a launch without the Monte-Carlo role stay to help frames whose hypothesis loop runs long
    int publish_all = 0;  // before the launch, put every frame whose final pose is available on the hand-over list
                          // (RS_RNG_REFERENCE: the host draws the Gaussian stream between the two halves of a solve)
    uint32_t seed;
    PoseIntrinsics K;
};

int pose_max_matches_supported();
// the three-launch chain (pose_chain.cu): RANSAC + final LM per frame, then the Monte-Carlo solves with the covariance folded in
bool pose_chain_supports(int max_matches);
int launch_pose_chain_ransac(const PoseBuffers& buf, const PoseLaunch& prm, cudaStream_t stream);
int launch_pose_chain_variance(const PoseBuffers& buf, const PoseLaunch& prm, cudaStream_t stream);   // longest match list whose staging fits the shared memory of one SM
// hundreds of hypotheses per frame: one hypothesis per lane, in chunks with the serial rule folded in between (pose_wide.cu)
bool pose_wide_supports(int max_matches);
int launch_pose_wide_hypotheses(const PoseBuffers& buf, const PoseLaunch& prm, cudaStream_t stream, int sm_count);
int launch_pose_prepare(const PoseBuffers& buf, const PoseLaunch& prm, cudaStream_t stream);
// RANSAC hypotheses, final LM, Monte-Carlo solves and covariance of a batch in ONE persistent kernel (prm.phase selects halves)
int launch_pose_fused(const PoseBuffers& buf, const PoseLaunch& prm, cudaStream_t stream);
// device-timer view of the last fused launch: ms from the first CTA's start to the last final LM / to the last CTA's exit
int pose_work_times(const PoseBuffers& buf, float* ransac_phase_ms, float* total_ms, cudaStream_t stream);
// fills normals[B][n_variance][M][4] with the Gaussian draws the RS_RNG_DEVICE variance kernel uses
int launch_pose_export_normals(const PoseBuffers& buf, const PoseLaunch& prm, double* normals, cudaStream_t stream);

int require_blackwell(int device);

}  // namespace rs
