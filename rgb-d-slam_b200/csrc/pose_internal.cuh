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
struct RansacFrame {
    double best_x[6];
    double max_score;
    int best_inliers, best_iteration, can_quit, started;
    int next_iter;    // next hypothesis index to hand out
    int applied;      // hypotheses whose bookkeeping has been applied, in iteration order
    int lock;         // guards the in-order bookkeeping
    int closed;       // the hypothesis stage is over (early stop, or every iteration applied): the final LM has an owner
    int joiners;      // CTAs working on this frame's hypotheses
    int opened;       // the frame has been put on the help list (the minimum of four hypotheses did not stop the loop)
    int pad[2];
    int done[kRansacRing];         // iteration + 1 once the slot's result is complete
    int hyp_ok[kRansacRing];
    int hyp_inliers[kRansacRing];
    double hyp_score[kRansacRing];
    double hyp_x[kRansacRing][6];
};
struct PoseWork {
    int join_ticket;   // frames handed to a first CTA so far
    int frames_done;   // frames whose RANSAC + final LM stage is over (whatever the outcome)
    int n_ready;       // frames published for their Monte-Carlo solves (completion order)
    int mc_head;       // next Monte-Carlo task: frame slot = mc_head / groups, sample group = mc_head % groups
    int n_open;        // frames on the help list
    int pad[3];
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
    int phase = 0;        // 0 = RANSAC + Monte-Carlo in one launch; 1 = RANSAC + final LM only; 2 = Monte-Carlo + covariance only
                          // (RS_RNG_REFERENCE: the host draws the Gaussian stream between the two halves)
    uint32_t seed;
    PoseIntrinsics K;
};

int pose_max_matches_supported();   // longest match list whose staging fits the shared memory of one SM
int launch_pose_prepare(const PoseBuffers& buf, const PoseLaunch& prm, cudaStream_t stream);
// RANSAC hypotheses, final LM, Monte-Carlo solves and covariance of a batch in ONE persistent kernel (prm.phase selects halves)
int launch_pose_fused(const PoseBuffers& buf, const PoseLaunch& prm, cudaStream_t stream);
// device-timer view of the last fused launch: ms from the first CTA's start to the last final LM / to the last CTA's exit
int pose_work_times(const PoseBuffers& buf, float* ransac_phase_ms, float* total_ms, cudaStream_t stream);
// fills normals[B][n_variance][M][4] with the Gaussian draws the RS_RNG_DEVICE variance kernel uses
int launch_pose_export_normals(const PoseBuffers& buf, const PoseLaunch& prm, double* normals, cudaStream_t stream);

int require_blackwell(int device);

}  // namespace rs
