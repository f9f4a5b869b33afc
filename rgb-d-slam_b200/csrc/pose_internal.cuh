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
    uint32_t seed;
    PoseIntrinsics K;
};

int pose_max_matches_supported();   // longest match list whose staging fits the shared memory of one SM
int launch_pose_prepare(const PoseBuffers& buf, const PoseLaunch& prm, cudaStream_t stream);
int launch_pose_ransac(const PoseBuffers& buf, const PoseLaunch& prm, cudaStream_t stream);
int launch_pose_variance(const PoseBuffers& buf, const PoseLaunch& prm, cudaStream_t stream);
int launch_pose_covariance(const PoseBuffers& buf, const PoseLaunch& prm, cudaStream_t stream);
// fills normals[B][n_variance][M][4] with the Gaussian draws the RS_RNG_DEVICE variance kernel uses
int launch_pose_export_normals(const PoseBuffers& buf, const PoseLaunch& prm, double* normals, cudaStream_t stream);

int require_blackwell(int device);

}  // namespace rs
