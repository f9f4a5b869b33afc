/*
 * rgbdslam_b200.h — C-ABI of the B200-native hot path of BaptisteHudyma/RGB-D-SLAM.
 *
 * The reference has no FFI layer: its seam is three C++ member functions
 * (SURVEY.md §8b). Each entry point below names the reference interface it replaces.
 * Everything is plain C: POD structs, caller-allocated buffers, int status (0 = ok).
 * No torch / Eigen / OpenCV types cross this boundary.
 *
 *   Depth_Map_Transformation::get_organized_cloud_array   src/features/primitives/depth_map_transformation.hpp:38-39
 *   Primitive_Detection::Primitive_Detection / find_primitives   src/features/primitives/primitive_detection.hpp:33,42-45
 *   Pose_Optimization::compute_optimized_pose               src/pose_optimization/pose_optimization.hpp:27-30
 *
 * Units follow the reference: depth and distances in millimetres, angles in radians,
 * pose = position (mm) + unit quaternion (w,x,y,z) of the camera in the world frame.
 */
#ifndef RGBDSLAM_B200_H
#define RGBDSLAM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------------ */
#define RS_OK 0
#define RS_ERR_INVALID_ARG 1   /* null pointer, bad size, batch > max_batch ...            */
#define RS_ERR_CUDA 2          /* a CUDA runtime/driver call failed (see rs_last_error())  */
#define RS_ERR_NO_DEVICE 3     /* no sm_100 device: the library never falls back to a CPU  */
#define RS_ERR_CAPACITY 4      /* an output capacity (planes, cylinders, boundary) was hit */

/* ---- capacities (compile-time, shared by kernels, oracle and callers) ---------------- */
#define RS_MAX_PLANES 128      /* plane segments per frame before merging                  */
#define RS_MAX_CYL_REGIONS 32  /* cylinder-branch regions per frame                        */
#define RS_MAX_CYL_SEGS 8      /* sequential-RANSAC sub-segments per region                */
#define RS_CYL_RANSAC_ITERS 43 /* uint(logf(0.2)/logf(1-0.33^3)), cylinder_segment.cpp:132 */

/* ---- CAPE (planes + cylinders) --------------------------------------------------------
 * Per-cell record written by the plane-fit kernel: 160 bytes, the "per-cell record" of
 * SURVEY.md §8d's algorithmic-bytes figure (4·W·H + 160·Ncells per frame).
 * Mirrors Plane_Segment (plane_segment.hpp:118-140) after init_plane_segment. */
typedef struct rs_cell_out {
    int32_t count;      /* _pointCount: pixels with z > 0 (0 if the cell failed the early tests) */
    int32_t planar;     /* _isPlanar                                                            */
    double S[9];        /* Sx Sy Sz Sxs Sys Szs Sxy Syz Szx (FP64 sums of FP32 values/products)   */
    double centroid[3];
    double normal[3];
    double d;
    double mse;         /* DBL_MAX when no fit was made                                           */
    double score;
    float tol;          /* _cellDistanceTols[cell], primitive_detection.cpp:201-220               */
    int32_t hist_bin;   /* Histogram bin of the normal (init_histogram, primitive_detection.cpp:239-265), -1 if not planar */
} rs_cell_out;

/* One plane segment (entry k of _planeSegments, primitive_detection.hpp:216) after merge_planes. */
typedef struct rs_plane_out {
    int32_t merge_label;  /* planeMergeLabels[k] (root id, 0-based)                               */
    int32_t planar;       /* is_planar() after merging                                            */
    int32_t is_final;     /* merge_label == k and planar: the plane survives to add_planes_to_primitives */
    int32_t count;        /* point count                                                          */
    double S[9];
    double centroid[3];
    double normal[3];
    double d;
    double mse;
    double score;
    int32_t n_boundary;      /* boundary centre points kept (compute_plane_segment_boundary); <3 = dropped by the reference */
    int32_t boundary_offset; /* first index into the frame's boundary arrays                      */
} rs_plane_out;

/* One cylinder-branch region (entry of _cylinderSegments, cylinder_segment.hpp). */
typedef struct rs_cyl_out {
    int32_t n_cells;     /* activated cells handed to the Cylinder_Segment ctor                  */
    int32_t n_segments;  /* _segmentCount (0 when the normal-PCA score is < 75)                   */
    double pca_score;    /* lambda2 / lambda0 of the +-normals covariance                         */
    double axis[3];
    double radius[RS_MAX_CYL_SEGS];
    double center[RS_MAX_CYL_SEGS][3];
    double mse[RS_MAX_CYL_SEGS];
    double plane_mse[RS_MAX_CYL_SEGS];   /* MSE of the plane refit over the segment's inliers      */
    int32_t n_inliers[RS_MAX_CYL_SEGS];
    int32_t assigned[RS_MAX_CYL_SEGS];   /* >0: cylinder id in cyl_labels; <0: -(plane id) it became; 0: nothing */
    int32_t kept[RS_MAX_CYL_SEGS];       /* cylinder id survives the opening test of add_cylinders_to_primitives */
} rs_cyl_out;

typedef struct rs_cape_frame_info {
    int32_t status;        /* RS_OK or RS_ERR_CAPACITY                                            */
    int32_t n_planar_cells;
    int32_t n_seeds;       /* iterations of the seed loop (grow_planes_and_cylinders)             */
    int32_t n_planes;      /* plane segments before merging                                       */
    int32_t n_final_planes;
    int32_t n_cyl_regions;
    int32_t n_cylinders;   /* cylinder2regionMap.size()                                           */
    int32_t n_boundary;    /* boundary points written for this frame                              */
} rs_cape_frame_info;

typedef struct rs_cape_ctx rs_cape_ctx;

/* Replaces Primitive_Detection::Primitive_Detection(w,h) + Depth_Map_Transformation ctor/init_matrices
 * (depth_map_transformation.cpp:147-173). cell_px = parameters::detection::depthMapPatchSize_px (20).
 * Intrinsics are the reference's camera-1 fx,fy,cx,cy (parameters.cpp:59-74). Returns NULL on error. */
rs_cape_ctx* rs_cape_create(int width, int height, int cell_px, double fx, double fy, double cx, double cy,
                            int max_batch, int device);
void rs_cape_destroy(rs_cape_ctx* ctx);

int rs_cape_cells_per_frame(const rs_cape_ctx* ctx);    /* hc*vc                                   */
int rs_cape_max_boundary(const rs_cape_ctx* ctx);       /* boundary-point capacity per frame        */

/* Outputs of one batched CAPE run; any pointer may be NULL to skip that copy.
 * Leading dimension is the batch; sizes per frame are in brackets. */
typedef struct rs_cape_outputs {
    rs_cell_out* cells;          /* [Ncells]                                                       */
    int32_t* plane_grid;         /* [Ncells] _gridPlaneSegmentMap: 0 none, k = k-th plane segment   */
    int32_t* plane_labels;       /* [Ncells] merged labels: 0 none, root+1 for final planes         */
    int32_t* cyl_labels;         /* [Ncells] _gridCylinderSegMap                                    */
    int32_t* cyl_region_seg;     /* [Ncells] 0 none, 1 + region*RS_MAX_CYL_SEGS + segment           */
    rs_plane_out* planes;        /* [RS_MAX_PLANES]                                                */
    rs_cyl_out* cyls;            /* [RS_MAX_CYL_REGIONS]                                           */
    double* boundary_xyz;        /* [max_boundary][3] camera-frame centre points                   */
    rs_cape_frame_info* info;    /* [1]                                                            */
} rs_cape_outputs;

/* Replaces get_organized_cloud_array + find_primitives for a batch of frames
 * (rgbd_slam.cpp:110-112,295). depth = B x H x W float32, row-major, mm, <=0 invalid, HOST memory.
 * seed = utils::Random::_seed of the per-frame thread-local engine (random.hpp:17-65; 0 under
 * MAKE_DETERMINISTIC): the cylinder RANSAC of EVERY frame restarts at mt19937(seed), as the
 * reference does because find_primitives runs on a fresh std::async thread per frame. */
int rs_cape_run(rs_cape_ctx* ctx, const float* depth_host, int batch, uint32_t seed, const rs_cape_outputs* out_host);

/* Same from the raw sensor image: CV_16U depth converted as cv::Mat::convertTo(CV_32F, alpha) does (float(src) *
 * float(alpha); examples/main_TUM.cpp:242 uses alpha = 1/5, main_CAPE.cpp:59 alpha = 1). The conversion runs on the
 * device, so the host-to-device copy is half the size of the float image. */
int rs_cape_run_u16(rs_cape_ctx* ctx, const uint16_t* depth_host, double alpha, int batch, uint32_t seed,
                    const rs_cape_outputs* out_host);

/* Same, with depth and outputs already resident in device memory (used for the HBM-resident
 * throughput number; `stream` is a cudaStream_t passed as void*). Asynchronous. */
int rs_cape_run_device(rs_cape_ctx* ctx, const float* depth_dev, int batch, uint32_t seed,
                       const rs_cape_outputs* out_dev, void* stream);

/* Only the per-cell plane fit (K1): get_organized_cloud_array + init_planar_cell_fitting.
 * cells_dev = B x Ncells records in device memory. Asynchronous. */
int rs_cape_cell_fit_device(rs_cape_ctx* ctx, const float* depth_dev, int batch, rs_cell_out* cells_dev, void* stream);

/* The other half of rs_cape_run_device: histogram seeding, region growing, cylinder RANSAC, merging and boundary points
 * (Primitive_Detection::find_primitives after init_planar_cell_fitting, primitive_detection.cpp:131-165) on the records a
 * preceding rs_cape_cell_fit_device call left in out_dev->cells (same depth image, same batch). Lets a caller place other
 * work between the two halves: bench.py holds the latency-bound segmentation back until the pose context's RANSAC kernel has
 * finished (rs_pose_stream_wait_ransac), so that the two do not compete for shared memory. Asynchronous. */
int rs_cape_segment_device(rs_cape_ctx* ctx, const float* depth_dev, int batch, uint32_t seed,
                           const rs_cape_outputs* out_dev, void* stream);

/* Depth_Map_Transformation::rectify_depth (src/features/primitives/depth_map_transformation.cpp:23-87, called through
 * RGBD_SLAM::rectify_depth, rgbd_slam.cpp:85-97, by examples/main_CAPE.cpp:186): the depth camera's image re-projected
 * into the colour camera's image; the last source pixel in raster order wins a destination pixel, untouched pixels are 0.
 * cam2_to_cam1 = Parameters::get_camera_2_to_camera_1_transformation(), row-major 4x4 (12 values are read).
 * With enable != 0 every rs_cape_run* call rectifies its input on the device before the plane fit (the boundary step
 * then reads the rectified image, as find_primitives does in the reference); rs_cape_rectify* expose the step alone. */
int rs_cape_set_rectification(rs_cape_ctx* ctx, const double* cam2_to_cam1, int enable);
int rs_cape_rectify(rs_cape_ctx* ctx, const float* depth_host, int batch, float* rectified_host);
int rs_cape_rectify_device(rs_cape_ctx* ctx, const float* depth_dev, int batch, float* rectified_dev, void* stream);

/* Makes `stream` wait until the streaming part of the most recent plane fit (K1a, the HBM-bound kernel) launched through
 * this context has finished.
 * The reference runs find_primitives on its own std::async thread beside the rest of the frame (rgbd_slam.cpp:288-300);
 * the device-side equivalent is a second stream: the HBM-bound K1 gets the GPU to itself and the latency-bound kernels
 * (cell-graph segmentation here, RANSAC in the pose context) then share the SMs. */
int rs_cape_stream_wait_fit(rs_cape_ctx* ctx, void* stream);

/* Device scratch owned by the context, for callers that keep data resident (bench, multi-frame pipelines). */
float* rs_cape_device_depth(rs_cape_ctx* ctx);                  /* max_batch x H x W                 */
const rs_cape_outputs* rs_cape_device_outputs(rs_cape_ctx* ctx); /* device pointers, max_batch frames */

/* Per-kernel device timing (replaces the cv::getTickCount accumulators of Primitive_Detection,
 * primitive_detection.hpp:233-239 / .cpp:124-165). n_slots > 0 allocates n_slots rings of CUDA events that
 * every following run records on its launching stream around each kernel (run i uses slot i % n_slots);
 * n_slots = 0 turns it off. rs_cape_kernel_ms waits for that slot's events and returns
 * ms[0] = plane-fit kernel (K1), ms[1] = segmentation kernel (K2-K4; 0 for cell-fit-only runs). */
int rs_cape_set_timing(rs_cape_ctx* ctx, int n_slots);
int rs_cape_kernel_ms(rs_cape_ctx* ctx, int slot, float ms[2]);

/* ---- pose solve ------------------------------------------------------------------------- */
#define RS_FEAT_POINT 0    /* PointOptimizationFeature  (map_point.cpp:16-65)    2 residuals, score 1/5 */
#define RS_FEAT_PLANE 1    /* PlaneOptimizationFeature  (map_primitive.cpp:15-85) 3 residuals, score 1/3 */
#define RS_FEAT_POINT2D 2  /* Point2dOptimizationFeature (map_point2d.cpp:15-81): inverse-depth map point, the "line" residual:
                              signed distance of the matched pixel to the screen line through the projections of the point's
                              furthest / closest depth estimates; 2 residuals, score 1/5, weight 0.3 / 2 */

/* One matched feature (an IOptimizationFeature flattened; matches_containers.hpp:122-180). */
typedef struct rs_match {
    int32_t type;
    int32_t reserved;
    double obs[4];    /* point: (u, v, -, -) screen px.  plane: camera-frame (nx, ny, nz, d).
                         point2d: (u, v, theta, phi): matched pixel + the map point's bearing angles (rad)          */
    double map[4];    /* point: world (X, Y, Z, -) mm.   plane: world-frame (nx, ny, nz, d).
                         point2d: (first observation X, Y, Z in mm, inverse depth in 1/mm)                           */
    double sigma[4];  /* standard deviation of the map side (Monte-Carlo variation).
                         point2d: (sigma inverse depth, sigma theta, sigma phi, -) = _mapPointStandardDev(3..5); the
                         first observation is not varied by the reference (map_point2d.cpp:50-52)                    */
} rs_match;

#define RS_RNG_REFERENCE 0 /* host std::mt19937 + libstdc++ shuffle/normal_distribution, one stream, as random.hpp/ransac.hpp */
#define RS_RNG_DEVICE 1    /* counter-based on-device generator (throughput mode)                    */

typedef struct rs_pose_opts {
    int32_t max_iterations;     /* RANSAC hypotheses; <=0 -> 119 (pose_optimization.cpp:129-132)     */
    int32_t n_variance;         /* Monte-Carlo LM solves for the covariance; <0 -> 100, 0 -> skip    */
    int32_t rng_mode;           /* RS_RNG_REFERENCE | RS_RNG_DEVICE                                   */
    uint32_t seed;
    double fx, fy, cx, cy;      /* camera-1 intrinsics; all 0 -> reference defaults 550,550,320,240  */
    int32_t lm_max_fev;         /* <=0 -> 400 (Eigen LevenbergMarquardt default)                     */
    int32_t sub_batches;        /* accepted for compatibility, no effect: the solve kernel hands every frame over from its final
                                   LM to its Monte-Carlo solves on its own (it used to split the batch into groups of frames
                                   whose kernel chains ran on separate streams)                                           */
    int32_t worker_ctas_per_sm; /* resident CTAs per SM the solve kernel is launched with; <= 0 -> as many as fit (4 at up to 400
                                   matches). A caller that runs another kernel beside the solve (bench.py: the cell-graph
                                   segmentation, 111 KB of shared memory per CTA) starts with fewer and adds the rest with
                                   rs_pose_add_workers once that kernel has drained. < 0 (rs_pose_solve_device only): launch
                                   the frame role alone (hypotheses + final LM; a CTA per frame that leaves when its frame is
                                   done) - the Monte-Carlo solves then run on the CTAs the caller adds with rs_pose_add_workers,
                                   and the solve is complete when those have finished                                      */
    int32_t solver;             /* 0 = chosen by shape, 1 = the three-launch chain (per-frame RANSAC kernel with its state in shared
                                   memory, then the Monte-Carlo kernel: small or short-lived CTAs that leave room for kernels the
                                   caller runs beside the solve; the default up to 256 hypotheses per frame), 2 = the fused
                                   persistent kernel (per-frame hand-over from the final LM to the Monte-Carlo solves, any number
                                   of CTAs per frame: fastest for a solve running alone - 1.2 against 1.5 ms per 256 frames - and
                                   for hundreds of hypotheses per frame), 3 = one hypothesis per LANE (every lane runs the whole
                                   LM of a minimal subset in its registers, a warp carries 32 hypotheses and refills lanes as
                                   solves end, the serial best-so-far / early-stop rule is folded over the results in iteration
                                   order; then the chain's final optimisation and Monte-Carlo kernels. The default beyond 256
                                   hypotheses per frame - 14.6 -> 2.0 ms for 64 x 1024 - for batches without RS_FEAT_POINT2D
                                   features, in a context created with max_iterations > 256)                               */
} rs_pose_opts;

typedef struct rs_pose_out {
    int32_t status;            /* 1 = pose and covariance valid (compute_optimized_pose returned true),
                                  0 = RANSAC failed, -1 = final LM failed, -2 = covariance failed     */
    int32_t n_inliers;
    int32_t iterations_run;    /* RANSAC iterations started before the early stop                   */
    int32_t best_iteration;
    int32_t n_variance_ok;     /* successful Monte-Carlo solves                                      */
    int32_t reserved;
    double score;              /* inlier score of the winning hypothesis                             */
    double pose[7];            /* x y z qw qx qy qz                                                  */
    double cov[36];            /* row-major 6x6 of [pos, eulerAngles(0,1,2)] + 1e-3 I                */
} rs_pose_out;

typedef struct rs_pose_ctx rs_pose_ctx;

/* Capacities of a solver context. max_matches: longest match list of one frame, 1..3175 (the RANSAC and Monte-Carlo kernels
 * stage one frame's matches in the 227 KB of shared memory of an SM; up to 370 matches eight Monte-Carlo samples share a
 * CTA, beyond that the launcher trades samples per CTA for room). NULL + rs_last_error() on invalid capacities. */
rs_pose_ctx* rs_pose_create(int max_batch, int max_matches, int max_iterations, int max_variance, int device);
void rs_pose_destroy(rs_pose_ctx* ctx);

/* Replaces Pose_Optimization::compute_optimized_pose for a batch of independent frames.
 * cur_pose = B x 7; matches = B x max_matches rs_match (frame b uses the first n_matches[b]);
 * out = B records; inlier_mask = B x max_matches bytes (may be NULL). HOST memory. */
int rs_pose_solve_batched(rs_pose_ctx* ctx, const double* cur_pose, const rs_match* matches, const int32_t* n_matches,
                          int batch, const rs_pose_opts* opts, rs_pose_out* out, uint8_t* inlier_mask);

/* The same call split in two, so that the solve overlaps other work of the frame (the reference runs find_primitives
 * on a std::async thread beside the rest of RGBD_SLAM::track, rgbd_slam.cpp:288-300): _begin enqueues the upload, the
 * kernels and the download on the context's stream and returns (RS_RNG_DEVICE; with RS_RNG_REFERENCE the host-side
 * random draws make it block until the covariance kernel is enqueued); _end waits. The host buffers, `out` and
 * `inlier_mask` must stay valid (and should be pinned) until _end returns. */
int rs_pose_solve_batched_begin(rs_pose_ctx* ctx, const double* cur_pose, const rs_match* matches, const int32_t* n_matches,
                                int batch, const rs_pose_opts* opts, rs_pose_out* out, uint8_t* inlier_mask);
int rs_pose_solve_batched_end(rs_pose_ctx* ctx);

/* Single-frame convenience with the reference's call shape (creates nothing: uses ctx, batch 1). */
int rs_pose_solve(rs_pose_ctx* ctx, const double cur_pose[7], const rs_match* matches, int n_matches,
                  const rs_pose_opts* opts, rs_pose_out* out, uint8_t* inlier_mask);

/* Device-resident variant (inputs uploaded once with rs_pose_upload; outputs stay on the device
 * until rs_pose_download). Asynchronous on `stream`. RS_RNG_DEVICE only. */
int rs_pose_upload(rs_pose_ctx* ctx, const double* cur_pose, const rs_match* matches, const int32_t* n_matches, int batch);
int rs_pose_solve_device(rs_pose_ctx* ctx, int batch, const rs_pose_opts* opts, void* stream);
/* The preparation step of rs_pose_solve_device alone (AoS -> SoA of the uploaded match lists, validity, reset of the work
 * state): a caller whose `stream` will be held up by other work (bench.py: the pose stream waits for the plane-fit kernel)
 * enqueues it ahead of that wait; the following rs_pose_solve_device call for the same batch then launches the solve kernel
 * only. Asynchronous. */
int rs_pose_prepare_device(rs_pose_ctx* ctx, int batch, const rs_pose_opts* opts, void* stream);
/* More CTAs for the solve kernel most recently launched through this context (its work lives in global-memory queues, any
 * number of launches may feed on them): `ctas_per_sm` <= 0 -> as many as fit. They leave at once when no work is left.
 * `stream` must not be one that runs ahead of that solve launch or behind the next one. Asynchronous. */
int rs_pose_add_workers(rs_pose_ctx* ctx, int ctas_per_sm, void* stream);
int rs_pose_download(rs_pose_ctx* ctx, int batch, rs_pose_out* out, uint8_t* inlier_mask);
/* Makes `stream` wait until the solve kernel of the most recent solve launched through this context has finished (with
 * RS_RNG_REFERENCE: its first half, hypotheses + final LM): the counterpart of rs_cape_stream_wait_fit. */
int rs_pose_stream_wait_ransac(rs_pose_ctx* ctx, void* stream);
double* rs_pose_device_poses(rs_pose_ctx* ctx); /* B x 7 doubles on the device (the all-gather payload) */

/* Per-kernel device timing, as rs_cape_set_timing (replaces the static timing doubles of Pose_Optimization,
 * pose_optimization.hpp:97-102). ms[0] = prepare, ms[1] = the solve kernel (RANSAC hypotheses, final LM, Monte-Carlo LM solves
 * and covariance of the batch in one persistent launch; with RS_RNG_REFERENCE its first half: hypotheses + final LM),
 * ms[2] = its second half with RS_RNG_REFERENCE (Monte-Carlo solves + covariance; the host's random draws between the halves
 * are not counted), else 0, ms[3] = 0. */
int rs_pose_set_timing(rs_pose_ctx* ctx, int n_slots);
int rs_pose_kernel_ms(rs_pose_ctx* ctx, int slot, float ms[4]);
/* Inside view of the most recent solve-kernel launch of this context, from the device's own timer (waits for the device):
 * ms[0] = first CTA in -> last frame's final LM out (the RANSAC phase; the Monte-Carlo solves of finished frames already run
 * beside it), ms[1] = first CTA in -> last CTA out. */
int rs_pose_phase_ms(rs_pose_ctx* ctx, float ms[2]);
/* Work counters of the most recent solve-kernel launch (waits for the device): [0] hypotheses run by a frame's first CTA,
 * [1] by CTAs that joined a frame later, [2] bookkeeping rounds, [3] hypotheses applied by the serial rule, [4] joins,
 * [5] Monte-Carlo tasks, [6] hypotheses dropped by the early stop, [7] reserved. */
int rs_pose_debug_counters(rs_pose_ctx* ctx, uint64_t out[8]);
/* Per-frame timeline of the most recent solve-kernel launch, ms since its first CTA started (-1: did not happen):
 * ms[b][0] hypotheses started, [1] hypothesis stage closed, [2] final LM done, [3] covariance done. */
int rs_pose_debug_frame_times(rs_pose_ctx* ctx, int batch, double* ms);

/* Debug/parity taps (RS_RNG_DEVICE): the random inputs the last solve used, so that the checker
 * can feed the same ones to the oracle. subsets = B x max_iterations x RS_MAX_SUBSET int32 (-1 padded);
 * normals = B x n_variance x max_matches x 4 doubles. Either may be NULL. */
#define RS_MAX_SUBSET 16
int rs_pose_export_random(rs_pose_ctx* ctx, int batch, int32_t* subsets, double* normals);

/* ---- misc ------------------------------------------------------------------------------- */
const char* rs_last_error(void);
const char* rs_version(void);
/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
uint64_t rs_launch_count(void);

/* ---- Kalman update of matched map features (the step after the pose solve) -------------------
 * Replaces tracking::SharedKalmanFilter<N, N>::get_new_state (src/tracking/kalman_filter.hpp:46-118) as called by
 *   tracking::Point::track  (src/tracking/point_with_tracking.cpp:32-84; N = 3, Q = process_noise * I, 0.001 there) and
 *   tracking::Plane::track  (src/tracking/plane_with_tracking.cpp:16-59,81-95; N = 4, 1e-6 there; without the polygon merge),
 * for n features at once (one thread each). Arrays are row-major, [n][N] / [n][N][N]; points: state = world position,
 * planes: state = (nx, ny, nz, d) with the filtered normal re-normalised. Per feature: out_status 0, or -1 / -2 when the
 * state / measurement covariance is not a valid covariance (is_covariance_valid, covariances.hpp:13-44), -4 when the
 * result is not a valid covariance (the reference throws). An innovation covariance whose determinant is within DBL_EPSILON of
 * zero takes the reference's pseudo-inverse branch (kalman_filter.hpp:73-77; Moore-Penrose inverse, Eigen's rank threshold); in those cases the feature is returned unchanged and
 * out_score = -1, as Point::track reports a refused update. out_score = |state - new state| otherwise; out_moving (points,
 * may be NULL) = the detection left the point's position by more than its own standard deviation on some axis.
 * Host-pointer entry points copy in and out on `device`; the _device variants take device pointers and are asynchronous. */
int rs_kalman_track_points(int device, int n, const double* state, const double* cov, const double* meas, const double* meas_cov,
                           double process_noise, double* out_state, double* out_cov, double* out_score, uint8_t* out_moving,
                           int32_t* out_status);
int rs_kalman_track_planes(int device, int n, const double* state, const double* cov, const double* meas, const double* meas_cov,
                           double process_noise, double* out_state, double* out_cov, double* out_score, int32_t* out_status);
int rs_kalman_track_points_device(int n, const double* state, const double* cov, const double* meas, const double* meas_cov,
                                  double process_noise, double* out_state, double* out_cov, double* out_score,
                                  uint8_t* out_moving, int32_t* out_status, void* stream);
int rs_kalman_track_planes_device(int n, const double* state, const double* cov, const double* meas, const double* meas_cov,
                                  double process_noise, double* out_state, double* out_cov, double* out_score,
                                  int32_t* out_status, void* stream);

/* ---- plane matching against the local map (the step between find_primitives and the pose solve) ---------------------
 * Replaces the body of MapPlane::find_matches (src/map_management/map_features/map_primitive.cpp:91-161) for every map plane
 * of every frame at once: the map plane and its boundary polygon go to camera space (PlaneWorldCoordinates::
 * to_camera_coordinates, plane_coordinates.cpp:20-24; WorldPolygon::to_camera_space, polygon_coordinates.cpp:135-162), every
 * detected plane of the frame that passes Plane::is_distance_similar / is_normal_similar (shape_primitives.cpp:66-84; 100 mm,
 * 20 degrees) gets the map polygon projected into its own frame (Polygon::project, polygon.cpp:349-382) and intersected with
 * its boundary polygon (Polygon::inter_area, polygon.cpp:542-561: the summed area of boost::geometry::intersection); the
 * detection with the greatest intersection area whose share of the detection's own area reaches
 * minimumPlaneOverlapToConsiderMatch (0.4f; half of it with advanced_search) is selected. As in the reference, detection 0 of
 * a frame can never be selected (`if (selectedIndex <= 0) return`, :146).
 * A plane with its polygon: the parametrisation (camera frame for detections, world frame for map planes), the polygon's
 * frame (Polygon::_center / _xAxis / _yAxis) and its ring in that frame (open or closed, either orientation, simple). The
 * polygons themselves are built on the host (CameraPolygon's concave hull / correct / simplify are flann + boost::geometry
 * third-party code) from the boundary points rs_cape_run returns. */
typedef struct rs_polygon_plane {
    double normal[3];
    double d;
    double center[3];
    double x_axis[3];
    double y_axis[3];
    int32_t first_vertex;   /* index of the ring's first (x, y) pair in the xy array */
    int32_t n_vertices;
} rs_polygon_plane;

/* Frame f owns detections [det_first[f], det_first[f+1]) and map planes [map_first[f], map_first[f+1]) (batched sequences:
 * every frame has its own local map). world_to_camera: n_frames row-major 4x4. det_matched (may be NULL): per detection, the
 * reference's isDetectedFeatureMatched when the call is made. Outputs per map plane: selected = index of the detection INSIDE
 * its frame's list or -1, inter_area = the winning intersection area (0 if none).
 * sequential != 0 reproduces the caller's loop as well (Feature_Map::get_matches, feature_map.hpp:652-669): the map planes of
 * a frame are served in list order and a detection taken by one is marked matched for those after it; det_matched_out (may be
 * NULL; per detection) returns the mask as the loop leaves it. sequential == 0: every map plane is matched against the mask as
 * given (what a single find_matches call sees). Host pointers; runs on `device`. */
int rs_plane_match(int device, int n_frames, const double* world_to_camera, const rs_polygon_plane* det,
                   const int32_t* det_first, const double* det_xy, const rs_polygon_plane* map, const int32_t* map_first,
                   const double* map_xy, const uint8_t* det_matched, int advanced_search, int sequential, int32_t* selected,
                   double* inter_area, uint8_t* det_matched_out);
/* The intersection area alone, for n_pairs polygon pairs given in a common 2-D frame (a = ring a_first[i]..a_first[i+1] of
 * a_xy, same for b): what Polygon::inter_area returns once `other` has been projected. */
int rs_polygon_inter_area(int device, int n_pairs, const double* a_xy, const int32_t* a_first, const double* b_xy,
                          const int32_t* b_first, double* area);

#ifdef __cplusplus
}
#endif
#endif /* RGBDSLAM_B200_H */
