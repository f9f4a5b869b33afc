// Internal declarations shared by the CAPE kernels and the C-ABI glue (not part of the public header).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace rs {

struct CellFitParams {
    int H, hc, vc, batch, cell;
    const double* kx;  // [W]  back-projection factor per image column (device)
    const double* ky;  // [H]  per image row
    const double* kxs; // [W]  kx * 2^896 (exact scaling; K1a widens floats without the 2^896, see cape_cell_fit.cu)
    const double* kys; // [H]
    int min_zero_point_count;   // floor(P * 0.7f), plane_segment.hpp:33-34
    float sin_merge;            // sinf(float(18 deg)), primitive_detection.cpp:189-190
    float merge_distance;       // 50 mm
    int items_per_strip, total_items;  // filled by the launcher
};

// `streamed` (may be null) is recorded after the streaming kernel K1a, before the per-cell fit kernel K1b
int launch_cape_cell_fit(const CUtensorMap& tmap, const CellFitParams& prm, rs_cell_out* cells, cudaStream_t stream,
                         cudaEvent_t streamed);
// TMA box of the plane-fit kernel: {cell px, cape_cell_fit_box_cells() cells, cape_cell_fit_box_rows(cell) rows}
int cape_cell_fit_box_rows(int cell);
int cape_cell_fit_box_cells();

struct SegmentParams {
    int W, H, hc, vc, cell, batch;
    const double* kx;
    const double* ky;
    double cos_merge;            // cos(18 deg) in FP64 (plane_segment.cpp:324)
    const double* uniforms;      // canonical doubles of mt19937(seed) (cylinder RANSAC draws), device
    int n_uniforms;
    int max_boundary;
};

struct SegmentBuffers {
    const float* depth;          // B x H x W
    const rs_cell_out* cells;    // B x Nc
    int32_t* plane_grid;
    int32_t* plane_labels;
    int32_t* cyl_labels;
    int32_t* cyl_region_seg;
    rs_plane_out* planes;        // B x RS_MAX_PLANES
    rs_cyl_out* cyls;            // B x RS_MAX_CYL_REGIONS
    double* boundary_xyz;        // B x max_boundary x 3
    rs_cape_frame_info* info;    // B
    double* scratch;             // B x cape_segment_scratch_doubles_per_frame(Nc): projected normals / centroids of the cylinder branch
};

struct RectifyParams {
    int W, H, batch;
    const float* preX;  // [W] the reference's _Xpre row: K1's back-projection factors as floats (init_matrices, :147-173)
    const float* preY;  // [H] _Ypre column
    double fx, fy, cx, cy;   // camera 1 intrinsics (the reference projects AND back-projects with camera 1's)
    double T[12];            // first three rows of the camera-2 -> camera-1 transformation, row-major
};
// out: batch*H*W floats (also the scratch the winners are found in)
int launch_rectify_depth(const RectifyParams& prm, const float* depth, float* out, cudaStream_t stream);

int launch_cape_segment(const SegmentParams& prm, const SegmentBuffers& buf, cudaStream_t stream);
int cape_segment_validate(int hc, int vc, int cell);   // RS_OK, or RS_ERR_INVALID_ARG + rs_last_error() for a grid the kernel cannot take
size_t cape_segment_scratch_doubles_per_frame(int n_cells);

}  // namespace rs
