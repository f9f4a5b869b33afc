"""ctypes mirror of include/rgbdslam_b200.h (POD structs and constants). Shared by the product binding (lib.py)
and by the test-side checker binding under tests/; contains no computation."""
import ctypes as C

import numpy as np

RS_OK, RS_ERR_INVALID_ARG, RS_ERR_CUDA, RS_ERR_NO_DEVICE, RS_ERR_CAPACITY = 0, 1, 2, 3, 4
RS_MAX_PLANES = 128
RS_MAX_CYL_REGIONS = 32
RS_MAX_CYL_SEGS = 8
RS_CYL_RANSAC_ITERS = 43
RS_FEAT_POINT, RS_FEAT_PLANE, RS_FEAT_POINT2D = 0, 1, 2
RS_RNG_REFERENCE, RS_RNG_DEVICE = 0, 1
RS_SOLVER_AUTO, RS_SOLVER_CHAIN, RS_SOLVER_FUSED, RS_SOLVER_WIDE = 0, 1, 2, 3
RS_MAX_SUBSET = 16

# numpy structured dtypes with the exact C layout (align=True reproduces the compiler's padding)
cell_dtype = np.dtype([("count", "<i4"), ("planar", "<i4"), ("S", "<f8", (9,)), ("centroid", "<f8", (3,)),
                       ("normal", "<f8", (3,)), ("d", "<f8"), ("mse", "<f8"), ("score", "<f8"), ("tol", "<f4"),
                       ("hist_bin", "<i4")], align=True)
plane_dtype = np.dtype([("merge_label", "<i4"), ("planar", "<i4"), ("is_final", "<i4"), ("count", "<i4"),
                        ("S", "<f8", (9,)), ("centroid", "<f8", (3,)), ("normal", "<f8", (3,)), ("d", "<f8"),
                        ("mse", "<f8"), ("score", "<f8"), ("n_boundary", "<i4"), ("boundary_offset", "<i4")], align=True)
cyl_dtype = np.dtype([("n_cells", "<i4"), ("n_segments", "<i4"), ("pca_score", "<f8"), ("axis", "<f8", (3,)),
                      ("radius", "<f8", (RS_MAX_CYL_SEGS,)), ("center", "<f8", (RS_MAX_CYL_SEGS, 3)),
                      ("mse", "<f8", (RS_MAX_CYL_SEGS,)), ("plane_mse", "<f8", (RS_MAX_CYL_SEGS,)),
                      ("n_inliers", "<i4", (RS_MAX_CYL_SEGS,)), ("assigned", "<i4", (RS_MAX_CYL_SEGS,)),
                      ("kept", "<i4", (RS_MAX_CYL_SEGS,))], align=True)
info_dtype = np.dtype([("status", "<i4"), ("n_planar_cells", "<i4"), ("n_seeds", "<i4"), ("n_planes", "<i4"),
                       ("n_final_planes", "<i4"), ("n_cyl_regions", "<i4"), ("n_cylinders", "<i4"),
                       ("n_boundary", "<i4")], align=True)
match_dtype = np.dtype([("type", "<i4"), ("reserved", "<i4"), ("obs", "<f8", (4,)), ("map", "<f8", (4,)),
                        ("sigma", "<f8", (4,))], align=True)
pose_out_dtype = np.dtype([("status", "<i4"), ("n_inliers", "<i4"), ("iterations_run", "<i4"),
                           ("best_iteration", "<i4"), ("n_variance_ok", "<i4"), ("reserved", "<i4"), ("score", "<f8"),
                           ("pose", "<f8", (7,)), ("cov", "<f8", (36,))], align=True)

polygon_plane_dtype = np.dtype([("normal", "<f8", (3,)), ("d", "<f8"), ("center", "<f8", (3,)), ("x_axis", "<f8", (3,)),
                                ("y_axis", "<f8", (3,)), ("first_vertex", "<i4"), ("n_vertices", "<i4")], align=True)

assert polygon_plane_dtype.itemsize == 112
assert cell_dtype.itemsize == 160, cell_dtype.itemsize
assert match_dtype.itemsize == 104
assert info_dtype.itemsize == 32


class CapeOutputs(C.Structure):
    _fields_ = [("cells", C.c_void_p), ("plane_grid", C.c_void_p), ("plane_labels", C.c_void_p),
                ("cyl_labels", C.c_void_p), ("cyl_region_seg", C.c_void_p), ("planes", C.c_void_p),
                ("cyls", C.c_void_p), ("boundary_xyz", C.c_void_p), ("info", C.c_void_p)]


class PoseOpts(C.Structure):
    _fields_ = [("max_iterations", C.c_int32), ("n_variance", C.c_int32), ("rng_mode", C.c_int32),
                ("seed", C.c_uint32), ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("lm_max_fev", C.c_int32), ("sub_batches", C.c_int32), ("worker_ctas_per_sm", C.c_int32), ("solver", C.c_int32)]


def alloc_cape_outputs(batch, n_cells, max_boundary):
    """Host buffers for one batched CAPE run, plus the CapeOutputs struct pointing at them."""
    arrs = {
        "cells": np.zeros((batch, n_cells), dtype=cell_dtype),
        "plane_grid": np.zeros((batch, n_cells), dtype=np.int32),
        "plane_labels": np.zeros((batch, n_cells), dtype=np.int32),
        "cyl_labels": np.zeros((batch, n_cells), dtype=np.int32),
        "cyl_region_seg": np.zeros((batch, n_cells), dtype=np.int32),
        "planes": np.zeros((batch, RS_MAX_PLANES), dtype=plane_dtype),
        "cyls": np.zeros((batch, RS_MAX_CYL_REGIONS), dtype=cyl_dtype),
        "boundary_xyz": np.zeros((batch, max_boundary, 3), dtype=np.float64),
        "info": np.zeros((batch,), dtype=info_dtype),
    }
    st = CapeOutputs(**{k: v.ctypes.data for k, v in arrs.items()})
    return arrs, st
