"""TEST INFRASTRUCTURE: ctypes binding of the CPU oracle (oracle/_build/liboracle.so). Only tests/, smoke() and
bench.py's CPU-baseline legs may import this; the product package never does."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "_build", "liboracle.so")

import rgbd_slam_b200 as rs  # noqa: E402  (abi structs only)

abi = rs.abi
_lib = None


def build():
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    return LIB


def load():
    global _lib
    if _lib is not None:
        return _lib
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".hpp", "Makefile"))]
    if not os.path.exists(LIB) or any(os.path.getmtime(f) > os.path.getmtime(LIB) for f in srcs):
        build()   # make rebuilds only what is stale
    lib = C.CDLL(LIB)
    vp, i32, u32, dbl = C.c_void_p, C.c_int, C.c_uint32, C.c_double
    lib.orc_cape_run.argtypes = [i32, i32, i32, dbl, dbl, dbl, dbl, vp, i32, u32, i32, C.POINTER(abi.CapeOutputs)]
    lib.orc_cape_cell_fit.argtypes = [i32, i32, i32, dbl, dbl, dbl, dbl, vp, i32, vp, vp]
    lib.orc_eigen3.argtypes = [vp, vp, vp]
    lib.orc_morphology.argtypes = [vp, i32, i32, i32, i32, i32, vp]
    lib.orc_world_to_camera.argtypes = [vp, vp, vp]
    lib.orc_pose_coefficients.argtypes = [vp, vp]
    lib.orc_pose_from_coefficients.argtypes = [vp, vp, vp]
    lib.orc_quaternion_from_euler.argtypes = [dbl, dbl, dbl, vp]
    lib.orc_residual_count.argtypes = [vp, i32]
    lib.orc_pose_residuals.argtypes = [vp, vp, i32, vp, vp]
    lib.orc_pose_inliers.argtypes = [vp, vp, vp, i32, vp]
    lib.orc_pose_lm.argtypes = [vp, vp, i32, vp, i32, vp]
    lib.orc_pose_solve.argtypes = [vp, vp, vp, i32, i32, i32, u32, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.orc_process_frames.restype = dbl
    lib.orc_process_frames.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, u32, i32, vp]
    lib.orc_kalman_new_state.argtypes = [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.orc_kalman_track_points.argtypes = [i32, vp, vp, vp, vp, dbl, vp, vp, vp, vp, vp]
    lib.orc_kalman_track_planes.argtypes = [i32, vp, vp, vp, vp, dbl, vp, vp, vp, vp]
    lib.orc_polygon_inter_area.restype = dbl
    lib.orc_polygon_inter_area.argtypes = [vp, i32, vp, i32]
    lib.orc_polygon_area.restype = dbl
    lib.orc_polygon_area.argtypes = [vp, i32]
    lib.orc_plane_match.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp]
    lib.orc_rectify_depth.argtypes = [i32, i32, dbl, dbl, dbl, dbl, vp, vp, i32, vp]
    lib.orc_ref_test_features.argtypes = [vp, dbl, dbl, dbl, dbl, vp, i32]
    _lib = lib
    return lib


def cape_run(depth, cell=20, K=(550.0, 550.0, 320.0, 240.0), seed=0):
    lib = load()
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    if depth.ndim == 2:
        depth = depth[None]
    B, H, W = depth.shape
    Nc = (W // cell) * (H // cell)
    arrs, st = abi.alloc_cape_outputs(B, Nc, 2 * Nc)
    lib.orc_cape_run(W, H, cell, *K, depth.ctypes.data, B, seed, 2 * Nc, C.byref(st))
    return arrs


def kalman_new_state(F, H, Q, x, P, z, R):
    """SharedKalmanFilter<N, M>::get_new_state restated (oracle/kalman.cpp). Returns (status, x_new, P_new)."""
    lib = load()
    F, H, Q, P, R = (np.ascontiguousarray(np.atleast_2d(a), dtype=np.float64) for a in (F, H, Q, P, R))
    x = np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64)
    z = np.ascontiguousarray(np.atleast_1d(z), dtype=np.float64)
    N, M = len(x), len(z)
    xo, Po = np.zeros(N), np.zeros((N, N))
    rc = lib.orc_kalman_new_state(N, M, F.ctypes.data, H.ctypes.data, Q.ctypes.data, x.ctypes.data, P.ctypes.data,
                                  z.ctypes.data, R.ctypes.data, xo.ctypes.data, Po.ctypes.data)
    return rc, xo, Po


def kalman_track_points(x, P, z, R, process_noise=0.001):
    lib = load()
    x, P, z, R = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, P, z, R))
    n = len(x)
    xo, Po, score = np.zeros((n, 3)), np.zeros((n, 3, 3)), np.zeros(n)
    moving, status = np.zeros(n, np.uint8), np.zeros(n, np.int32)
    lib.orc_kalman_track_points(n, x.ctypes.data, P.ctypes.data, z.ctypes.data, R.ctypes.data, process_noise, xo.ctypes.data,
                                Po.ctypes.data, score.ctypes.data, moving.ctypes.data, status.ctypes.data)
    return xo, Po, score, moving, status


def kalman_track_planes(x, P, z, R, process_noise=1e-6):
    lib = load()
    x, P, z, R = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, P, z, R))
    n = len(x)
    xo, Po, score, status = np.zeros((n, 4)), np.zeros((n, 4, 4)), np.zeros(n), np.zeros(n, np.int32)
    lib.orc_kalman_track_planes(n, x.ctypes.data, P.ctypes.data, z.ctypes.data, R.ctypes.data, process_noise, xo.ctypes.data,
                                Po.ctypes.data, score.ctypes.data, status.ctypes.data)
    return xo, Po, score, status


def rectify_depth(depth, cam2_to_cam1=None, K=(550.0, 550.0, 320.0, 240.0)):
    """Depth_Map_Transformation::rectify_depth restated (oracle/cape.cpp): depth [B,H,W] float32 -> rectified [B,H,W]."""
    lib = load()
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    if depth.ndim == 2:
        depth = depth[None]
    B, H, W = depth.shape
    T = np.ascontiguousarray(np.eye(4) if cam2_to_cam1 is None else cam2_to_cam1, dtype=np.float64).reshape(16)
    out = np.zeros_like(depth)
    lib.orc_rectify_depth(W, H, *K, T.ctypes.data, depth.ctypes.data, B, out.ctypes.data)
    return out


def cape_cell_fit(depth, cell=20, K=(550.0, 550.0, 320.0, 240.0), want_cloud=False):
    lib = load()
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    if depth.ndim == 2:
        depth = depth[None]
    B, H, W = depth.shape
    Nc = (W // cell) * (H // cell)
    cells = np.zeros((B, Nc), dtype=abi.cell_dtype)
    cloud = np.zeros((B, 3, W * H), dtype=np.float32) if want_cloud else None
    lib.orc_cape_cell_fit(W, H, cell, *K, depth.ctypes.data, B, cells.ctypes.data, cloud.ctypes.data if want_cloud else None)
    return (cells, cloud) if want_cloud else cells


def pose_solve(cur_pose, matches, K=(550.0, 550.0, 320.0, 240.0), max_iterations=0, n_variance=-1, seed=0, subsets=None,
               normals=None, max_matches=0, lm_max_fev=0, taps=False):
    lib = load()
    Kc = np.asarray(K, dtype=np.float64)
    cur = np.ascontiguousarray(cur_pose, dtype=np.float64)
    m = np.ascontiguousarray(matches)
    n = len(m)
    out = np.zeros((1,), dtype=abi.pose_out_dtype)
    mask = np.zeros((n,), dtype=np.uint8)
    iters = max_iterations if max_iterations > 0 else lib.orc_ransac_default_iterations()
    cand_poses = np.zeros((iters, 7))
    cand_ok = np.zeros((iters,), dtype=np.int32)
    cand_scores = np.zeros((iters,))
    subsets_out = np.full((iters, abi.RS_MAX_SUBSET), -1, dtype=np.int32)
    if subsets is not None:
        subsets = np.ascontiguousarray(subsets, dtype=np.int32)
    if normals is not None:
        normals = np.ascontiguousarray(normals, dtype=np.float64)
    lib.orc_pose_solve(Kc.ctypes.data, cur.ctypes.data, m.ctypes.data, n, max_iterations, n_variance, seed,
                       subsets.ctypes.data if subsets is not None else None,
                       normals.ctypes.data if normals is not None else None, max_matches, lm_max_fev,
                       out.ctypes.data, mask.ctypes.data, cand_poses.ctypes.data, cand_ok.ctypes.data,
                       cand_scores.ctypes.data, subsets_out.ctypes.data)
    if taps:
        return out[0], mask, dict(cand_poses=cand_poses, cand_ok=cand_ok, cand_scores=cand_scores, subsets=subsets_out)
    return out[0], mask


def pose_residuals(matches, x6, K=(550.0, 550.0, 320.0, 240.0)):
    """Global_Pose_Estimator::operator() as the oracle restates it (features taken as given)."""
    lib = load()
    Kc = np.asarray(K, dtype=np.float64)
    m = np.ascontiguousarray(matches)
    x = np.ascontiguousarray(x6, dtype=np.float64)
    out = np.zeros(lib.orc_residual_count(m.ctypes.data, len(m)))
    lib.orc_pose_residuals(Kc.ctypes.data, m.ctypes.data, len(m), x.ctypes.data, out.ctypes.data)
    return out


def pose_lm_evaluations(cur_pose, matches, K=(550.0, 550.0, 320.0, 240.0), maxfev=400):
    """Function evaluations the oracle's LM (compute_optimized_global_pose) spends on `matches` from `cur_pose`, and its status."""
    lib = load()
    lib.orc_pose_lm.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    lib.orc_pose_coefficients.argtypes = [C.c_void_p, C.c_void_p]
    Kc = np.asarray(K, dtype=np.float64)
    m = normalized_planes(matches)
    x = np.zeros(6)
    lib.orc_pose_coefficients(np.ascontiguousarray(cur_pose, dtype=np.float64).ctypes.data, x.ctypes.data)
    nfev = C.c_int(0)
    status = lib.orc_pose_lm(Kc.ctypes.data, m.ctypes.data, len(m), x.ctypes.data, maxfev, C.byref(nfev))
    return nfev.value, status


def normalized_planes(matches):
    """Plane normals normalised once, as pose_solve does on entry (normalize_features)."""
    m = np.array(matches, copy=True)
    for f in m:
        if f["type"] == abi.RS_FEAT_PLANE:
            for key in ("obs", "map"):
                v = f[key][:3]
                z = v[0] * v[0] + v[1] * v[1] + v[2] * v[2]
                if z > 0:
                    f[key][:3] = v / np.sqrt(z)
    return np.ascontiguousarray(m)


def pose_inliers(pose7, matches, K=(550.0, 550.0, 320.0, 240.0)):
    lib = load()
    Kc = np.asarray(K, dtype=np.float64)
    p = np.ascontiguousarray(pose7, dtype=np.float64)
    m = np.ascontiguousarray(matches)
    mask = np.zeros(len(m), np.uint8)
    lib.orc_pose_inliers(Kc.ctypes.data, p.ctypes.data, m.ctypes.data, len(m), mask.ctypes.data)
    return mask


def ref_test_features(true_pose, point_error=5.0, point_outliers=0.0, plane_error=5.0, plane_outliers=-1.0):
    """Scenario builder of tests/test_pose_optimization.cpp. outliers < 0 disables that feature kind."""
    lib = load()
    tp = np.ascontiguousarray(true_pose, dtype=np.float64)
    buf = np.zeros((512,), dtype=abi.match_dtype)
    n = lib.orc_ref_test_features(tp.ctypes.data, point_error, point_outliers, plane_error, plane_outliers, buf.ctypes.data, 512)
    return buf[:n].copy()


def quaternion_from_euler(yaw, pitch, roll):
    q = np.zeros(4)
    load().orc_quaternion_from_euler(yaw, pitch, roll, q.ctypes.data)
    return q


def morphology(mask, erode, cross, border_zero):
    """The oracle's restatement of cv::erode / cv::dilate with a 3x3 square or cross kernel (oracle/cape.cpp: morph)."""
    lib = load()
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    out = np.empty_like(m)
    lib.orc_morphology(m.ctypes.data, m.shape[0], m.shape[1], int(erode), int(cross), int(border_zero), out.ctypes.data)
    return out


def polygon_inter_area(a, b):
    """Polygon::inter_area once `other` is projected (polygon.cpp:542-561): area of the intersection of two simple rings."""
    lib = load()
    a, b = np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64)
    return lib.orc_polygon_inter_area(a.ctypes.data, len(a), b.ctypes.data, len(b))


def polygon_area(a):
    lib = load()
    a = np.ascontiguousarray(a, dtype=np.float64)
    return lib.orc_polygon_area(a.ctypes.data, len(a))


def plane_match(w2c, det, det_first, det_xy, mp, map_first, map_xy, det_matched=None, advanced_search=False, sequential=False,
                return_matched=False):
    """MapPlane::find_matches for every map plane of every frame (oracle/polygon.cpp). Same arguments as rs.plane_match."""
    lib = load()
    w2c = np.ascontiguousarray(w2c, dtype=np.float64)
    det, mp = np.ascontiguousarray(det, dtype=abi.polygon_plane_dtype), np.ascontiguousarray(mp, dtype=abi.polygon_plane_dtype)
    det_first, map_first = np.ascontiguousarray(det_first, dtype=np.int32), np.ascontiguousarray(map_first, dtype=np.int32)
    det_xy, map_xy = np.ascontiguousarray(det_xy, dtype=np.float64), np.ascontiguousarray(map_xy, dtype=np.float64)
    n_frames = len(det_first) - 1
    sel, inter = np.full(len(mp), -1, np.int32), np.zeros(len(mp))
    dm = None if det_matched is None else np.ascontiguousarray(det_matched, dtype=np.uint8)
    mout = np.zeros(max(len(det), 1), np.uint8)
    lib.orc_plane_match(n_frames, w2c.ctypes.data, det.ctypes.data, det_first.ctypes.data, det_xy.ctypes.data, mp.ctypes.data,
                        map_first.ctypes.data, map_xy.ctypes.data, None if dm is None else dm.ctypes.data, int(advanced_search),
                        int(sequential), sel.ctypes.data, inter.ctypes.data, mout.ctypes.data)
    return (sel, inter, mout[:len(det)]) if return_matched else (sel, inter)


# ---- the reference's own CAPE sources, compiled against stand-in third-party headers (oracle/ref_shim -> oracle/_ref) ----
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_cape.so")
_ref = None


def ref_available():
    """True when oracle/_ref/libref_cape.so exists; it is (re)built from the reference tree wherever that is present
    (/root/reference: this container - the GPU box gets the prebuilt file with the snapshot)."""
    if os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-C", os.path.join(ORACLE_DIR, "ref_shim")], capture_output=True)
    return os.path.exists(REF_LIB)


def _ref_load():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_LIB)
        vp, i32 = C.c_void_p, C.c_int
        lib.ref_cape_run.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp, i32, vp, vp, i32, vp, vp, i32]
        lib.ref_rectify_depth.argtypes = [vp, i32, i32, vp, vp]
        _ref = lib
    return _ref


def ref_cape_run(depth):
    """One 640x480 frame through the REFERENCE's get_organized_cloud_array + find_primitives (MAKE_DETERMINISTIC: seed 0)."""
    lib = _ref_load()
    d = np.ascontiguousarray(depth, dtype=np.float32)
    H, W = d.shape
    cell = lib.ref_cape_cell_px()
    Nc = (W // cell) * (H // cell)
    out = dict(plane_grid=np.zeros(Nc, np.int32), cyl_grid=np.zeros(Nc, np.int32), planar=np.zeros(Nc, np.int32),
               cell=np.zeros((Nc, 5)), count=np.zeros(Nc, np.int32))
    planes, cyls, boundary = np.zeros((128, 6)), np.zeros((64, 4)), np.zeros((4 * Nc, 3))
    npl, ncy = C.c_int32(0), C.c_int32(0)
    rc = lib.ref_cape_run(d.ctypes.data, W, H, out["plane_grid"].ctypes.data, out["cyl_grid"].ctypes.data, out["planar"].ctypes.data,
                          out["cell"].ctypes.data, out["count"].ctypes.data, planes.ctypes.data, 128, C.byref(npl), cyls.ctypes.data,
                          64, C.byref(ncy), boundary.ctypes.data, len(boundary))
    if rc != 0:
        raise RuntimeError("reference get_organized_cloud_array failed")
    out["planes"], out["cyls"] = planes[:npl.value], cyls[:ncy.value]
    out["boundary"] = boundary[:int(planes[:npl.value, 5].sum())]
    return out


def ref_rectify_depth(depth, cam2_to_cam1):
    lib = _ref_load()
    d = np.ascontiguousarray(depth, dtype=np.float32)
    T = np.ascontiguousarray(cam2_to_cam1, dtype=np.float64)
    out = np.zeros_like(d)
    if lib.ref_rectify_depth(d.ctypes.data, d.shape[1], d.shape[0], T.ctypes.data, out.ctypes.data) != 0:
        raise RuntimeError("reference rectify_depth failed")
    return out


# ---- the reference's own pose-solve sources, compiled the same way (oracle/_ref/libref_pose.so) ----
REF_POSE_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_pose.so")
_ref_pose = None


def ref_pose_available():
    ref_available()          # runs the recipe where /root/reference exists
    return os.path.exists(REF_POSE_LIB)


def _ref_pose_load():
    global _ref_pose
    if _ref_pose is None:
        lib = C.CDLL(REF_POSE_LIB)
        vp, i32 = C.c_void_p, C.c_int
        lib.ref_pose_solve.argtypes = [vp, vp, i32, vp, vp, vp]
        lib.ref_pose_base.argtypes = [vp, vp]
        lib.ref_pose_residuals.argtypes = [vp, i32, vp, vp]
        lib.ref_pose_lm.argtypes = [vp, vp, i32, vp]
        lib.ref_pose_inliers.argtypes = [vp, vp, i32, vp]
        lib.ref_pose_trace_lm.argtypes = [vp, vp, i32, vp, vp, vp]
        lib.ref_pose_coefficients.argtypes = [vp, vp]
        lib.ref_pose_from_coefficients.argtypes = [vp, vp]
        _ref_pose = lib
    return _ref_pose


def ref_pose_solve(cur_pose, matches):
    """Pose_Optimization::compute_optimized_pose of the REFERENCE (MAKE_DETERMINISTIC: thread-local mt19937 from seed 0).
    -> (ok, pose7, cov 6x6, inlier mask)"""
    lib = _ref_pose_load()
    cur = np.ascontiguousarray(cur_pose, dtype=np.float64)
    m = np.ascontiguousarray(matches)
    pose, cov, mask = np.zeros(7), np.zeros(36), np.zeros(len(m), np.uint8)
    ok = lib.ref_pose_solve(cur.ctypes.data, m.ctypes.data, len(m), pose.ctypes.data, cov.ctypes.data, mask.ctypes.data)
    return bool(ok), pose, cov.reshape(6, 6), mask


def ref_pose_base(pose7):
    """The pose as utils::PoseBase(position, orientation) of the reference holds it: set_parameters normalises the quaternion
    (pose.cpp:16-22). ref_pose_solve / ref_pose_lm build exactly one PoseBase from the pose they are given, so the oracle has
    to be handed ref_pose_base(pose) for the same problem. (Iterating to a fixed point does not work: q / |q| can alternate
    between two neighbouring values for ever - problem 1097 of tools/sweep_reference_pose_build.py does.)"""
    lib = _ref_pose_load()
    p = np.ascontiguousarray(pose7, dtype=np.float64)
    q = np.zeros(7)
    lib.ref_pose_base(p.ctypes.data, q.ctypes.data)
    return q


def ref_pose_residuals(matches, x6):
    lib = _ref_pose_load()
    m = np.ascontiguousarray(matches)
    x = np.ascontiguousarray(x6, dtype=np.float64)
    out = np.zeros(3 * len(m))
    n = lib.ref_pose_residuals(m.ctypes.data, len(m), x.ctypes.data, out.ctypes.data)
    return out[:n]


def ref_pose_lm(cur_pose, matches):
    lib = _ref_pose_load()
    cur = np.ascontiguousarray(cur_pose, dtype=np.float64)
    m = np.ascontiguousarray(matches)
    pose = np.zeros(7)
    ok = lib.ref_pose_lm(cur.ctypes.data, m.ctypes.data, len(m), pose.ctypes.data)
    return bool(ok), pose


def ref_pose_inliers(pose7, matches):
    lib = _ref_pose_load()
    p = np.ascontiguousarray(pose7, dtype=np.float64)
    m = np.ascontiguousarray(matches)
    mask = np.zeros(len(m), np.uint8)
    lib.ref_pose_inliers(p.ctypes.data, m.ctypes.data, len(m), mask.ctypes.data)
    return mask


def stable_plane_normals(matches):
    """The reference re-normalises a plane's normal in every PlaneCoordinates constructor, COPY constructor and assignment
    (plane_coordinates.hpp:19-38), so a normal travels through an unknown number of normalisations before it is used, and
    x / |x| need not be idempotent in floating point (it can even alternate between two neighbours). The library and the oracle
    normalise once. For a comparison that does not depend on the number of copies, plane normals are moved to a fixed point of
    the normalisation first (nudged by an ulp where the iteration alternates)."""
    m = np.array(matches, copy=True)
    for f in m:
        if f["type"] != abi.RS_FEAT_PLANE:
            continue
        for key in ("obs", "map"):
            v = f[key][:3].copy()
            for attempt in range(64):
                for _ in range(8):
                    w = v / np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
                    if np.array_equal(w, v):
                        break
                    v = w
                else:
                    v[attempt % 3] = np.nextafter(v[attempt % 3], 2.0)
                    continue
                break
            else:
                raise RuntimeError("no fixed point of the normalisation")
            f[key][:3] = v
    return m
