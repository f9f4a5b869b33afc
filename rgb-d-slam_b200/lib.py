"""ctypes binding of librgbdslam_b200.so — the drop-in C-ABI of the CUDA hot path (include/rgbdslam_b200.h).

Host-side mirror of the reference's call surface:
  PrimitiveDetection.find_primitives   <-> Primitive_Detection::find_primitives (primitive_detection.hpp:42-45)
                                            + Depth_Map_Transformation::get_organized_cloud_array (fused)
  PoseOptimization.compute_optimized_pose <-> Pose_Optimization::compute_optimized_pose (pose_optimization.hpp:27-30)
There is no CPU fallback: if the shared library is missing or no sm_100 GPU is visible these raise."""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librgbdslam_b200.so")
_lib = None


class RsError(RuntimeError):
    pass


def load():
    """Loads the CUDA library (built by build.py). Raises if it is missing — never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RsError("librgbdslam_b200.so is not built (run `python rgb-d-slam_b200/build.py`); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, i32, u32, dbl = C.c_void_p, C.c_int, C.c_uint32, C.c_double
    lib.rs_cape_create.restype = vp
    lib.rs_cape_create.argtypes = [i32, i32, i32, dbl, dbl, dbl, dbl, i32, i32]
    lib.rs_cape_destroy.argtypes = [vp]
    lib.rs_cape_cells_per_frame.argtypes = [vp]
    lib.rs_cape_max_boundary.argtypes = [vp]
    lib.rs_cape_run.argtypes = [vp, vp, i32, u32, C.POINTER(abi.CapeOutputs)]
    lib.rs_cape_run_u16.argtypes = [vp, vp, C.c_double, i32, u32, C.POINTER(abi.CapeOutputs)]
    lib.rs_cape_run_device.argtypes = [vp, vp, i32, u32, C.POINTER(abi.CapeOutputs), vp]
    lib.rs_cape_cell_fit_device.argtypes = [vp, vp, i32, vp, vp]
    lib.rs_cape_stream_wait_fit.argtypes = [vp, vp]
    lib.rs_cape_segment_device.argtypes = [vp, vp, i32, u32, C.POINTER(abi.CapeOutputs), vp]
    lib.rs_pose_stream_wait_ransac.argtypes = [vp, vp]
    lib.rs_cape_set_rectification.argtypes = [vp, vp, i32]
    lib.rs_cape_rectify.argtypes = [vp, vp, i32, vp]
    lib.rs_cape_rectify_device.argtypes = [vp, vp, i32, vp, vp]
    lib.rs_cape_device_depth.restype = vp
    lib.rs_cape_device_depth.argtypes = [vp]
    lib.rs_cape_device_outputs.restype = C.POINTER(abi.CapeOutputs)
    lib.rs_cape_device_outputs.argtypes = [vp]
    lib.rs_cape_set_timing.argtypes = [vp, i32]
    lib.rs_cape_kernel_ms.argtypes = [vp, i32, vp]
    lib.rs_kalman_track_points.argtypes = [i32, i32, vp, vp, vp, vp, C.c_double, vp, vp, vp, vp, vp]
    lib.rs_kalman_track_planes.argtypes = [i32, i32, vp, vp, vp, vp, C.c_double, vp, vp, vp, vp]
    lib.rs_kalman_track_points_device.argtypes = [i32, vp, vp, vp, vp, C.c_double, vp, vp, vp, vp, vp, vp]
    lib.rs_kalman_track_planes_device.argtypes = [i32, vp, vp, vp, vp, C.c_double, vp, vp, vp, vp, vp]
    lib.rs_plane_match.argtypes = [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp]
    lib.rs_polygon_inter_area.argtypes = [i32, i32, vp, vp, vp, vp, vp]
    lib.rs_last_error.restype = C.c_char_p
    lib.rs_version.restype = C.c_char_p
    lib.rs_launch_count.restype = C.c_uint64
    if hasattr(lib, "rs_pose_create"):
        lib.rs_pose_create.restype = vp
        lib.rs_pose_create.argtypes = [i32, i32, i32, i32, i32]
        lib.rs_pose_solve_batched_begin.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp]
        lib.rs_pose_solve_batched_end.argtypes = [vp]
        lib.rs_pose_destroy.argtypes = [vp]
        lib.rs_pose_solve_batched.argtypes = [vp, vp, vp, vp, i32, C.POINTER(abi.PoseOpts), vp, vp]
        lib.rs_pose_solve.argtypes = [vp, vp, vp, i32, C.POINTER(abi.PoseOpts), vp, vp]
        lib.rs_pose_upload.argtypes = [vp, vp, vp, vp, i32]
        lib.rs_pose_solve_device.argtypes = [vp, i32, C.POINTER(abi.PoseOpts), vp]
        lib.rs_pose_prepare_device.argtypes = [vp, i32, C.POINTER(abi.PoseOpts), vp]
        lib.rs_pose_add_workers.argtypes = [vp, i32, vp]
        lib.rs_pose_download.argtypes = [vp, i32, vp, vp]
        lib.rs_pose_device_poses.restype = vp
        lib.rs_pose_device_poses.argtypes = [vp]
        lib.rs_pose_export_random.argtypes = [vp, i32, vp, vp]
        lib.rs_pose_set_timing.argtypes = [vp, i32]
        lib.rs_pose_kernel_ms.argtypes = [vp, i32, vp]
        lib.rs_pose_phase_ms.argtypes = [vp, vp]
        lib.rs_pose_debug_counters.argtypes = [vp, vp]
        lib.rs_pose_debug_frame_times.argtypes = [vp, i32, vp]
    _lib = lib
    return lib


def last_error():
    return load().rs_last_error().decode()


def _check(rc, what):
    if rc != abi.RS_OK:
        raise RsError("%s failed (status %d): %s" % (what, rc, last_error()))


def launch_count():
    return int(load().rs_launch_count())


class PrimitiveDetection:
    """Batched CAPE plane/cylinder extraction on one GPU.

    Mirrors Primitive_Detection(width, height) + find_primitives(...) of the reference; the organized point cloud is
    never materialised (the back-projection is fused into the plane-fit kernel)."""

    def __init__(self, width=640, height=480, cell_px=20, fx=550.0, fy=550.0, cx=320.0, cy=240.0, max_batch=1, device=0):
        lib = load()
        self._lib = lib
        self.width, self.height, self.cell_px, self.max_batch, self.device = width, height, cell_px, max_batch, device
        self._ctx = lib.rs_cape_create(width, height, cell_px, fx, fy, cx, cy, max_batch, device)
        if not self._ctx:
            raise RsError("rs_cape_create failed: " + last_error())
        self.n_cells = lib.rs_cape_cells_per_frame(self._ctx)
        self.max_boundary = lib.rs_cape_max_boundary(self._ctx)
        self.hc, self.vc = width // cell_px, height // cell_px

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.rs_cape_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def find_primitives(self, depth, seed=0, cells_only=False, out=None):
        """depth: float32 [B,H,W] (or [H,W]) host array in mm. Returns a dict of numpy outputs (see abi.py)."""
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        if depth.ndim == 2:
            depth = depth[None]
        B = depth.shape[0]
        if depth.shape[1:] != (self.height, self.width):
            raise ValueError("depth must be [B,%d,%d]" % (self.height, self.width))
        if out is None:
            arrs, st = abi.alloc_cape_outputs(B, self.n_cells, self.max_boundary)
        else:
            arrs, st = out
        if cells_only:
            st = abi.CapeOutputs(cells=arrs["cells"].ctypes.data)
        _check(self._lib.rs_cape_run(self._ctx, depth.ctypes.data, B, seed, C.byref(st)), "rs_cape_run")
        return arrs

    def find_primitives_u16(self, depth16, alpha=1.0, seed=0, out=None):
        """depth16: uint16 [B,H,W] raw sensor image; depth in mm = float32(depth16) * float32(alpha), as
        cv::Mat::convertTo(CV_32F, alpha) in the reference's examples. Same outputs as find_primitives."""
        depth16 = np.ascontiguousarray(depth16, dtype=np.uint16)
        if depth16.ndim == 2:
            depth16 = depth16[None]
        B = depth16.shape[0]
        if depth16.shape[1:] != (self.height, self.width):
            raise ValueError("depth must be [B,%d,%d]" % (self.height, self.width))
        arrs, st = abi.alloc_cape_outputs(B, self.n_cells, self.max_boundary) if out is None else out
        _check(self._lib.rs_cape_run_u16(self._ctx, depth16.ctypes.data, float(alpha), B, seed, C.byref(st)), "rs_cape_run_u16")
        return arrs

    # --- device-resident entry points (pointers are raw CUDA device addresses, e.g. torch.Tensor.data_ptr()) ---
    def device_depth_ptr(self):
        return self._lib.rs_cape_device_depth(self._ctx)

    def device_outputs(self):
        return self._lib.rs_cape_device_outputs(self._ctx).contents

    def run_device(self, depth_ptr, batch, seed=0, outputs=None, stream=0):
        o = outputs if outputs is not None else self.device_outputs()
        _check(self._lib.rs_cape_run_device(self._ctx, depth_ptr, batch, seed, C.byref(o), stream), "rs_cape_run_device")

    def set_rectification(self, cam2_to_cam1=None, enable=True):
        """rectify_depth in front of every find_primitives call (cam2_to_cam1: 4x4, identity by default)."""
        T = np.ascontiguousarray(np.eye(4) if cam2_to_cam1 is None else cam2_to_cam1, dtype=np.float64).reshape(16)
        _check(self._lib.rs_cape_set_rectification(self._ctx, T.ctypes.data, 1 if enable else 0), "rs_cape_set_rectification")

    def rectify_depth(self, depth):
        """Depth_Map_Transformation::rectify_depth alone: float32 [B,H,W] -> rectified float32 [B,H,W]."""
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        if depth.ndim == 2:
            depth = depth[None]
        out = np.empty_like(depth)
        _check(self._lib.rs_cape_rectify(self._ctx, depth.ctypes.data, depth.shape[0], out.ctypes.data), "rs_cape_rectify")
        return out

    def rectify_device(self, depth_ptr, batch, out_ptr, stream=0):
        _check(self._lib.rs_cape_rectify_device(self._ctx, depth_ptr, batch, out_ptr, stream), "rs_cape_rectify_device")

    def stream_wait_fit(self, stream):
        """`stream` (cudaStream_t as int) waits for the latest plane-fit kernel of this context."""
        _check(self._lib.rs_cape_stream_wait_fit(self._ctx, stream), "rs_cape_stream_wait_fit")

    def set_timing(self, n_slots):
        _check(self._lib.rs_cape_set_timing(self._ctx, n_slots), "rs_cape_set_timing")

    def kernel_ms(self, slot):
        """(plane-fit kernel ms, segmentation kernel ms) of the run that used this timing slot."""
        ms = (C.c_float * 2)()
        _check(self._lib.rs_cape_kernel_ms(self._ctx, slot, ms), "rs_cape_kernel_ms")
        return float(ms[0]), float(ms[1])

    def cell_fit_device(self, depth_ptr, batch, cells_ptr=None, stream=0):
        if cells_ptr is None:
            cells_ptr = self.device_outputs().cells
        _check(self._lib.rs_cape_cell_fit_device(self._ctx, depth_ptr, batch, cells_ptr, stream), "rs_cape_cell_fit_device")

    def segment_device(self, depth_ptr, batch, seed=0, outputs=None, stream=0):
        """The second half of run_device (K2-K4) on the records a preceding cell_fit_device call left in outputs.cells."""
        o = outputs if outputs is not None else self.device_outputs()
        _check(self._lib.rs_cape_segment_device(self._ctx, depth_ptr, batch, seed, C.byref(o), stream), "rs_cape_segment_device")


def kalman_track_points(state, cov, meas, meas_cov, process_noise=0.001, device=0):
    """tracking::Point::track for n matched map points at once (point_with_tracking.cpp:32-84).
    state / meas [n,3], cov / meas_cov [n,3,3]. Returns (new_state, new_cov, score, is_moving, status)."""
    lib = load()
    state, cov, meas, meas_cov = (np.ascontiguousarray(a, dtype=np.float64) for a in (state, cov, meas, meas_cov))
    n = state.shape[0]
    if state.shape != (n, 3) or cov.shape != (n, 3, 3) or meas.shape != (n, 3) or meas_cov.shape != (n, 3, 3):
        raise ValueError("expected state/meas [n,3] and cov/meas_cov [n,3,3]")
    out_state, out_cov, score = np.zeros((n, 3)), np.zeros((n, 3, 3)), np.zeros(n)
    moving, status = np.zeros(n, np.uint8), np.zeros(n, np.int32)
    _check(lib.rs_kalman_track_points(device, n, state.ctypes.data, cov.ctypes.data, meas.ctypes.data, meas_cov.ctypes.data,
                                      process_noise, out_state.ctypes.data, out_cov.ctypes.data, score.ctypes.data,
                                      moving.ctypes.data, status.ctypes.data), "rs_kalman_track_points")
    return out_state, out_cov, score, moving, status


def kalman_track_planes(state, cov, meas, meas_cov, process_noise=1e-6, device=0):
    """tracking::Plane::track's filter step for n matched map planes (plane_with_tracking.cpp:16-59): state = (n, d).
    Returns (new_state with unit normal, new_cov, score, status)."""
    lib = load()
    state, cov, meas, meas_cov = (np.ascontiguousarray(a, dtype=np.float64) for a in (state, cov, meas, meas_cov))
    n = state.shape[0]
    if state.shape != (n, 4) or cov.shape != (n, 4, 4) or meas.shape != (n, 4) or meas_cov.shape != (n, 4, 4):
        raise ValueError("expected state/meas [n,4] and cov/meas_cov [n,4,4]")
    out_state, out_cov, score, status = np.zeros((n, 4)), np.zeros((n, 4, 4)), np.zeros(n), np.zeros(n, np.int32)
    _check(lib.rs_kalman_track_planes(device, n, state.ctypes.data, cov.ctypes.data, meas.ctypes.data, meas_cov.ctypes.data,
                                      process_noise, out_state.ctypes.data, out_cov.ctypes.data, score.ctypes.data,
                                      status.ctypes.data), "rs_kalman_track_planes")
    return out_state, out_cov, score, status


def plane_match(w2c, det, det_first, det_xy, map_planes, map_first, map_xy, det_matched=None, advanced_search=False, device=0,
                sequential=False, return_matched=False):
    """MapPlane::find_matches (map_primitive.cpp:91-161) for every map plane of every frame at once.
    w2c [F,4,4]; det / map_planes: abi.polygon_plane_dtype records; det_first / map_first [F+1] offsets; det_xy / map_xy [V,2]
    polygon vertices. sequential: also the caller's loop (feature_map.hpp:652-669) - the map planes of a frame are served in
    order and a detection taken by one is unavailable to those after it. Returns (selected [n_map] = index of the matched
    detection inside its frame's list or -1, inter_area) and, with return_matched, the matched mask as the loop leaves it."""
    lib = load()
    w2c = np.ascontiguousarray(w2c, dtype=np.float64)
    det = np.ascontiguousarray(det, dtype=abi.polygon_plane_dtype)
    mp = np.ascontiguousarray(map_planes, dtype=abi.polygon_plane_dtype)
    det_first, map_first = np.ascontiguousarray(det_first, dtype=np.int32), np.ascontiguousarray(map_first, dtype=np.int32)
    det_xy, map_xy = np.ascontiguousarray(det_xy, dtype=np.float64), np.ascontiguousarray(map_xy, dtype=np.float64)
    n_frames = len(det_first) - 1
    if len(map_first) != n_frames + 1 or w2c.size != 16 * n_frames or det_first[-1] != len(det) or map_first[-1] != len(mp):
        raise ValueError("offset arrays do not describe the plane arrays")
    sel, inter = np.full(len(mp), -1, np.int32), np.zeros(len(mp))
    dm = None if det_matched is None else np.ascontiguousarray(det_matched, dtype=np.uint8)
    mout = np.zeros(max(len(det), 1), np.uint8)
    _check(lib.rs_plane_match(device, n_frames, w2c.ctypes.data, det.ctypes.data, det_first.ctypes.data, det_xy.ctypes.data,
                              mp.ctypes.data, map_first.ctypes.data, map_xy.ctypes.data, None if dm is None else dm.ctypes.data,
                              int(advanced_search), int(sequential), sel.ctypes.data, inter.ctypes.data,
                              mout.ctypes.data if return_matched else None), "rs_plane_match")
    return (sel, inter, mout[:len(det)]) if return_matched else (sel, inter)


def polygon_inter_area(a_rings, b_rings, device=0):
    """Polygon::inter_area (polygon.cpp:542-561) for pairs of rings given in a common 2-D frame: lists of [n,2] arrays."""
    lib = load()
    n = len(a_rings)
    if len(b_rings) != n:
        raise ValueError("need as many b rings as a rings")
    a_first = np.zeros(n + 1, np.int32)
    b_first = np.zeros(n + 1, np.int32)
    a_first[1:] = np.cumsum([len(r) for r in a_rings])
    b_first[1:] = np.cumsum([len(r) for r in b_rings])
    a_xy = np.ascontiguousarray(np.concatenate([np.asarray(r, np.float64).reshape(-1, 2) for r in a_rings]) if n else np.zeros((0, 2)))
    b_xy = np.ascontiguousarray(np.concatenate([np.asarray(r, np.float64).reshape(-1, 2) for r in b_rings]) if n else np.zeros((0, 2)))
    area = np.zeros(n)
    if a_xy.size == 0:
        a_xy = np.zeros((1, 2))
    if b_xy.size == 0:
        b_xy = np.zeros((1, 2))
    _check(lib.rs_polygon_inter_area(device, n, a_xy.ctypes.data, a_first.ctypes.data, b_xy.ctypes.data, b_first.ctypes.data,
                                     area.ctypes.data), "rs_polygon_inter_area")
    return area


def make_matches(n):
    return np.zeros((n,), dtype=abi.match_dtype)


class PoseOptimization:
    """Batched RANSAC + Levenberg-Marquardt pose solve (+ Monte-Carlo covariance) on one GPU."""

    def __init__(self, max_batch=1, max_matches=512, max_iterations=119, max_variance=100, device=0):
        lib = load()
        if not hasattr(lib, "rs_pose_create"):
            raise RsError("this build of librgbdslam_b200.so has no pose solver")
        self._lib = lib
        self.max_batch, self.max_matches = max_batch, max_matches
        self.max_iterations, self.max_variance = max_iterations, max_variance
        self._ctx = lib.rs_pose_create(max_batch, max_matches, max_iterations, max_variance, device)
        if not self._ctx:
            raise RsError("rs_pose_create failed: " + last_error())

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.rs_pose_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def options(max_iterations=0, n_variance=-1, rng_mode=abi.RS_RNG_REFERENCE, seed=0, intrinsics=None, lm_max_fev=0,
                sub_batches=0, worker_ctas_per_sm=0, solver=0):
        o = abi.PoseOpts()
        o.solver = solver                           # abi.RS_SOLVER_AUTO / _CHAIN / _FUSED / _WIDE
        o.worker_ctas_per_sm = worker_ctas_per_sm   # <= 0: as many resident CTAs per SM as fit
        o.max_iterations, o.n_variance, o.rng_mode, o.seed, o.lm_max_fev = max_iterations, n_variance, rng_mode, seed, lm_max_fev
        o.sub_batches = sub_batches   # RS_RNG_DEVICE: frame groups on separate streams (same results)
        if intrinsics is not None:
            o.fx, o.fy, o.cx, o.cy = intrinsics
        return o

    def compute_optimized_pose(self, cur_pose, matches, n_matches=None, opts=None):
        """cur_pose [B,7] (x y z qw qx qy qz); matches [B,max_matches] of abi.match_dtype (or a list of 1-D arrays).
        Returns (pose_out[B] structured array, inlier_mask[B,max_matches] uint8)."""
        self.compute_optimized_pose_begin(cur_pose, matches, n_matches, opts)
        return self.compute_optimized_pose_end()

    def compute_optimized_pose_begin(self, cur_pose, matches, n_matches=None, opts=None, out=None, mask=None):
        """Enqueues the solve and returns; compute_optimized_pose_end() waits and hands back (pose_out, inlier_mask).
        `out` / `mask` may be caller-provided (e.g. pinned) arrays; all host arrays must stay alive until _end."""
        cur_pose = np.ascontiguousarray(np.atleast_2d(cur_pose), dtype=np.float64)
        B = cur_pose.shape[0]
        if isinstance(matches, (list, tuple)):
            n_matches = np.array([len(m) for m in matches], dtype=np.int32)
            packed = np.zeros((B, self.max_matches), dtype=abi.match_dtype)
            for b, m in enumerate(matches):
                packed[b, :len(m)] = m
            matches = packed
        matches = np.ascontiguousarray(matches)
        if matches.ndim == 1:
            matches = matches[None]
        if matches.shape[1] != self.max_matches:
            raise ValueError("matches must be [B, max_matches=%d]" % self.max_matches)
        if n_matches is None:
            n_matches = np.full((B,), matches.shape[1], dtype=np.int32)
        n_matches = np.ascontiguousarray(n_matches, dtype=np.int32)
        if out is None:
            out = np.zeros((B,), dtype=abi.pose_out_dtype)
        if mask is None:
            mask = np.zeros((B, self.max_matches), dtype=np.uint8)
        o = opts if opts is not None else self.options()
        self._pending = (out, mask, cur_pose, matches, n_matches, o)  # keep the host buffers alive until _end
        _check(self._lib.rs_pose_solve_batched_begin(self._ctx, cur_pose.ctypes.data, matches.ctypes.data,
                                                     n_matches.ctypes.data, B, C.byref(o), out.ctypes.data, mask.ctypes.data),
               "rs_pose_solve_batched_begin")

    def compute_optimized_pose_end(self):
        if getattr(self, "_pending", None) is None:
            raise RuntimeError("compute_optimized_pose_end without a pending compute_optimized_pose_begin")
        _check(self._lib.rs_pose_solve_batched_end(self._ctx), "rs_pose_solve_batched_end")
        out, mask = self._pending[0], self._pending[1]
        self._pending = None
        return out, mask

    def upload(self, cur_pose, matches, n_matches):
        cur_pose = np.ascontiguousarray(cur_pose, dtype=np.float64)
        matches = np.ascontiguousarray(matches)
        n_matches = np.ascontiguousarray(n_matches, dtype=np.int32)
        _check(self._lib.rs_pose_upload(self._ctx, cur_pose.ctypes.data, matches.ctypes.data, n_matches.ctypes.data,
                                        cur_pose.shape[0]), "rs_pose_upload")

    def solve_device(self, batch, opts, stream=0):
        _check(self._lib.rs_pose_solve_device(self._ctx, batch, C.byref(opts), stream), "rs_pose_solve_device")

    def prepare_device(self, batch, opts, stream=0):
        """The preparation kernel of solve_device alone; the next solve_device(batch, ...) then launches the solve kernel only."""
        _check(self._lib.rs_pose_prepare_device(self._ctx, batch, C.byref(opts), stream), "rs_pose_prepare_device")

    def add_workers(self, ctas_per_sm=0, stream=0):
        """More CTAs for the solve kernel most recently launched through this context (they leave at once when no work is left)."""
        _check(self._lib.rs_pose_add_workers(self._ctx, ctas_per_sm, stream), "rs_pose_add_workers")

    def stream_wait_ransac(self, stream):
        """`stream` (cudaStream_t as int) waits for the RANSAC + final LM kernel of the latest solve of this context."""
        _check(self._lib.rs_pose_stream_wait_ransac(self._ctx, stream), "rs_pose_stream_wait_ransac")

    def download(self, batch):
        out = np.zeros((batch,), dtype=abi.pose_out_dtype)
        mask = np.zeros((batch, self.max_matches), dtype=np.uint8)
        _check(self._lib.rs_pose_download(self._ctx, batch, out.ctypes.data, mask.ctypes.data), "rs_pose_download")
        return out, mask

    def set_timing(self, n_slots):
        _check(self._lib.rs_pose_set_timing(self._ctx, n_slots), "rs_pose_set_timing")

    def kernel_ms(self, slot):
        """(prepare, solve kernel, its second half with RS_RNG_REFERENCE, 0) ms of the run that used this slot."""
        ms = (C.c_float * 4)()
        _check(self._lib.rs_pose_kernel_ms(self._ctx, slot, ms), "rs_pose_kernel_ms")
        return tuple(float(v) for v in ms)

    def phase_ms(self):
        """(RANSAC phase ms, whole solve kernel ms) of the most recent solve-kernel launch, from the device's own timer."""
        ms = (C.c_float * 2)()
        _check(self._lib.rs_pose_phase_ms(self._ctx, ms), "rs_pose_phase_ms")
        return float(ms[0]), float(ms[1])

    def work_counters(self):
        """Work counters of the most recent solve-kernel launch (see rs_pose_debug_counters)."""
        out = (C.c_uint64 * 8)()
        _check(self._lib.rs_pose_debug_counters(self._ctx, out), "rs_pose_debug_counters")
        names = ("hyp_first_cta", "hyp_helper_ctas", "bookkeeping_rounds", "hyp_applied", "helper_joins", "mc_tasks", "hyp_dropped", "reserved")
        return {k: int(v) for k, v in zip(names, out)}

    def frame_times(self, batch):
        """[batch, 4] ms since the solve kernel's first CTA: hypotheses started, stage closed, final LM done, covariance done."""
        ms = np.zeros((batch, 4))
        _check(self._lib.rs_pose_debug_frame_times(self._ctx, batch, ms.ctypes.data), "rs_pose_debug_frame_times")
        return ms

    def device_poses_ptr(self):
        return self._lib.rs_pose_device_poses(self._ctx)

    def export_random(self, batch, n_iterations, n_variance):
        subsets = np.full((batch, n_iterations, abi.RS_MAX_SUBSET), -1, dtype=np.int32)
        normals = np.zeros((batch, max(n_variance, 1), self.max_matches, 4), dtype=np.float64)
        _check(self._lib.rs_pose_export_random(self._ctx, batch, subsets.ctypes.data, normals.ctypes.data), "rs_pose_export_random")
        return subsets, normals
