"""Batched per-frame hot path of RGBD_SLAM::track (src/rgbd_slam.cpp:99-125,163-259) on one GPU per process:
depth -> CAPE planes/cylinders (find_primitives) and (current pose, matches) -> optimised pose + covariance
(compute_optimized_pose), frames sharded over the ranks and the poses all-gathered (sharding.py).
Feature matching / map update (Local_Map) sit between the two calls in the reference and stay on the host
(out of scope); this class therefore takes the match lists as an input."""
import numpy as np
import torch

from . import abi, sharding
from .lib import PoseOptimization, PrimitiveDetection


class _DevicePtr:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3}


class FramePipeline:
    def __init__(self, width=640, height=480, cell_px=20, intrinsics=(550.0, 550.0, 320.0, 240.0), max_frames=32,
                 max_matches=320, max_iterations=119, n_variance=100, device=0):
        self.device = device
        self.max_frames, self.max_matches = max_frames, max_matches
        self.intrinsics = tuple(intrinsics)
        self.detector = PrimitiveDetection(width, height, cell_px, *intrinsics, max_batch=max_frames, device=device)
        self.solver = PoseOptimization(max_batch=max_frames, max_matches=max_matches, max_iterations=max_iterations,
                                       max_variance=max(n_variance, 1), device=device)
        self.max_iterations, self.n_variance = max_iterations, n_variance
        # Pinned staging owned by the pipeline: cudaMemcpyAsync from / into pageable memory blocks the host until the
        # copy (and everything before it on the stream) is done, which would serialise the pose solve and the depth
        # upload that track_batch overlaps. Inputs are copied in here, results land here and are handed out as copies.
        self._wanted = ("cells", "plane_grid", "plane_labels", "cyl_labels", "cyl_region_seg", "planes", "cyls", "boundary_xyz", "info")
        self._pin = {}
        self._h_depth = self._pinned("depth", (max_frames, height, width), np.float32)
        self._h_cur = self._pinned("cur", (max_frames, 7), np.float64)
        self._h_matches = self._pinned("matches", (max_frames, max_matches), abi.match_dtype)
        self._h_n = self._pinned("n", (max_frames,), np.int32)
        self._h_out = self._pinned("out", (max_frames,), abi.pose_out_dtype)
        self._h_mask = self._pinned("mask", (max_frames, max_matches), np.uint8)
        arrs, _ = abi.alloc_cape_outputs(max_frames, self.detector.n_cells, self.detector.max_boundary)
        self._h_prims = {k: self._pinned("prims_" + k, arrs[k].shape, arrs[k].dtype) for k in self._wanted}

    def _pinned(self, name, shape, dtype):
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        t = torch.empty(max(nbytes, 1), dtype=torch.uint8).pin_memory()
        self._pin[name] = t   # keeps the allocation alive
        return t.numpy()[:nbytes].view(dtype).reshape(shape)

    def close(self):
        self.detector.close()
        self.solver.close()
        self._pin.clear()

    def track_batch(self, depth, cur_pose, matches, n_matches, seed=0, rng_mode=abi.RS_RNG_REFERENCE, n_frames_global=None):
        """This rank's shard: depth [F,H,W] float32, cur_pose [F,7], matches [F,max_matches], n_matches [F] (host).
        Returns (primitives dict, pose_out[F], inlier_mask[F,max_matches], all_poses [n_frames_global,7] tensor)."""
        F = len(depth)
        if F > self.max_frames:
            raise ValueError("batch of %d frames exceeds max_frames=%d" % (F, self.max_frames))
        if isinstance(matches, (list, tuple)):
            n_matches = np.array([len(m) for m in matches], dtype=np.int32)
            self._h_matches[:F] = np.zeros((), dtype=abi.match_dtype)
            for b, m in enumerate(matches):
                self._h_matches[b, :len(m)] = m
        else:
            self._h_matches[:F] = matches
        self._h_depth[:F] = depth
        self._h_cur[:F] = cur_pose
        self._h_n[:F] = n_matches
        opts = self.solver.options(max_iterations=self.max_iterations, n_variance=self.n_variance, rng_mode=rng_mode,
                                   seed=seed, intrinsics=self.intrinsics)
        # the reference runs find_primitives on a std::async thread beside the rest of the frame (rgbd_slam.cpp:288-300):
        # the pose solve is enqueued first, the depth batch streams through the GPU meanwhile, then the solve is joined
        self.solver.compute_optimized_pose_begin(self._h_cur[:F], self._h_matches[:F], self._h_n[:F], opts,
                                                 out=self._h_out[:F], mask=self._h_mask[:F])
        views = {k: v[:F] for k, v in self._h_prims.items()}
        st = abi.CapeOutputs(**{k: v.ctypes.data for k, v in views.items()})
        self.detector.find_primitives(self._h_depth[:F], seed=seed, out=(views, st))
        out, mask = self.solver.compute_optimized_pose_end()
        prims = {k: v.copy() for k, v in views.items()}
        out, mask = out.copy(), mask.copy()
        poses = torch.as_tensor(_DevicePtr(self.solver.device_poses_ptr(), (F, 7), "<f8"), device="cuda:%d" % self.device)
        total = F if n_frames_global is None else n_frames_global
        all_poses = sharding.gather_poses(poses, total)
        return prims, out, mask, all_poses


def poses_from_out(out):
    return np.ascontiguousarray(out["pose"], dtype=np.float64)
