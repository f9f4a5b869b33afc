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

    def close(self):
        self.detector.close()
        self.solver.close()

    def track_batch(self, depth, cur_pose, matches, n_matches, seed=0, rng_mode=abi.RS_RNG_REFERENCE, n_frames_global=None):
        """This rank's shard: depth [F,H,W] float32, cur_pose [F,7], matches [F,max_matches], n_matches [F] (host).
        Returns (primitives dict, pose_out[F], inlier_mask[F,max_matches], all_poses [n_frames_global,7] tensor)."""
        F = len(depth)
        opts = self.solver.options(max_iterations=self.max_iterations, n_variance=self.n_variance, rng_mode=rng_mode,
                                   seed=seed, intrinsics=self.intrinsics)
        # the reference runs find_primitives on a std::async thread beside the rest of the frame (rgbd_slam.cpp:288-300):
        # the pose solve is enqueued first, the depth batch streams through the GPU meanwhile, then the solve is joined
        self.solver.compute_optimized_pose_begin(cur_pose, matches, n_matches, opts)
        prims = self.detector.find_primitives(depth, seed=seed)
        out, mask = self.solver.compute_optimized_pose_end()
        poses = torch.as_tensor(_DevicePtr(self.solver.device_poses_ptr(), (F, 7), "<f8"), device="cuda:%d" % self.device)
        total = F if n_frames_global is None else n_frames_global
        all_poses = sharding.gather_poses(poses, total)
        return prims, out, mask, all_poses


def poses_from_out(out):
    return np.ascontiguousarray(out["pose"], dtype=np.float64)
