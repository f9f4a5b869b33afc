"""rgb-d-slam_b200 — B200-native (sm_100a CUDA) hot path of BaptisteHudyma/RGB-D-SLAM:
CAPE depth-cell plane/cylinder segmentation + RANSAC/Levenberg-Marquardt pose solve, behind a C-ABI
(include/rgbdslam_b200.h). The directory name contains hyphens; import it through the root shim `rgbd_slam_b200`."""
from . import abi, sharding, synth  # noqa: F401
from .pipeline import FramePipeline  # noqa: F401
from .lib import (PoseOptimization, PrimitiveDetection, RsError, kalman_track_planes, kalman_track_points, last_error,  # noqa: F401
                  launch_count, load, make_matches, plane_match, polygon_inter_area)

__all__ = ["abi", "sharding", "synth", "FramePipeline", "PrimitiveDetection", "PoseOptimization", "RsError", "load", "last_error", "launch_count",
           "make_matches", "kalman_track_points", "kalman_track_planes", "plane_match", "polygon_inter_area"]
