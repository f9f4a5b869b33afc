"""Small run of the round-2 kernels for `compute-sanitizer --tool racecheck` (GPU box): the one-hypothesis-per-lane solver
(pose_hypotheses_kernel, pose_fold_kernel, the chain's final-only RANSAC kernel), the CTA-batched Monte-Carlo kernel and plane
matching. The chain's full RANSAC kernel is left out on purpose: its ring of hypothesis results is handed between warps through
volatile flags, fences and an atomic lock, which racecheck (a barrier-based checker) reports as hazards by design."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rgbd_slam_b200 as rs  # noqa: E402

M = 320
truth, cur, matches, n = rs.synth.pose_batch(900, 3, M, outlier_frac=0.3)
n[2] = 150
s = rs.PoseOptimization(max_batch=3, max_matches=M, max_iterations=512, max_variance=100)
out, mask = s.compute_optimized_pose(cur, matches, n, s.options(max_iterations=96, seed=1, rng_mode=rs.abi.RS_RNG_DEVICE,
                                                                solver=rs.abi.RS_SOLVER_WIDE, n_variance=24))
print("wide solver:", out["status"], out["iterations_run"], out["n_variance_ok"])
s.close()
pm = rs.synth.plane_match_problem(3, n_frames=2)
sel, inter = rs.plane_match(*pm[:-1], det_matched=pm[-1])
print("plane matching:", sel)
