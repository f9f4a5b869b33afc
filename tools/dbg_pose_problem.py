"""Prints the device's and the oracle's answers for one problem of tools/sweep_reference_direct.py (GPU box):
python tools/dbg_pose_problem.py <index>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, oracle_lib as ol, rgbd_slam_b200 as rs, parity
import test_reference_direct_gpu as trd
i = int(sys.argv[1])
if i % 4 == 3:
    _, guess, matches = rs.synth.pose_correspondences(i, n_points=150, n_planes=10, n_points2d=60, outlier_frac=0.05 * (i % 7))
else:
    _, guess, matches = rs.synth.random_pose_problem(i)
m = ol.stable_plane_normals(matches)
cur = ol.ref_pose_base(guess)
rout, rmask, taps = ol.pose_solve(cur, m, seed=0, taps=True)
M = 400
for name, sv in (("chain", rs.abi.RS_SOLVER_CHAIN), ("fused", rs.abi.RS_SOLVER_FUSED)):
    s = rs.PoseOptimization(1, M)
    mm = np.zeros((1, M), dtype=rs.abi.match_dtype); mm[0, :len(m)] = m
    out, mask = s.compute_optimized_pose(cur[None], mm, np.array([len(m)], np.int32), s.options(seed=0, rng_mode=rs.abi.RS_RNG_REFERENCE, solver=sv))
    out, mask = out[0], mask[0][:len(m)]
    print(name, "status", out["status"], rout["status"], "iters", out["iterations_run"], rout["iterations_run"], "best", out["best_iteration"], rout["best_iteration"],
          "inliers", out["n_inliers"], rout["n_inliers"], "score", out["score"], rout["score"], "mask diff at", np.nonzero(mask != rmask)[0],
          "pose close", parity.pose_close(rout["pose"], out["pose"]))
    s.close()
