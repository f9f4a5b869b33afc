"""Parity sweep of the pose solve on synth.random_pose_problem (random sizes, outlier fractions, poses, guesses, noise;
a third 'hard'): C-ABI (RS_RNG_REFERENCE) vs the CPU oracle with the same std::mt19937 stream.
Usage (GPU box): python tools/sweep_random_pose.py [first] [count] [hypotheses] [solver: auto|chain|fused|wide].
Prints every frame that differs."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
rs = importlib.import_module("rgb-d-slam_b200")
sys.modules.setdefault("rgbd_slam_b200", rs)
import oracle_lib as ol  # noqa: E402
import parity  # noqa: E402
import test_pose_gpu as tp  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
count = int(sys.argv[2]) if len(sys.argv) > 2 else 256
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 119
which = {"auto": rs.abi.RS_SOLVER_AUTO, "chain": rs.abi.RS_SOLVER_CHAIN, "fused": rs.abi.RS_SOLVER_FUSED,
         "wide": rs.abi.RS_SOLVER_WIDE}[sys.argv[4] if len(sys.argv) > 4 else "auto"]
M, B = 336, 32
solver = rs.PoseOptimization(max_batch=B, max_matches=M, max_iterations=max(iters, 257), max_variance=100)


def oracle_solve(cur, matches, seed):
    return ol.pose_solve(cur, matches, seed=seed, max_iterations=iters)


bad, illcond, stat = [], [], {"ok": 0, "failed": 0, "cov": 0}
for s0 in range(first, first + count, B):
    nb = min(B, first + count - s0)
    cur = np.zeros((nb, 7))
    matches = np.zeros((nb, M), dtype=rs.abi.match_dtype)
    n = np.zeros((nb,), np.int32)
    for b in range(nb):
        _, g, m = rs.synth.random_pose_problem(s0 + b)
        if which == rs.abi.RS_SOLVER_WIDE:
            m = m[m["type"] != rs.abi.RS_FEAT_POINT2D]   # the one-hypothesis-per-lane kernel knows points and planes only
        cur[b], n[b] = g, len(m)
        matches[b, :len(m)] = m
    opts = solver.options(seed=1234 + s0, rng_mode=rs.abi.RS_RNG_REFERENCE, max_iterations=iters, solver=which)
    out, mask = solver.compute_optimized_pose(cur, matches, n, opts)
    for b in range(nb):
        rout, rmask = oracle_solve(cur[b], matches[b][:n[b]], 1234 + s0 + b)
        stat["ok" if rout["status"] == 1 else "failed"] += 1
        stat["cov"] += int(rout["n_variance_ok"] > 0)
        try:
            tp.assert_out_match(rout, out[b], rmask, mask[b], n[b], cov_rtol=2e-2)
        except AssertionError as e:
            determined, why = parity.oracle_pose_is_determined(oracle_solve, cur[b], matches[b][:n[b]], 1234 + s0 + b)
            if not determined and out[b]["status"] == rout["status"]:
                illcond.append((s0 + b, int(n[b]), why))
            elif not determined:
                # an undetermined frame on which even the verdict differs: the oracle accepts a consensus of a handful of
                # features around a pose that moves by metres to kilometres under a 1e-12 input change, the device rejects it
                illcond.append((s0 + b, int(n[b]), "status %d vs oracle %d; %s" % (out[b]["status"], rout["status"], why)))
            else:
                bad.append((s0 + b, int(n[b]), str(e)[:160]))
solver.close()
print("frames %d: oracle ok %d, rejected %d, with covariance %d" % (count, stat["ok"], stat["failed"], stat["cov"]))
print("frames whose oracle answer is itself undetermined (parity.oracle_pose_is_determined): %d" % len(illcond))
for b in illcond[:8]:
    print("  frame %d (n=%d): %s" % b)
print("mismatching frames: %d" % len(bad))
for b in bad[:40]:
    print("  frame %d (n=%d): %s" % b)
