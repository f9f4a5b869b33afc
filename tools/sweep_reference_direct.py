"""Sweep of the CUDA path against the reference's own compiled sources (oracle/_ref, see oracle/ref_shim), no oracle in between:
python tools/sweep_reference_direct.py [first] [rooms] [pose problems]   (GPU box). Per random room: every integer of
find_primitives identical, values within tolerance (tests/test_reference_direct_gpu.py::compare_cape); per pose problem:
success flag and inlier mask identical, pose / covariance within tolerance (::compare_pose; frames whose answer the reference
algorithm itself does not determine are counted apart)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
import rgbd_slam_b200 as rs  # noqa: E402
import test_reference_direct_gpu as trd  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
rooms = int(sys.argv[2]) if len(sys.argv) > 2 else 256
problems = int(sys.argv[3]) if len(sys.argv) > 3 else 256
assert ol.ref_available() and ol.ref_pose_available(), "oracle/_ref libraries are not built"

det = rs.PrimitiveDetection(640, 480, 20, max_batch=8)
bad, planes, cyls = [], 0, 0
for s in range(first, first + rooms):
    try:
        info = trd.compare_cape(det, rs.synth.random_scene_depth(s))
        planes += int(info["n_final_planes"])
        cyls += int(info["n_cylinders"])
    except AssertionError as e:
        bad.append((s, str(e)[:120]))
det.close()
print("CAPE, CUDA path vs compiled reference sources: rooms %d (final planes %d, cylinders %d), rooms that differ: %d" % (rooms, planes, cyls, len(bad)))
for b in bad[:20]:
    print("  seed %d: %s" % b)

M = 400
solver = rs.PoseOptimization(1, M)
bad, solved, undetermined, with2d = [], 0, 0, 0
for i in range(first, first + problems):
    if i % 4 == 3:
        _, guess, matches = rs.synth.pose_correspondences(i, n_points=150, n_planes=10, n_points2d=60, outlier_frac=0.05 * (i % 7))
        with2d += 1
    else:
        _, guess, matches = rs.synth.random_pose_problem(i)
    try:
        r = trd.compare_pose(solver, guess, matches, M)
        undetermined += r is None
        solved += bool(r)
    except AssertionError as e:
        bad.append((i, str(e)[:120]))
solver.close()
print("pose solve (RS_RNG_REFERENCE), CUDA path vs compiled reference sources: problems %d (%d with point2d features), solved %d, "
      "undetermined by the reference algorithm itself %d, problems that differ: %d" % (problems, with2d, solved, undetermined, len(bad)))
for b in bad[:20]:
    print("  index %d: %s" % b)
