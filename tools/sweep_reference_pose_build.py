"""Sweep of the pose oracle against the reference's own compiled pose-solve sources (oracle/_ref/libref_pose.so, see
oracle/ref_shim) on random problems: python tools/sweep_reference_pose_build.py [first index] [count]. CPU only; success flag,
inlier mask, pose and the 6 x 6 Monte-Carlo covariance must be bit-identical (tests/test_reference_pose_build.py::compare)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
import rgbd_slam_b200 as rs  # noqa: E402
import test_reference_pose_build as trp  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 256
assert ol.ref_pose_available(), "oracle/_ref/libref_pose.so is not built"
bad, solved, inliers, with2d = [], 0, 0, 0
for i in range(first, first + count):
    # every fourth problem carries inverse-depth (Point2dOptimizationFeature) matches beside points and planes
    if i % 4 == 3:
        _, guess, matches = rs.synth.pose_correspondences(i, n_points=150, n_planes=10, n_points2d=60, outlier_frac=0.05 * (i % 7))
        with2d += 1
    else:
        _, guess, matches = rs.synth.random_pose_problem(i)
    try:
        ok, n_in = trp.compare(guess, matches)
        solved += ok
        inliers += n_in
    except AssertionError as e:
        bad.append((i, str(e)[:120]))
print("problems %d (%d with point2d features): reference returned true on %d, %d inliers in all" % (count, with2d, solved, inliers))
print("problems on which the oracle differs from the compiled reference sources (flag / inlier mask / pose / covariance bits): %d" % len(bad))
for b in bad[:20]:
    print("  index %d: %s" % b)
