"""The pose oracle against the REFERENCE'S OWN POSE-SOLVE SOURCES, compiled here.

oracle/ref_shim builds /root/reference/src/pose_optimization/{pose_optimization, levenberg_marquardt_functors}.cpp, utils/{pose,
camera_transformation}.cpp, the coordinate classes, ransac.hpp / random.hpp - where they lie, unmodified - and the three
optimisation-feature classes cut out of map_management/map_features/map_{point, primitive, point2d}.{hpp, cpp} at build time, into
oracle/_ref/libref_pose.so, against the stand-in headers for Eigen (Eigen::LevenbergMarquardt is the oracle's restated MINPACK
lmdif: third-party code, pinned against scipy's MINPACK elsewhere). What this pins is everything the reference wrote around the
LM: the RANSAC loop with its early stop and overload rule, the std::shuffle subset draws from the thread-local mt19937 (seed 0
under MAKE_DETERMINISTIC), the residual functor, the three is_inlier tests, the final re-optimisation, the per-feature random
variations and the 100-solve Monte-Carlo covariance - BIT FOR BIT: status, inlier mask, pose and the 6 x 6 covariance.

Inputs are made independent of how many times the reference copies (= re-normalises) a value on its way in: the oracle gets the
start pose as the one utils::PoseBase the reference entry builds holds it, plane normals sit on a fixed point of x / |x|
(oracle_lib.stable_plane_normals).
Skipped when the library has not been built (no /root/reference at build time)."""
import numpy as np
import pytest

import oracle_lib as ol
import ref_scenarios as scn
import rgbd_slam_b200 as rs

pytestmark = pytest.mark.skipif(not ol.ref_pose_available(), reason="oracle/_ref/libref_pose.so not built (needs /root/reference)")


def compare(guess, matches, expect_ok=None):
    m = ol.stable_plane_normals(matches)
    ok, pose, cov, mask = ol.ref_pose_solve(guess, m)                  # builds one PoseBase from the guess
    out, omask = ol.pose_solve(ol.ref_pose_base(guess), m, seed=0)     # RS_RNG_REFERENCE semantics: one mt19937 stream from seed 0
    assert ok == (out["status"] == 1)
    if expect_ok is not None:
        assert ok == expect_ok
    if ok:
        assert np.array_equal(mask, omask)
        assert np.array_equal(pose, out["pose"])
        assert np.array_equal(cov.ravel(), out["cov"])
    return ok, int(mask.sum())


@pytest.mark.parametrize("scenario", scn.SCENARIOS, ids=[s[0] for s in scn.SCENARIOS])
def test_reference_scenarios(scenario):
    """The 40 scenarios of the reference's tests/test_pose_optimization.cpp."""
    _, guess, matches = scn.build(scenario)
    compare(guess, matches, expect_ok=True)


@pytest.mark.parametrize("first", [0, 8, 16])
def test_random_problems(first):
    """Frames of the synthetic workload: 300 points + 20 planes with 10 % outliers, varying guesses."""
    solved = 0
    for i in range(first, first + 8):
        _, guess, matches = rs.synth.random_pose_problem(i)
        ok, _ = compare(guess, matches)
        solved += ok
    assert solved >= 4   # the generator mixes in hopeless frames on purpose


def test_point2d_features():
    """Point2dOptimizationFeature (inverse-depth points, the "line" residual) mixed with points and planes."""
    for i in range(4):
        _, guess, matches = rs.synth.pose_correspondences(i, n_points=120, n_planes=8, n_points2d=60)
        assert (matches["type"] == rs.abi.RS_FEAT_POINT2D).sum() == 60
        ok, _ = compare(guess, matches)
        assert ok


def test_failures_agree():
    """Too few features / hopeless outlier sets: the reference returns false exactly where the oracle reports a failure."""
    _, guess, matches = rs.synth.pose_correspondences(3, n_points=4, n_planes=0)
    compare(guess, matches, expect_ok=False)
    _, guess, matches = rs.synth.pose_correspondences(5, n_points=40, n_planes=0, outlier_frac=0.95)
    compare(guess, matches)


def test_residual_functor_and_inlier_tests():
    """Global_Pose_Estimator::operator() and IOptimizationFeature::is_inlier alone, at perturbed coefficient vectors."""
    rng = np.random.default_rng(5)
    for i in range(6):
        truth, guess, matches = rs.synth.pose_correspondences(i, n_points=50, n_planes=6, n_points2d=20)
        m = ol.stable_plane_normals(matches)
        ok, pose = ol.ref_pose_lm(guess, m)
        assert ok
        # inlier tests of every feature under the reference's own LM result and under the true pose
        for p in (pose, truth):
            want = ol.ref_pose_inliers(p, m)
            got = ol.pose_inliers(p, m)
            assert np.array_equal(want, got)
        x = np.concatenate([pose[:3] + rng.normal(0, 3.0, 3), rng.normal(0, 0.3, 3)])
        assert np.array_equal(ol.ref_pose_residuals(m, x), ol.pose_residuals(m, x))
