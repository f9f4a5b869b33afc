"""CPU: pins of the pose oracle.
(1) the 40 scenarios of the reference's own tests/test_pose_optimization.cpp with the reference's tolerances;
(2) the restated Eigen LevenbergMarquardt (MINPACK lmdif) against scipy.optimize.leastsq (an independent MINPACK);
(3) the frame conventions against a numpy restatement; (4) the committed golden fixture."""
import os

import numpy as np
import pytest
from scipy.optimize import leastsq

import oracle_lib as ol
import ref_scenarios as scn
import rgbd_slam_b200 as rs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
K = np.array([550.0, 550.0, 320.0, 240.0])


@pytest.mark.parametrize("scenario", scn.SCENARIOS, ids=[s[0] for s in scn.SCENARIOS])
def test_reference_scenarios(scenario):
    truth, guess, feats = scn.build(scenario)
    seeds = range(5)
    failures = []
    for seed in seeds:
        out, _ = ol.pose_solve(guess, feats, seed=seed)
        err = "status %d" % out["status"] if out["status"] != 1 else scn.check_reference_tolerance(truth, out["pose"])
        if err:
            failures.append((seed, err))
    if scenario[0].endswith("100PercentOutliers") and scenario[0].startswith("multi"):
        # 64 of the 136 matches are outliers whose observations all sit within 2 px of pixel (0,0)
        # (vector2::Random(), test_pose_optimization.cpp:136): a far-away pose that projects everything there scores
        # as many inliers as the true pose, so the reference's randomised test is seed-dependent by construction.
        assert len(failures) <= 2, failures
    else:
        assert not failures, failures


def _residual_fn(feats):
    lib = ol.load()
    m = lib.orc_residual_count(feats.ctypes.data, len(feats))

    def f(x):
        out = np.zeros(m)
        xx = np.ascontiguousarray(x, dtype=np.float64)
        lib.orc_pose_residuals(K.ctypes.data, feats.ctypes.data, len(feats), xx.ctypes.data, out.ctypes.data)
        return out
    return f, m


@pytest.mark.parametrize("guess_factor,expected_nfev", [(1.0, 8), (0.9, 29), (0.5, 36), (0.1, 43)])
def test_lm_matches_minpack(guess_factor, expected_nfev):
    """Same exit code, same iterate count and same minimiser as MINPACK lmdif (scipy) on the reference's cube scene.
    Eigen's NumericalDiff re-evaluates f(x) once per Jacobian, so its nfev = MINPACK's + (number of Jacobians)."""
    lib = ol.load()
    truth = scn.pose7(scn.PE, scn.RE)
    guess = scn.pose7([guess_factor * v for v in scn.PE], [guess_factor * v for v in scn.RE])
    feats = ol.ref_test_features(truth, 5.0, 0.0, 5.0, -1.0)
    f, m = _residual_fn(feats)
    x0 = np.zeros(6)
    lib.orc_pose_coefficients(guess.ctypes.data, x0.ctypes.data)
    xs, _, info, _, ier = leastsq(f, x0.copy(), full_output=True, maxfev=400)
    x = x0.copy()
    import ctypes as C
    nfev = C.c_int(0)
    status = lib.orc_pose_lm(K.ctypes.data, feats.ctypes.data, len(feats), x.ctypes.data, 400, C.byref(nfev))
    # the cube scene is a zero-residual problem (the noisy world points are the ones projected): both stop on xtol
    assert status == ier == 2
    assert info["nfev"] == expected_nfev
    n_jacobians = (info["nfev"] - 1) // 7  # each MINPACK outer iteration: 6 Jacobian evaluations + 1 trial here
    assert nfev.value == info["nfev"] + n_jacobians
    np.testing.assert_allclose(x, xs, rtol=0, atol=1e-9)


def test_transforms_match_numpy():
    lib = ol.load()
    rng = np.random.default_rng(0)
    for _ in range(20):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        pose = np.concatenate([rng.uniform(-500, 500, 3), q])
        w2c, pw2c = np.zeros(16), np.zeros(16)
        lib.orc_world_to_camera(pose.ctypes.data, w2c.ctypes.data, pw2c.ctypes.data)
        c2w = rs.synth.camera_to_world(pose)
        np.testing.assert_allclose(w2c.reshape(4, 4), np.linalg.inv(c2w), atol=1e-9)
        # plane world->camera = (c2w)^T restricted as in camera_transformation.cpp:52-71
        Rp, tp = c2w[:3, :3], c2w[:3, 3]
        M = np.eye(4)
        M[:3, :3] = Rp.T
        M[3, :3] = tp
        np.testing.assert_allclose(pw2c.reshape(4, 4), M, atol=1e-9)


def test_coefficient_round_trip():
    lib = ol.load()
    rng = np.random.default_rng(1)
    for _ in range(50):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        if q[3] < -0.9:
            continue
        pose = np.concatenate([rng.uniform(-100, 100, 3), q])
        x, back, v6 = np.zeros(6), np.zeros(7), np.zeros(6)
        lib.orc_pose_coefficients(pose.ctypes.data, x.ctypes.data)
        lib.orc_pose_from_coefficients(x.ctypes.data, back.ctypes.data, v6.ctypes.data)
        np.testing.assert_allclose(back, pose, atol=1e-12)


def test_camera_axes_convention():
    """C = [[0,0,1],[-1,0,0],[0,-1,0]]: the camera looks along world +x at identity (camera_transformation.cpp:11-17)."""
    c2w = rs.synth.camera_to_world(np.array([0, 0, 0, 1.0, 0, 0, 0]))
    np.testing.assert_allclose(c2w[:3, :3] @ np.array([0, 0, 1.0]), [1, 0, 0], atol=1e-12)
    lib = ol.load()
    w2c, pw = np.zeros(16), np.zeros(16)
    pose = np.array([0, 0, 0, 1.0, 0, 0, 0])
    lib.orc_world_to_camera(pose.ctypes.data, w2c.ctypes.data, pw.ctypes.data)
    np.testing.assert_allclose(w2c.reshape(4, 4)[:3, :3] @ np.array([1.0, 0, 0]), [0, 0, 1], atol=1e-12)


def test_ransac_constants():
    assert ol.load().orc_ransac_default_iterations() == 119  # pose_optimization.cpp:129-132


def test_synthetic_workload_and_golden():
    g = np.load(os.path.join(GOLDEN, "pose_synth_v0.npz"))
    for f in range(4):
        truth, guess, m = rs.synth.pose_correspondences(f)
        out, mask = ol.pose_solve(guess, m, seed=f)
        assert out["status"] == g["status"][f] == 1
        assert out["n_inliers"] == g["n_inliers"][f]
        assert out["iterations_run"] == g["iterations_run"][f]
        assert out["best_iteration"] == g["best_iteration"][f]
        np.testing.assert_allclose(out["pose"], g["pose"][f], rtol=0, atol=1e-9)
        np.testing.assert_allclose(out["cov"], g["cov"][f], rtol=1e-6, atol=1e-12)
        assert np.array_equal(mask, g["mask"][f])
        # 270 true point matches + 18 true planes must be the inliers, the appended outliers must not
        assert mask[:270].sum() >= 260 and mask[270:300].sum() == 0 and mask[318:].sum() == 0
        assert np.linalg.norm(out["pose"][:3] - truth[:3]) < 3.0


def test_explicit_random_inputs_are_honoured():
    """Feeding back the subsets the oracle drew must reproduce the run (this is how the RS_RNG_DEVICE parity test
    hands the library's on-device draws to the oracle)."""
    truth, guess, m = rs.synth.pose_correspondences(7)
    out, mask, taps = ol.pose_solve(guess, m, seed=3, n_variance=0, taps=True)
    out2, mask2 = ol.pose_solve(guess, m, seed=99, n_variance=0, subsets=taps["subsets"])
    assert out2["iterations_run"] == out["iterations_run"] and out2["best_iteration"] == out["best_iteration"]
    np.testing.assert_array_equal(out2["pose"], out["pose"])
    assert np.array_equal(mask, mask2)


def test_failure_modes():
    truth, guess, m = rs.synth.pose_correspondences(0)
    out, _ = ol.pose_solve(guess, m[:4])                     # 4 points: score 0.8 < 1 -> RANSAC refuses
    assert out["status"] == 0
    np.testing.assert_array_equal(out["pose"], guess)
    bad = m.copy()
    bad["map"][5, 0] = np.nan
    out, _ = ol.pose_solve(guess, bad)                       # invalid feature -> compute_optimized_pose returns false
    assert out["status"] == 0


def test_lm_matches_minpack_noisy_mixed():
    """Non-zero-residual problem (0.5 px noise, points + planes): same exit code as MINPACK, same minimiser."""
    lib = ol.load()
    import ctypes as C
    truth, guess, m = rs.synth.pose_correspondences(0)
    feats = np.concatenate([m[:270], m[300:318]])
    f, _ = _residual_fn(feats)
    x0 = np.zeros(6)
    lib.orc_pose_coefficients(guess.ctypes.data, x0.ctypes.data)
    xs, _, info, _, ier = leastsq(f, x0.copy(), full_output=True, maxfev=400)
    x = x0.copy()
    nfev = C.c_int(0)
    status = lib.orc_pose_lm(K.ctypes.data, feats.ctypes.data, len(feats), x.ctypes.data, 400, C.byref(nfev))
    assert status == ier
    assert nfev.value == info["nfev"] + (info["nfev"] - 1) // 7 or abs(nfev.value - info["nfev"]) <= 8
    np.testing.assert_allclose(x, xs, rtol=0, atol=2e-6)
