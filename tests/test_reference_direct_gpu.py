"""The CUDA path against the REFERENCE'S OWN SOURCES compiled here, in one hop (no oracle in between).

oracle/_ref/libref_cape.so and libref_pose.so are the reference's CAPE and pose-solve translation units, unmodified, compiled
against stand-in third-party headers (oracle/ref_shim; built in the container that has /root/reference, shipped to the GPU box
as files). These tests drive librgbdslam_b200.so through the C-ABI on the same inputs and require: every integer the reference
produces - per-cell point counts and planar flags, both label grids, the number of surviving planes, their boundary point
counts, the cylinder count; the pose solve's success flag and inlier mask - IDENTICAL, and the floating-point values within the
tolerances of north_star (normals n.n_ref >= 1 - 1e-8, d / pose 1e-4 relative, boundary points 1e-12).
Skipped when the libraries are absent."""
import numpy as np
import pytest

import oracle_lib as ol
import parity
import rgbd_slam_b200 as rs
from test_independent_cape import edge_cases

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (ol.ref_available() and ol.ref_pose_available()), reason="oracle/_ref libraries not built")]


@pytest.fixture(scope="module")
def det():
    d = rs.PrimitiveDetection(640, 480, 20, max_batch=8)
    yield d
    d.close()


def compare_cape(det, depth):
    ref = ol.ref_cape_run(depth)
    got = det.find_primitives(depth[None], seed=0)   # MAKE_DETERMINISTIC: engine seeded with 0, restarted per frame
    cells, info = got["cells"][0], got["info"][0]
    assert np.array_equal(ref["count"], cells["count"])
    assert np.array_equal(ref["planar"], cells["planar"])
    fitted = cells["planar"] == 1
    dots = np.sum(ref["cell"][fitted, :3] * cells["normal"][fitted], axis=-1)
    assert np.all(dots >= 1 - 1e-8)
    np.testing.assert_allclose(cells["d"][fitted], ref["cell"][fitted, 3], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(cells["mse"][fitted], ref["cell"][fitted, 4], rtol=1e-6, atol=1e-9)
    assert np.array_equal(ref["plane_grid"], got["plane_grid"][0])
    assert np.array_equal(ref["cyl_grid"], got["cyl_labels"][0])
    planes = got["planes"][0][:info["n_planes"]]
    final = planes[planes["is_final"] == 1]
    assert len(ref["planes"]) == len(final)
    cursor = 0
    for r, p in zip(ref["planes"], final):
        assert float(np.dot(r[:3], p["normal"])) >= 1 - 1e-8
        np.testing.assert_allclose(p["d"], r[3], rtol=1e-4)
        nb = int(r[5])
        assert nb == p["n_boundary"]
        want = got["boundary_xyz"][0][p["boundary_offset"]:p["boundary_offset"] + nb]
        np.testing.assert_allclose(want, ref["boundary"][cursor:cursor + nb], rtol=1e-12)
        cursor += nb
    cyls = got["cyls"][0][:info["n_cyl_regions"]]
    kept_axes = [c["axis"] for c in cyls for s in range(c["n_segments"]) if c["kept"][s]]
    assert len(ref["cyls"]) == len(kept_axes)
    for r, axis in zip(ref["cyls"], kept_axes):
        assert abs(float(np.dot(r[:3], axis))) >= 1 - 1e-8
    return info


def test_cape_scene_v0(det):
    for i in range(4):
        info = compare_cape(det, rs.synth.scene_v0_depth(i))
        assert info["n_final_planes"] >= 7 and info["n_cyl_regions"] >= 1


@pytest.mark.parametrize("name,depth", list(edge_cases()), ids=[n for n, _ in edge_cases()])
def test_cape_edge_cases(det, name, depth):
    compare_cape(det, depth)


def test_cape_random_rooms(det):
    planes = cylinders = 0
    for s in range(300, 324):
        info = compare_cape(det, rs.synth.random_scene_depth(s))
        planes += info["n_final_planes"]
        cylinders += info["n_cylinders"]
    assert planes > 30 and cylinders > 5


def test_rectify_depth(det):
    depth = rs.synth.scene_v0_depth(2)
    T = np.eye(4)
    T[:3, 3] = (25.0, -3.0, 4.0)
    c, s = np.cos(0.02), np.sin(0.02)
    T[:3, :3] = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    for ext in (np.eye(4), T):
        det.set_rectification(ext, enable=True)
        got = det.rectify_depth(depth)[0]
        det.set_rectification(None, enable=False)
        assert np.array_equal(got.view(np.uint32), ol.ref_rectify_depth(depth, ext).view(np.uint32))


def solve_one(solver, guess, matches, M):
    """One frame as frame 0 of its own batch with seed 0: RS_RNG_REFERENCE draws from std::mt19937(seed + frame)."""
    m = np.zeros((1, M), dtype=rs.abi.match_dtype)
    m[0, :len(matches)] = matches
    n = np.array([len(matches)], np.int32)
    out, mask = solver.compute_optimized_pose(guess[None], m, n, solver.options(seed=0, rng_mode=rs.abi.RS_RNG_REFERENCE))
    return out[0], mask[0][:len(matches)]


def compare_pose(solver, guess, matches, M, expect_ok=None):
    m = ol.stable_plane_normals(matches)
    ok, pose, cov, mask = ol.ref_pose_solve(guess, m)
    cur = ol.ref_pose_base(guess)
    out, gmask = solve_one(solver, cur, m, M)
    try:
        assert ok == (out["status"] == 1)
        if ok:
            assert np.array_equal(mask, gmask)
            close, dt, qd = parity.pose_close(pose, out["pose"])
            assert close, (dt, qd)
            # the 100-solve Monte-Carlo covariance: same Gaussian stream, so the same samples up to the LM's rounding; compared
            # as tests/test_pose_gpu.py does in this RNG mode, and only when no LM coefficient is ~0 (there the reference's own
            # forward-difference step underflows and its covariance is rounding noise)
            d = 1.0 / max(1.0 + pose[6], 0.001)
            coeff = np.array([pose[0], pose[1], pose[2], pose[3] * d, pose[4] * d, pose[5] * d])
            gc = out["cov"].reshape(6, 6)
            assert np.all(np.isfinite(gc)) and np.allclose(gc, gc.T)
            if np.abs(coeff).min() > 1e-3:
                scale = np.sqrt(np.outer(np.diag(cov), np.diag(cov)))
                assert (np.abs(gc - cov) / (2e-2 * scale + 1e-12)).max() <= 1.0
    except AssertionError:
        # frames whose answer the reference algorithm itself does not determine (parity.oracle_pose_is_determined)
        determined, why = parity.oracle_pose_is_determined(ol.pose_solve, cur, m, 0)
        if determined:
            # ... or whose winning hypothesis (the device's or the reference's) is an LM cut off by the evaluation budget
            rout, _ = ol.pose_solve(cur, m, seed=0)
            if parity.winning_hypotheses_converged(ol, cur, m, 0, (int(out["best_iteration"]), int(rout["best_iteration"]))):
                raise
        return None
    if expect_ok is not None:
        assert ok == expect_ok
    return ok


def test_pose_reference_scenarios():
    """The 40 scenarios of the reference's tests/test_pose_optimization.cpp, device against the compiled reference."""
    import ref_scenarios as scn
    M = 160
    solver = rs.PoseOptimization(1, M)
    undetermined = 0
    for scenario in scn.SCENARIOS:
        _, guess, matches = scn.build(scenario)
        if compare_pose(solver, guess, matches, M) is None:
            undetermined += 1
    solver.close()
    assert undetermined <= 2, undetermined   # the two multi*_100PercentOutliers scenarios are seed-dependent by construction


def test_pose_random_problems():
    M = 400
    solver = rs.PoseOptimization(1, M)
    solved = undetermined = 0
    for i in range(2000, 2032):
        if i % 4 == 3:
            _, guess, matches = rs.synth.pose_correspondences(i, n_points=150, n_planes=10, n_points2d=60, outlier_frac=0.05 * (i % 7))
        else:
            _, guess, matches = rs.synth.random_pose_problem(i)
        r = compare_pose(solver, guess, matches, M)
        undetermined += r is None
        solved += bool(r)
    solver.close()
    assert solved >= 16 and undetermined <= 3, (solved, undetermined)
