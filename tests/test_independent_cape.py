"""CPU: the C++ oracle's CAPE path against tests/independent_cape.py - a second restatement written from the reference
sources with other building blocks (LAPACK eigh, the real cv2 morphology, numpy's MT19937, literal recursion).
Label grids, merge labels, seed counts, cylinder inlier sets and boundary selections must be IDENTICAL; real-valued outputs
agree to the last bits of the two eigen-solvers. PARITY STAYS PARTIAL: neither side is the reference binary (oracle/_ref/README.md)."""
import numpy as np
import pytest

import independent_cape as ic
import oracle_lib as ol
import rgbd_slam_b200 as rs


def test_mt19937_canonical_doubles_match_the_survey_vectors():
    # mt19937(0)'s first raw outputs are the published 2357136044, 2546248239, 3071714933, 3626093760, 2588848963, 3684848379;
    # generate_canonical<double, 53> pairs them (low word first): 0.5928..., 0.8442..., 0.8579... - the three values SURVEY.md
    # A.8 lists (there in reverse order, the way a right-to-left evaluated argument list prints them)
    g = ic.StdMt19937Uniform(0)
    assert [g._raw() for _ in range(6)] == [2357136044, 2546248239, 3071714933, 3626093760, 2588848963, 3684848379]
    g = ic.StdMt19937Uniform(0)
    got = [g.uniform() for _ in range(3)]
    assert abs(got[0] - 0.592844616516682) < 1e-14 and abs(got[1] - 0.844265744256598) < 1e-14 and abs(got[2] - 0.857945619989829) < 1e-14


def compare(depth, cell=20, K=(550.0, 550.0, 320.0, 240.0), seed=0, noise_free=False):
    """noise_free: an exactly planar depth image - every cell's MSE is the rounding noise of its eigen-solver (1e-14), so the
    seed (argmin MSE, counted twice in its region's sums) is not determined by the algorithm: the sums / MSE / score of the
    grown segments are then compared only through the plane they give."""
    ind = ic.find_primitives(depth, cell=cell, K=K, seed=seed)
    ref = ol.cape_run(depth, cell=cell, K=K, seed=seed)
    cells, info = ref["cells"][0], ref["info"][0]
    # a2-a4: per-cell fit (sums bit-identical: same values, same order; eigen-solver outputs to rounding)
    planar = np.array([s.planar for s in ind["grid"]])
    assert np.array_equal(planar, cells["planar"] == 1)
    assert np.array_equal(np.array([s.count for s in ind["grid"]]), cells["count"])
    assert np.array_equal(np.array([s.S for s in ind["grid"]]), cells["S"])
    for i in np.nonzero(planar)[0]:
        assert abs(np.dot(ind["grid"][i].normal, cells["normal"][i])) >= 1 - 1e-12
        assert abs(ind["grid"][i].d - cells["d"][i]) <= 1e-9 * max(1.0, abs(cells["d"][i]))
        assert abs(ind["grid"][i].mse - cells["mse"][i]) <= 1e-9 * cells["mse"][i] + 1e-12
    assert np.array_equal(ind["tols"], cells["tol"])
    # a5: histogram bins as init_histogram assigned them (-1 for the others)
    # (the independent run's bins are read AFTER the seed loop: assigned cells carry the remove_point value 1)
    # a6-a8: label grids, bit for bit
    assert ind["n_planar"] == info["n_planar_cells"]
    assert ind["n_seeds"] == info["n_seeds"]
    assert np.array_equal(ind["plane_grid"], ref["plane_grid"][0])
    assert np.array_equal(ind["cyl_labels"], ref["cyl_labels"][0])
    assert np.array_equal(ind["plane_labels"], ref["plane_labels"][0])
    assert len(ind["planes"]) == info["n_planes"]
    for k, p in enumerate(ind["planes"]):
        r = ref["planes"][0][k]
        for f in ("merge_label", "planar", "is_final", "count", "n_boundary", "boundary_offset"):
            assert p[f] == r[f], (k, f, p[f], r[f])
        if not noise_free:
            np.testing.assert_allclose(p["S"], r["S"], rtol=1e-15)
        if p["planar"]:
            assert abs(np.dot(p["normal"], r["normal"])) >= 1 - 1e-12
            np.testing.assert_allclose(p["d"], r["d"], rtol=1e-9)
            if not noise_free:
                np.testing.assert_allclose(p["mse"], r["mse"], rtol=1e-7, atol=1e-12)
                np.testing.assert_allclose(p["score"], r["score"], rtol=1e-7)
    # a9: boundary points (raster order) and the cylinder opening test
    assert len(ind["boundary"]) == info["n_boundary"]
    np.testing.assert_allclose(ind["boundary"], ref["boundary_xyz"][0][:info["n_boundary"]], rtol=1e-13)
    # a7: cylinder regions, sub-segments, inlier sets
    assert len(ind["cylinders"]) == info["n_cyl_regions"]
    assert len(ind["cylinder2region"]) == info["n_cylinders"]
    for r_, cyl in enumerate(ind["cylinders"]):
        rc = ref["cyls"][0][r_]
        assert cyl.n_cells == rc["n_cells"] and len(cyl.radius) == rc["n_segments"]
        np.testing.assert_allclose(cyl.pca_score, rc["pca_score"], rtol=1e-7)
        if len(cyl.radius):
            assert abs(np.dot(cyl.axis, rc["axis"])) >= 1 - 1e-12
        for s in range(len(cyl.radius)):
            assert sum(cyl.inliers[s]) == rc["n_inliers"][s]
            assert ind["cyl_assigned"][r_][s] == rc["assigned"][s]
            np.testing.assert_allclose(cyl.radius[s], rc["radius"][s], rtol=1e-8)
            np.testing.assert_allclose(cyl.centers[s], rc["center"][s], rtol=1e-7, atol=1e-6)
            np.testing.assert_allclose(cyl.mse[s], rc["mse"][s], rtol=1e-6, atol=1e-9)
            kept = bool(rc["assigned"][s] > 0 and ind["cyl_kept"][rc["assigned"][s] - 1])
            assert kept == bool(rc["kept"][s])
    return ind, ref


def test_scene_v0_frames():
    for frame in (0, 1):
        ind, ref = compare(rs.synth.scene_v0_depth(frame))
        assert ref["info"][0]["n_planes"] >= 7 and ref["info"][0]["n_cyl_regions"] >= 1


def edge_cases():
    H, W = 480, 640
    u, v = np.meshgrid(np.arange(W), np.arange(H))
    dx, dy = (u - 320.0) / 550.0, (v - 240.0) / 550.0
    rng = np.random.default_rng(11)
    flat = np.full((H, W), 1500.0, dtype=np.float32)
    yield "empty", np.zeros((H, W), dtype=np.float32)
    yield "exact plane", flat
    half = rs.synth.scene_v0_depth(3).copy()
    half[:, ::2] = 0
    yield "50% invalid columns", half
    step = flat.copy()
    step[:, 333:] = 2100.0
    yield "depth step", (step + rng.normal(0, 1.0, step.shape)).astype(np.float32)
    yield "pure noise", rng.uniform(500, 4000, (H, W)).astype(np.float32)
    ragged = rs.synth.scene_v0_depth(4).copy()
    ragged[rng.random((H, W)) < 0.35] = 0
    yield "ragged validity", ragged
    n = np.array([np.sin(0.25), 0.0, -np.cos(0.25)])           # the histogram's bin-1 quirk (remove_point), see test_oracle_cape
    z = -2000.0 / (n[0] * dx + n[1] * dy + n[2])
    yield "bin-1 floor", (z + rng.normal(0, 1.0, z.shape)).astype(np.float32)


@pytest.mark.parametrize("name,depth", list(edge_cases()), ids=[n for n, _ in edge_cases()])
def test_edge_cases(name, depth):
    compare(depth, noise_free=(name == "exact plane"))


def test_other_seed_and_geometry():
    compare(rs.synth.scene_v0_depth(7), seed=12345)
    K2 = rs.synth.intrinsics(2)
    compare(rs.synth.scene_v0_depth(8, 1280, 960), cell=40, K=K2)


@pytest.mark.parametrize("first", range(0, 104, 8))
def test_random_rooms(first):
    """104 randomly furnished rooms (slanted cylinders, spheres, small regions, planes that nearly merge)."""
    stats = np.zeros(3, dtype=int)
    for seed in range(first, first + 8):
        ind, ref = compare(rs.synth.random_scene_depth(seed))
        stats += (ref["info"][0]["n_planes"], ref["info"][0]["n_cyl_regions"], ref["info"][0]["n_seeds"])
    assert stats[0] > 0
