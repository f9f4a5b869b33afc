"""GPU checks at BASELINE.json's full sizes (256-frame batches, 1280x960 / 40 px cells, 1024 hypotheses), where the CPU
oracle would take minutes: size-independent properties of the outputs plus oracle parity on a sample of the batch."""
import numpy as np
import pytest

import oracle_lib as ol
import parity
import rgbd_slam_b200 as rs

pytestmark = pytest.mark.gpu


def _check_frame_invariants(out, b, n_cells):
    info = out["info"][b]
    cells = out["cells"][b]
    labels = out["plane_labels"][b]
    grid = out["plane_grid"][b]
    cyl = out["cyl_labels"][b]
    assert info["status"] == 0
    assert info["n_planar_cells"] == int((cells["planar"] != 0).sum())
    # only planar cells carry labels; a cell is a plane cell or a cylinder cell, never both
    assert not labels[cells["planar"] == 0].any() and not cyl[cells["planar"] == 0].any()
    assert not ((grid > 0) & (cyl > 0)).any()
    # final labels are the merge roots of the segment grid: label set == {k + 1 : plane k is final}
    planes = out["planes"][b][: info["n_planes"]]
    finals = {k + 1 for k in range(info["n_planes"]) if planes["is_final"][k]}
    assert set(np.unique(labels[labels > 0]).tolist()) <= finals
    assert info["n_final_planes"] == len(finals)
    for k in range(info["n_planes"]):
        assert planes["merge_label"][k] <= k                       # merges always point at an earlier plane
        assert abs(np.linalg.norm(planes["normal"][k]) - 1.0) < 1e-12
        assert planes["d"][k] > 0                                  # fit_plane flips the normal so that d > 0
    # boundary points: offsets tile [0, n_boundary) and every kept point lies on its plane within 3 sqrt(MSE)
    pos = 0
    for k in range(info["n_planes"]):
        if not planes["is_final"][k]:
            continue
        assert planes["boundary_offset"][k] == pos
        pts = out["boundary_xyz"][b][pos:pos + planes["n_boundary"][k]]
        dist = np.abs(pts @ planes["normal"][k] + planes["d"][k])
        assert (dist < 3 * np.sqrt(planes["mse"][k]) + 1e-9).all()
        pos += planes["n_boundary"][k]
    assert pos == info["n_boundary"]


def test_cape_batch_256_invariants_permutation_and_sampled_parity():
    F = 256
    depth = rs.synth.scene_v0_batch(100, F)
    det = rs.PrimitiveDetection(640, 480, 20, max_batch=F)
    out = det.find_primitives(depth, seed=5)
    for b in range(0, F, 7):
        _check_frame_invariants(out, b, det.n_cells)
    # frames are independent: reversing the batch reverses the outputs, byte for byte
    rev = det.find_primitives(depth[::-1].copy(), seed=5)
    assert out["cells"][::-1].tobytes() == rev["cells"].tobytes()
    assert np.array_equal(out["plane_labels"][::-1], rev["plane_labels"])
    assert np.array_equal(out["cyl_labels"][::-1], rev["cyl_labels"])
    assert out["info"][::-1].tobytes() == rev["info"].tobytes()
    # and running twice is idempotent
    again = det.find_primitives(depth, seed=5)
    assert out["planes"].tobytes() == again["planes"].tobytes()
    # oracle parity on a sample of the batch
    sample = list(range(0, F, 8))                                  # 32 of the 256 frames
    ref = ol.cape_run(depth[sample], seed=5)
    for i, b in enumerate(sample):
        got_b = {k: v[b:b + 1] for k, v in out.items()}
        ref_b = {k: v[i:i + 1] for k, v in ref.items()}
        parity.assert_cells_match(ref_b["cells"][0], got_b["cells"][0])
        parity.assert_frame_match(ref_b, got_b, 0)
    det.close()


def test_cape_1280x960_batch_invariants():
    F = 16
    K = rs.synth.intrinsics(2)
    depth = rs.synth.scene_v0_batch(200, F, 1280, 960)
    det = rs.PrimitiveDetection(1280, 960, 40, *K, max_batch=F)
    out = det.find_primitives(depth, seed=0)
    for b in range(F):
        _check_frame_invariants(out, b, det.n_cells)
    sample = list(range(0, F, 2))                                  # 8 of the 16 frames
    ref = ol.cape_run(depth[sample], cell=40, K=K, seed=0)
    for i, b in enumerate(sample):
        parity.assert_cells_match(ref["cells"][i], out["cells"][b])
        parity.assert_frame_match({k: v[i:i + 1] for k, v in ref.items()}, {k: v[b:b + 1] for k, v in out.items()}, 0)
    det.close()


def test_config5_full_frames_through_frame_pipeline_match_oracle():
    """BASELINE configs[4] end to end: 1280x960 depth / 40 px cells / 1024 RANSAC hypotheses per frame, CAPE and the pose
    solve together through FramePipeline.track_batch on 8 frames, every output against the oracle (reference RNG streams).
    Half of the frames carry 30-45 % outliers so that the hypothesis loop runs well past the fourth iteration."""
    F, M = 8, 320
    K = rs.synth.intrinsics(2)
    depth = rs.synth.scene_v0_batch(700, F, 1280, 960)
    cur, matches, n = np.zeros((F, 7)), np.zeros((F, M), dtype=rs.abi.match_dtype), np.zeros(F, dtype=np.int32)
    for b in range(F):
        frac = 0.1 if b % 2 == 0 else 0.3 + 0.05 * (b // 2)
        truth, guess, m = rs.synth.pose_correspondences(700 + b, outlier_frac=frac, scale=2)
        cur[b], matches[b, :len(m)], n[b] = guess, m, len(m)
    pipe = rs.FramePipeline(1280, 960, 40, intrinsics=K, max_frames=F, max_matches=M, max_iterations=1024, n_variance=100)
    prims, out, mask, all_poses = pipe.track_batch(depth, cur, matches, n, seed=40, rng_mode=rs.abi.RS_RNG_REFERENCE)
    ref = ol.cape_run(depth, cell=40, K=K, seed=40)
    iterations = []
    for b in range(F):
        parity.assert_cells_match(ref["cells"][b], prims["cells"][b])
        parity.assert_frame_match({k: v[b:b + 1] for k, v in ref.items()}, {k: v[b:b + 1] for k, v in prims.items()}, 0)
        rout, rmask = ol.pose_solve(cur[b], matches[b][:n[b]], K=K, max_iterations=1024, seed=40 + b)
        assert out[b]["status"] == rout["status"] == 1
        for f in ("n_inliers", "iterations_run", "best_iteration", "n_variance_ok"):
            assert out[b][f] == rout[f], (b, f, out[b][f], rout[f])
        assert np.array_equal(mask[b][:n[b]], rmask)
        ok, dt, qd = parity.pose_close(rout["pose"], out[b]["pose"])
        assert ok, (b, dt, qd)
        iterations.append(int(out[b]["iterations_run"]))
    assert max(iterations) > 4, iterations                          # the loop really ran past the early-stop minimum somewhere
    assert np.array_equal(all_poses.cpu().numpy(), out["pose"])
    pipe.close()


def test_pose_batch_256_properties_and_determinism():
    F, M = 256, 320
    truth, cur, m, n = rs.synth.pose_batch(300, F, M, n_points=300, n_planes=20)
    solver = rs.PoseOptimization(max_batch=F, max_matches=M, max_iterations=119, max_variance=100)
    opts = solver.options(seed=77, rng_mode=rs.abi.RS_RNG_DEVICE)
    out, mask = solver.compute_optimized_pose(cur, m, n, opts)
    assert (out["status"] == 1).all()
    err = np.linalg.norm(out["pose"][:, :3] - truth[:, :3], axis=1)
    assert np.median(err) < 2.0 and err.max() < 10.0               # mm, 0.5 px / 5 mm noise, 10 % outliers
    assert np.allclose(np.linalg.norm(out["pose"][:, 3:], axis=1), 1.0, atol=1e-12)
    assert (out["n_inliers"] == mask.sum(axis=1)).all()
    assert (out["n_inliers"] >= 0.8 * n).all() and (out["iterations_run"] >= 4).all()
    assert (out["best_iteration"] < out["iterations_run"]).all()
    # the injected outliers (last 30 points, last 2 planes) are rejected, the rest mostly kept
    assert mask[:, 270:300].mean() < 0.05 and mask[:, :270].mean() > 0.97
    cov = out["cov"].reshape(F, 6, 6)
    assert np.allclose(cov, cov.transpose(0, 2, 1), rtol=0, atol=1e-12 * np.abs(cov).max())
    assert (np.linalg.eigvalsh(cov) > 0).all()
    assert (out["n_variance_ok"] >= 50).all()
    # same seed, same bytes
    out2, mask2 = solver.compute_optimized_pose(cur, m, n, opts)
    assert out.tobytes() == out2.tobytes() and mask.tobytes() == mask2.tobytes()
    solver.close()


def test_frame_pipeline_track_batch_matches_separate_calls():
    """FramePipeline.track_batch (pose solve enqueued, depth streamed meanwhile, solve joined) == the two blocking calls."""
    F, M = 6, 320
    depth = rs.synth.scene_v0_batch(500, F)
    truth, cur, m, n = rs.synth.pose_batch(500, F, M)
    pipe = rs.FramePipeline(max_frames=F, max_matches=M)
    prims, out, mask, all_poses = pipe.track_batch(depth, cur, m, n, seed=3, rng_mode=rs.abi.RS_RNG_DEVICE)
    det = rs.PrimitiveDetection(640, 480, 20, max_batch=F)
    solver = rs.PoseOptimization(max_batch=F, max_matches=M)
    want_prims = det.find_primitives(depth, seed=3)
    want_out, want_mask = solver.compute_optimized_pose(cur, m, n, solver.options(seed=3, rng_mode=rs.abi.RS_RNG_DEVICE))
    assert prims["cells"].tobytes() == want_prims["cells"].tobytes()
    assert np.array_equal(prims["plane_labels"], want_prims["plane_labels"])
    assert out.tobytes() == want_out.tobytes() and mask.tobytes() == want_mask.tobytes()
    assert np.array_equal(all_poses.cpu().numpy(), out["pose"])
    pipe.close(), det.close(), solver.close()
