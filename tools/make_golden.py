"""Generates the golden fixtures under tests/golden/ from the CPU oracle (oracle/_build/liboracle.so).

The reference itself cannot be compiled in this environment (Eigen / OpenCV C++ / TBB / boost / flann are absent and
there is no network) and has no test, golden vector or dataset for the CAPE path, so these vectors are
outputs of the restated oracle, NOT of the reference binary: they pin the oracle (and the CUDA path) against
regressions, nothing more. Re-run after an intentional oracle change:  python tools/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
import rgbd_slam_b200 as rs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    # CAPE: frames 0..3 of scene v0 (640x480, 20 px cells), seed 0
    depth = rs.synth.scene_v0_batch(0, 4)
    r = ol.cape_run(depth, seed=0)
    np.savez_compressed(
        os.path.join(OUT, "cape_scene_v0.npz"),
        frames=np.arange(4), plane_labels=r["plane_labels"], plane_grid=r["plane_grid"], cyl_labels=r["cyl_labels"],
        info=r["info"], cell_count=r["cells"]["count"], cell_planar=r["cells"]["planar"],
        cell_normal=r["cells"]["normal"], cell_d=r["cells"]["d"], cell_mse=r["cells"]["mse"],
        plane_normal=r["planes"]["normal"][:, :16], plane_d=r["planes"]["d"][:, :16],
        plane_is_final=r["planes"]["is_final"][:, :16], cyl_axis=r["cyls"]["axis"][:, :4],
        cyl_radius=r["cyls"]["radius"][:, :4], cyl_n_segments=r["cyls"]["n_segments"][:, :4])
    # pose: frames 0..3 of the 300-point / 20-plane correspondence sets, reference RNG stream seeded with 0
    poses, status, n_inl, iters, best_it, covs, masks = [], [], [], [], [], [], []
    for f in range(4):
        truth, guess, m = rs.synth.pose_correspondences(f)
        out, mask = ol.pose_solve(guess, m, seed=f)
        poses.append(out["pose"]), status.append(out["status"]), n_inl.append(out["n_inliers"])
        iters.append(out["iterations_run"]), best_it.append(out["best_iteration"]), covs.append(out["cov"])
        masks.append(mask)
    np.savez_compressed(os.path.join(OUT, "pose_synth_v0.npz"), frames=np.arange(4), pose=np.array(poses),
                        status=np.array(status), n_inliers=np.array(n_inl), iterations_run=np.array(iters),
                        best_iteration=np.array(best_it), cov=np.array(covs), mask=np.array(masks))
    # rectify_depth (depth_map_transformation.cpp:23-87): identity and offset camera pair on frame 0; the images are kept
    # as SHA-256 digests plus a few statistics (1.2 MB each otherwise), the histogram bins of frames 0..3 in full
    import hashlib
    import json
    T = np.eye(4)
    c, s_ = np.cos(0.01), np.sin(0.01)
    T[:3, :3] = np.array([[c, -s_, 0], [s_, c, 0], [0, 0, 1]])
    T[:3, 3] = (25.0, -3.0, 4.0)
    ident = ol.rectify_depth(depth[:1])
    moved = ol.rectify_depth(depth[:1], T)
    meta = {"transform_row_major": T.reshape(-1).tolist(),
            "identity": {"sha256": hashlib.sha256(ident.tobytes()).hexdigest(), "valid": int((ident > 0).sum()),
                         "sum": float(ident.astype(np.float64).sum())},
            "offset": {"sha256": hashlib.sha256(moved.tobytes()).hexdigest(), "valid": int((moved > 0).sum()),
                       "sum": float(moved.astype(np.float64).sum())},
            "hist_bin_sha256": hashlib.sha256(np.ascontiguousarray(r["cells"]["hist_bin"]).tobytes()).hexdigest()}
    with open(os.path.join(OUT, "rectify_scene_v0.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
