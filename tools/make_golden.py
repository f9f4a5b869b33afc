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
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
