"""Parity sweep: CAPE through the C-ABI vs the CPU oracle on N random scenes (synth.random_scene_depth).
Usage (GPU box): python tools/sweep_random_scenes.py [first] [count] [seed]. Prints the frames that differ."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
rs = importlib.import_module("rgb-d-slam_b200")
import oracle_lib as ol  # noqa: E402
import parity  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
count = int(sys.argv[2]) if len(sys.argv) > 2 else 512
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 0
B = 64
det = rs.PrimitiveDetection(640, 480, 20, max_batch=B)
bad = []
stats = np.zeros(4, np.int64)
for s0 in range(first, first + count, B):
    n = min(B, first + count - s0)
    depth = rs.synth.random_scene_batch(s0, n)
    got = det.find_primitives(depth, seed=seed)
    ref = ol.cape_run(depth, seed=seed)
    stats += [ref["info"]["n_seeds"].sum(), ref["info"]["n_final_planes"].sum(), ref["info"]["n_cylinders"].sum(),
              (ref["info"]["n_planes"] != ref["info"]["n_final_planes"]).sum()]
    for b in range(n):
        try:
            assert np.array_equal(ref["plane_labels"][b], got["plane_labels"][b]), "plane labels"
            assert np.array_equal(ref["cyl_labels"][b], got["cyl_labels"][b]), "cylinder labels"
            parity.assert_cells_match(ref["cells"][b], got["cells"][b])
            parity.assert_frame_match(ref, got, b)
        except AssertionError as e:
            bad.append((s0 + b, str(e)[:200]))
det.close()
print("scenes %d  seeds %d  planes %d  cylinders %d  frames with a merge %d" % ((count,) + tuple(stats)))
print("mismatching frames: %d" % len(bad))
for b in bad[:40]:
    print("  scene %d: %s" % b)
