"""Parity sweep of plane matching on synth.plane_match_problem: rs_plane_match (GPU, fan-triangle clipping) vs the oracle (slab
sweep) - selections exact, intersection areas to 1e-9 of the polygons' area. Usage (GPU box): python tools/sweep_plane_match.py
[first seed] [problems] [frames per problem]."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
import rgbd_slam_b200 as rs  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
count = int(sys.argv[2]) if len(sys.argv) > 2 else 64
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 32
n_map = n_sel = n_bad = 0
worst = 0.0
for seed in range(first, first + count):
    args = rs.synth.plane_match_problem(seed, n_frames=frames, n_det=8, n_extra_map=3, max_vertices=48)
    matched = args[-1] if seed & 1 else None
    for adv in (False, True):
        sel, inter = rs.plane_match(*args[:-1], det_matched=matched, advanced_search=adv)
        rsel, rinter = ol.plane_match(*args[:-1], det_matched=matched, advanced_search=adv)
        n_map += len(sel)
        n_sel += int((sel >= 0).sum())
        if not np.array_equal(sel, rsel):
            n_bad += int((sel != rsel).sum())
            print("seed %d advanced %d: selections differ at map planes %s" % (seed, adv, np.nonzero(sel != rsel)[0][:8]))
        ok = sel == rsel
        rel = np.abs(inter[ok] - rinter[ok]) / np.maximum(rinter[ok], 1.0)
        worst = max(worst, float(rel.max()) if len(rel) else 0.0)
print("map planes %d (x2 search modes counted), matched %d, selections that differ %d, worst relative area difference %.3g"
      % (n_map, n_sel, n_bad, worst))
