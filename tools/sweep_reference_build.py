"""Sweep of the oracle against the reference's own compiled CAPE sources (oracle/_ref/libref_cape.so, see oracle/ref_shim) on
random rooms: python tools/sweep_reference_build.py [first seed] [count]. CPU only; every field tests/test_reference_build.py
compares must be bit-identical."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
import rgbd_slam_b200 as rs  # noqa: E402
import test_reference_build as trb  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 256
assert ol.ref_available(), "oracle/_ref/libref_cape.so is not built"
bad, seeds, planes, cyls, merges = [], 0, 0, 0, 0
for s in range(first, first + count):
    try:
        info = trb.compare(rs.synth.random_scene_depth(s))
        seeds += int(info["n_seeds"])
        planes += int(info["n_final_planes"])
        cyls += int(info["n_cylinders"])
        merges += int(info["n_planes"] != info["n_final_planes"])
    except AssertionError as e:
        bad.append((s, str(e)[:120]))
print("rooms %d: seeds %d, final planes %d, cylinders %d, frames with a merge %d" % (count, seeds, planes, cyls, merges))
print("rooms on which the oracle differs from the compiled reference sources: %d" % len(bad))
for b in bad[:20]:
    print("  seed %d: %s" % b)
