"""Inputs and packing of the reference-made CAPE fixtures (oracle/ref_fixture/README.md).
  python tools/ref_fixture.py inputs DIR          # the parity tests' synthetic depth frames as raw float32 files
  python tools/ref_fixture.py pack DIR OUT.npz    # the generator's .labels files -> tests/golden/reference_cape.npz"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import rgbd_slam_b200 as rs  # noqa: E402

FRAMES = [("v0", i) for i in range(4)] + [("random", 100 + i) for i in range(12)]


def frame_depth(kind, index):
    return rs.synth.scene_v0_depth(index) if kind == "v0" else rs.synth.random_scene_depth(index)


def main():
    cmd, d = sys.argv[1], sys.argv[2]
    if cmd == "inputs":
        os.makedirs(d, exist_ok=True)
        for kind, index in FRAMES:
            frame_depth(kind, index).astype(np.float32).tofile(os.path.join(d, "%s_%04d.f32" % (kind, index)))
    elif cmd == "pack":
        out = {}
        for kind, index in FRAMES:
            raw = np.fromfile(os.path.join(d, "%s_%04d.f32.labels" % (kind, index)), dtype=np.uint8)
            vc, hc = np.frombuffer(raw[:8], dtype=np.int32)
            o, n = 8, int(vc) * int(hc)
            key = "%s_%04d" % (kind, index)
            out[key + "_plane_grid"] = np.frombuffer(raw[o:o + 4 * n], dtype=np.int32).reshape(vc, hc)
            o += 4 * n
            out[key + "_cyl_grid"] = np.frombuffer(raw[o:o + 4 * n], dtype=np.int32).reshape(vc, hc)
            o += 4 * n
            npl = int(np.frombuffer(raw[o:o + 4], dtype=np.int32)[0])
            o += 4
            out[key + "_planes"] = np.frombuffer(raw[o:o + 32 * npl], dtype=np.float64).reshape(npl, 4)
            o += 32 * npl
            ncy = int(np.frombuffer(raw[o:o + 4], dtype=np.int32)[0])
            o += 4
            out[key + "_cylinders"] = np.frombuffer(raw[o:o + 32 * ncy], dtype=np.float64).reshape(ncy, 4)
        np.savez_compressed(sys.argv[3], **out)
    else:
        raise SystemExit(__doc__)


if __name__ == "__main__":
    main()
