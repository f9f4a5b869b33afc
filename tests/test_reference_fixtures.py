"""The oracle against fixtures made by the REFERENCE ITSELF (oracle/ref_fixture/README.md: the reference's own sources compiled
into a generator on a machine that has Eigen / OpenCV / boost). Skipped until tests/golden/reference_cape.npz exists - the
round's image cannot build the reference, so the CAPE oracle's parity stays "partial" (DESIGN.md §2)."""
import os
import sys

import numpy as np
import pytest

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURE = os.path.join(ROOT, "tests", "golden", "reference_cape.npz")
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.skipif(not os.path.exists(FIXTURE), reason="no reference-made fixture (needs a reference build)")


def test_oracle_reproduces_the_reference_label_grids():
    import ref_fixture

    fx = np.load(FIXTURE)
    for kind, index in ref_fixture.FRAMES:
        key = "%s_%04d" % (kind, index)
        got = ol.cape_run(ref_fixture.frame_depth(kind, index), seed=0)
        vc, hc = fx[key + "_plane_grid"].shape
        # the reference's private grids: plane-segment index + 1 per cell (before the merge labels are applied, which only
        # add_planes_to_primitives does on masks) and cylinder index + 1 per cell (primitive_detection.cpp:402,454,471)
        assert np.array_equal(got["plane_grid"][0].reshape(vc, hc), fx[key + "_plane_grid"]), key
        assert np.array_equal(got["cyl_labels"][0].reshape(vc, hc), fx[key + "_cyl_grid"]), key
        planes = got["planes"][0]
        final = planes[planes["is_final"] == 1]
        ref = fx[key + "_planes"]
        assert len(final) == len(ref), key
        for p, r in zip(final, ref):
            assert np.allclose(p["normal"], r[:3], rtol=0, atol=1e-9) and abs(p["d"] - r[3]) <= 1e-9 * max(1.0, abs(r[3])), key
