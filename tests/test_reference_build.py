"""The oracle against the REFERENCE'S OWN CAPE SOURCES, compiled here.

oracle/ref_shim builds /root/reference/src/features/primitives/{plane_segment, cylinder_segment, primitive_detection,
depth_map_transformation}.cpp and coordinates/{point, plane}_coordinates.cpp - where they lie, unmodified - into
oracle/_ref/libref_cape.so against stand-in headers for the third-party libraries this image lacks (a small dense-matrix class
instead of Eigen, an image class + 3x3 morphology instead of OpenCV; the 3x3 eigen-solver / inverse are oracle/linalg.hpp's
restatements of Eigen's algorithms). What that pins is everything the reference itself wrote: the cell fit and its continuity /
planarity tests, histogram seeding, region growing, cylinder RANSAC (std::mt19937 from seed 0), plane merging, boundary
extraction, rectify_depth. The tests are skipped when the library has not been built (no /root/reference at build time)."""
import numpy as np
import pytest

import oracle_lib as ol
import rgbd_slam_b200 as rs
from test_independent_cape import edge_cases

pytestmark = pytest.mark.skipif(not ol.ref_available(), reason="oracle/_ref/libref_cape.so not built (needs /root/reference)")


def compare(depth, exact_cells=True):
    ref = ol.ref_cape_run(depth)
    got = ol.cape_run(depth, seed=0)     # MAKE_DETERMINISTIC: utils::Random::_seed = 0, engine restarted per frame (fresh thread)
    cells, info = got["cells"][0], got["info"][0]
    # init_planar_cell_fitting: per cell point count, planar flag and - where a plane was fitted - its parameters, BIT FOR BIT
    assert np.array_equal(ref["count"], cells["count"])
    assert np.array_equal(ref["planar"], cells["planar"])
    fitted = cells["planar"] == 1
    assert np.array_equal(ref["cell"][fitted, :3], cells["normal"][fitted])
    assert np.array_equal(ref["cell"][fitted, 3], cells["d"][fitted])
    assert np.array_equal(ref["cell"][fitted, 4], cells["mse"][fitted])
    # the private label grids of Primitive_Detection after find_primitives: plane segment index + 1, cylinder index + 1
    assert np.array_equal(ref["plane_grid"], got["plane_grid"][0])
    assert np.array_equal(ref["cyl_grid"], got["cyl_labels"][0])
    # plane_container: the merged planes that survive, in order, with the refitted parameters and their ordered boundary points
    planes = got["planes"][0][:info["n_planes"]]
    final = planes[planes["is_final"] == 1]
    assert len(ref["planes"]) == len(final)
    cursor = 0
    for r, p in zip(ref["planes"], final):
        assert np.array_equal(r[:3], p["normal"]) and r[3] == p["d"] and r[4] == p["mse"]
        nb = int(r[5])
        assert nb == p["n_boundary"]
        want = got["boundary_xyz"][0][p["boundary_offset"]:p["boundary_offset"] + nb]
        assert np.array_equal(ref["boundary"][cursor:cursor + nb], want)
        cursor += nb
    # cylinder_container: one entry per kept (region, sub-segment) pair, the region's axis; the radius is NaN in the reference
    # (Cylinder_Segment's copy constructor resets the segment count, Cylinder::Cylinder divides by it: INTEGRATION.md)
    cyls = got["cyls"][0][:info["n_cyl_regions"]]
    kept_axes = [c["axis"] for c in cyls for s in range(c["n_segments"]) if c["kept"][s]]
    assert len(ref["cyls"]) == len(kept_axes)
    for r, axis in zip(ref["cyls"], kept_axes):
        assert np.array_equal(r[:3], axis)
        assert np.isnan(r[3])
    return info


def test_scene_v0_frames():
    for i in range(4):
        info = compare(rs.synth.scene_v0_depth(i))
        assert info["n_final_planes"] >= 7 and info["n_cyl_regions"] >= 1


@pytest.mark.parametrize("name,depth", list(edge_cases()), ids=[n for n, _ in edge_cases()])
def test_edge_cases(name, depth):
    compare(depth)


@pytest.mark.parametrize("first", [100, 132, 164])
def test_random_rooms(first):
    seeds = planes = cylinders = 0
    for s in range(first, first + 32):
        info = compare(rs.synth.random_scene_depth(s))
        seeds += info["n_seeds"]
        planes += info["n_final_planes"]
        cylinders += info["n_cylinders"]
    assert seeds > 60 and planes > 40 and cylinders > 10   # the rooms do exercise every branch


def _cam2_to_cam1(rx=0.0, ry=0.0, rz=0.0, t=(0.0, 0.0, 0.0)):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = t
    return T


def test_rectify_depth():
    """Depth_Map_Transformation::rectify_depth: identity, a pure offset and random depth -> colour extrinsics, byte for byte."""
    rng = np.random.default_rng(5)
    transforms = [np.eye(4), _cam2_to_cam1(t=(25.0, -3.0, 1.5))] + [
        _cam2_to_cam1(*rng.uniform(-0.05, 0.05, 3), t=rng.uniform(-80, 80, 3)) for _ in range(6)]
    for k, T in enumerate(transforms):
        depth = rs.synth.random_scene_depth(400 + k) if k & 1 else rs.synth.scene_v0_depth(k)
        want = ol.ref_rectify_depth(depth, T)
        got = ol.rectify_depth(depth, T)[0]
        assert want.tobytes() == got.tobytes(), k
