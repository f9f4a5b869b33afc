"""GPU parity of plane matching (csrc/plane_match.cu) against the oracle (oracle/polygon.cpp), through the C-ABI:
rs_polygon_inter_area = Polygon::inter_area (polygon.cpp:542-561), rs_plane_match = MapPlane::find_matches
(map_primitive.cpp:91-161). The device computes the area by signed fan-triangle clipping, the oracle by a slab sweep:
two different constructions of the same number, compared at 1e-9 of the polygons' scale."""
import numpy as np
import pytest

import oracle_lib as ol
import rgbd_slam_b200 as rs

pytestmark = pytest.mark.gpu


def test_inter_area_matches_oracle_on_random_concave_rings():
    rng = np.random.default_rng(3)
    a_rings, b_rings = [], []
    for k in range(600):
        a_rings.append(rs.synth.star_polygon(rng, int(rng.integers(3, 64)), clockwise=bool(k & 1), closed=bool(k & 4)))
        b_rings.append(rs.synth.star_polygon(rng, int(rng.integers(3, 64)), 100, 1200, center=rng.uniform(-900, 900, 2),
                                             clockwise=bool(k & 2), closed=bool(k & 8)))
    # edge cases: disjoint, contained, identical, shared edge, degenerate rings
    sq = np.array([[0, 0], [4, 0], [4, 3], [0, 3]], dtype=np.float64)
    a_rings += [sq, sq, sq, sq, sq[:2], sq]
    b_rings += [sq + 10, sq * 0.25 + 1, sq, sq + [4, 0], sq, np.zeros((3, 2))]
    got = rs.polygon_inter_area(a_rings, b_rings)
    for i, (a, b) in enumerate(zip(a_rings, b_rings)):
        want = ol.polygon_inter_area(a, b)
        scale = max(ol.polygon_area(a), ol.polygon_area(b), 1.0)
        assert abs(got[i] - want) <= 1e-9 * scale, (i, got[i], want)
    assert got[-6] == 0.0 and got[-3] == 0.0 and got[-2] == 0.0 and got[-1] == 0.0
    assert got[-4] == pytest.approx(12.0, rel=1e-12) and got[-5] == pytest.approx(0.75, rel=1e-12)


@pytest.mark.parametrize("advanced", [False, True])
def test_plane_match_matches_oracle(advanced):
    n_sel = 0
    for seed in range(8):
        args = rs.synth.plane_match_problem(100 + seed, n_frames=16, n_det=8, n_extra_map=3, max_vertices=40)
        matched = args[-1] if seed & 1 else None
        sel, inter = rs.plane_match(*args[:-1], det_matched=matched, advanced_search=advanced)
        rsel, rinter = ol.plane_match(*args[:-1], det_matched=matched, advanced_search=advanced)
        assert np.array_equal(sel, rsel), seed
        assert np.allclose(inter, rinter, rtol=1e-9, atol=1e-6)
        n_sel += int((sel >= 0).sum())
    assert n_sel > 100


def test_plane_match_frames_are_independent():
    """A frame's result does not depend on the batch it is in (one warp per map plane, no cross-frame state)."""
    args = rs.synth.plane_match_problem(7, n_frames=8)
    w2c, det, df, dxy, mp, mf, mxy, matched = args
    sel, inter = rs.plane_match(w2c, det, df, dxy, mp, mf, mxy, matched)
    for f in (0, 3, 7):
        s1, i1 = rs.plane_match(w2c[f:f + 1], det[df[f]:df[f + 1]], [0, df[f + 1] - df[f]], dxy, mp[mf[f]:mf[f + 1]],
                                [0, mf[f + 1] - mf[f]], mxy, matched[df[f]:df[f + 1]])
        assert np.array_equal(s1, sel[mf[f]:mf[f + 1]])
        assert np.array_equal(i1, inter[mf[f]:mf[f + 1]])


def test_plane_match_rejects_oversized_map_polygon():
    args = list(rs.synth.plane_match_problem(1, n_frames=1))
    args[4] = args[4].copy()
    args[4]["n_vertices"][0] = 300
    with pytest.raises(rs.RsError):
        rs.plane_match(*args[:-1])
    args = list(rs.synth.plane_match_problem(1, n_frames=2))
    args[2] = np.array([0, 5, 3], np.int32)     # offsets that go backwards
    with pytest.raises((rs.RsError, ValueError)):
        rs.plane_match(*args[:-1])


def test_sequential_matching_matches_oracle():
    """The caller's loop on the device (plane_select_kernel): same selections and matched mask as the oracle's sequential walk."""
    for seed in range(6):
        w2c, det, df, dxy, mp, mf, mxy, matched = rs.synth.plane_match_problem(300 + seed, n_frames=12, n_extra_map=1)
        mp2, mf2 = [], [0]
        for f in range(len(df) - 1):
            mp2 += list(mp[mf[f]:mf[f + 1]]) * 2   # competing copies
            mf2.append(len(mp2))
        mp2 = np.array(mp2, dtype=mp.dtype)
        dm = matched if seed & 1 else None
        sel, inter, mout = rs.plane_match(w2c, det, df, dxy, mp2, mf2, mxy, det_matched=dm, sequential=True, return_matched=True)
        rsel, rinter, rmout = ol.plane_match(w2c, det, df, dxy, mp2, mf2, mxy, det_matched=dm, sequential=True, return_matched=True)
        assert np.array_equal(sel, rsel) and np.array_equal(mout, rmout), seed
        assert np.allclose(inter, rinter, rtol=1e-9, atol=1e-6)
        one, _ = rs.plane_match(w2c, det, df, dxy, mp2, mf2, mxy, det_matched=dm)
        assert (one != sel).any()
