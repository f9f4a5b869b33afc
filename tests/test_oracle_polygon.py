"""CPU tests of the plane-matching oracle (oracle/polygon.cpp): the intersection area behind Polygon::inter_area
(/root/reference/src/utils/polygon.cpp:542-561 = summed boost::geometry::area of boost::geometry::intersection) against closed
forms, the real OpenCV (cv2.intersectConvexConvex), a raster count (cv2.fillPoly) for concave rings, and MapPlane::find_matches
(/root/reference/src/map_management/map_features/map_primitive.cpp:91-161) against an independent numpy restatement.
boost::geometry itself is absent from this machine: parity with it is unpinned beyond these semantics."""
import numpy as np
import pytest

import oracle_lib as ol
import rgbd_slam_b200 as rs

cv2 = pytest.importorskip("cv2")


def rect(x0, y0, x1, y1):
    return np.array([[x0, y0], [x1, y0], [x1, y1], [x0, y1]], dtype=np.float64)


def test_closed_forms():
    a = rect(0, 0, 4, 3)
    assert ol.polygon_area(a) == 12.0
    assert ol.polygon_inter_area(a, rect(1, 1, 2, 2)) == pytest.approx(1.0, rel=1e-14)            # contained
    assert ol.polygon_inter_area(a, rect(2, 1, 10, 2)) == pytest.approx(2.0, rel=1e-14)           # partial
    assert ol.polygon_inter_area(a, rect(5, 0, 6, 3)) == 0.0                                      # disjoint
    assert ol.polygon_inter_area(a, rect(4, 0, 6, 3)) == 0.0                                      # shared edge only
    assert ol.polygon_inter_area(a, a) == pytest.approx(12.0, rel=1e-14)                          # identical
    assert ol.polygon_inter_area(a, a[::-1]) == pytest.approx(12.0, rel=1e-14)                    # orientation irrelevant
    assert ol.polygon_inter_area(a, np.concatenate([a, a[:1]])) == pytest.approx(12.0, rel=1e-14)  # closed ring
    tri = np.array([[0, 0], [4, 0], [0, 4]], dtype=np.float64)
    assert ol.polygon_inter_area(tri, rect(0, 0, 2, 2)) == pytest.approx(4.0, rel=1e-14)
    assert ol.polygon_inter_area(tri, rect(1, 1, 3, 3)) == pytest.approx(2.0, rel=1e-14)          # corner cut by the hypotenuse
    # a U shape against a bar across its two prongs: two disjoint pieces, areas summed (the loop of polygon.cpp:555-559)
    u = np.array([[0, 0], [5, 0], [5, 4], [4, 4], [4, 1], [1, 1], [1, 4], [0, 4]], dtype=np.float64)
    assert ol.polygon_inter_area(u, rect(-1, 2, 6, 3)) == pytest.approx(2.0, rel=1e-14)
    assert ol.polygon_inter_area(u[:2], a) == 0.0                                                  # degenerate ring


def test_convex_pairs_against_opencv():
    rng = np.random.default_rng(5)
    for _ in range(200):
        pts_a = rng.uniform(-100, 100, (12, 2)).astype(np.float32)
        pts_b = (rng.uniform(-100, 100, (9, 2)) + rng.uniform(-60, 60, 2)).astype(np.float32)
        ha, hb = cv2.convexHull(pts_a).reshape(-1, 2), cv2.convexHull(pts_b).reshape(-1, 2)
        want, _ = cv2.intersectConvexConvex(ha, hb)
        got = ol.polygon_inter_area(ha.astype(np.float64), hb.astype(np.float64))
        assert got == pytest.approx(want, rel=2e-4, abs=0.05)   # OpenCV works in float32


def test_concave_pairs_against_raster():
    rng = np.random.default_rng(9)
    S = 2048
    for k in range(40):
        a = rs.synth.star_polygon(rng, int(rng.integers(3, 30)), 150, 900, clockwise=bool(k & 1))
        b = rs.synth.star_polygon(rng, int(rng.integers(3, 30)), 150, 900, center=rng.uniform(-500, 500, 2), closed=bool(k & 2))
        lo, hi = -1500.0, 1500.0
        scale = S / (hi - lo) * 16   # fillPoly with 4 fractional bits
        ma, mb = np.zeros((S, S), np.uint8), np.zeros((S, S), np.uint8)
        cv2.fillPoly(ma, [np.round((a - lo) * scale).astype(np.int32)], 1, shift=4)
        cv2.fillPoly(mb, [np.round((b - lo) * scale).astype(np.int32)], 1, shift=4)
        px = ((hi - lo) / S) ** 2
        want = float(np.count_nonzero(ma & mb)) * px
        got = ol.polygon_inter_area(a, b)
        perim = (np.abs(np.diff(a, axis=0)).sum() + np.abs(np.diff(b, axis=0)).sum()) * (hi - lo) / S
        assert abs(got - want) <= 1.5 * perim + 1e-6, (k, got, want)
        assert got <= min(ol.polygon_area(a), ol.polygon_area(b)) * (1 + 1e-12)
        assert got == pytest.approx(ol.polygon_inter_area(b, a), rel=1e-10, abs=1e-6)            # symmetric


def numpy_find_matches(w2c, det, det_xy, mp, map_xy, det_matched, advanced, sequential=False):
    """MapPlane::find_matches written from the reference with numpy; the intersection area comes from the oracle.
    sequential: with the caller's loop (feature_map.hpp:652-669) marking taken detections between two map planes."""
    R, t = w2c[:3, :3], w2c[:3, 3]
    out = []
    det_matched = np.zeros(len(det), bool) if det_matched is None else np.array(det_matched, bool)
    thr = np.float64(np.float32(0.4)) / (2 if advanced else 1)
    for m in mp:
        pw = np.array([*m["normal"], m["d"]])
        # plane world->camera 4x4 = inverse of [[R_cw, 0], [-t_cw^T R_cw, 1]] (camera_transformation.cpp:41-71)
        c2w = np.linalg.inv(w2c)
        Pcw = np.eye(4)
        Pcw[:3, :3] = c2w[:3, :3]
        Pcw[3, :3] = -c2w[:3, 3] @ c2w[:3, :3]
        pc = np.linalg.inv(Pcw) @ pw
        nc, dc = pc[:3] / np.linalg.norm(pc[:3]), pc[3]
        ring = map_xy[m["first_vertex"]:m["first_vertex"] + m["n_vertices"]]
        nC = R @ m["center"] + t
        nX = R @ m["x_axis"]
        nX /= np.linalg.norm(nX)
        nY = R @ m["y_axis"]
        nY /= np.linalg.norm(nY)
        world = m["center"] + ring[:, :1] * m["x_axis"] + ring[:, 1:] * m["y_axis"]
        camp = world @ R.T + t - nC
        cam = np.stack([camp @ nX, camp @ nY], axis=1)
        sel, best = -1, 0.0
        if ol.polygon_area(cam) > 0:
            for k, dpl in enumerate(det):
                if det_matched[k]:
                    continue
                if not abs(dpl["d"] - dc) < 100.0 or not abs(dpl["normal"] @ nc) > abs(np.cos(np.deg2rad(20.0))):
                    continue
                p3 = nC + cam[:, :1] * nX + cam[:, 1:] * nY - dpl["center"]
                prj = np.stack([p3 @ dpl["x_axis"], p3 @ dpl["y_axis"]], axis=1)
                dring = det_xy[dpl["first_vertex"]:dpl["first_vertex"] + dpl["n_vertices"]]
                inter = ol.polygon_inter_area(dring, prj)
                if inter > best and inter / ol.polygon_area(dring) >= thr:
                    sel, best = k, inter
        out.append((sel, best) if sel > 0 else (-1, 0.0))
        if sequential and sel > 0:
            det_matched[sel] = True
    return out


@pytest.mark.parametrize("advanced", [False, True])
def test_plane_match_against_numpy_restatement(advanced):
    n_sel = 0
    for seed in range(6):
        w2c, det, df, dxy, mp, mf, mxy, matched = rs.synth.plane_match_problem(seed, n_frames=3)
        sel, inter = ol.plane_match(w2c, det, df, dxy, mp, mf, mxy, matched, advanced)
        for f in range(len(df) - 1):
            want = numpy_find_matches(w2c[f], det[df[f]:df[f + 1]], dxy, mp[mf[f]:mf[f + 1]], mxy, matched[df[f]:df[f + 1]], advanced)
            for i, (ws, wi) in enumerate(want):
                assert sel[mf[f] + i] == ws, (seed, f, i)
                assert inter[mf[f] + i] == pytest.approx(wi, rel=1e-9, abs=1e-6)
        n_sel += int((sel >= 0).sum())
        assert not (sel == 0).any()   # the reference's `selectedIndex <= 0` quirk
    assert n_sel > 10   # the scenario does produce matches


def test_sequential_matching_reproduces_the_callers_loop():
    """sequential=True: the map planes of a frame are served in order and a taken detection is marked matched for the next ones
    (Feature_Map::get_matches). Map planes that compete for one detection must come out differently from the one-shot mode."""
    differs = 0
    for seed in range(8):
        w2c, det, df, dxy, mp, mf, mxy, matched = rs.synth.plane_match_problem(40 + seed, n_frames=3, n_extra_map=0)
        # duplicate every frame's map planes: the copies compete with the originals for the same detections
        mp2, mf2 = [], [0]
        for f in range(len(df) - 1):
            mp2 += list(mp[mf[f]:mf[f + 1]]) * 2
            mf2.append(len(mp2))
        mp2 = np.array(mp2, dtype=mp.dtype)
        sel, inter, mout = ol.plane_match(w2c, det, df, dxy, mp2, mf2, mxy, matched, sequential=True, return_matched=True)
        one, _ = ol.plane_match(w2c, det, df, dxy, mp2, mf2, mxy, matched)
        differs += int((sel != one).sum())
        for f in range(len(df) - 1):
            want = numpy_find_matches(w2c[f], det[df[f]:df[f + 1]], dxy, mp2[mf2[f]:mf2[f + 1]], mxy, matched[df[f]:df[f + 1]], False,
                                      sequential=True)
            got = sel[mf2[f]:mf2[f + 1]]
            assert [w[0] for w in want] == list(got), (seed, f)
            taken = got[got >= 0]
            assert len(set(taken)) == len(taken)                          # no detection is handed out twice
            expect = np.array(matched[df[f]:df[f + 1]], bool)
            expect[taken] = True
            assert np.array_equal(mout[df[f]:df[f + 1]].astype(bool), expect)
    assert differs > 0
