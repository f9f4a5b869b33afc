"""CPU: pins of the CAPE oracle. The reference has no test / golden vector for this path, so the pins are
(1) numpy restatements of the per-cell arithmetic, (2) numpy.linalg for the 3x3 eigen-solver, (3) the scene statistics
recorded in SURVEY.md Appendix C(5), and (4) the committed golden fixtures (tools/make_golden.py)."""
import os

import numpy as np
import pytest

import oracle_lib as ol
import rgbd_slam_b200 as rs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def numpy_cell_sums(depth, cell=20, K=(550.0, 550.0, 320.0, 240.0)):
    """A.2 + A.3 step 4 in numpy: FP64 back-projection cast to float, FP32 products, FP64 sums."""
    H, W = depth.shape
    fx, fy, cx, cy = K
    invdet = 1.0 / (fx * fy)
    kx = (fy * invdet) * np.arange(W, dtype=np.float64) + (-cx * fy) * invdet
    ky = (fx * invdet) * np.arange(H, dtype=np.float64) + (-fx * cy) * invdet
    z = depth.astype(np.float32)
    x = (z.astype(np.float64) * kx[None, :]).astype(np.float32)
    y = (z.astype(np.float64) * ky[:, None]).astype(np.float32)
    valid = z > 0
    vc, hc = H // cell, W // cell
    out = np.zeros((vc * hc, 10))
    for r in range(vc):
        for c in range(hc):
            sl = (slice(r * cell, (r + 1) * cell), slice(c * cell, (c + 1) * cell))
            v = valid[sl]
            xs, ys, zs = x[sl][v], y[sl][v], z[sl][v]
            f64 = lambda a: a.astype(np.float64).sum()  # noqa: E731
            out[r * hc + c] = [v.sum(), f64(xs), f64(ys), f64(zs), f64(xs * xs), f64(ys * ys), f64(zs * zs), f64(xs * ys),
                               f64(ys * zs), f64(zs * xs)]
    return out


def test_cell_sums_match_numpy():
    depth = rs.synth.scene_v0_depth(0)
    cells = ol.cape_cell_fit(depth)[0]
    ref = numpy_cell_sums(depth)
    fitted = cells["count"] > 0
    assert fitted.sum() > 700
    assert np.array_equal(cells["count"][fitted], ref[fitted, 0].astype(np.int32))
    np.testing.assert_allclose(cells["S"][fitted], ref[fitted, 1:], rtol=1e-12)


def test_organized_cloud_layout():
    """get_organized_cloud_array: rows of a cell contiguous, raster order inside the cell (A.2)."""
    depth = rs.synth.scene_v0_depth(1)
    _, cloud = ol.cape_cell_fit(depth, want_cloud=True)
    cloud = cloud[0]  # [3, W*H] column-major cloud: x column, y column, z column
    H, W, cs = 480, 640, 20
    hc = W // cs
    for (r, c) in [(0, 0), (17, 33), (240, 320), (479, 639), (100, 619)]:
        idx = ((r // cs) * hc + c // cs) * cs * cs + (r % cs) * cs + (c % cs)
        z = depth[r, c]
        if z > 0:
            assert cloud[2, idx] == z
            assert cloud[0, idx] == np.float32(np.float64(z) * ((c - 320.0) / 550.0)) or abs(
                cloud[0, idx] - z * (c - 320.0) / 550.0) <= 1e-4 * abs(z)
        else:
            assert cloud[0, idx] == 0 and cloud[1, idx] == 0 and cloud[2, idx] == 0


def test_eigen3_matches_numpy():
    lib = ol.load()
    rng = np.random.default_rng(3)
    for _ in range(200):
        a = rng.normal(size=(3, 3)) * 10 ** rng.uniform(-3, 6)
        a = a @ a.T
        ev, q = np.zeros(3), np.zeros(9)
        lib.orc_eigen3(np.ascontiguousarray(a).ctypes.data, ev.ctypes.data, q.ctypes.data)
        w, v = np.linalg.eigh(a)
        np.testing.assert_allclose(ev, w, rtol=1e-9, atol=1e-12 * abs(w).max())
        q = q.reshape(3, 3)
        for k in range(3):
            if k == 0 or (w[k] - w[k - 1]) > 1e-6 * abs(w).max():
                if k == 2 or (w[k + 1] - w[k]) > 1e-6 * abs(w).max():
                    assert abs(np.dot(q[:, k], v[:, k])) > 1 - 1e-8


def test_plane_fit_recovers_known_plane():
    """Every planar cell of the back wall (n = (0,0,-1), d0 = 2500) must carry that plane within the noise."""
    depth = rs.synth.scene_v0_depth(2)
    r = ol.cape_run(depth)
    lab = r["plane_labels"][0].reshape(24, 32)
    wall = lab[2, 10]
    assert wall > 0
    p = r["planes"][0][wall - 1]
    n, d = p["normal"], p["d"]
    assert abs(abs(n[2]) - 1) < 1e-4 and abs(d - 2500) < 2.0


def test_scene_v0_statistics():
    """719/768 planar cells, 10 seeds -> 8 plane regions + 2 cylinder-branch regions."""
    r = ol.cape_run(rs.synth.scene_v0_depth(0))
    info = r["info"][0]
    assert info["n_planar_cells"] == 719
    assert info["n_seeds"] == 10
    assert info["n_planes"] == 8
    assert info["n_cyl_regions"] == 2
    cyl = r["cyls"][0][0]
    assert cyl["n_segments"] == 1 and abs(cyl["radius"][0] - 250) < 10          # the synthetic cylinder: r = 250 mm
    assert abs(abs(cyl["axis"][1]) - 1) < 1e-3                                     # axis parallel to y


def test_histogram_bin1_quirk_is_reproduced():
    """remove_point sets the bin to 1 instead of -1 (histogram.hpp:110-112): the oracle must keep running and stay
    deterministic when bin 1 becomes the fullest bin (a floor-like plane seen from above: theta in [9.5, 18.9) deg)."""
    H, W = 480, 640
    u, v = np.meshgrid(np.arange(W), np.arange(H))
    dx, dy = (u - 320.0) / 550.0, (v - 240.0) / 550.0
    n = np.array([np.sin(0.25), 0.0, -np.cos(0.25)])
    z = -2000.0 / (n[0] * dx + n[1] * dy + n[2])
    rng = np.random.default_rng(5)
    depth = (z + rng.normal(0, 1.0, z.shape)).astype(np.float32)
    a = ol.cape_run(depth)
    b = ol.cape_run(depth)
    assert a["info"][0]["n_planes"] >= 1
    assert np.array_equal(a["plane_labels"], b["plane_labels"])


def test_golden_fixture():
    g = np.load(os.path.join(GOLDEN, "cape_scene_v0.npz"))
    depth = rs.synth.scene_v0_batch(0, 4)
    r = ol.cape_run(depth, seed=0)
    assert np.array_equal(r["plane_labels"], g["plane_labels"])
    assert np.array_equal(r["plane_grid"], g["plane_grid"])
    assert np.array_equal(r["cyl_labels"], g["cyl_labels"])
    assert r["info"].tobytes() == g["info"].tobytes()
    assert np.array_equal(r["cells"]["count"], g["cell_count"])
    assert np.array_equal(r["cells"]["planar"], g["cell_planar"])
    np.testing.assert_allclose(r["cells"]["normal"], g["cell_normal"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(r["cells"]["d"], g["cell_d"], rtol=1e-12)
    np.testing.assert_allclose(r["planes"]["normal"][:, :16], g["plane_normal"], atol=1e-12)
    np.testing.assert_allclose(r["planes"]["d"][:, :16], g["plane_d"], rtol=1e-12)


def test_golden_rectify_and_bins():
    import hashlib
    import json
    meta = json.load(open(os.path.join(GOLDEN, "rectify_scene_v0.json")))
    depth = rs.synth.scene_v0_batch(0, 4)
    T = np.array(meta["transform_row_major"]).reshape(4, 4)
    ident = ol.rectify_depth(depth[:1])
    moved = ol.rectify_depth(depth[:1], T)
    assert hashlib.sha256(ident.tobytes()).hexdigest() == meta["identity"]["sha256"]
    assert hashlib.sha256(moved.tobytes()).hexdigest() == meta["offset"]["sha256"]
    assert int((moved > 0).sum()) == meta["offset"]["valid"]
    r = ol.cape_run(depth, seed=0)
    assert hashlib.sha256(np.ascontiguousarray(r["cells"]["hist_bin"]).tobytes()).hexdigest() == meta["hist_bin_sha256"]


@pytest.mark.parametrize("case", ["empty", "flat", "noise"])
def test_degenerate_frames(case):
    H, W = 480, 640
    if case == "empty":
        depth = np.zeros((H, W), np.float32)
    elif case == "flat":
        depth = np.full((H, W), 1500.0, np.float32)  # exact plane: the det == 0 guard rejects cells (A.4)
    else:
        depth = np.random.default_rng(0).uniform(500, 4000, (H, W)).astype(np.float32)
    r = ol.cape_run(depth)
    info = r["info"][0]
    assert info["status"] == 0
    if case != "flat":
        assert info["n_planes"] == 0 and not r["plane_labels"].any()


def test_rectify_depth_restatement_properties():
    """depth_map_transformation.cpp:23-87 restated: untouched pixels are 0, first row / column are never written
    (the reference's strict > 0 test), values are depths of the source image, and an offset camera shifts the image."""
    import rgbd_slam_b200 as rs
    depth = rs.synth.scene_v0_batch(0, 1)
    r = ol.rectify_depth(depth)
    assert r.shape == depth.shape and r.dtype == np.float32
    assert not r[0, 0, :].any() and not r[0, :, 0].any()
    assert np.isin(r[r > 0], depth[depth > 0]).all()          # identity transform: z is carried over unchanged
    assert 0.5 < (r > 0).mean() < (depth > 0).mean()          # float tables: some pixels collide, some stay empty
    assert not ol.rectify_depth(np.zeros_like(depth)).any()
    T = np.eye(4)
    T[0, 3] = 100.0                                            # 100 mm to the right: the scene moves right in the image
    shifted = ol.rectify_depth(depth, T)
    cols = np.arange(640)[None, :]
    assert (shifted[0] > 0).sum() > 0
    assert ((shifted[0] > 0) * cols).sum() / (shifted[0] > 0).sum() > ((r[0] > 0) * cols).sum() / (r[0] > 0).sum()


def test_veltkamp_split_is_the_float_rounding_k1_needs():
    """K1 rounds z*kx, z*ky and three of the FP32 products to float precision inside the FP64 pipe (cape_cell_fit.cu,
    round_to_float): p = v (2^29 + 1), hi = p - (p - v). It must equal double(float(v)) bit for bit, ties included."""
    rng = np.random.default_rng(0)
    C = np.float64(2 ** 29 + 1)

    def veltkamp(v):
        p = v * C
        return p - (p - v)

    m = rng.integers(2 ** 23, 2 ** 24, size=200000).astype(np.float64)
    e = rng.integers(-20, 30, size=m.size)
    ties = (m + 0.5) * np.exp2(e - 23.0)                       # exactly halfway between two floats
    ties = np.concatenate([ties, -ties, np.nextafter(ties, np.inf), np.nextafter(ties, -np.inf), [0.0, -0.0]])
    assert np.array_equal(veltkamp(ties), ties.astype(np.float32).astype(np.float64))
    z = rng.uniform(300, 9000, 500000).astype(np.float32).astype(np.float64)
    k = (rng.integers(0, 640, z.size) - 320.0) / 550.0
    v = z * k                                                   # the back-projection products
    assert np.array_equal(veltkamp(v), v.astype(np.float32).astype(np.float64))
    a = v.astype(np.float32)
    b = (z * ((rng.integers(0, 480, z.size) - 240.0) / 550.0)).astype(np.float32)
    exact = a.astype(np.float64) * b.astype(np.float64)        # exact: 24-bit x 24-bit fits the 53-bit significand
    assert np.array_equal(veltkamp(exact), (a * b).astype(np.float64))


def test_shifted_float_bits_are_the_scaled_double_k1_accumulates():
    """K1 widens a float f >= 0 without a conversion (cape_cell_fit.cu, widen_scaled): the 64-bit pattern
    (bits(f) >> 3, bits(f) << 29) read as a double is exactly f * 2^-896 - zero and subnormal floats included - so
    fma(D, 2^896, S) equals S + double(f), and D * (k * 2^896) equals double(f) * k, bit for bit."""
    rng = np.random.default_rng(1)
    f = np.concatenate([
        rng.uniform(0, 1e4, 100000).astype(np.float32),
        np.exp2(rng.uniform(-149, 127, 100000)).astype(np.float32),            # the whole exponent range
        np.array([0.0, np.float32(1e-45), np.float32(1.1754942e-38), np.float32(1.17549435e-38), np.finfo(np.float32).max],
                 np.float32)])
    b = f.view(np.uint32).astype(np.uint64)
    hi = b >> np.uint64(3)
    lo = (b << np.uint64(29)) & np.uint64(0xFFFFFFFF)
    D = ((hi << np.uint64(32)) | lo).view(np.float64)
    assert np.array_equal(D * np.float64(2.0 ** 896), f.astype(np.float64))
    assert np.all(np.isfinite(D)) and np.all(D >= 0)
    # pre-scaled back-projection factor: same product, same rounding
    k = (rng.integers(0, 640, f.size) - 319.5) / 550.0
    with np.errstate(over="ignore"):
        assert np.array_equal(D * np.ldexp(k, 896), f.astype(np.float64) * k)
    # the sign of a product restored on the widened operand or on the multiplier
    s = rng.choice([-1.0, 1.0], f.size)
    Dneg = (D.view(np.uint64) | np.where(s < 0, np.uint64(1) << np.uint64(63), np.uint64(0))).view(np.float64)
    assert np.array_equal(Dneg * np.float64(2.0 ** 896), s * f.astype(np.float64))
    assert np.array_equal(D * (s * np.float64(2.0 ** 896)), s * f.astype(np.float64))


def test_morphology_restatement_matches_opencv():
    """The boundary and cylinder-opening steps call cv::erode / cv::dilate on the 24x32 cell masks
    (primitive_detection.cpp:48-54,596-598,678-680,719-721). OpenCV's C++ headers are not on this machine but its Python wheel
    is: the oracle's restatement is pinned against the real cv2 calls with the reference's kernels, anchors and borders."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    cross = np.ones((3, 3), np.uint8)
    cross[0, 0] = cross[2, 2] = cross[0, 2] = cross[2, 0] = 0
    square = np.ones((3, 3), np.uint8)
    masks = [(rng.random((24, 32)) < p).astype(np.uint8) for p in (0.05, 0.3, 0.5, 0.7, 0.95) for _ in range(8)]
    masks += [np.zeros((24, 32), np.uint8), np.ones((24, 32), np.uint8)]
    blob = np.zeros((24, 32), np.uint8)
    blob[0:6, 0:9] = 1          # touches the image border: this is where the two border conventions differ
    blob[10:20, 25:32] = 1
    masks.append(blob)
    for m in masks:
        # compute_plane_segment_boundary: erode(cross, BORDER_CONSTANT, Scalar(0)), dilate(square)
        want = cv2.erode(m, cross, anchor=(-1, -1), iterations=1, borderType=cv2.BORDER_CONSTANT, borderValue=0)
        assert np.array_equal(ol.morphology(m, erode=True, cross=True, border_zero=True), want)
        assert np.array_equal(ol.morphology(m, erode=False, cross=False, border_zero=False), cv2.dilate(m, square))
        # add_cylinders_to_primitives: dilate(cross), erode(cross), erode(cross), default border
        d = cv2.dilate(m, cross)
        assert np.array_equal(ol.morphology(m, erode=False, cross=True, border_zero=False), d)
        e = cv2.erode(d, cross)
        assert np.array_equal(ol.morphology(d, erode=True, cross=True, border_zero=False), e)
        assert np.array_equal(ol.morphology(e, erode=True, cross=True, border_zero=False), cv2.erode(e, cross))
    # the two erode borders do differ on a mask that touches the frame (so the test can tell them apart)
    assert not np.array_equal(ol.morphology(blob, True, True, True), ol.morphology(blob, True, True, False))


def test_u16_conversion_formula_matches_opencv_convertTo():
    """rs_cape_run_u16 replaces cv::Mat::convertTo(CV_32F, alpha) of the examples (main_TUM.cpp:242) with
    float(src) * float(alpha) on the device (api_cape.cu: depth_u16_to_f32_kernel; tests/test_cape_gpu.py builds its
    expectation with the same formula). cv2 does not bind convertTo itself, but cv2.normalize(NORM_MINMAX, dtype=CV_32F) ends
    in src.convertTo(dst, CV_32F, scale, shift) with scale = beta / max(src) and shift = 0 when min(src) = 0: with
    max(src) = 65535 and beta = 13107 the scale is exactly the reference's 1/5. The real OpenCV agrees with the FP32
    formula on every pixel and disagrees with an FP64 product rounded to float."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    d16 = rng.integers(0, 65536, (480, 640), dtype=np.uint16)
    d16[0, 0], d16[0, 1] = 0, 65535
    assert 13107.0 / 65535.0 == 1.0 / 5.0
    want = cv2.normalize(d16, None, 0, 13107.0, cv2.NORM_MINMAX, dtype=cv2.CV_32F)
    assert np.array_equal(want, d16.astype(np.float32) * np.float32(1.0 / 5.0))
    assert not np.array_equal(want, (d16.astype(np.float64) * (1.0 / 5.0)).astype(np.float32))
