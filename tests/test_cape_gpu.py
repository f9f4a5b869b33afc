"""GPU parity: CAPE through the C-ABI (librgbdslam_b200.so) vs the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

import oracle_lib as ol
import parity
import rgbd_slam_b200 as rs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def det():
    d = rs.PrimitiveDetection(640, 480, 20, max_batch=8)
    yield d
    d.close()


def test_cell_fit_matches_oracle(det):
    depth = rs.synth.scene_v0_batch(0, 4)
    got = det.find_primitives(depth, cells_only=True)["cells"]
    ref = ol.cape_cell_fit(depth)
    for b in range(4):
        frac = parity.assert_cells_match(ref[b], got[b])
        # sums of FP32 values in FP64 are exact unless a cell straddles the optical axis; expect most cells bit-identical
        assert frac > 0.9, "only %.3f of the cells are bit-identical" % frac


def test_find_primitives_matches_oracle(det):
    depth = rs.synth.scene_v0_batch(0, 8)
    got = det.find_primitives(depth, seed=0)
    ref = ol.cape_run(depth, seed=0)
    for b in range(8):
        parity.assert_frame_match(ref, got, b)
    assert np.array_equal(ref["plane_labels"], got["plane_labels"])
    assert np.array_equal(ref["cyl_labels"], got["cyl_labels"])


def test_seed_changes_cylinder_draws_consistently(det):
    depth = rs.synth.scene_v0_batch(3, 2)
    for seed in (1, 12345):
        got = det.find_primitives(depth, seed=seed)
        ref = ol.cape_run(depth, seed=seed)
        for b in range(2):
            parity.assert_frame_match(ref, got, b)


def test_edge_cases(det):
    H, W = 480, 640
    rng = np.random.default_rng(7)
    frames = []
    frames.append(np.zeros((H, W), np.float32))                                    # empty depth
    frames.append(np.full((H, W), 1500.0, np.float32))                             # exact plane: singular scatter
    f = rs.synth.scene_v0_depth(11)
    f[:, ::2] = 0                                                                  # 50 % invalid columns
    frames.append(f)
    f = rs.synth.scene_v0_depth(12)
    f[200:280, :] += 400.0                                                         # depth step through cell rows
    frames.append(f)
    frames.append((rng.uniform(500, 4000, (H, W))).astype(np.float32))             # pure noise: nothing planar
    f = rs.synth.scene_v0_depth(13)
    f[rng.random((H, W)) < 0.35] = 0                                               # ragged validity around the 280 limit
    frames.append(f)
    depth = np.stack(frames)
    got = det.find_primitives(depth, seed=0)
    ref = ol.cape_run(depth, seed=0)
    for b in range(len(frames)):
        parity.assert_cells_match(ref["cells"][b], got["cells"][b])
        parity.assert_frame_match(ref, got, b)


@pytest.mark.parametrize("first,seed", [(0, 0), (48, 7), (96, 2024)])
def test_random_scenes_match_oracle(first, seed):
    """48 randomly furnished rooms per case (synth.random_scene_depth: 0-6 planes, slanted cylinders, a sphere, varying
    noise and holes): labels bit-exact, every plane / cylinder / boundary point within tolerance."""
    depth = rs.synth.random_scene_batch(first, 48)
    d = rs.PrimitiveDetection(640, 480, 20, max_batch=48)
    got = d.find_primitives(depth, seed=seed)
    d.close()
    ref = ol.cape_run(depth, seed=seed)
    assert np.array_equal(ref["plane_labels"], got["plane_labels"])
    assert np.array_equal(ref["cyl_labels"], got["cyl_labels"])
    for b in range(48):
        parity.assert_cells_match(ref["cells"][b], got["cells"][b])
        parity.assert_frame_match(ref, got, b)
    assert ref["info"]["n_cylinders"].sum() > 10 and ref["info"]["n_seeds"].max() >= 8   # the sweep is not trivial


def test_large_cells_1280x960():
    d = rs.PrimitiveDetection(1280, 960, 40, *rs.synth.intrinsics(2), max_batch=2)
    depth = rs.synth.scene_v0_batch(0, 2, 1280, 960)
    got = d.find_primitives(depth, seed=0)
    ref = ol.cape_run(depth, cell=40, K=rs.synth.intrinsics(2), seed=0)
    for b in range(2):
        parity.assert_cells_match(ref["cells"][b], got["cells"][b])
        parity.assert_frame_match(ref, got, b)
    d.close()


def test_batch_size_independent(det):
    depth = rs.synth.scene_v0_batch(20, 8)
    full = det.find_primitives(depth, seed=0)
    one = det.find_primitives(depth[5:6], seed=0)
    assert full["cells"][5].tobytes() == one["cells"][0].tobytes()
    assert np.array_equal(full["plane_labels"][5], one["plane_labels"][0])


def test_invalid_arguments(det):
    with pytest.raises(rs.RsError):
        det.find_primitives(rs.synth.scene_v0_batch(0, 9))   # batch > max_batch
    with pytest.raises(rs.RsError):
        rs.PrimitiveDetection(640, 480, 18)                    # cell side not a multiple of 4


def test_u16_depth_matches_float_path_and_oracle(det):
    # the raw sensor image of the reference's examples: CV_16U, 1/5 mm units, convertTo(CV_32F, 1/5) (main_TUM.cpp:242)
    depth = rs.synth.scene_v0_batch(30, 3)
    d16 = np.clip(np.rint(depth * 5.0), 0, 65535).astype(np.uint16)
    as_float = d16.astype(np.float32) * np.float32(1.0 / 5.0)   # OpenCV's cvtScale 16u -> 32f works in float
    got = det.find_primitives_u16(d16, alpha=1.0 / 5.0, seed=0)
    same = det.find_primitives(as_float, seed=0)
    ref = ol.cape_run(as_float, seed=0)
    assert got["cells"].tobytes() == same["cells"].tobytes()
    assert np.array_equal(got["plane_labels"], same["plane_labels"])
    for b in range(3):
        parity.assert_cells_match(ref["cells"][b], got["cells"][b])
        parity.assert_frame_match(ref, got, b)


def test_chunked_host_run_covers_ragged_batches():
    # rs_cape_run streams 32-frame chunks: a batch that is not a multiple of the chunk, against per-frame runs
    d = rs.PrimitiveDetection(640, 480, 20, max_batch=70)
    depth = rs.synth.scene_v0_batch(40, 70)
    full = d.find_primitives(depth, seed=3)
    for b in (0, 31, 32, 63, 64, 69):
        one = d.find_primitives(depth[b:b + 1], seed=3)
        assert full["cells"][b].tobytes() == one["cells"][0].tobytes()
        assert np.array_equal(full["plane_labels"][b], one["plane_labels"][0])
        assert np.array_equal(full["cyl_labels"][b], one["cyl_labels"][0])
        assert full["info"][b].tobytes() == one["info"][0].tobytes()
    d.close()


def _cam2_to_cam1(rx=0.01, ry=-0.02, rz=0.005, t=(25.0, -3.0, 4.0)):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = t
    return T


@pytest.mark.parametrize("transform", [None, _cam2_to_cam1()])
def test_rectify_depth_matches_oracle_bit_exact(det, transform):
    depth = rs.synth.scene_v0_batch(50, 3)
    depth[1, 100:140, :] = 0           # an empty band
    det.set_rectification(transform, enable=True)
    try:
        got = det.rectify_depth(depth)
        ref = ol.rectify_depth(depth, transform)
        assert got.tobytes() == ref.tobytes()
        assert (ref > 0).mean() > 0.3   # the scatter really moved data (and, with float tables, really loses some)
        # the whole path behind it: find_primitives on the raw image == oracle CAPE on the oracle-rectified image
        full = det.find_primitives(depth, seed=0)
        want = ol.cape_run(ref, seed=0)
        for b in range(3):
            parity.assert_cells_match(want["cells"][b], full["cells"][b])
            parity.assert_frame_match(want, full, b)
    finally:
        det.set_rectification(None, enable=False)
    # and switching it off restores the plain path
    plain = det.find_primitives(depth[:1], seed=0)
    ref_plain = ol.cape_run(depth[:1], seed=0)
    parity.assert_frame_match(ref_plain, plain, 0)


def test_rectify_depth_random_extrinsics_bit_exact(det):
    """Eight random depth->colour extrinsics (up to 3 degrees per axis, 80 mm baseline, both signs) on random rooms:
    the order-free scatter must reproduce the oracle's raster-order last-writer-wins image bit for bit."""
    rng = np.random.default_rng(99)
    depth = rs.synth.random_scene_batch(300, 4)
    try:
        for _ in range(8):
            T = _cam2_to_cam1(*rng.uniform(-0.05, 0.05, 3), t=rng.uniform(-80, 80, 3))
            det.set_rectification(T, enable=True)
            got = det.rectify_depth(depth)
            ref = ol.rectify_depth(depth, T)
            assert got.tobytes() == ref.tobytes()
    finally:
        det.set_rectification(None, enable=False)


def test_rectify_depth_extreme_depths_bit_exact(det):
    """Depths a sensor never produces but a float image can hold: subnormal, tiny, huge (finite), negative. The kernels widen
    float -> double with integer shifts for zero / normal values and fall back to the conversion otherwise (rectify.cu: widen);
    the image must still equal the oracle's byte for byte. (NaN / infinite depths make the reference call exit(-1).)"""
    depth = rs.synth.scene_v0_batch(70, 2)
    rng = np.random.default_rng(4)
    specials = np.array([1e-45, 1e-40, 1.1754942e-38, 1.1754944e-38, 1e-30, 1e-10, 1e10, 1e30, 3.4e38, -1.0, -0.0], np.float32)
    rows, cols = rng.integers(0, 480, 4000), rng.integers(0, 640, 4000)
    depth[0, rows, cols] = specials[rng.integers(0, len(specials), 4000)]
    depth[1, 200:203, :] = np.float32(1e-41)          # whole rows of subnormals
    T = _cam2_to_cam1(0.01, -0.02, 0.015, t=(30.0, -5.0, 8.0))
    try:
        for ext in (None, T):
            det.set_rectification(ext, enable=True)
            got = det.rectify_depth(depth)
            ref = ol.rectify_depth(depth, ext)
            assert got.tobytes() == ref.tobytes()
    finally:
        det.set_rectification(None, enable=False)


def test_new_entry_points_fail_loudly(det):
    depth = rs.synth.scene_v0_batch(0, 1)
    fresh = rs.PrimitiveDetection(640, 480, 20, max_batch=1)
    with pytest.raises(rs.RsError):
        fresh.rectify_depth(depth)                             # rs_cape_set_rectification has never been called
    fresh.close()
    with pytest.raises(rs.RsError):
        det.find_primitives_u16(np.zeros((9, 480, 640), np.uint16))   # batch > max_batch
    with pytest.raises(ValueError):
        det.find_primitives_u16(np.zeros((1, 100, 100), np.uint16))   # wrong geometry
    # an all-zero sensor image is an empty frame, not an error
    out = det.find_primitives_u16(np.zeros((1, 480, 640), np.uint16), alpha=0.2)
    assert out["info"][0]["n_planar_cells"] == 0 and not out["plane_labels"].any()


def test_rectify_depth_against_committed_golden(det):
    import hashlib
    import json
    import os
    meta = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rectify_scene_v0.json")))
    depth = rs.synth.scene_v0_batch(0, 1)
    T = np.array(meta["transform_row_major"]).reshape(4, 4)
    try:
        det.set_rectification(None, enable=True)
        assert hashlib.sha256(det.rectify_depth(depth).tobytes()).hexdigest() == meta["identity"]["sha256"]
        det.set_rectification(T, enable=True)
        assert hashlib.sha256(det.rectify_depth(depth).tobytes()).hexdigest() == meta["offset"]["sha256"]
    finally:
        det.set_rectification(None, enable=False)


def test_two_contexts_on_two_host_threads_match_serial_runs():
    """Two batches in flight (bench.py's e2e two_lanes mode): the blocking host API called from two host threads, each on
    its own context, returns what the same calls return one after the other."""
    import threading
    a = rs.PrimitiveDetection(640, 480, 20, max_batch=8)
    b = rs.PrimitiveDetection(640, 480, 20, max_batch=8)
    da, db = rs.synth.scene_v0_batch(0, 8), rs.synth.scene_v0_batch(8, 8)
    want = [a.find_primitives(da, seed=3), b.find_primitives(db, seed=4)]
    got = [None, None]
    gate = threading.Barrier(2)

    def work(i, det_, d, seed):
        gate.wait()
        for _ in range(4):
            got[i] = det_.find_primitives(d, seed=seed)

    th = [threading.Thread(target=work, args=(0, a, da, 3)), threading.Thread(target=work, args=(1, b, db, 4))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for w, g in zip(want, got):
        for k in ("plane_labels", "cyl_labels", "plane_grid", "cyl_region_seg"):
            assert np.array_equal(w[k], g[k]), k
        assert w["cells"].tobytes() == g["cells"].tobytes()
        assert w["info"].tobytes() == g["info"].tobytes()
    a.close()
    b.close()


def test_split_fit_and_segment_equal_the_fused_device_run(det):
    """rs_cape_cell_fit_device + rs_cape_segment_device (with a pose solve and rs_pose_stream_wait_ransac between them, as a
    scheduler would place them) produce what rs_cape_run_device produces."""
    import torch
    depth = rs.synth.scene_v0_batch(50, 4)
    want = det.find_primitives(depth, seed=9)
    d = torch.from_numpy(depth).cuda()
    truth, cur, m = rs.synth.pose_correspondences(0)
    solver = rs.PoseOptimization(max_batch=1, max_matches=len(m))
    solver.upload(cur[None], m[None], np.array([len(m)], np.int32))
    opts = solver.options(seed=0, rng_mode=rs.abi.RS_RNG_DEVICE)
    s = torch.cuda.current_stream().cuda_stream
    p = torch.cuda.Stream()
    det.cell_fit_device(d.data_ptr(), 4, stream=s)
    det.stream_wait_fit(p.cuda_stream)
    solver.solve_device(1, opts, stream=p.cuda_stream)
    solver.stream_wait_ransac(s)
    det.segment_device(d.data_ptr(), 4, seed=9, stream=s)
    torch.cuda.synchronize()

    class _DevPtr:   # torch view of a device buffer owned by the library
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
    host, _ = rs.abi.alloc_cape_outputs(4, det.n_cells, det.max_boundary)
    dev = det.device_outputs()
    got = {}
    for k in ("cells", "plane_labels", "cyl_labels", "plane_grid", "cyl_region_seg", "info"):
        raw = torch.as_tensor(_DevPtr(getattr(dev, k), host[k].nbytes), device="cuda").cpu().numpy()
        got[k] = raw.view(host[k].dtype).reshape(host[k].shape)
    for k in ("plane_labels", "cyl_labels", "plane_grid", "cyl_region_seg"):
        assert np.array_equal(want[k], got[k]), k
    assert want["cells"].tobytes() == got["cells"].tobytes()
    assert want["info"].tobytes() == got["info"].tobytes()
    out, _ = solver.download(1)
    assert out[0]["status"] == 1
    solver.close()


@pytest.mark.parametrize("W,H,cell", [(400, 300, 20), (320, 240, 20), (640, 480, 40), (1000, 720, 40), (180, 100, 20)])
def test_other_geometries_match_oracle(W, H, cell):
    """Grids whose width is not a multiple of K1a's 8-cell work item (20, 16, 25 and 9 cells per row: the last item of a
    cell row hangs over the image and its TMA box is zero-filled), other aspect ratios, 40 px cells on a small image."""
    K = rs.synth.intrinsics(W / 640.0)
    d = rs.PrimitiveDetection(W, H, cell, *K, max_batch=3)
    depth = rs.synth.scene_v0_batch(70, 3, width=W, height=H)
    got = d.find_primitives(depth, seed=2)
    ref = ol.cape_run(depth, cell=cell, K=K, seed=2)
    for b in range(3):
        parity.assert_cells_match(ref["cells"][b], got["cells"][b])
        parity.assert_frame_match(ref, got, b)
    assert np.array_equal(ref["plane_labels"], got["plane_labels"])
    assert np.array_equal(ref["cyl_labels"], got["cyl_labels"])
    d.close()
