"""GPU parity: batched Kalman update of matched map features through the C-ABI vs the CPU oracle (which is pinned by the
reference's own tests/test_kalman_filtering.cpp, see tests/test_oracle_kalman.py)."""
import numpy as np
import pytest

import oracle_lib as ol
import rgbd_slam_b200 as rs

pytestmark = pytest.mark.gpu


def _spd(rng, n, d, scale, floor):
    A = rng.standard_normal((n, d, d)) * scale
    return A @ A.transpose(0, 2, 1) + np.eye(d) * floor


def test_points_match_oracle_bit_exact():
    rng = np.random.default_rng(1)
    n = 5000                                                      # ~ a 16-frame batch of 300 matched points
    x = rng.uniform(-3000, 3000, (n, 3))
    P = _spd(rng, n, 3, 2.0, 0.1)
    z = x + rng.standard_normal((n, 3)) * 4
    R = _spd(rng, n, 3, 2.0, 0.1)
    got = rs.kalman_track_points(x, P, z, R)
    ref = ol.kalman_track_points(x, P, z, R)
    for g, r in zip(got, ref):
        assert g.tobytes() == r.tobytes()
    assert (got[4] == 0).all() and got[3].any() and not got[3].all()


def test_planes_match_oracle_bit_exact():
    rng = np.random.default_rng(2)
    n = 700
    nrm = rng.standard_normal((n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    x = np.concatenate([nrm, rng.uniform(500, 3000, (n, 1))], axis=1)
    z = x + np.concatenate([rng.standard_normal((n, 3)) * 0.01, rng.standard_normal((n, 1)) * 5], axis=1)
    P = _spd(rng, n, 4, 0.05, 1e-3)
    R = _spd(rng, n, 4, 0.05, 1e-3)
    got = rs.kalman_track_planes(x, P, z, R)
    ref = ol.kalman_track_planes(x, P, z, R)
    for g, r in zip(got, ref):
        assert g.tobytes() == r.tobytes()
    np.testing.assert_allclose(np.linalg.norm(got[0][:, :3], axis=1), 1.0, atol=1e-12)


def test_refused_updates_and_edge_cases():
    rng = np.random.default_rng(3)
    n = 8
    x = rng.uniform(-100, 100, (n, 3))
    P = _spd(rng, n, 3, 1.0, 0.1)
    z = x + 1.0
    R = _spd(rng, n, 3, 1.0, 0.1)
    P[1] = np.array([[1.0, 2.0, 0], [2.0, 1.0, 0], [0, 0, 1.0]])      # indefinite state covariance
    R[2][0, 1] += 0.5                                                  # asymmetric measurement covariance
    P[3][1, 1] = np.nan
    P[4] = 0.0
    R[4] = 0.0                                                         # singular innovation (only the process noise left)
    got = rs.kalman_track_points(x, P, z, R, process_noise=0.0)
    ref = ol.kalman_track_points(x, P, z, R, process_noise=0.0)
    assert np.array_equal(got[4], ref[4])
    assert got[4][1] == -1 and got[4][2] == -2 and got[4][3] == -1 and got[4][4] == -3
    for i in (1, 2, 4):                                                # refused: feature unchanged, score -1
        assert np.array_equal(got[0][i], x[i]) and got[2][i] == -1.0
    ok = got[4] == 0
    assert got[0][ok].tobytes() == ref[0][ok].tobytes() and got[1][ok].tobytes() == ref[1][ok].tobytes()
    # empty batch is a no-op; wrong shapes are rejected on the host
    e = rs.kalman_track_points(np.zeros((0, 3)), np.zeros((0, 3, 3)), np.zeros((0, 3)), np.zeros((0, 3, 3)))
    assert e[0].shape == (0, 3)
    with pytest.raises(ValueError):
        rs.kalman_track_planes(np.zeros((2, 3)), np.zeros((2, 3, 3)), np.zeros((2, 3)), np.zeros((2, 3, 3)))


def test_repeated_updates_converge_like_the_reference_filter():
    # tests/test_kalman_filtering.cpp:58-104 (temperature in a tank) on every axis of 3-D points, through the GPU path
    ms = (49.95, 49.967, 50.1, 50.106, 49.992, 49.819, 49.933, 50.007, 50.023, 49.99)
    x = np.full((4, 3), 10.0)
    P = np.tile(np.eye(3) * 100.0 * 100.0, (4, 1, 1))
    R = np.tile(np.eye(3) * 0.1 * 0.1, (4, 1, 1))
    for m in ms:
        x, P, score, moving, status = rs.kalman_track_points(x, P, np.full((4, 3), m), R, process_noise=0.0)
        assert (status == 0).all()
    assert np.abs(x - 50).max() < 0.05
