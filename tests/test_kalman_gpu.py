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
    R[4] = 0.0                                                         # zero innovation covariance: pseudo-inverse 0, nothing moves
    got = rs.kalman_track_points(x, P, z, R, process_noise=0.0)
    ref = ol.kalman_track_points(x, P, z, R, process_noise=0.0)
    assert np.array_equal(got[4], ref[4])
    assert got[4][1] == -1 and got[4][2] == -2 and got[4][3] == -1 and got[4][4] == 0
    for i in (1, 2):                                                   # refused: feature unchanged, score -1
        assert np.array_equal(got[0][i], x[i]) and got[2][i] == -1.0
    assert np.array_equal(got[0][4], x[4]) and got[2][4] == 0.0        # accepted, gain 0
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


def test_pseudo_inverse_branch_matches_the_oracle_bit_for_bit():
    """kalman_filter.hpp:73-77: |det(innovation)| <= DBL_EPSILON -> completeOrthogonalDecomposition().pseudoInverse().
    Rank-deficient innovations (an axis with no uncertainty at all) and well-conditioned ones with small entries (plane
    normals) both go there; GPU == oracle bit for bit, oracle == numpy.linalg.pinv in tests/test_oracle_kalman.py."""
    rng = np.random.default_rng(9)
    n = 256
    x = rng.uniform(-500, 500, (n, 3))
    z = x + rng.standard_normal((n, 3))
    P, R = np.zeros((n, 3, 3)), np.zeros((n, 3, 3))
    for i in range(n):
        rank = 1 + i % 2                                                # rank-1 and rank-2 covariances, null space on coordinate axes
        a = rng.standard_normal((rank, rank))
        axes = rng.permutation(3)[:rank]
        blk = a @ a.T + np.eye(rank) * 0.5
        P[i][np.ix_(axes, axes)] = blk
        R[i][np.ix_(axes, axes)] = blk * rng.uniform(0.5, 2.0)
    got = rs.kalman_track_points(x, P, z, R, process_noise=0.0)
    ref = ol.kalman_track_points(x, P, z, R, process_noise=0.0)
    assert np.array_equal(got[4], ref[4])
    ok = ref[4] == 0
    assert ok.all()
    assert got[0][ok].tobytes() == ref[0][ok].tobytes() and got[1][ok].tobytes() == ref[1][ok].tobytes()
    # planes with realistic (small) covariances: det ~ 1e-20, full rank
    npl = 512
    nrm = rng.standard_normal((npl, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    xp = np.concatenate([nrm, rng.uniform(500, 3000, (npl, 1))], axis=1)
    zp = xp + np.concatenate([rng.standard_normal((npl, 3)) * 1e-3, rng.standard_normal((npl, 1))], axis=1)
    C = rng.standard_normal((npl, 4, 4)) * np.array([5e-4, 5e-4, 5e-4, 1.0])[None, :, None]
    Pq = C @ C.transpose(0, 2, 1) + np.diag([1e-7, 1e-7, 1e-7, 1e-2])
    assert (np.abs(np.linalg.det(2 * Pq)) < 2.2e-16).mean() > 0.5     # most of them take the branch
    gp = rs.kalman_track_planes(xp, Pq, zp, Pq, process_noise=1e-9)
    rp = ol.kalman_track_planes(xp, Pq, zp, Pq, process_noise=1e-9)
    assert np.array_equal(gp[3], rp[3]) and (rp[3] == 0).mean() > 0.95
    okp = rp[3] == 0
    assert gp[0][okp].tobytes() == rp[0][okp].tobytes() and gp[1][okp].tobytes() == rp[1][okp].tobytes()
