"""CPU: the Kalman-update oracle (oracle/kalman.cpp) pinned by the reference's own known-answer tests,
tests/test_kalman_filtering.cpp, restated with the reference's inputs and tolerances."""
import numpy as np

import oracle_lib as ol


class KalmanFilter:
    """tracking::KalmanFilter<N, M> (kalman_filter.hpp:124-170): init + update around get_new_state."""

    def __init__(self, F, H, Q):
        self.F, self.H, self.Q = (np.atleast_2d(np.asarray(a, dtype=np.float64)) for a in (F, H, Q))

    def init(self, P0, x0):
        self.P, self.x = np.atleast_2d(np.asarray(P0, np.float64)), np.atleast_1d(np.asarray(x0, np.float64))

    def update(self, z, R):
        rc, self.x, self.P = ol.kalman_new_state(self.F, self.H, self.Q, self.x, self.P, z, R)
        assert rc == 0, rc


def test_building_height_guess():            # :10-56
    kf = KalmanFilter([[1]], [[1]], [[0]])
    kf.init([[15.0 * 15.0]], [60.0])
    for m in (48.54, 47.11, 55.01, 55.15, 49.89, 40.85, 46.72, 50.05, 51.27, 49.95):
        kf.update([m], [[25.0]])
    assert abs(kf.x[0] - 50) < 0.5


def test_temperature_in_tank():              # :58-104
    kf = KalmanFilter([[1]], [[1]], [[0]])
    kf.init([[100.0 * 100.0]], [10.0])
    for m in (49.95, 49.967, 50.1, 50.106, 49.992, 49.819, 49.933, 50.007, 50.023, 49.99):
        kf.update([m], [[0.1 * 0.1]])
    assert abs(kf.x[0] - 50) < 0.05


def test_temperature_in_heating_tank():      # :106-156
    ms = (50.45, 50.967, 51.6, 52.106, 52.492, 52.819, 53.433, 54.007, 54.523, 54.99)
    kf = KalmanFilter([[1]], [[1]], [[0.15]])
    kf.init([[100.0 * 100.0]], [10.0])
    for m in ms:
        kf.update([m], [[0.1 * 0.1]])
    assert abs(kf.x[0] - ms[-1]) < 0.05


def test_vehicule_location_estimation():     # :158-218
    dt, acc, err = 1.0, 0.2, 3.0
    blk = np.array([[1, dt, 0.5 * dt * dt], [0, 1, dt], [0, 0, 1]])
    F = np.zeros((6, 6))
    F[:3, :3] = blk
    F[3:, 3:] = blk
    H = np.zeros((2, 6))
    H[0, 0] = 1
    H[1, 3] = 1
    q = np.array([[dt ** 4 / 4, dt ** 3 / 2, dt ** 2 / 2], [dt ** 3 / 2, dt * dt, dt], [dt ** 2 / 2, dt, 1]])
    Q = np.zeros((6, 6))
    Q[:3, :3] = q
    Q[3:, 3:] = q
    Q *= acc * acc
    kf = KalmanFilter(F, H, Q)
    kf.init(np.eye(6) * 500, np.zeros(6))
    meas = [(-393.66, 300.4), (-375.93, 301.78), (-351.04, 295.1), (-328.96, 305.19), (-299.35, 301.06), (-273.36, 302.05),
            (-245.89, 300), (-222.58, 303.57), (-198.03, 296.33), (-174.17, 297.65), (-146.32, 297.41), (-123.72, 299.61),
            (-103.47, 299.6), (-78.23, 302.39), (-52.63, 295.04), (-23.34, 300.09), (25.96, 294.72), (49.72, 298.61),
            (76.94, 294.64), (95.38, 284.88), (119.83, 272.82), (144.01, 264.93), (161.84, 251.46), (180.56, 241.27),
            (201.42, 222.98), (222.62, 203.73), (239.4, 184.1), (252.51, 166.12), (266.26, 138.71), (271.75, 119.71),
            (277.4, 100.41), (294.12, 79.76), (301.23, 50.62), (291.8, 32.99), (299.89, 2.14)]
    for m in meas:
        kf.update(m, np.eye(2) * err * err)
    assert abs(kf.x[0] - meas[-1][0]) < 1.5 and abs(kf.x[3] - meas[-1][1]) < 1.5


def test_1d_projectile_motion_no_noise():    # :251-312 (measurement noise 0: the trajectory is deterministic)
    dt, g = 1.0 / 30, -9.81
    F = [[1, dt, 0], [0, 1, dt], [0, 0, 1]]
    H = [[1, 0, 0]]
    Q = [[.05, .05, .0], [.05, .05, .0], [.0, .0, .0]]
    P0 = [[.1, .1, .1], [.1, 10000, 10], [.1, 10, 100]]
    traj, last = [0.0], 0.0
    for _ in range(100):
        last = last + 0.5 * g * dt * dt
        traj.append(last)
    kf = KalmanFilter(F, H, Q)
    kf.init(P0, [0, 0, g])
    for m in traj:
        kf.update([m], [[0.01 * 0.01]])
    assert abs(kf.x[0] - traj[-1]) < 0.001


def test_invalid_covariances_are_refused():
    F = np.eye(3)
    bad = np.array([[1.0, 2.0, 0], [2.0, 1.0, 0], [0, 0, 1.0]])      # indefinite
    assert ol.kalman_new_state(F, F, F * 1e-3, np.zeros(3), bad, np.zeros(3), np.eye(3))[0] == -1
    assert ol.kalman_new_state(F, F, F * 1e-3, np.zeros(3), np.eye(3), np.zeros(3), bad)[0] == -2
    asym = np.eye(3)
    asym[0, 1] = 0.5
    assert ol.kalman_new_state(F, F, F * 1e-3, np.zeros(3), asym, np.zeros(3), np.eye(3))[0] == -1


def test_point_and_plane_tracking_semantics():
    rng = np.random.default_rng(0)
    n = 64
    x = rng.uniform(-2000, 2000, (n, 3))
    A = rng.standard_normal((n, 3, 3))
    P = A @ A.transpose(0, 2, 1) + np.eye(3) * 0.5
    z = x + rng.standard_normal((n, 3)) * 3
    B = rng.standard_normal((n, 3, 3))
    R = B @ B.transpose(0, 2, 1) + np.eye(3) * 0.5
    xo, Po, score, moving, status = ol.kalman_track_points(x, P, z, R)
    assert (status == 0).all() and (score >= 0).all()
    # identity dynamics / output: the textbook update
    for i in range(n):
        Pp = P[i] + np.eye(3) * 0.001
        K = Pp @ np.linalg.inv(Pp + R[i])
        np.testing.assert_allclose(xo[i], x[i] + K @ (z[i] - x[i]), rtol=1e-10, atol=1e-9)
        np.testing.assert_allclose(Po[i], (np.eye(3) - K) @ Pp, rtol=1e-8, atol=1e-9)
        assert moving[i] == int(((x[i] - z[i]) > np.sqrt(np.diag(R[i]))).any())
        assert abs(score[i] - np.linalg.norm(x[i] - xo[i])) < 1e-9
    # planes: the filtered normal comes back unit length
    nrm = rng.standard_normal((n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    xp = np.concatenate([nrm, rng.uniform(500, 3000, (n, 1))], axis=1)
    zp = xp + np.concatenate([rng.standard_normal((n, 3)) * 0.01, rng.standard_normal((n, 1)) * 5], axis=1)
    C = rng.standard_normal((n, 4, 4)) * 0.05
    Pq = C @ C.transpose(0, 2, 1) + np.eye(4) * 1e-3
    xo, Po, score, status = ol.kalman_track_planes(xp, Pq, zp, Pq)
    assert (status == 0).all()
    np.testing.assert_allclose(np.linalg.norm(xo[:, :3], axis=1), 1.0, atol=1e-12)


def _rank_deficient(rng, rank):
    """3x3 covariance of the given rank whose null space is spanned by coordinate axes: the determinant is an exact 0 (a
    rotated null space leaves det ~ 1e-14 > DBL_EPSILON, and the reference then inverts a numerically singular matrix)."""
    a = rng.standard_normal((rank, rank))
    blk = a @ a.T + np.eye(rank) * 0.5
    m = np.zeros((3, 3))
    axes = rng.permutation(3)[:rank]
    m[np.ix_(axes, axes)] = blk
    return (m + m.T) / 2, axes


def test_pseudo_inverse_branch_against_numpy():
    """kalman_filter.hpp:73-77. (a) a rank-deficient innovation covariance: the update equals the textbook one with
    numpy.linalg.pinv; (b) a well-conditioned covariance with small entries (|det| <= DBL_EPSILON all the same): the
    pseudo-inverse is the inverse."""
    rng = np.random.default_rng(4)
    I3 = np.eye(3)
    for trial in range(50):
        P, axes = _rank_deficient(rng, 1 + trial % 2)
        R = P * 1.7
        x, z = rng.uniform(-10, 10, 3), rng.uniform(-10, 10, 3)
        rc, xo, Po = ol.kalman_new_state(I3, I3, I3 * 0.0, x, P, z, R)
        assert rc == 0
        S = P + R
        assert abs(np.linalg.det(S)) <= 2.2e-16
        K = P @ np.linalg.pinv(S, rcond=3 * 2.2e-16, hermitian=True)
        np.testing.assert_allclose(xo, x + K @ (z - x), rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(Po, (I3 - K) @ P, rtol=1e-9, atol=1e-9)
    I4 = np.eye(4)
    in_branch = 0
    for trial in range(50):
        C = rng.standard_normal((4, 4)) * np.array([5e-4, 5e-4, 5e-4, 1.0])[:, None]
        P = C @ C.T + np.diag([1e-7, 1e-7, 1e-7, 1e-2])
        x, z = rng.standard_normal(4), rng.standard_normal(4)
        rc, xo, Po = ol.kalman_new_state(I4, I4, I4 * 1e-9, x, P, z, P)
        assert rc == 0
        Pp = P + I4 * 1e-9
        S = Pp + P
        in_branch += abs(np.linalg.det(S)) <= 2.2e-16
        K = Pp @ np.linalg.inv(S)
        np.testing.assert_allclose(xo, x + K @ (z - x), rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(Po, (I4 - K) @ Pp, rtol=1e-6, atol=1e-14)
    assert in_branch > 25
