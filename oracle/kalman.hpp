// TEST INFRASTRUCTURE — CPU oracle (see linalg.hpp header). Restates the reference's Kalman update of matched map features:
//   src/tracking/kalman_filter.hpp:46-118 (SharedKalmanFilter<N, M>::get_new_state), src/utils/covariances.hpp:13-64
//   (is_covariance_valid, propagate_covariance), src/tracking/point_with_tracking.cpp:32-84 (Point::track, N = M = 3),
//   src/tracking/plane_with_tracking.cpp:16-59,81-95 (Plane::track without the polygon merge, N = M = 4).
// Pinned by the reference's own known-answer tests tests/test_kalman_filtering.cpp (restated in tests/test_oracle_kalman.py).
// Eigen's fixed-size inverse() (cofactors up to 4x4, partial-pivot LU above) is un-vendored: restated as Gauss-Jordan with
// partial pivoting, so agreement with the reference binary is to rounding, not bit for bit. The pseudo-inverse branch taken
// when det(innovation) == 0 is not restated: such an update reports status -3.
#pragma once
#include <cstdint>

namespace oracle {

constexpr int KF_MAX = 6;

// is_covariance_valid<N> (covariances.hpp:13-44)
bool covariance_valid_n(const double* c, int N);

// get_new_state: x [N], P [N x N], z [M], R [M x M], F [N x N], H [M x N], Q [N x N], row-major. Returns 0, or a negative
// status: -1 invalid state covariance, -2 invalid measurement covariance, -3 singular innovation, -4 invalid result.
int kalman_new_state(int N, int M, const double* F, const double* H, const double* Q, const double* x, const double* P,
                     const double* z, const double* R, double* x_out, double* P_out);

// Point::track (process noise 0.001 I, identity dynamics / output). score < 0: the update was refused (invalid covariance).
void kalman_track_point(const double x[3], const double P[9], const double z[3], const double R[9], double process_noise,
                        double x_out[3], double P_out[9], double* score, uint8_t* moving, int32_t* status);
// Plane::track's filter step (process noise 1e-6 I): the new (normal, d) with the normal re-normalised by PlaneCoordinates.
void kalman_track_plane(const double x[4], const double P[16], const double z[4], const double R[16], double process_noise,
                        double x_out[4], double P_out[16], double* score, int32_t* status);

}  // namespace oracle
