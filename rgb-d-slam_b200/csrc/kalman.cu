// KF `kalman_track` — batched Kalman update of matched map features, one thread per feature.
//
// Replaces  tracking::SharedKalmanFilter<N, N>::get_new_state (src/tracking/kalman_filter.hpp:46-118) as instantiated by
//           tracking::Point::track (point_with_tracking.cpp:32-84: N = 3, identity dynamics / output, Q = 0.001 I) and
//           tracking::Plane::track (plane_with_tracking.cpp:16-59,81-95: N = 4, Q = 1e-6 I, normal re-normalised), with
//           utils::is_covariance_valid / propagate_covariance (covariances.hpp:13-64).
// The reference updates one feature at a time on the host after the pose solve (a few hundred 3x3 / 4x4 problems per frame);
// here a frame's (or a batch of frames') matched features are one launch. FP64, same operation order as the restated
// CPU checker used by the tests, compiled with -fmad=false, so results are bit-identical to it.
#include <float.h>

#include "common.cuh"
#include "kalman_internal.cuh"

namespace rs {

namespace {

template <int N>
__device__ bool covariance_valid_n(const double* c)
{
    for (int i = 0; i < N * N; ++i)
        if (!isfinite(c[i])) return false;
    double diff2 = 0, n2 = 0;
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            const double dd = c[i * N + j] - c[j * N + i];
            diff2 += dd * dd;
            n2 += c[i * N + j] * c[i * N + j];
        }
    if (!(diff2 <= 1e-12 * 1e-12 * n2)) return false;
    double a[N][N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) a[i][j] = c[(i < j ? i : j) * N + (i < j ? j : i)];
    bool neg = false;
    for (int k = 0; k < N; ++k) {
        int p = k;
        double best = fabs(a[k][k]);
        for (int i = k + 1; i < N; ++i)
            if (fabs(a[i][i]) > best) best = fabs(a[i][i]), p = i;
        if (p != k) {
            for (int j = 0; j < N; ++j) {
                const double t = a[k][j];
                a[k][j] = a[p][j], a[p][j] = t;
            }
            for (int i = 0; i < N; ++i) {
                const double t = a[i][k];
                a[i][k] = a[i][p], a[i][p] = t;
            }
        }
        const double dkk = a[k][k];
        if (dkk < 0) neg = true;
        if (fabs(dkk) <= DBL_MIN) break;
        for (int i = k + 1; i < N; ++i) {
            const double l = a[i][k] / dkk;
            for (int j = k + 1; j < N; ++j) a[i][j] -= l * a[k][j];
        }
    }
    return !neg;
}

template <int N>
__device__ __forceinline__ double symL(const double* a, int i, int j)
{
    return i >= j ? a[i * N + j] : a[j * N + i];
}

// (J symL(C) J^T) read back through its lower triangle, J = identity: the reference still performs the two products
template <int N>
__device__ void propagate_identity(const double* C, double* out)
{
    double t[N * N], full[N * N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0;
            for (int k = 0; k < N; ++k) s += ((i == k) ? 1.0 : 0.0) * symL<N>(C, k, j);
            t[i * N + j] = s;
        }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0;
            for (int k = 0; k < N; ++k) s += t[i * N + k] * ((j == k) ? 1.0 : 0.0);
            full[i * N + j] = s;
        }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) out[i * N + j] = symL<N>(full, i, j);
}

// Gauss-Jordan with partial pivoting; returns the determinant
template <int N>
__device__ double invert(const double* a, double* inv)
{
    double w[N][2 * N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) w[i][j] = a[i * N + j], w[i][N + j] = (i == j) ? 1.0 : 0.0;
    double det = 1.0;
    for (int k = 0; k < N; ++k) {
        int p = k;
        for (int i = k + 1; i < N; ++i)
            if (fabs(w[i][k]) > fabs(w[p][k])) p = i;
        if (w[p][k] == 0.0) return 0.0;
        if (p != k) {
            for (int j = 0; j < 2 * N; ++j) {
                const double t = w[k][j];
                w[k][j] = w[p][j], w[p][j] = t;
            }
            det = -det;
        }
        det *= w[k][k];
        const double ip = 1.0 / w[k][k];
        for (int j = 0; j < 2 * N; ++j) w[k][j] *= ip;
        for (int i = 0; i < N; ++i)
            if (i != k) {
                const double f = w[i][k];
                if (f != 0.0)
                    for (int j = 0; j < 2 * N; ++j) w[i][j] -= f * w[k][j];
            }
    }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) inv[i * N + j] = w[i][N + j];
    return det;
}

// Moore-Penrose inverse of the symmetric matrix read through a's lower triangle: the reference's pseudo-inverse branch
// (kalman_filter.hpp:73-77, taken when |det| <= DBL_EPSILON - singular innovations, and well-conditioned ones with small
// entries such as plane normals). Spectral decomposition by cyclic Jacobi, Eigen's relative rank threshold; operation for
// operation what the oracle's pinv_sym does.
template <int N>
__device__ void pinv_sym(const double* a, double* out)
{
    // cyclic Jacobi on the symmetric matrix read through the lower triangle; V accumulates the rotations
    double A[N][N], V[N][N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) A[i][j] = symL<N>(a, i, j), V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 16; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < N; ++p)
            for (int q = p + 1; q < N; ++q) off += A[p][q] * A[p][q];
        if (off == 0.0) break;
        for (int p = 0; p < N; ++p)
            for (int q = p + 1; q < N; ++q) {
                const double apq = A[p][q];
                if (apq == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < N; ++k) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < N; ++k) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < N; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    double lmax = 0.0;
    for (int i = 0; i < N; ++i) lmax = fmax(lmax, fabs(A[i][i]));
    const double tol = DBL_EPSILON * double(N) * lmax;   // Eigen's default rank threshold: epsilon * size, relative to the largest pivot
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0.0;
            for (int k = 0; k < N; ++k)
                if (fabs(A[k][k]) > tol) s += V[i][k] * (1.0 / A[k][k]) * V[j][k];
            out[i * N + j] = s;
        }
}

// get_new_state with F = H = I, Q = q I. Returns 0 or a negative status (see the header).
template <int N>
__device__ int new_state_identity(const double* x, const double* P, const double* z, const double* R, const double q,
                                  double* x_out, double* P_out)
{
    if (!covariance_valid_n<N>(P)) return -1;
    if (!covariance_valid_n<N>(R)) return -2;
    double xe[N], Pp[N * N], S[N * N], Si[N * N];
    for (int i = 0; i < N; ++i) {
        double s = 0;
        for (int k = 0; k < N; ++k) s += ((i == k) ? 1.0 : 0.0) * x[k];
        xe[i] = s;
    }
    propagate_identity<N>(P, Pp);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) Pp[i * N + j] += (i == j) ? q : 0.0;
    propagate_identity<N>(Pp, S);
    for (int i = 0; i < N * N; ++i) S[i] += R[i];
    const double det = invert<N>(S, Si);
    if (fabs(det - 0.0) <= DBL_EPSILON) pinv_sym<N>(S, Si);   // utils::double_equal(det, 0): the pseudo-inverse branch
    double PHt[N * N], Kg[N * N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0;
            for (int k = 0; k < N; ++k) s += symL<N>(Pp, i, k) * ((j == k) ? 1.0 : 0.0);
            PHt[i * N + j] = s;
        }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0;
            for (int k = 0; k < N; ++k) s += PHt[i * N + k] * Si[k * N + j];
            Kg[i * N + j] = s;
        }
    double innov[N];
    for (int i = 0; i < N; ++i) {
        double s = 0;
        for (int k = 0; k < N; ++k) s += ((i == k) ? 1.0 : 0.0) * xe[k];
        innov[i] = z[i] - s;
    }
    for (int i = 0; i < N; ++i) {
        double s = 0;
        for (int k = 0; k < N; ++k) s += Kg[i * N + k] * innov[k];
        x_out[i] = xe[i] + s;
    }
    double IKH[N * N], full[N * N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0;
            for (int k = 0; k < N; ++k) s += Kg[i * N + k] * ((k == j) ? 1.0 : 0.0);
            IKH[i * N + j] = ((i == j) ? 1.0 : 0.0) - s;
        }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0;
            for (int k = 0; k < N; ++k) s += IKH[i * N + k] * symL<N>(Pp, k, j);
            full[i * N + j] = s;
        }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) P_out[i * N + j] = symL<N>(full, i, j);
    return covariance_valid_n<N>(P_out) ? 0 : -4;
}

template <int N>
__global__ void __launch_bounds__(128) kalman_track_kernel(const KalmanBatch b)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n) return;
    double x[N], P[N * N], z[N], R[N * N], xo[N], Po[N * N];
    for (int k = 0; k < N; ++k) x[k] = b.state[size_t(i) * N + k], z[k] = b.meas[size_t(i) * N + k];
    for (int k = 0; k < N * N; ++k) P[k] = b.cov[size_t(i) * N * N + k], R[k] = b.meas_cov[size_t(i) * N * N + k];
    const int rc = new_state_identity<N>(x, P, z, R, b.process_noise, xo, Po);
    double score = -1.0;
    unsigned char moving = 0;
    if (rc != 0) {   // the reference logs and leaves the feature untouched
        for (int k = 0; k < N; ++k) xo[k] = x[k];
        for (int k = 0; k < N * N; ++k) Po[k] = P[k];
    }
    else {
        if (N == 3) {   // Point::track: moved beyond the detection's own uncertainty (:51-54)
            bool mv = false;
            for (int k = 0; k < N; ++k) mv = mv || ((x[k] - z[k]) > sqrt(R[k * N + k]));
            moving = mv ? 1 : 0;
        }
        else {          // Plane::track: PlaneWorldCoordinates re-normalises the normal (plane_coordinates.hpp:23-40)
            const double n2 = (xo[0] * xo[0] + xo[1] * xo[1]) + xo[2] * xo[2];
            if (n2 > 0) {
                const double n = sqrt(n2);
                xo[0] /= n, xo[1] /= n, xo[2] /= n;
            }
        }
        double s = 0;
        for (int k = 0; k < N; ++k) s += (x[k] - xo[k]) * (x[k] - xo[k]);
        score = sqrt(s);
    }
    for (int k = 0; k < N; ++k) b.out_state[size_t(i) * N + k] = xo[k];
    for (int k = 0; k < N * N; ++k) b.out_cov[size_t(i) * N * N + k] = Po[k];
    b.out_score[i] = score;
    if (b.out_moving) b.out_moving[i] = moving;
    b.out_status[i] = rc;
}

}  // namespace

int launch_kalman_track(const KalmanBatch& b, int dim, cudaStream_t stream)
{
    if (b.n <= 0) return RS_OK;
    const int grid = (b.n + 127) / 128;
    if (dim == 3)
        kalman_track_kernel<3><<<grid, 128, 0, stream>>>(b);
    else if (dim == 4)
        kalman_track_kernel<4><<<grid, 128, 0, stream>>>(b);
    else {
        set_last_error("kalman_track: state dimension must be 3 (points) or 4 (planes)");
        return RS_ERR_INVALID_ARG;
    }
    RS_LAUNCH_CHECK();
    return RS_OK;
}

}  // namespace rs
