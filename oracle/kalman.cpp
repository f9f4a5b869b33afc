// TEST INFRASTRUCTURE — see kalman.hpp.
#include "kalman.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace oracle {

bool covariance_valid_n(const double* c, const int N)
{
    for (int i = 0; i < N * N; ++i)
        if (!std::isfinite(c[i])) return false;
    if (N == 1) return c[0] >= 0;
    double diff2 = 0, n2 = 0;   // Eigen isApprox: |a - b|^2 <= prec^2 min(|a|^2, |b|^2), prec = 1e-12
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            const double dd = c[i * N + j] - c[j * N + i];
            diff2 += dd * dd;
            n2 += c[i * N + j] * c[i * N + j];
        }
    if (!(diff2 <= 1e-12 * 1e-12 * n2)) return false;
    double a[KF_MAX][KF_MAX];   // selfadjointView<Upper>().ldlt(): diagonally pivoted LDL^T, no negative pivot
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) a[i][j] = c[std::min(i, j) * N + std::max(i, j)];
    bool neg = false;
    for (int k = 0; k < N; ++k) {
        int p = k;
        double best = std::fabs(a[k][k]);
        for (int i = k + 1; i < N; ++i)
            if (std::fabs(a[i][i]) > best) best = std::fabs(a[i][i]), p = i;
        if (p != k) {
            for (int j = 0; j < N; ++j) std::swap(a[k][j], a[p][j]);
            for (int i = 0; i < N; ++i) std::swap(a[i][k], a[i][p]);
        }
        const double dkk = a[k][k];
        if (dkk < 0) neg = true;
        if (std::fabs(dkk) <= DBL_MIN) break;
        for (int i = k + 1; i < N; ++i) {
            const double l = a[i][k] / dkk;
            for (int j = k + 1; j < N; ++j) a[i][j] -= l * a[k][j];
        }
    }
    return !neg;
}

namespace {
// symmetric matrix read through its lower triangle (selfadjointView<Lower>)
inline double symL(const double* a, int N, int i, int j) { return i >= j ? a[i * N + j] : a[j * N + i]; }

// propagate_covariance (covariances.hpp:55-64): (J symL(C) J^T) read back through its lower triangle. J is [Mo x N].
void propagate(const double* C, const double* J, int N, int Mo, double* out)
{
    double t[KF_MAX * KF_MAX];
    for (int i = 0; i < Mo; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0;
            for (int k = 0; k < N; ++k) s += J[i * N + k] * symL(C, N, k, j);
            t[i * N + j] = s;
        }
    double full[KF_MAX * KF_MAX];
    for (int i = 0; i < Mo; ++i)
        for (int j = 0; j < Mo; ++j) {
            double s = 0;
            for (int k = 0; k < N; ++k) s += t[i * N + k] * J[j * N + k];
            full[i * Mo + j] = s;
        }
    for (int i = 0; i < Mo; ++i)
        for (int j = 0; j < Mo; ++j) out[i * Mo + j] = symL(full, Mo, i, j);
}

// inverse by Gauss-Jordan with partial pivoting; returns the determinant
double invert(const double* a, int M, double* inv)
{
    double w[KF_MAX][2 * KF_MAX];
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < M; ++j) w[i][j] = a[i * M + j], w[i][M + j] = (i == j) ? 1.0 : 0.0;
    double det = 1.0;
    for (int k = 0; k < M; ++k) {
        int p = k;
        for (int i = k + 1; i < M; ++i)
            if (std::fabs(w[i][k]) > std::fabs(w[p][k])) p = i;
        if (w[p][k] == 0.0) return 0.0;
        if (p != k) {
            for (int j = 0; j < 2 * M; ++j) std::swap(w[k][j], w[p][j]);
            det = -det;
        }
        det *= w[k][k];
        const double ip = 1.0 / w[k][k];
        for (int j = 0; j < 2 * M; ++j) w[k][j] *= ip;
        for (int i = 0; i < M; ++i)
            if (i != k) {
                const double f = w[i][k];
                if (f != 0.0)
                    for (int j = 0; j < 2 * M; ++j) w[i][j] -= f * w[k][j];
            }
    }
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < M; ++j) inv[i * M + j] = w[i][M + j];
    return det;
}

// Moore-Penrose inverse of the symmetric M x M matrix read through a's lower triangle. The reference calls
// completeOrthogonalDecomposition().pseudoInverse() (kalman_filter.hpp:73-77) when the innovation covariance has a determinant
// within DBL_EPSILON of zero - which covers the truly singular case and every well-conditioned covariance whose entries are
// simply small (plane normals: variances of 1e-6 give a 4 x 4 determinant below 1e-16). The pseudo-inverse is unique once the
// rank is decided; here the rank is decided on the eigenvalues (spectral decomposition by cyclic Jacobi) with Eigen's
// relative threshold, which agrees with the pivoted-QR decision except for eigenvalues within a factor ~M of the threshold.
// The CUDA kernel (rgb-d-slam_b200/csrc/kalman.cu) runs the same operations in the same order.
void pinv_sym(const double* a, int M, double* out)
{
    // cyclic Jacobi on the symmetric matrix read through the lower triangle; V accumulates the rotations
    double A[KF_MAX][KF_MAX], V[KF_MAX][KF_MAX];
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < M; ++j) A[i][j] = symL(a, M, i, j), V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 16; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < M; ++p)
            for (int q = p + 1; q < M; ++q) off += A[p][q] * A[p][q];
        if (off == 0.0) break;
        for (int p = 0; p < M; ++p)
            for (int q = p + 1; q < M; ++q) {
                const double apq = A[p][q];
                if (apq == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < M; ++k) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < M; ++k) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < M; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    double lmax = 0.0;
    for (int i = 0; i < M; ++i) lmax = std::fmax(lmax, std::fabs(A[i][i]));
    const double tol = DBL_EPSILON * double(M) * lmax;   // Eigen's default rank threshold: epsilon * size, relative to the largest pivot
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < M; ++j) {
            double s = 0.0;
            for (int k = 0; k < M; ++k)
                if (std::fabs(A[k][k]) > tol) s += V[i][k] * (1.0 / A[k][k]) * V[j][k];
            out[i * M + j] = s;
        }
}
}  // namespace

int kalman_new_state(const int N, const int M, const double* F, const double* H, const double* Q, const double* x,
                     const double* P, const double* z, const double* R, double* x_out, double* P_out)
{
    if (!covariance_valid_n(P, N)) return -1;
    if (!covariance_valid_n(R, M)) return -2;
    double xe[KF_MAX], Pp[KF_MAX * KF_MAX], S[KF_MAX * KF_MAX], Si[KF_MAX * KF_MAX];
    for (int i = 0; i < N; ++i) {
        double s = 0;
        for (int k = 0; k < N; ++k) s += F[i * N + k] * x[k];
        xe[i] = s;
    }
    propagate(P, F, N, N, Pp);
    for (int i = 0; i < N * N; ++i) Pp[i] += Q[i];
    propagate(Pp, H, N, M, S);
    for (int i = 0; i < M * M; ++i) S[i] += R[i];
    const double det = invert(S, M, Si);
    if (std::fabs(det - 0.0) <= DBL_EPSILON) pinv_sym(S, M, Si);   // utils::double_equal(det, 0): the pseudo-inverse branch
    // K = symL(Pp) H^T S^-1
    double PHt[KF_MAX * KF_MAX], Kg[KF_MAX * KF_MAX];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < M; ++j) {
            double s = 0;
            for (int k = 0; k < N; ++k) s += symL(Pp, N, i, k) * H[j * N + k];
            PHt[i * M + j] = s;
        }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < M; ++j) {
            double s = 0;
            for (int k = 0; k < M; ++k) s += PHt[i * M + k] * Si[k * M + j];
            Kg[i * M + j] = s;
        }
    double innov[KF_MAX];
    for (int i = 0; i < M; ++i) {
        double s = 0;
        for (int k = 0; k < N; ++k) s += H[i * N + k] * xe[k];
        innov[i] = z[i] - s;
    }
    for (int i = 0; i < N; ++i) {
        double s = 0;
        for (int k = 0; k < M; ++k) s += Kg[i * M + k] * innov[k];
        x_out[i] = xe[i] + s;
    }
    // P' = (I - K H) symL(Pp), then read back through its lower triangle
    double IKH[KF_MAX * KF_MAX], full[KF_MAX * KF_MAX];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0;
            for (int k = 0; k < M; ++k) s += Kg[i * M + k] * H[k * N + j];
            IKH[i * N + j] = ((i == j) ? 1.0 : 0.0) - s;
        }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0;
            for (int k = 0; k < N; ++k) s += IKH[i * N + k] * symL(Pp, N, k, j);
            full[i * N + j] = s;
        }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) P_out[i * N + j] = symL(full, N, i, j);
    return covariance_valid_n(P_out, N) ? 0 : -4;
}

namespace {
void identity_filter(int N, double q, double* F, double* Q)
{
    std::memset(F, 0, sizeof(double) * N * N);
    std::memset(Q, 0, sizeof(double) * N * N);
    for (int i = 0; i < N; ++i) F[i * N + i] = 1.0, Q[i * N + i] = q;
}
}  // namespace

void kalman_track_point(const double x[3], const double P[9], const double z[3], const double R[9], const double q,
                        double x_out[3], double P_out[9], double* score, uint8_t* moving, int32_t* status)
{
    double F[9], Q[9];
    identity_filter(3, q, F, Q);
    std::memcpy(x_out, x, sizeof(double) * 3);
    std::memcpy(P_out, P, sizeof(double) * 9);
    *moving = 0;
    const int rc = kalman_new_state(3, 3, F, F, Q, x, P, z, R, x_out, P_out);
    *status = rc;
    if (rc != 0) {   // Point::track logs and returns -1 without touching the point
        std::memcpy(x_out, x, sizeof(double) * 3);
        std::memcpy(P_out, P, sizeof(double) * 9);
        *score = -1.0;
        return;
    }
    bool mv = false;
    for (int i = 0; i < 3; ++i) mv = mv || ((x[i] - z[i]) > std::sqrt(R[i * 3 + i]));
    *moving = mv ? 1 : 0;
    double s = 0;
    for (int i = 0; i < 3; ++i) s += (x[i] - x_out[i]) * (x[i] - x_out[i]);
    *score = std::sqrt(s);
}

void kalman_track_plane(const double x[4], const double P[16], const double z[4], const double R[16], const double q,
                        double x_out[4], double P_out[16], double* score, int32_t* status)
{
    double F[16], Q[16];
    identity_filter(4, q, F, Q);
    const int rc = kalman_new_state(4, 4, F, F, Q, x, P, z, R, x_out, P_out);
    *status = rc;
    if (rc != 0) {
        std::memcpy(x_out, x, sizeof(double) * 4);
        std::memcpy(P_out, P, sizeof(double) * 16);
        *score = -1.0;
        return;
    }
    // PlaneWorldCoordinates(res.first): the normal is re-normalised, d kept (plane_coordinates.hpp:23-40)
    const double n2 = (x_out[0] * x_out[0] + x_out[1] * x_out[1]) + x_out[2] * x_out[2];
    if (n2 > 0) {
        const double n = std::sqrt(n2);
        x_out[0] /= n, x_out[1] /= n, x_out[2] /= n;
    }
    double s = 0;
    for (int i = 0; i < 4; ++i) s += (x[i] - x_out[i]) * (x[i] - x_out[i]);
    *score = std::sqrt(s);
}

}  // namespace oracle
