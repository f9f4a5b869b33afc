// TEST INFRASTRUCTURE — CPU oracle (see pose.hpp). Every function cites the reference lines it restates.
#include "pose.hpp"

#include <algorithm>
#include <array>
#include <cassert>
#include <numeric>

namespace oracle {

namespace {
// parameters.hpp:23-44
constexpr float kMaxRetroprojectionErrorForPointInliers_px = 3.0f;
constexpr float kMaxRetroprojectionErrorForPlaneInliers_mm = 50.0f;
constexpr float kMaxRetroprojectionErrorForPlaneInliersNormal = 0.2f;
constexpr double kMinimumInliersProportionForEarlyStop = 0.80f;  // double initialised from a float literal
constexpr float kProbabilityOfSuccess = 0.8f;
constexpr float kInlierProportion = 0.65f;
constexpr float kFeatureTrustCount = 10.0f;
constexpr unsigned kMinimumPointForOptimization = 5;
constexpr unsigned kMinimumPlanesForOptimization = 3;
constexpr unsigned kMinimumPoint2dForOptimization = 5;                    // parameters.hpp:41-42
constexpr float kMaxRetroprojectionErrorForPoint2DInliers_px = 3.0f;      // parameters.hpp:23-24

// ---- small dense helpers for the m x 6 problem ------------------------------------------------
constexpr int N = 6;

double naive_norm(const double* v, int n)
{
    double s = 0;
    for (int i = 0; i < n; ++i) s += v[i] * v[i];
    return std::sqrt(s);
}
// Eigen/src/Core/StableNorm.h: stableNorm() for a single block (scale by max |coeff|).
double stable_norm(const double* v, int n)
{
    if (n == 1) return std::fabs(v[0]);
    double maxCoeff = 0;
    for (int i = 0; i < n; ++i) maxCoeff = std::max(maxCoeff, std::fabs(v[i]));
    if (!(maxCoeff > 0)) return maxCoeff != maxCoeff ? maxCoeff : 0.0;
    if (maxCoeff > DBL_MAX) return maxCoeff;
    const double invScale = 1.0 / maxCoeff;
    double ssq = 0;
    for (int i = 0; i < n; ++i) {
        const double t = v[i] * invScale;
        ssq += t * t;
    }
    return maxCoeff * std::sqrt(ssq);
}

// Eigen/src/QR/ColPivHouseholderQR.h computeInPlace() on a column-major m x 6 matrix.
struct ColPivQR {
    int m = 0;
    std::vector<double> qr;  // column-major m x N
    double hcoeffs[N];
    int perm[N];             // colsPermutation().indices()
    int nonzero_pivots = N;
    double maxpivot = 0;
    double& at(int r, int c) { return qr[size_t(c) * m + r]; }
    double at(int r, int c) const { return qr[size_t(c) * m + r]; }

    void compute(const std::vector<double>& a, int rows)
    {
        m = rows;
        qr = a;
        const int cols = N, size = std::min(rows, cols);
        int transpositions[N];
        double colNormsUpdated[N], colNormsDirect[N];
        for (int k = 0; k < cols; ++k) {
            colNormsDirect[k] = naive_norm(&qr[size_t(k) * m], m);
            colNormsUpdated[k] = colNormsDirect[k];
        }
        double maxNorm = colNormsUpdated[0];
        for (int k = 1; k < cols; ++k) maxNorm = std::max(maxNorm, colNormsUpdated[k]);
        const double th = maxNorm * DBL_EPSILON;
        const double threshold_helper = (th * th) / double(rows);
        const double norm_downdate_threshold = std::sqrt(DBL_EPSILON);
        nonzero_pivots = size;
        maxpivot = 0;
        for (int k = 0; k < size; ++k) {
            int biggest = k;
            double best = colNormsUpdated[k];
            for (int j = k + 1; j < cols; ++j)
                if (colNormsUpdated[j] > best) {
                    best = colNormsUpdated[j];
                    biggest = j;
                }
            const double biggest_col_sq_norm = best * best;
            if (nonzero_pivots == size && biggest_col_sq_norm < threshold_helper * double(rows - k)) nonzero_pivots = k;
            transpositions[k] = biggest;
            if (k != biggest) {
                for (int r = 0; r < m; ++r) std::swap(at(r, k), at(r, biggest));
                std::swap(colNormsUpdated[k], colNormsUpdated[biggest]);
                std::swap(colNormsDirect[k], colNormsDirect[biggest]);
            }
            // makeHouseholderInPlace on col(k).tail(rows-k)  (Householder.h)
            double tailSqNorm = 0;
            for (int r = k + 1; r < m; ++r) tailSqNorm += at(r, k) * at(r, k);
            const double c0 = at(k, k);
            double tau, beta;
            if (rows - k == 1 || tailSqNorm <= DBL_MIN) {
                tau = 0;
                beta = c0;
                for (int r = k + 1; r < m; ++r) at(r, k) = 0;
            }
            else {
                beta = std::sqrt(c0 * c0 + tailSqNorm);
                if (c0 >= 0) beta = -beta;
                const double den = c0 - beta;
                for (int r = k + 1; r < m; ++r) at(r, k) = at(r, k) / den;
                tau = (beta - c0) / beta;
            }
            hcoeffs[k] = tau;
            at(k, k) = beta;
            if (std::fabs(beta) > maxpivot) maxpivot = std::fabs(beta);
            // applyHouseholderOnTheLeft to bottomRightCorner(rows-k, cols-k-1)
            if (rows - k == 1) {
                for (int j = k + 1; j < cols; ++j) at(k, j) *= (1.0 - tau);
            }
            else if (tau != 0) {
                for (int j = k + 1; j < cols; ++j) {
                    double t = 0;
                    for (int r = k + 1; r < m; ++r) t += at(r, k) * at(r, j);
                    t += at(k, j);
                    at(k, j) -= tau * t;
                    for (int r = k + 1; r < m; ++r) at(r, j) -= tau * at(r, k) * t;
                }
            }
            // norm downdate (LAPACK xGEQPF style)
            for (int j = k + 1; j < cols; ++j) {
                if (colNormsUpdated[j] != 0) {
                    double t = std::fabs(at(k, j)) / colNormsUpdated[j];
                    t = (1.0 + t) * (1.0 - t);
                    t = t < 0 ? 0 : t;
                    const double r2 = colNormsUpdated[j] / colNormsDirect[j];
                    const double t2 = t * (r2 * r2);
                    if (t2 <= norm_downdate_threshold) {
                        colNormsDirect[j] = (m - k - 1 > 0) ? naive_norm(&at(k + 1, j), m - k - 1) : 0.0;
                        colNormsUpdated[j] = colNormsDirect[j];
                    }
                    else {
                        colNormsUpdated[j] *= std::sqrt(t);
                    }
                }
            }
        }
        for (int k = 0; k < cols; ++k) perm[k] = k;
        for (int k = 0; k < size; ++k) std::swap(perm[k], perm[transpositions[k]]);
    }
    int rank() const
    {
        const double premultiplied = std::fabs(maxpivot) * (DBL_EPSILON * double(std::min(m, N)));
        int r = 0;
        for (int i = 0; i < nonzero_pivots; ++i) r += (std::fabs(at(i, i)) > premultiplied) ? 1 : 0;
        return r;
    }
    // b <- Q^T b (householderQ().adjoint() applied on the left)
    void apply_qt(double* b) const
    {
        const int size = std::min(m, N);
        for (int k = 0; k < size; ++k) {
            const double tau = hcoeffs[k];
            if (m - k == 1) {
                b[k] *= (1.0 - tau);
            }
            else if (tau != 0) {
                double t = 0;
                for (int r = k + 1; r < m; ++r) t += at(r, k) * b[r];
                t += b[k];
                b[k] -= tau * t;
                for (int r = k + 1; r < m; ++r) b[r] -= tau * at(r, k) * t;
            }
        }
    }
};

// unsupported/Eigen/src/NonLinearOptimization/qrsolv.h. s = n x n working copy (row-major s[i][j]).
void qrsolv(double s[N][N], const int* ipvt, const double* diag, const double* qtb, double* x, double* sdiag)
{
    double wa[N];
    for (int j = 0; j < N; ++j) {
        x[j] = s[j][j];
        wa[j] = qtb[j];
    }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < i; ++j) s[i][j] = s[j][i];
    for (int j = 0; j < N; ++j) {
        const int l = ipvt[j];
        if (diag[l] == 0.) break;
        for (int k = j; k < N; ++k) sdiag[k] = 0;
        sdiag[j] = diag[l];
        double qtbpj = 0.;
        for (int k = j; k < N; ++k) {
            const Givens g = make_givens(-s[k][k], sdiag[k]);
            s[k][k] = g.c * s[k][k] + g.s * sdiag[k];
            const double temp = g.c * wa[k] + g.s * qtbpj;
            qtbpj = -g.s * wa[k] + g.c * qtbpj;
            wa[k] = temp;
            for (int i = k + 1; i < N; ++i) {
                const double t = g.c * s[i][k] + g.s * sdiag[i];
                sdiag[i] = -g.s * s[i][k] + g.c * sdiag[i];
                s[i][k] = t;
            }
        }
    }
    int nsing;
    for (nsing = 0; nsing < N && sdiag[nsing] != 0; nsing++) {
    }
    for (int j = nsing; j < N; ++j) wa[j] = 0;
    // solve S^T (upper, stored in the lower triangle of s) * z = wa : back substitution
    for (int i = nsing - 1; i >= 0; --i) {
        double sum = wa[i];
        for (int j = i + 1; j < nsing; ++j) sum -= s[j][i] * wa[j];
        wa[i] = sum / s[i][i];
    }
    for (int j = 0; j < N; ++j) {
        sdiag[j] = s[j][j];
        s[j][j] = x[j];
    }
    for (int j = 0; j < N; ++j) x[ipvt[j]] = wa[j];
}

// unsupported/Eigen/src/NonLinearOptimization/lmpar.h: lmpar2 (ColPivHouseholderQR flavour).
void lmpar2(const ColPivQR& qr, const double* diag, const double* qtb, double delta, double& par, double* x)
{
    const double dwarf = DBL_MIN;
    double wa1[N], wa2[N];
    const int rank = qr.rank();
    for (int j = 0; j < N; ++j) wa1[j] = (j < rank) ? qtb[j] : 0.0;
    for (int i = rank - 1; i >= 0; --i) {
        double sum = wa1[i];
        for (int j = i + 1; j < rank; ++j) sum -= qr.at(i, j) * wa1[j];
        wa1[i] = sum / qr.at(i, i);
    }
    for (int j = 0; j < N; ++j) x[qr.perm[j]] = wa1[j];

    int iter = 0;
    for (int j = 0; j < N; ++j) wa2[j] = diag[j] * x[j];
    double dxnorm = naive_norm(wa2, N);
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) {
        par = 0;
        return;
    }
    double parl = 0.;
    if (rank == N) {
        for (int j = 0; j < N; ++j) wa1[j] = diag[qr.perm[j]] * wa2[qr.perm[j]] / dxnorm;
        // R^T (lower) solve, forward substitution
        for (int i = 0; i < N; ++i) {
            double sum = wa1[i];
            for (int j = 0; j < i; ++j) sum -= qr.at(j, i) * wa1[j];
            wa1[i] = sum / qr.at(i, i);
        }
        const double temp = naive_norm(wa1, N);
        parl = fp / delta / temp / temp;
    }
    for (int j = 0; j < N; ++j) {
        double sum = 0;
        for (int i = 0; i <= j; ++i) sum += qr.at(i, j) * qtb[i];
        wa1[j] = sum / diag[qr.perm[j]];
    }
    const double gnorm = stable_norm(wa1, N);
    double paru = gnorm / delta;
    if (paru == 0.) paru = dwarf / std::min(delta, 0.1);
    par = std::max(par, parl);
    par = std::min(par, paru);
    if (par == 0.) par = gnorm / dxnorm;

    double s[N][N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) s[i][j] = qr.at(i, j);
    while (true) {
        ++iter;
        if (par == 0.) par = std::max(dwarf, .001 * paru);
        const double sq = std::sqrt(par);
        for (int j = 0; j < N; ++j) wa1[j] = sq * diag[j];
        double sdiag[N];
        qrsolv(s, qr.perm, wa1, qtb, x, sdiag);
        for (int j = 0; j < N; ++j) wa2[j] = diag[j] * x[j];
        dxnorm = naive_norm(wa2, N);
        double temp = fp;
        fp = dxnorm - delta;
        if (std::fabs(fp) <= 0.1 * delta || (parl == 0. && fp <= temp && temp < 0.) || iter == 10) break;
        for (int j = 0; j < N; ++j) wa1[j] = diag[qr.perm[j]] * (wa2[qr.perm[j]] / dxnorm);
        for (int j = 0; j < N; ++j) {
            wa1[j] /= sdiag[j];
            temp = wa1[j];
            for (int i = j + 1; i < N; ++i) wa1[i] -= s[i][j] * temp;
        }
        temp = naive_norm(wa1, N);
        const double parc = fp / delta / temp / temp;
        if (fp > 0.) parl = std::max(parl, par);
        if (fp < 0.) paru = std::min(paru, par);
        par = std::max(parl, par + parc);
    }
    if (iter == 0) par = 0.;
}

}  // namespace

// unsupported/Eigen/src/NonLinearOptimization/LevenbergMarquardt.h: minimize() = minimizeInit + minimizeOneStep
// loop, with NumericalDiff<Functor, Forward>::df (unsupported/Eigen/src/NumericalDiff/NumericalDiff.h).
LMResult lm_minimize(const ResidualFn& f, const int m, double x[6], const int maxfev)
{
    LMResult res;
    const double factor = 100., ftol = std::sqrt(DBL_EPSILON), xtol = std::sqrt(DBL_EPSILON), gtol = 0.;
    const int n = N;
    if (m < n || maxfev <= 0) {
        res.status = 0;  // ImproperInputParameters
        return res;
    }
    std::vector<double> fvec(m), wa4(m), fjac(size_t(m) * n), val1(m), val2(m);
    double diag[N], qtf[N], wa1[N], wa2[N], wa3[N];
    int nfev = 1;
    f(x, fvec.data());
    double fnorm = stable_norm(fvec.data(), m);
    double par = 0., delta = 0., xnorm = 0.;
    int iter = 1;
    ColPivQR qr;
    const double eps = std::sqrt(std::max(0.0, DBL_EPSILON));

    while (true) {
        // NumericalDiff::df, Forward mode: re-evaluates f(x), then one evaluation per column
        {
            double xx[N];
            for (int j = 0; j < n; ++j) xx[j] = x[j];
            f(xx, val1.data());
            int dfev = 1;
            for (int j = 0; j < n; ++j) {
                double h = eps * std::fabs(xx[j]);
                if (h == 0.) h = eps;
                xx[j] += h;
                f(xx, val2.data());
                dfev++;
                xx[j] = x[j];
                for (int i = 0; i < m; ++i) fjac[size_t(j) * m + i] = (val2[i] - val1[i]) / h;
            }
            nfev += dfev;
        }
        for (int j = 0; j < n; ++j) wa2[j] = naive_norm(&fjac[size_t(j) * m], m);  // colwise().blueNorm()
        qr.compute(fjac, m);
        const int* perm = qr.perm;

        if (iter == 1) {
            for (int j = 0; j < n; ++j) diag[j] = (wa2[j] == 0.) ? 1. : wa2[j];
            double t[N];
            for (int j = 0; j < n; ++j) t[j] = diag[j] * x[j];
            xnorm = stable_norm(t, n);
            delta = factor * xnorm;
            if (delta == 0.) delta = factor;
        }
        wa4 = fvec;
        qr.apply_qt(wa4.data());
        for (int j = 0; j < n; ++j) qtf[j] = wa4[j];

        double gnorm = 0.;
        if (fnorm != 0.)
            for (int j = 0; j < n; ++j)
                if (wa2[perm[j]] != 0.) {
                    double sum = 0;
                    for (int i = 0; i <= j; ++i) sum += qr.at(i, j) * (qtf[i] / fnorm);
                    gnorm = std::max(gnorm, std::fabs(sum / wa2[perm[j]]));
                }
        if (gnorm <= gtol) {
            res.status = 4;  // CosinusTooSmall
            break;
        }
        for (int j = 0; j < n; ++j) diag[j] = std::max(diag[j], wa2[j]);

        double ratio = 0;
        int status = -100;  // still running
        do {
            lmpar2(qr, diag, qtf, delta, par, wa1);
            for (int j = 0; j < n; ++j) {
                wa1[j] = -wa1[j];
                wa2[j] = x[j] + wa1[j];
            }
            double t[N];
            for (int j = 0; j < n; ++j) t[j] = diag[j] * wa1[j];
            const double pnorm = stable_norm(t, n);
            if (iter == 1) delta = std::min(delta, pnorm);

            f(wa2, wa4.data());
            ++nfev;
            const double fnorm1 = stable_norm(wa4.data(), m);

            double actred = -1.;
            if (.1 * fnorm1 < fnorm) actred = 1. - (fnorm1 / fnorm) * (fnorm1 / fnorm);

            // wa3 = R * (P^-1 * wa1)
            for (int i = 0; i < n; ++i) {
                double sum = 0;
                for (int j = i; j < n; ++j) sum += qr.at(i, j) * wa1[perm[j]];
                wa3[i] = sum;
            }
            const double r1 = stable_norm(wa3, n) / fnorm;
            const double temp1 = r1 * r1;
            const double r2 = std::sqrt(par) * pnorm / fnorm;
            const double temp2 = r2 * r2;
            const double prered = temp1 + temp2 / .5;
            const double dirder = -(temp1 + temp2);
            ratio = 0.;
            if (prered != 0.) ratio = actred / prered;

            if (ratio <= .25) {
                double temp = 0;
                if (actred >= 0.) temp = .5;
                if (actred < 0.) temp = .5 * dirder / (dirder + .5 * actred);
                if (.1 * fnorm1 >= fnorm || temp < .1) temp = .1;
                delta = temp * std::min(delta, pnorm / .1);
                par /= temp;
            }
            else if (!(par != 0. && ratio < .75)) {
                delta = pnorm / .5;
                par = .5 * par;
            }

            if (ratio >= 1e-4) {
                for (int j = 0; j < n; ++j) x[j] = wa2[j];
                for (int j = 0; j < n; ++j) wa2[j] = diag[j] * x[j];
                fvec = wa4;
                xnorm = stable_norm(wa2, n);
                fnorm = fnorm1;
                ++iter;
            }

            if (std::fabs(actred) <= ftol && prered <= ftol && .5 * ratio <= 1. && delta <= xtol * xnorm)
                status = 3;
            else if (std::fabs(actred) <= ftol && prered <= ftol && .5 * ratio <= 1.)
                status = 1;
            else if (delta <= xtol * xnorm)
                status = 2;
            else if (nfev >= maxfev)
                status = 5;
            else if (std::fabs(actred) <= DBL_EPSILON && prered <= DBL_EPSILON && .5 * ratio <= 1.)
                status = 6;
            else if (delta <= DBL_EPSILON * xnorm)
                status = 7;
            else if (gnorm <= DBL_EPSILON)
                status = 8;
            if (status != -100) break;
        } while (ratio < 1e-4);
        if (status != -100) {
            res.status = status;
            break;
        }
    }
    res.nfev = nfev;
    res.iterations = iter;
    res.fnorm = fnorm;
    return res;
}

// angle_utils.cpp:6-11
void quaternion_from_euler(const double yaw, const double pitch, const double roll, double q[4])
{
    const double qx[4] = {std::cos(roll * 0.5), std::sin(roll * 0.5), 0, 0};
    const double qy[4] = {std::cos(pitch * 0.5), 0, std::sin(pitch * 0.5), 0};
    const double qz[4] = {std::cos(yaw * 0.5), 0, 0, std::sin(yaw * 0.5)};
    double t[4];
    quat_mul(qx, qy, t);
    quat_mul(t, qz, q);
}

namespace {
Mat4 transformation_matrix(const Mat3& R, const double t[3])
{
    Mat4 m;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) m(i, j) = R(i, j);
        m(i, 3) = t[i];
    }
    m(3, 0) = 0, m(3, 1) = 0, m(3, 2) = 0, m(3, 3) = 1;
    return m;
}
// camera_transformation.cpp:11-17 : static CameraToWorld
const Mat4& camera_to_world_static()
{
    static const Mat4 c = [] {
        const double EulerToRadian = M_PI / 180.0;
        double q[4];
        quaternion_from_euler(0.0, 90.0 * EulerToRadian, -90.0 * EulerToRadian, q);
        const double z[3] = {0, 0, 0};
        return transformation_matrix(quat_to_rot(q), z);
    }();
    return c;
}
}  // namespace

Mat4 world_to_camera(const double q[4], const double t[3])
{
    const Mat4 c2w = matmul(camera_to_world_static(), transformation_matrix(quat_to_rot(q), t));
    return inverse4(c2w);
}

Mat4 plane_world_to_camera(const Mat4& w2c)
{
    const Mat4 c2w = inverse4(w2c);
    Mat4 pc2w;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) pc2w(i, j) = c2w(i, j);
        pc2w(i, 3) = 0;
    }
    for (int j = 0; j < 3; ++j)
        pc2w(3, j) = ((-c2w(0, 3)) * c2w(0, j) + (-c2w(1, 3)) * c2w(1, j)) + (-c2w(2, 3)) * c2w(2, j);
    pc2w(3, 3) = 1;
    return inverse4(pc2w);
}

void coefficients_from_pose(const Pose7& p, double x[6])
{
    x[0] = p.t[0], x[1] = p.t[1], x[2] = p.t[2];
    const double divider = 1.0 / std::max(1.0 + p.q[3], 0.001);
    x[3] = p.q[0] * divider;
    x[4] = p.q[1] * divider;
    x[5] = p.q[2] * divider;
}

Pose7 pose_from_coefficients(const double x[6])
{
    Pose7 p;
    p.t[0] = x[0], p.t[1] = x[1], p.t[2] = x[2];
    const double alpha = x[3] * x[3] + x[4] * x[4] + x[5] * x[5];
    const double divider = 1.0 / (alpha + 1);
    double q[4] = {2.0 * x[3] * divider, 2.0 * x[4] * divider, 2.0 * x[5] * divider, (1 - alpha) * divider};
    // PoseBase::set_parameters normalises (pose.cpp:18-22)
    const double nn = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (nn > 0)
        for (double& v : q) v /= nn;
    for (int i = 0; i < 4; ++i) p.q[i] = q[i];
    return p;
}

void pose_vector6(const Pose7& p, double v[6])
{
    v[0] = p.t[0], v[1] = p.t[1], v[2] = p.t[2];
    euler_angles_012(quat_to_rot(p.q), v + 3);
}

double feature_score(const rs_match& f)
{
    if (f.type == RS_FEAT_POINT) return 1.0 / kMinimumPointForOptimization;
    if (f.type == RS_FEAT_POINT2D) return 1.0 / kMinimumPoint2dForOptimization;  // map_point2d.cpp:27-31
    return 1.0 / kMinimumPlanesForOptimization;
}

// get_feature_part_count (map_point.cpp:25, map_point2d.cpp:25, map_primitive.cpp:25)
int feature_parts(const rs_match& f) { return f.type == RS_FEAT_PLANE ? 3 : 2; }

int residual_count(const std::vector<rs_match>& feats)
{
    int m = 0;
    for (const rs_match& f : feats) m += feature_parts(f);
    return m;
}

namespace {

// WorldCoordinate::get_signed_distance_2D_px (point_coordinates.cpp:245-260)
void point_signed_distance(const Intrinsics& K, const rs_match& f, const Mat4& w2c, double out[2])
{
    const double X = f.map[0], Y = f.map[1], Z = f.map[2];
    double h[4];
    for (int i = 0; i < 4; ++i) h[i] = ((w2c(i, 0) * X + w2c(i, 1) * Y) + w2c(i, 2) * Z) + w2c(i, 3) * 1.0;
    const double xc = h[0] / h[3], yc = h[1] / h[3], zc = h[2] / h[3];
    // CameraCoordinate::to_screen_coordinates: 1/z * (K * p).head<2>()
    const double inv = 1.0 / zc;
    const double u = inv * ((K.fx * xc + 0.0 * yc) + K.cx * zc);
    const double v = inv * ((0.0 * xc + K.fy * yc) + K.cy * zc);
    if (u != u || v != v) {
        out[0] = DBL_MAX;
        out[1] = DBL_MAX;
        return;
    }
    out[0] = f.obs[0] - u;
    out[1] = f.obs[1] - v;
}

// WorldCoordinate::to_screen_coordinates (point_coordinates.cpp:227-231,201-210): false when the projection has a NaN
bool world_to_screen(const Intrinsics& K, const double P[3], const Mat4& w2c, double uv[2])
{
    double h[4];
    for (int i = 0; i < 4; ++i) h[i] = ((w2c(i, 0) * P[0] + w2c(i, 1) * P[1]) + w2c(i, 2) * P[2]) + w2c(i, 3) * 1.0;
    const double xc = h[0] / h[3], yc = h[1] / h[3], zc = h[2] / h[3];
    const double inv = 1.0 / zc;
    uv[0] = inv * ((K.fx * xc + 0.0 * yc) + K.cx * zc);
    uv[1] = inv * ((0.0 * xc + K.fy * yc) + K.cy * zc);
    return !(uv[0] != uv[0] || uv[1] != uv[1]);
}

// InverseDepthWorldPoint::compute_signed_screen_distance (inverse_depth_coordinates.cpp:58-68) as called by
// Point2dOptimizationFeature::get_distance (map_point2d.cpp:40-45) with the inverse-depth STANDARD DEVIATION in the place
// of the covariance (so its square root is taken once more, :159), get_furthest/closest_estimation with the
// std::min(., 1e-9) of :146,153, Segment<2>::distance (line.hpp:27-41,95-99).
void point2d_signed_distance(const Intrinsics& K, const rs_match& f, const Mat4& w2c, double out[2])
{
    const double theta = f.obs[2], phi = f.obs[3];
    // _bearingVector = Cartesian::from(Spherical(1.0, theta, phi)) (basis_changes.cpp:5-10)
    const double sinTheta = std::sin(theta);
    const double b[3] = {1.0 * sinTheta * std::cos(phi), 1.0 * sinTheta * std::sin(phi), 1.0 * std::cos(theta)};
    const double depthStandardDev = std::sqrt(f.sigma[0]);
    const double depthVariation = depthStandardDev * 3;
    const double dFar = std::min(f.map[3] - depthVariation, 1e-9);
    const double dNear = std::min(f.map[3] + depthVariation, 1e-9);
    const double pFar[3] = {f.map[0] + b[0] / dFar, f.map[1] + b[1] / dFar, f.map[2] + b[2] / dFar};
    const double pNear[3] = {f.map[0] + b[0] / dNear, f.map[1] + b[1] / dNear, f.map[2] + b[2] / dNear};
    double s[2], e[2];
    if (!(world_to_screen(K, pFar, w2c, s) and world_to_screen(K, pNear, w2c, e))) {
        out[0] = DBL_MAX;
        out[1] = DBL_MAX;
        return;
    }
    // normal = (end - start).normalized(); closest = start + normal * ((p - start) . normal); distance = p - closest
    double n[2] = {e[0] - s[0], e[1] - s[1]};
    const double z = n[0] * n[0] + n[1] * n[1];
    if (z > 0) {
        const double l = std::sqrt(z);
        n[0] /= l, n[1] /= l;
    }
    const double along = (f.obs[0] - s[0]) * n[0] + (f.obs[1] - s[1]) * n[1];
    out[0] = f.obs[0] - (s[0] + n[0] * along);
    out[1] = f.obs[1] - (s[1] + n[1] * along);
}

// PlaneWorldCoordinates::to_camera_coordinates (plane_coordinates.cpp:20-24) + PlaneCoordinates(vector4) ctor
void plane_to_camera(const rs_match& f, const Mat4& M, Vec3& n, double& d)
{
    double h[4];
    for (int i = 0; i < 4; ++i)
        h[i] = ((M(i, 0) * f.map[0] + M(i, 1) * f.map[1]) + M(i, 2) * f.map[2]) + M(i, 3) * f.map[3];
    n = normalized(Vec3{h[0], h[1], h[2]});
    d = h[3];
}

double angle_distance(double a, double b) { return std::atan2(std::sin(a - b), std::cos(a - b)); }

void normalize_features(std::vector<rs_match>& feats)
{
    for (rs_match& f : feats)
        if (f.type == RS_FEAT_PLANE) {
            const Vec3 a = normalized(Vec3{f.obs[0], f.obs[1], f.obs[2]});
            const Vec3 b = normalized(Vec3{f.map[0], f.map[1], f.map[2]});
            for (int i = 0; i < 3; ++i) {
                f.obs[i] = a[i];
                f.map[i] = b[i];
            }
        }
}

bool has_nan(const double* v, int n)
{
    for (int i = 0; i < n; ++i)
        if (v[i] != v[i]) return true;
    return false;
}

}  // namespace

// levenberg_marquardt_functors.cpp:128-169
void pose_residuals(const Intrinsics& K, const std::vector<rs_match>& feats, const double x[6], double* fvec)
{
    if (has_nan(x, 6)) return;
    const Pose7 pose = pose_from_coefficients(x);
    const Mat4 w2c = world_to_camera(pose.q, pose.t);
    Mat4 planeM;
    bool havePlaneM = false;
    int idx = 0;
    for (const rs_match& f : feats) {
        if (f.type == RS_FEAT_POINT) {
            double dist[2];
            point_signed_distance(K, f, w2c, dist);
            if (!has_nan(dist, 2)) {
                fvec[idx] = dist[0] * 1.0 / 2.0;
                fvec[idx + 1] = dist[1] * 1.0 / 2.0;
            }
            idx += 2;
        }
        else if (f.type == RS_FEAT_POINT2D) {
            double dist[2];
            point2d_signed_distance(K, f, w2c, dist);
            if (!has_nan(dist, 2)) {   // get_alpha_reduction() = 0.3 (map_point2d.cpp:47), 2 parts
                fvec[idx] = dist[0] * 0.3 / 2.0;
                fvec[idx + 1] = dist[1] * 0.3 / 2.0;
            }
            idx += 2;
        }
        else {
            if (!havePlaneM) {
                planeM = plane_world_to_camera(w2c);  // identical for every plane (map_primitive.cpp:51-62)
                havePlaneM = true;
            }
            Vec3 np;
            double dp;
            plane_to_camera(f, planeM, np, dp);
            // get_reduced_signed_distance (plane_coordinates.cpp:49-56)
            const double r[3] = {f.obs[3] * f.obs[0] - dp * np.x, f.obs[3] * f.obs[1] - dp * np.y,
                                 f.obs[3] * f.obs[2] - dp * np.z};
            if (!has_nan(r, 3))
                for (int i = 0; i < 3; ++i) fvec[idx + i] = r[i] * 1.0 / 3.0;
            idx += 3;
        }
    }
}

bool feature_is_inlier(const Intrinsics& K, const rs_match& f, const Mat4& w2c, const Mat4& planeW2c)
{
    if (f.type == RS_FEAT_POINT) {
        // map_point.cpp:34-38 + get_distance_px (point_coordinates.cpp:262-278)
        double dist[2];
        point_signed_distance(K, f, w2c, dist);
        double distance;
        if (dist[0] >= DBL_MAX or dist[1] >= DBL_MAX)
            distance = DBL_MAX;
        else
            distance = std::fabs(dist[0]) + std::fabs(dist[1]);
        return distance <= kMaxRetroprojectionErrorForPointInliers_px;
    }
    if (f.type == RS_FEAT_POINT2D) {
        // map_point2d.cpp:33-38: (get_distance(w2c).array() <= threshold).all() on the SIGNED distance
        double dist[2];
        point2d_signed_distance(K, f, w2c, dist);
        return dist[0] <= kMaxRetroprojectionErrorForPoint2DInliers_px and dist[1] <= kMaxRetroprojectionErrorForPoint2DInliers_px;
    }
    // map_primitive.cpp:33-49 + get_signed_distance (plane_coordinates.cpp:26-37)
    Vec3 np;
    double dp;
    plane_to_camera(f, planeW2c, np, dp);
    const double e0 = std::fabs(angle_distance(f.obs[0], np.x));
    const double e1 = std::fabs(angle_distance(f.obs[1], np.y));
    const double e2 = std::fabs(angle_distance(f.obs[2], np.z));
    const double e3 = std::fabs(f.obs[3] - dp);
    const double tn = kMaxRetroprojectionErrorForPlaneInliersNormal;
    const double td = kMaxRetroprojectionErrorForPlaneInliers_mm;
    return e0 <= tn and e1 <= tn and e2 <= tn and e3 <= td;
}

// pose_optimization.cpp:302-359
bool optimized_global_pose(const Intrinsics& K, const Pose7& cur, const std::vector<rs_match>& feats, Pose7& out,
                           const int lm_max_fev, LMResult* info)
{
    double x[6];
    coefficients_from_pose(cur, x);
    for (double v : x)
        if (!std::isfinite(v)) return false;
    double optimizationScore = 0.0;
    size_t optiParts = 0;
    for (const rs_match& f : feats) {
        optiParts += size_t(feature_parts(f));
        optimizationScore += feature_score(f);
    }
    if (optiParts <= 1) return false;  // `optiParts <= input.cols()` with a column vector (:322)
    if (optimizationScore < 1.0) return false;
    const int m = int(optiParts);
    const ResidualFn fn = [&](const double* xx, double* fvec) { pose_residuals(K, feats, xx, fvec); };
    const LMResult r = lm_minimize(fn, m, x, lm_max_fev);
    if (info) *info = r;
    if (r.status <= 0) return false;
    const Pose7 p = pose_from_coefficients(x);
    double v6[6];
    pose_vector6(p, v6);
    if (has_nan(v6, 6)) return false;
    out = p;
    // optimizedPose.set_parameters(outputPose.get_position(), outputPose.get_orientation_quaternion()) (:357) normalises the
    // quaternion the PoseBase constructor has already normalised (pose.cpp:16-22): up to one more ulp on q. Observed against
    // the compiled reference sources (oracle/ref_shim, tests/test_reference_pose_build.py).
    const double nn = std::sqrt(p.q[0] * p.q[0] + p.q[1] * p.q[1] + p.q[2] * p.q[2] + p.q[3] * p.q[3]);
    if (nn > 0)
        for (double& v : out.q) v /= nn;
    return true;
}

int ransac_default_iterations()
{
    // pose_optimization.cpp:129-132 — float log/pow overloads
    const float num = std::log(1.0f - kProbabilityOfSuccess);
    const float den = std::log(1.0f - std::pow(kInlierProportion, kFeatureTrustCount));
    return int(static_cast<unsigned>(std::ceil(num / den)));
}

namespace {

// get_features_inliers_outliers (pose_optimization.cpp:33-72)
double inliers_outliers(const Intrinsics& K, const std::vector<rs_match>& feats, const Pose7& pose,
                        std::vector<uint8_t>& mask, int& nInliers)
{
    const Mat4 w2c = world_to_camera(pose.q, pose.t);
    const Mat4 pm = plane_world_to_camera(w2c);
    double score = 0.0;
    nInliers = 0;
    mask.assign(feats.size(), 0);
    for (size_t i = 0; i < feats.size(); ++i) {
        if (feature_is_inlier(K, feats[i], w2c, pm)) {
            mask[i] = 1;
            ++nInliers;
            score += feature_score(feats[i]);
        }
    }
    return score;
}

// ransac::get_random_subset_with_score (ransac.hpp:77-103): shuffled prefix until the score reaches 1, each pick
// PREPENDED to the output list.
std::vector<int> random_subset(const std::vector<rs_match>& feats, std::mt19937& engine, bool& ok)
{
    std::vector<int> order(feats.size());
    std::iota(order.begin(), order.end(), 0);
    std::shuffle(order.begin(), order.end(), engine);
    double cumulatedScore = 0.0;
    std::vector<int> out;
    ok = false;
    for (int idx : order) {
        cumulatedScore += feature_score(feats[idx]);
        out.insert(out.begin(), idx);
        if (cumulatedScore >= 1.0) {
            ok = true;
            break;
        }
    }
    return out;
}

// is_covariance_valid (covariances.hpp:13-44): finite, isApprox-symmetric, LDLT positive semi-definite.
bool covariance_valid(const double c[36])
{
    for (int i = 0; i < 36; ++i)
        if (!std::isfinite(c[i])) return false;
    double diff2 = 0, n2 = 0;
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            const double dd = c[i * 6 + j] - c[j * 6 + i];
            diff2 += dd * dd;
            n2 += c[i * 6 + j] * c[i * 6 + j];
        }
    if (!(diff2 <= 1e-12 * 1e-12 * n2)) return false;
    // pivoted LDL^T on a copy; negative pivot -> not PSD
    double a[6][6];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) a[i][j] = c[std::min(i, j) * 6 + std::max(i, j)];
    bool pos = false, neg = false;
    for (int k = 0; k < 6; ++k) {
        int p = k;
        double best = std::fabs(a[k][k]);
        for (int i = k + 1; i < 6; ++i)
            if (std::fabs(a[i][i]) > best) {
                best = std::fabs(a[i][i]);
                p = i;
            }
        if (p != k) {
            for (int j = 0; j < 6; ++j) std::swap(a[k][j], a[p][j]);
            for (int i = 0; i < 6; ++i) std::swap(a[i][k], a[i][p]);
        }
        const double dkk = a[k][k];
        if (dkk > 0) pos = true;
        if (dkk < 0) neg = true;
        if (std::fabs(dkk) <= DBL_MIN) break;
        for (int i = k + 1; i < 6; ++i) {
            const double l = a[i][k] / dkk;
            for (int j = k + 1; j < 6; ++j) a[i][j] -= l * a[k][j];
        }
    }
    (void)pos;
    return !neg;
}

}  // namespace

PoseSolveResult pose_solve(const Intrinsics& K, const Pose7& cur, const std::vector<rs_match>& featsIn,
                           int max_iterations, int n_variance, PoseRandom& rnd, const int lm_max_fev)
{
    PoseSolveResult R;
    rs_pose_out& out = R.out;
    out = rs_pose_out{};
    out.best_iteration = -1;
    std::vector<rs_match> feats = featsIn;
    normalize_features(feats);
    const size_t Nf = feats.size();
    R.inlier_mask.assign(Nf, 0);
    for (int i = 0; i < 3; ++i) out.pose[i] = cur.t[i];
    for (int i = 0; i < 4; ++i) out.pose[3 + i] = cur.q[i];
    if (max_iterations <= 0) max_iterations = ransac_default_iterations();
    if (n_variance < 0) n_variance = 100;

    // compute_optimized_pose: every feature must be valid (:269-282)
    for (const rs_match& f : feats) {
        // is_valid: map_point.cpp:60-64, map_primitive.cpp:79-83, map_point2d.cpp:75-79 (bearing = f(theta, phi))
        const int k = (f.type == RS_FEAT_POINT) ? 3 : 4;
        const int ko = (f.type == RS_FEAT_POINT) ? 2 : 4;
        const int ks = (f.type == RS_FEAT_POINT2D) ? 3 : k;
        if (has_nan(f.obs, ko) or has_nan(f.map, k) or has_nan(f.sigma, ks)) return R;
        for (int i = 0; i < ks; ++i)
            if (!(f.sigma[i] >= 0)) return R;
    }

    // compute_pose_with_ransac (:107-262)
    double initialFeatureScore = 0;
    for (const rs_match& f : feats) initialFeatureScore += feature_score(f);
    if (initialFeatureScore < 1.0) return R;

    const size_t inliersToStop = size_t(std::ceil(double(Nf) * kMinimumInliersProportionForEarlyStop));
    double maxScore = 1.0;
    Pose7 bestPose = cur;
    std::vector<uint8_t> bestMask(Nf, 0);
    int bestInliers = 0;
    bool canQuit = false;
    int started = 0;
    out.best_iteration = -1;
    for (int iteration = 0; iteration < max_iterations; ++iteration) {
        if (canQuit) break;
        ++started;
        std::vector<int> subset;
        bool ok = true;
        if (rnd.subsets) {
            for (int k = 0; k < RS_MAX_SUBSET; ++k) {
                const int idx = rnd.subsets[size_t(iteration) * RS_MAX_SUBSET + k];
                if (idx >= 0) subset.push_back(idx);
            }
        }
        else {
            subset = random_subset(feats, rnd.engine, ok);
        }
        R.subsets.push_back(subset);
        R.candidate_poses.emplace_back();
        R.candidate_ok.push_back(0);
        R.candidate_scores.push_back(0.0);
        if (!ok) continue;
        std::vector<rs_match> selected;
        for (int idx : subset) selected.push_back(feats[idx]);
        Pose7 candidate;
        if (not optimized_global_pose(K, cur, selected, candidate, lm_max_fev)) continue;
        R.candidate_poses.back() = candidate;
        R.candidate_ok.back() = 1;

        std::vector<uint8_t> mask;
        int nIn = 0;
        const double featureInlierScore = inliers_outliers(K, feats, candidate, mask, nIn);
        R.candidate_scores.back() = featureInlierScore;
        if (featureInlierScore < 1.0) continue;
        const bool canOverload = (featureInlierScore > maxScore) or
                                 (std::fabs(featureInlierScore - maxScore) <= 0.1 and bestInliers < nIn);
        if (canOverload) {
            maxScore = featureInlierScore;
            bestPose = candidate;
            bestMask.swap(mask);
            bestInliers = nIn;
            out.best_iteration = iteration;
        }
        if (iteration >= 3 and size_t(bestInliers) > inliersToStop) canQuit = true;
    }
    out.iterations_run = started;
    R.ransac_best = bestPose;

    double inlierScore = 0;
    std::vector<rs_match> inliers;
    for (size_t i = 0; i < Nf; ++i)
        if (bestMask[i]) {
            inlierScore += feature_score(feats[i]);
            inliers.push_back(feats[i]);
        }
    out.n_inliers = bestInliers;
    out.score = maxScore;
    if (inlierScore < 1.0) {
        out.status = 0;
        return R;
    }
    Pose7 finalPose;
    if (not optimized_global_pose(K, bestPose, inliers, finalPose, lm_max_fev)) {
        out.status = -1;
        return R;
    }
    R.inlier_mask = bestMask;
    for (int i = 0; i < 3; ++i) out.pose[i] = finalPose.t[i];
    for (int i = 0; i < 4; ++i) out.pose[3 + i] = finalPose.q[i];

    // compute_pose_variance (:361-437)
    if (n_variance == 0) {
        out.status = 1;
        return R;
    }
    std::vector<std::array<double, 6>> poses;
    double medium[6] = {0, 0, 0, 0, 0, 0};
    std::vector<int> inlierIdx;
    for (size_t i = 0; i < Nf; ++i)
        if (bestMask[i]) inlierIdx.push_back(int(i));
    for (int it = 0; it < n_variance; ++it) {
        // compute_random_variation_of_pose (:482-501)
        std::vector<rs_match> variated;
        variated.reserve(inliers.size());
        for (size_t k = 0; k < inliers.size(); ++k) {
            rs_match f = inliers[k];
            double g[4] = {0, 0, 0, 0};
            const int nd = (f.type == RS_FEAT_POINT) ? 3 : (f.type == RS_FEAT_POINT2D ? 2 : 4);
            for (int i = 0; i < nd; ++i) {
                if (rnd.normals)
                    g[i] = rnd.normals[(size_t(it) * rnd.max_matches + inlierIdx[k]) * 4 + i];
                else
                    g[i] = rnd.normal(rnd.engine);
            }
            if (f.type == RS_FEAT_POINT) {
                // map_point.cpp:49-58
                for (int i = 0; i < 3; ++i) f.map[i] += g[i] * f.sigma[i];
            }
            else if (f.type == RS_FEAT_POINT2D) {
                // map_point2d.cpp:49-73: theta then phi, clamped; observation point and inverse depth are kept
                f.obs[2] = std::clamp(f.obs[2] + g[0] * f.sigma[1], 0.0, M_PI);
                f.obs[3] = std::clamp(f.obs[3] + g[1] * f.sigma[2], -M_PI, M_PI);
            }
            else {
                // map_primitive.cpp:66-77 (+ the PlaneCoordinates copy ctor normalisation in make_shared)
                Vec3 n{f.map[0] + g[0] * f.sigma[0], f.map[1] + g[1] * f.sigma[1], f.map[2] + g[2] * f.sigma[2]};
                n = normalized(n);
                n = normalized(n);
                f.map[0] = n.x, f.map[1] = n.y, f.map[2] = n.z;
                f.map[3] += g[3] * f.sigma[3];
            }
            variated.push_back(f);
        }
        Pose7 newPose;
        if (optimized_global_pose(K, finalPose, variated, newPose, lm_max_fev)) {
            std::array<double, 6> v;
            pose_vector6(newPose, v.data());
            for (int i = 0; i < 6; ++i) medium[i] += v[i];
            poses.push_back(v);
        }
    }
    out.n_variance_ok = int(poses.size());
    if (poses.size() < size_t(unsigned(n_variance) / 2)) {
        out.status = -2;
        return R;
    }
    for (double& v : medium) v /= static_cast<double>(poses.size());
    double cov[36];
    for (double& v : cov) v = 0;
    for (const auto& p : poses) {
        double def[6];
        for (int i = 0; i < 6; ++i) def[i] = p[i] - medium[i];
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) cov[i * 6 + j] += def[i] * def[j];
    }
    for (double& v : cov) v /= static_cast<double>(poses.size() - 1);
    for (int i = 0; i < 6; ++i) cov[i * 6 + i] += 0.001;
    for (int i = 0; i < 36; ++i) out.cov[i] = cov[i];
    if (not covariance_valid(cov)) {
        out.status = -2;
        return R;
    }
    out.status = 1;
    return R;
}

}  // namespace oracle
