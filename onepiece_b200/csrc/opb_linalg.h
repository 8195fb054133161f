// Small dense linear algebra in double, usable on host and device: the 6x6 normal-equation solves, the SE(3)
// exponential and the 3x3 SVD the solver loops need once per iteration.
//
// Reference semantics restated (file:line relative to the reference tree):
//   JacobiSVD(JTJ).solve(-JTr)      src/Registration/ICP.cpp:137-138        -> solve_sym6_pinv  (min-norm LS)
//   JTJ.ldlt().solve(-JTr)          src/Odometry/DenseOdometryFunction.cpp:402 -> solve_sym6_pinv (same answer
//                                                                               for full-rank systems)
//   geometry::Se3ToSE3 -> Sophus::SE3Group<scalar>::exp   src/Geometry/Geometry.cpp:9-13,
//                                   3rdparty/Sophus/sophus/se3.hpp:468-489, so3.hpp:388-412 -> se3_exp
//   geometry::EstimateRigidTransformation (Kabsch)  src/Geometry/Geometry.cpp:107-151 -> kabsch_from_sums
// The reference evaluates these in float32 (geometry::scalar); here they are evaluated in double from
// double-accumulated sums.  The difference is below the float32 reference's own deviation from its
// -DUSING_FLOAT64 build, which is what pose parity is gated on.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define OPB_HD __host__ __device__ inline
#else
#define OPB_HD inline
#endif

namespace opb
{
namespace linalg
{
// 1 / x, correctly rounded on both sides (the device intrinsic avoids the generic division routine: the solves below run
// on a single thread at the end of every solver iteration, so their latency is on the critical path)
OPB_HD double recip(double x)
{
#ifdef __CUDA_ARCH__
    return __drcp_rn(x);
#else
    return 1.0 / x;
#endif
}
// cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (row-major, n <= 6); A is destroyed,
// eigenvalues end up on its diagonal, V holds the eigenvectors as columns
template <int N>
OPB_HD void jacobi_eigen(double *A, double *V)
{
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) V[i * N + j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; ++sweep)
    {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < N; ++i)
        {
            diag += A[i * N + i] * A[i * N + i];
            for (int j = i + 1; j < N; ++j) off += A[i * N + j] * A[i * N + j];
        }
        if (off <= 1e-30 * diag || off == 0.0) break;
        for (int p = 0; p < N - 1; ++p)
            for (int q = p + 1; q < N; ++q)
            {
                const double apq = A[p * N + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * N + q] - A[p * N + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < N; ++k)
                {
                    const double akp = A[k * N + p], akq = A[k * N + q];
                    A[k * N + p] = c * akp - s * akq;
                    A[k * N + q] = s * akp + c * akq;
                }
                for (int k = 0; k < N; ++k)
                {
                    const double apk = A[p * N + k], aqk = A[q * N + k];
                    A[p * N + k] = c * apk - s * aqk;
                    A[q * N + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < N; ++k)
                {
                    const double vkp = V[k * N + p], vkq = V[k * N + q];
                    V[k * N + p] = c * vkp - s * vkq;
                    V[k * N + q] = s * vkp + c * vkq;
                }
            }
    }
}

// x = pinv(A) * b for symmetric positive semi-definite A (6x6, row-major).  Eigenvalues below
// rel_threshold * max eigenvalue are treated as zero, like JacobiSVD::solve's rank decision
// (Eigen default threshold: epsilon * max(rows, cols), here float32 epsilon * 6).
OPB_HD void solve_sym6_pinv(const double *A_in, const double *b, double *x, double rel_threshold = 6.0 * 1.1920929e-7)
{
    double A[36], V[36];
    for (int i = 0; i < 36; ++i) A[i] = A_in[i];
    jacobi_eigen<6>(A, V);
    double lmax = 0.0;
    for (int i = 0; i < 6; ++i) lmax = fmax(lmax, fabs(A[i * 6 + i]));
    for (int i = 0; i < 6; ++i) x[i] = 0.0;
    for (int k = 0; k < 6; ++k)
    {
        const double l = A[k * 6 + k];
        if (!(fabs(l) > rel_threshold * lmax) || l == 0.0) continue;
        double proj = 0.0;
        for (int i = 0; i < 6; ++i) proj += V[i * 6 + k] * b[i];
        proj /= l;
        for (int i = 0; i < 6; ++i) x[i] += V[i * 6 + k] * proj;
    }
}

// x = A^-1 b by LDL^T when A (6x6 symmetric, row-major) is comfortably positive definite; returns false (x
// untouched) when a pivot ratio suggests the rank decision of solve_sym6_pinv could matter.  Every loop has a
// compile-time trip count and is unrolled, so on the device L, D and y live in registers (this runs on ONE thread at
// the end of every solver iteration: its latency is on the critical path of the whole loop).
OPB_HD bool solve_sym6_ldlt(const double *A, const double *b, double *x)
{
    double L[36], D[6], invD[6];
    double dmax = 0.0, dmin = 1e300;
#pragma unroll
    for (int j = 0; j < 6; ++j)
    {
        double d = A[j * 6 + j];
#pragma unroll
        for (int k = 0; k < 6; ++k)
            if (k < j) d -= L[j * 6 + k] * L[j * 6 + k] * D[k];
        if (!(d > 0.0)) return false;
        D[j] = d;
        dmax = fmax(dmax, d);
        dmin = fmin(dmin, d);
        const double inv_d = recip(d);
        invD[j] = inv_d;
#pragma unroll
        for (int i = 0; i < 6; ++i)
            if (i > j)
            {
                double v = A[i * 6 + j];
#pragma unroll
                for (int k = 0; k < 6; ++k)
                    if (k < j) v -= L[i * 6 + k] * L[j * 6 + k] * D[k];
                L[i * 6 + j] = v * inv_d;
            }
    }
    if (!(dmin > 1e-4 * dmax)) return false;
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i)
    {
        double v = b[i];
#pragma unroll
        for (int k = 0; k < 6; ++k)
            if (k < i) v -= L[i * 6 + k] * y[k];
        y[i] = v;
    }
#pragma unroll
    for (int i = 5; i >= 0; --i)
    {
        double v = y[i] * invD[i];
#pragma unroll
        for (int k = 0; k < 6; ++k)
            if (k > i) v -= L[k * 6 + i] * x[k];
        x[i] = v;
    }
    return true;
}
// the solve both solver loops use: direct when well conditioned, pseudo-inverse otherwise (kept out of line on the
// device so that its dynamically indexed arrays do not drag the common path into local memory)
#ifdef __CUDACC__
__host__ __device__ __noinline__
#endif
inline void solve_sym6_pinv_cold(const double *A, const double *b, double *x) { solve_sym6_pinv(A, b, x); }
OPB_HD void solve_normal_equations6(const double *A, const double *b, double *x)
{
    if (!solve_sym6_ldlt(A, b, x)) solve_sym6_pinv_cold(A, b, x);
}

// 4x4 row-major helpers
OPB_HD void mat4_mul(const double *A, const double *B, double *C)
{
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
        {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += A[i * 4 + k] * B[k * 4 + j];
            C[i * 4 + j] = s;
        }
}

// Sophus::SE3Group::exp for the tangent (upsilon = x[0..2], omega = x[3..5]) -> 4x4 row-major
OPB_HD void se3_exp(const double *x, double *T)
{
    const double wx = x[3], wy = x[4], wz = x[5];
    const double theta_sq = wx * wx + wy * wy + wz * wz;
    const double theta = sqrt(theta_sq);
    // SO3Group::expAndTheta: unit quaternion from the half angle, Taylor expansion near zero
    double imag, real;
    if (theta < 1e-10)
    {
        const double t4 = theta_sq * theta_sq;
        imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * t4;
        real = 1.0 - 0.5 * theta_sq + (1.0 / 384.0) * t4;
    }
    else
    {
        double sh, ch;
        sincos(0.5 * theta, &sh, &ch);
        imag = sh * recip(theta);
        real = ch;
    }
    double qw = real, qx = imag * wx, qy = imag * wy, qz = imag * wz;
    const double inv_qn = recip(sqrt(qw * qw + qx * qx + qy * qy + qz * qz));
    qw *= inv_qn; qx *= inv_qn; qy *= inv_qn; qz *= inv_qn;
    double R[9] = {1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw),
                   2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw),
                   2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)};
    // V = I + (1-cos)/theta^2 * Omega + (theta - sin)/theta^3 * Omega^2   (se3.hpp:468-489)
    const double O[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    double O2[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
        {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += O[i * 3 + k] * O[k * 3 + j];
            O2[i * 3 + j] = s;
        }
    double V[9];
    if (theta < 1e-10)
        for (int i = 0; i < 9; ++i) V[i] = R[i]; // Sophus uses V = R in the small-angle branch
    else
    {
        double st, ct;
        sincos(theta, &st, &ct);
        const double inv_t2 = recip(theta_sq);
        const double a = (1.0 - ct) * inv_t2, b = (theta - st) * inv_t2 * recip(theta);
        for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0 ? 1.0 : 0.0) + a * O[i] + b * O2[i];
    }
    for (int i = 0; i < 3; ++i)
    {
        for (int j = 0; j < 3; ++j) T[i * 4 + j] = R[i * 3 + j];
        T[i * 4 + 3] = V[i * 3] * x[0] + V[i * 3 + 1] * x[1] + V[i * 3 + 2] * x[2];
    }
    T[12] = T[13] = T[14] = 0.0;
    T[15] = 1.0;
}

// Kabsch from accumulated sums over n pairs (s_i, t_i): sum_s[3], sum_t[3], sum_st[9] = sum s_i t_i^T
// (row-major).  Mirrors EstimateRigidTransformation: W = sum (s - mean_s)(t - mean_t)^T, SVD W = U S V^T,
// R = V U^T with the sign fix on V's last column when det(R) < 0, t = mean_t - R mean_s.  T is 4x4 row-major.
OPB_HD void kabsch_from_sums(double n, const double *sum_s, const double *sum_t, const double *sum_st, double *T)
{
    double ms[3], mt[3], W[9];
    for (int i = 0; i < 3; ++i) { ms[i] = sum_s[i] / n; mt[i] = sum_t[i] / n; }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) W[i * 3 + j] = sum_st[i * 3 + j] - n * ms[i] * mt[j];
    // eigen-decomposition of W^T W gives V and the squared singular values; U = W V / sigma
    double WtW[9], V[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
        {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += W[k * 3 + i] * W[k * 3 + j];
            WtW[i * 3 + j] = s;
        }
    jacobi_eigen<3>(WtW, V);
    // sort singular values descending (JacobiSVD order), carrying the columns of V
    int order[3] = {0, 1, 2};
    for (int a = 0; a < 2; ++a)
        for (int b = a + 1; b < 3; ++b)
            if (WtW[order[b] * 3 + order[b]] > WtW[order[a] * 3 + order[a]]) { const int t = order[a]; order[a] = order[b]; order[b] = t; }
    double Vs[9], U[9], sig[3];
    for (int c = 0; c < 3; ++c)
    {
        const int k = order[c];
        sig[c] = sqrt(fmax(WtW[k * 3 + k], 0.0));
        for (int i = 0; i < 3; ++i) Vs[i * 3 + c] = V[i * 3 + k];
    }
    for (int c = 0; c < 3; ++c)
    {
        double u[3];
        for (int i = 0; i < 3; ++i) u[i] = W[i * 3] * Vs[0 * 3 + c] + W[i * 3 + 1] * Vs[1 * 3 + c] + W[i * 3 + 2] * Vs[2 * 3 + c];
        if (sig[c] > 1e-14 * (sig[0] + 1e-300))
            for (int i = 0; i < 3; ++i) U[i * 3 + c] = u[i] / sig[c];
        else
        {   // rank-deficient W: complete U to a right-handed orthonormal basis
            const double *a = &U[0], *b = &U[1];
            if (c == 2)
            {
                U[0 * 3 + 2] = a[1 * 3] * b[2 * 3] - a[2 * 3] * b[1 * 3];
                U[1 * 3 + 2] = a[2 * 3] * b[0 * 3] - a[0 * 3] * b[2 * 3];
                U[2 * 3 + 2] = a[0 * 3] * b[1 * 3] - a[1 * 3] * b[0 * 3];
            }
            else
                for (int i = 0; i < 3; ++i) U[i * 3 + c] = i == c ? 1.0 : 0.0;
        }
    }
    double R[9];
    for (int pass = 0; pass < 2; ++pass)
    {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
            {
                double s = 0.0;
                for (int k = 0; k < 3; ++k) s += Vs[i * 3 + k] * U[j * 3 + k];
                R[i * 3 + j] = s;
            }
        const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
        if (det >= 0 || pass == 1) break;
        for (int i = 0; i < 3; ++i) Vs[i * 3 + 2] = -Vs[i * 3 + 2];
    }
    for (int i = 0; i < 3; ++i)
    {
        for (int j = 0; j < 3; ++j) T[i * 4 + j] = R[i * 3 + j];
        T[i * 4 + 3] = mt[i] - (R[i * 3] * ms[0] + R[i * 3 + 1] * ms[1] + R[i * 3 + 2] * ms[2]);
    }
    T[12] = T[13] = T[14] = 0.0;
    T[15] = 1.0;
}

} // namespace linalg
} // namespace opb
