// Plane fit of geometry::FitPlane (src/Geometry/Geometry.cpp:172-199) in the reference's float operation order, and the
// Eigen 3.3.7 JacobiSVD<MatrixXf> restatement for a 3x3 matrix it needs (SVD/JacobiSVD.h:660-780, misc/RealSvd2x2.h:19-50,
// Jacobi/Jacobi.h:83-113).  Shared by the grid-based and the kd-tree-based normal estimation kernels.
#pragma once
#include <cfloat>

#include "opb_common.cuh"

namespace opb
{
struct Rot { float c, s; };
// JacobiRotation::makeJacobi(x, y, z) (Jacobi.h:83-113)
__device__ __forceinline__ Rot make_jacobi(float x, float y, float z)
{
    Rot r;
    const float deno = fmul(2.0f, fabsf(y));
    if (deno < FLT_MIN) { r.c = 1.0f; r.s = 0.0f; return r; }
    const float tau = fdiv(fsub(x, z), deno);
    const float w = __fsqrt_rn(fadd(fmul(tau, tau), 1.0f));
    const float t = tau > 0.0f ? fdiv(1.0f, fadd(tau, w)) : fdiv(1.0f, fsub(tau, w));
    const float sign_t = t > 0.0f ? 1.0f : -1.0f;
    const float n = fdiv(1.0f, __fsqrt_rn(fadd(fmul(t, t), 1.0f)));
    r.s = fmul(fmul(fmul(-sign_t, fdiv(y, fabsf(y))), fabsf(t)), n);
    r.c = n;
    return r;
}
// apply_rotation_in_the_plane: x' = c x + s y, y' = -s x + c y
__device__ __forceinline__ void rot_apply(float &x, float &y, Rot j)
{
    const float xi = x, yi = y;
    x = fadd(fmul(j.c, xi), fmul(j.s, yi));
    y = fadd(fmul(-j.s, xi), fmul(j.c, yi));
}
// third column of U of JacobiSVD(W) after the descending sort; W row-major
static __device__ void svd3_smallest_direction(const float *Win, float *normal)
{
    float W[3][3], U[3][3];
    float scale = 0.0f;
#pragma unroll
    for (int i = 0; i < 9; ++i) scale = fmaxf(scale, fabsf(Win[i]));
    if (scale == 0.0f) scale = 1.0f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) { W[r][c] = fdiv(Win[r * 3 + c], scale); U[r][c] = r == c ? 1.0f : 0.0f; }
    const float precision = 2.0f * FLT_EPSILON;
    float max_diag = fmaxf(fabsf(W[0][0]), fmaxf(fabsf(W[1][1]), fabsf(W[2][2])));
    bool finished = false;
    for (int sweep = 0; sweep < 64 && !finished; ++sweep) // converges in a handful of sweeps; the cap only guards NaN input
    {
        finished = true;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq)
        {
            const int p = pq == 0 ? 1 : 2, q = pq == 2 ? 1 : 0; // (1,0), (2,0), (2,1)
            const float threshold = fmaxf(FLT_MIN, fmul(precision, max_diag));
            if (fabsf(W[p][q]) > threshold || fabsf(W[q][p]) > threshold)
            {
                finished = false;
                // real_2x2_jacobi_svd (RealSvd2x2.h:19-50)
                float m00 = W[p][p], m01 = W[p][q], m10 = W[q][p], m11 = W[q][q];
                Rot rot1;
                const float t = fadd(m00, m11), d = fsub(m10, m01);
                if (fabsf(d) < FLT_MIN) { rot1.s = 0.0f; rot1.c = 1.0f; }
                else
                {
                    const float u = fdiv(t, d);
                    const float tmp = __fsqrt_rn(fadd(1.0f, fmul(u, u)));
                    rot1.s = fdiv(1.0f, tmp);
                    rot1.c = fdiv(u, tmp);
                }
                if (!(rot1.c == 1.0f && rot1.s == 0.0f)) { rot_apply(m00, m10, rot1); rot_apply(m01, m11, rot1); }
                const Rot j_right = make_jacobi(m00, m01, m11);
                const Rot jrt = {j_right.c, -j_right.s};
                const Rot j_left = {fsub(fmul(rot1.c, jrt.c), fmul(rot1.s, jrt.s)), fadd(fmul(rot1.c, jrt.s), fmul(rot1.s, jrt.c))};
                if (!(j_left.c == 1.0f && j_left.s == 0.0f))
                {
#pragma unroll
                    for (int c = 0; c < 3; ++c) rot_apply(W[p][c], W[q][c], j_left); // rows p, q
#pragma unroll
                    for (int r = 0; r < 3; ++r) rot_apply(U[r][p], U[r][q], j_left); // U.applyOnTheRight(p, q, j_left^T) applies (j_left^T)^T
                }
                if (!(jrt.c == 1.0f && jrt.s == 0.0f))
                {
#pragma unroll
                    for (int r = 0; r < 3; ++r) rot_apply(W[r][p], W[r][q], jrt);    // columns p, q with j_right^T
                }
                max_diag = fmaxf(max_diag, fmaxf(fabsf(W[p][p]), fabsf(W[q][q])));
            }
        }
    }
    float sv[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
        const float a = W[i][i];
        sv[i] = fmul(fabsf(a), scale);
        if (a < 0.0f)
#pragma unroll
            for (int r = 0; r < 3; ++r) U[r][i] = -U[r][i];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
        int pos = i;
#pragma unroll
        for (int k = i + 1; k < 3; ++k) if (sv[k] > sv[pos]) pos = k;
        if (sv[pos] == 0.0f) break;
        if (pos != i)
        {
            const float tmp = sv[i]; sv[i] = sv[pos]; sv[pos] = tmp;
#pragma unroll
            for (int r = 0; r < 3; ++r) { const float u = U[r][i]; U[r][i] = U[r][pos]; U[r][pos] = u; }
        }
    }
    normal[0] = U[0][2]; normal[1] = U[1][2]; normal[2] = U[2][2];
}

// FitPlane over the points pts[index(0..count)]: float mean, float covariance / count, smallest singular direction,
// normalize(); fewer than three points leave the zero vector (the reference's warning path)
template <class IndexOf>
__device__ __forceinline__ void fit_plane_normal(const float *__restrict__ pts, int count, IndexOf index_of, float *nrm)
{
    nrm[0] = nrm[1] = nrm[2] = 0.0f;
    if (count < 3) return;
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
    for (int k = 0; k < count; ++k)
    {
        const int j = index_of(k);
        sx = fadd(sx, pts[3 * j]); sy = fadd(sy, pts[3 * j + 1]); sz = fadd(sz, pts[3 * j + 2]);
    }
    const float cnt = (float)count;
    const float mx = fdiv(sx, cnt), my = fdiv(sy, cnt), mz = fdiv(sz, cnt);
    float W[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < count; ++k)
    {
        const int j = index_of(k);
        const float d[3] = {fsub(pts[3 * j], mx), fsub(pts[3 * j + 1], my), fsub(pts[3 * j + 2], mz)};
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) W[a * 3 + b] = fadd(W[a * 3 + b], fmul(d[a], d[b]));
    }
#pragma unroll
    for (int a = 0; a < 9; ++a) W[a] = fdiv(W[a], cnt);
    svd3_smallest_direction(W, nrm);
    // normal.normalize(): squaredNorm in Eigen's order a0 + (a1 + a2), division by the root if positive
    const float n2 = fadd(fmul(nrm[0], nrm[0]), fadd(fmul(nrm[1], nrm[1]), fmul(nrm[2], nrm[2])));
    if (n2 > 0.0f)
    {
        const float nn = __fsqrt_rn(n2);
        nrm[0] = fdiv(nrm[0], nn); nrm[1] = fdiv(nrm[1], nn); nrm[2] = fdiv(nrm[2], nn);
    }
}
} // namespace opb
