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
// JacobiSVD<MatrixXf>(W 3x3, ComputeThinU | ComputeThinV): U and (WITH_V) V, columns sorted by descending singular value
template <bool WITH_V>
static __device__ void jacobi_svd3(const float *Win, float (&U)[3][3], float (&V)[3][3])
{
    float W[3][3];
    float scale = 0.0f;
#pragma unroll
    for (int i = 0; i < 9; ++i) scale = fmaxf(scale, fabsf(Win[i]));
    if (scale == 0.0f) scale = 1.0f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) { W[r][c] = fdiv(Win[r * 3 + c], scale); U[r][c] = r == c ? 1.0f : 0.0f; if (WITH_V) V[r][c] = U[r][c]; }
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
                    if (WITH_V)
#pragma unroll
                        for (int r = 0; r < 3; ++r) rot_apply(V[r][p], V[r][q], jrt); // m_matrixV.applyOnTheRight(p, q, j_right)
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
            if (WITH_V)
#pragma unroll
                for (int r = 0; r < 3; ++r) { const float v = V[r][i]; V[r][i] = V[r][pos]; V[r][pos] = v; }
        }
    }
}
// third column of U of JacobiSVD(W) after the descending sort; W row-major
static __device__ void svd3_smallest_direction(const float *Win, float *normal)
{
    float U[3][3], V[3][3];
    jacobi_svd3<false>(Win, U, V);
    normal[0] = U[0][2]; normal[1] = U[1][2]; normal[2] = U[2][2];
}
// geometry::EstimateRigidTransformation (Geometry.cpp:107-151) over `count` pairs a(k) -> b(k), operation for operation: float
// means, W += (a - ma)(b - mb)^T, JacobiSVD, R = V U^T (Eigen's coefficient-based product of dynamic matrices: sequential sum),
// cofactor determinant, V's last column flipped when it is negative, t = mb - R ma (fixed-size product: x + (y + z)).
// R row-major, t.  PairOf(k, a, b) fills the k-th pair.
template <class PairOf>
static __device__ void kabsch_f32(int count, PairOf pair_of, float *R, float *t)
{
    float ma[3] = {0.0f, 0.0f, 0.0f}, mb[3] = {0.0f, 0.0f, 0.0f};
    for (int k = 0; k < count; ++k)
    {
        float a[3], b[3];
        pair_of(k, a, b);
#pragma unroll
        for (int c = 0; c < 3; ++c) { ma[c] = fadd(ma[c], a[c]); mb[c] = fadd(mb[c], b[c]); }
    }
    const float cnt = (float)count;
#pragma unroll
    for (int c = 0; c < 3; ++c) { ma[c] = fdiv(ma[c], cnt); mb[c] = fdiv(mb[c], cnt); }
    float W[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < count; ++k)
    {
        float a[3], b[3];
        pair_of(k, a, b);
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = 0; q < 3; ++q) W[p * 3 + q] = fadd(W[p * 3 + q], fmul(fsub(a[p], ma[p]), fsub(b[q], mb[q])));
    }
    float U[3][3], V[3][3];
    jacobi_svd3<true>(W, U, V);
#pragma unroll
    for (int pass = 0; pass < 2; ++pass)
    {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) R[i * 3 + j] = fadd(fadd(fmul(V[i][0], U[j][0]), fmul(V[i][1], U[j][1])), fmul(V[i][2], U[j][2]));
        const float c0 = fmul(R[0], fsub(fmul(R[4], R[8]), fmul(R[5], R[7])));
        const float c1 = fmul(R[1], fsub(fmul(R[3], R[8]), fmul(R[5], R[6])));
        const float c2 = fmul(R[2], fsub(fmul(R[3], R[7]), fmul(R[4], R[6])));
        const float det = fadd(fsub(c0, c1), c2);
        if (!(det < 0.0f) || pass) break;
        V[0][2] = -V[0][2]; V[1][2] = -V[1][2]; V[2][2] = -V[2][2];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = fsub(mb[i], fadd(fmul(R[i * 3], ma[0]), fadd(fmul(R[i * 3 + 1], ma[1]), fmul(R[i * 3 + 2], ma[2]))));
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
