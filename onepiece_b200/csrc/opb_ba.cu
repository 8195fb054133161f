// optimization::SimpleBA = Optimizer::FastBA (reference src/Optimization/SimpleBA.cpp:18-157, Optimizer.h:22-25): the pose-graph
// refinement DenseSlam runs over its submaps.
//
// Per frame pair (s, t) and point pair (p1, p2): q1 = R_s p1 + t_s, q2 = R_t p2 + t_t, r = q1 - q2,
// J_s = [I | -skew(q1)], J_t = [-I | skew(q2)]; the reference sums four 6x6 blocks J^T J and two 6-vectors -J^T r per frame pair
// (:40-66), assembles them into sparse normal equations over poses 1 .. n-1 (:109-137), solves (SimplicialLDLT, :139-142) and
// updates pose_i <- Se3ToSE3(delta_i) * pose_i (:144-152).
//
// All 156 block entries are linear in 28 sums per frame pair: N, sum q1, sum q2, sum q1 q1^T (6), sum q2 q2^T (6),
// sum q1 q2^T (9) -- skew(a)^T skew(a) = |a|^2 I - a a^T, skew(a) skew(b) = b a^T - (a . b) I, skew(q1) r = -(q1 x q2).
// The device reduces those (one CTA per frame pair; products in float as in the reference, accumulation in double, as in the
// ICP reduction) and the host does the rest: 6 (n - 1) unknowns, a dense LDL^T in double.  The reference sums the entries in
// float; the result is gated by tolerance against the oracle (itself pinned to the compiled reference), like the ICP pose.
//
// STATUS: the kernel runs on the host-thread emulator against numpy and the host half against the oracle in the CPU test suite
// (tests/test_ba_cpu.py); it has not been run on a B200 yet (this round's GPU budget was spent before it was written).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/onepiece_b200.h"
#include "opb_common.cuh"
#include "opb_linalg.h"

namespace opb
{
constexpr int kBaThreads = 128;
constexpr int kBaSums = 28;

// sums[k * 28 ..]: N, q1 (3), q2 (3), q1 q1^T upper (xx xy xz yy yz zz), q2 q2^T upper, q1 q2^T row-major (9)
__global__ void __launch_bounds__(kBaThreads) ba_pair_sums_kernel(const float *__restrict__ poses_rm, const int *__restrict__ src_id,
                                                                  const int *__restrict__ tgt_id, const long long *__restrict__ offset,
                                                                  const float *__restrict__ a, const float *__restrict__ b,
                                                                  double *__restrict__ sums)
{
    __shared__ double part[kBaThreads / 32][kBaSums];
    const int k = blockIdx.x;
    const float *Ps = poses_rm + 16 * src_id[k], *Pt = poses_rm + 16 * tgt_id[k];
    double acc[kBaSums];
#pragma unroll
    for (int e = 0; e < kBaSums; ++e) acc[e] = 0.0;
    for (long long j = offset[k] + threadIdx.x; j < offset[k + 1]; j += kBaThreads)
    {
        float q1[3], q2[3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
            // TransformPoint in the reference's order: (R row . p as x + (y + z)) + t
            q1[i] = fadd(fadd(fmul(Ps[4 * i], a[3 * j]), fadd(fmul(Ps[4 * i + 1], a[3 * j + 1]), fmul(Ps[4 * i + 2], a[3 * j + 2]))), Ps[4 * i + 3]);
            q2[i] = fadd(fadd(fmul(Pt[4 * i], b[3 * j]), fadd(fmul(Pt[4 * i + 1], b[3 * j + 1]), fmul(Pt[4 * i + 2], b[3 * j + 2]))), Pt[4 * i + 3]);
        }
        acc[0] += 1.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) { acc[1 + i] += (double)q1[i]; acc[4 + i] += (double)q2[i]; }
        int e = 7;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int c = i; c < 3; ++c) { acc[e] += (double)fmul(q1[i], q1[c]); acc[e + 6] += (double)fmul(q2[i], q2[c]); ++e; }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) acc[19 + 3 * i + c] += (double)fmul(q1[i], q2[c]);
    }
#pragma unroll
    for (int e = 0; e < kBaSums; ++e)
    {
        double v = acc[e];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5][e] = v;
    }
    __syncthreads();
    if (threadIdx.x < kBaSums)
    {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kBaThreads / 32; ++w) v += part[w][threadIdx.x]; // fixed order: the result does not depend on scheduling
        sums[(size_t)k * kBaSums + threadIdx.x] = v;
    }
}
} // namespace opb

using namespace opb;

namespace
{
// the six outputs of ComputeJTJAndJTr (SimpleBA.cpp:18-78) from the 28 sums: ss, tt, st, ts (row-major 6x6), rs, rt
void ba_blocks_from_sums(const double *S, double *out)
{
    const double N = S[0], *s1 = S + 1, *s2 = S + 4;
    auto sym = [](const double *u, int i, int j) { // upper-triangular storage xx xy xz yy yz zz
        static const int idx[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
        return u[idx[i][j]];
    };
    const double *M11 = S + 7, *M22 = S + 13, *M12 = S + 19;
    const double tr11 = M11[0] + M11[3] + M11[5], tr22 = M22[0] + M22[3] + M22[5], tr12 = M12[0] + M12[4] + M12[8];
    auto skew = [](const double *v, int i, int j) { // skew(v)(i, j)
        if (i == j) return 0.0;
        const int k = 3 - i - j;
        const double sign = ((j - i + 3) % 3 == 1) ? -1.0 : 1.0; // (0,1) (1,2) (2,0) carry the minus sign
        return sign * v[k];
    };
    double *ss = out, *tt = out + 36, *st = out + 72, *ts = out + 108, *rs = out + 144, *rt = out + 150;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
        {
            const double I = i == j ? 1.0 : 0.0;
            ss[6 * i + j] = N * I;
            ss[6 * i + 3 + j] = -skew(s1, i, j);                  // sum S1, S1 = -skew(q1)
            ss[6 * (3 + i) + j] = skew(s1, i, j);                 // sum S1^T
            ss[6 * (3 + i) + 3 + j] = tr11 * I - sym(M11, i, j);  // sum skew(q1)^T skew(q1)
            tt[6 * i + j] = N * I;
            tt[6 * i + 3 + j] = -skew(s2, i, j);                  // -sum S2, S2 = skew(q2)
            tt[6 * (3 + i) + j] = skew(s2, i, j);                 // -sum S2^T
            tt[6 * (3 + i) + 3 + j] = tr22 * I - sym(M22, i, j);
            st[6 * i + j] = -N * I;
            st[6 * i + 3 + j] = skew(s2, i, j);                   // sum S2
            st[6 * (3 + i) + j] = -skew(s1, i, j);                // -sum S1^T
            st[6 * (3 + i) + 3 + j] = M12[3 * j + i] - tr12 * I;  // sum skew(q1) skew(q2) = sum q2 q1^T - (q1 . q2) I
        }
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) ts[6 * i + j] = st[6 * j + i];
    const double c[3] = {M12[5] - M12[7], M12[6] - M12[2], M12[1] - M12[3]}; // sum q1 x q2
    for (int i = 0; i < 3; ++i)
    {
        rs[i] = s2[i] - s1[i]; rs[3 + i] = c[i];
        rt[i] = s1[i] - s2[i]; rt[3 + i] = -c[i];
    }
}
// dense LDL^T of the symmetric positive definite normal equations, in place; false if a pivot vanishes
bool ba_ldlt_solve(std::vector<double> &A, std::vector<double> &g, int n)
{
    for (int j = 0; j < n; ++j)
    {
        double d = A[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k] * A[(size_t)k * n + k];
        if (!(std::fabs(d) > 0.0)) return false;
        A[(size_t)j * n + j] = d;
        for (int i = j + 1; i < n; ++i)
        {
            double v = A[(size_t)i * n + j];
            for (int k = 0; k < j; ++k) v -= A[(size_t)i * n + k] * A[(size_t)j * n + k] * A[(size_t)k * n + k];
            A[(size_t)i * n + j] = v / d;
        }
    }
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < i; ++k) g[i] -= A[(size_t)i * n + k] * g[k];
    for (int i = 0; i < n; ++i) g[i] /= A[(size_t)i * n + i];
    for (int i = n - 1; i >= 0; --i)
        for (int k = i + 1; k < n; ++k) g[i] -= A[(size_t)k * n + i] * g[k];
    return true;
}
// one iteration's host half (SimpleBA.cpp:109-152): assemble, solve, update poses (row-major floats) in place
int ba_solve_and_update(int n_poses, float *poses_rm, int n_corr, const int32_t *src_id, const int32_t *tgt_id, const double *sums)
{
    const int nv = 6 * (n_poses - 1);
    std::vector<double> A((size_t)nv * nv, 0.0), g((size_t)nv, 0.0);
    for (int k = 0; k < n_corr; ++k)
    {
        double blk[156];
        ba_blocks_from_sums(sums + (size_t)k * kBaSums, blk);
        const int s = src_id[k], t = tgt_id[k];
        for (int i = 0; i < 6; ++i)
        {
            for (int j = 0; j < 6; ++j)
            {
                if (s != 0) // the first pose is fixed (:123-131)
                {
                    A[(size_t)((s - 1) * 6 + i) * nv + (s - 1) * 6 + j] += blk[6 * i + j];
                    A[(size_t)((s - 1) * 6 + i) * nv + (t - 1) * 6 + j] += blk[72 + 6 * i + j];
                    A[(size_t)((t - 1) * 6 + i) * nv + (s - 1) * 6 + j] += blk[108 + 6 * i + j];
                }
                A[(size_t)((t - 1) * 6 + i) * nv + (t - 1) * 6 + j] += blk[36 + 6 * i + j];
            }
            if (s != 0) g[(s - 1) * 6 + i] += blk[144 + i];
            g[(t - 1) * 6 + i] += blk[150 + i];
        }
    }
    if (!ba_ldlt_solve(A, g, nv)) { set_error("SimpleBA: the normal equations are singular (unconnected or degenerate pose graph)"); return OPB_ERR_INVALID; }
    for (int i = 1; i < n_poses; ++i)
    {
        double x[6], D[16], P[16], Nw[16];
        for (int e = 0; e < 6; ++e) x[e] = (double)(float)g[(i - 1) * 6 + e]; // the reference's delta is float
        linalg::se3_exp(x, D);
        for (int e = 0; e < 16; ++e) P[e] = poses_rm[16 * i + e];
        linalg::mat4_mul(D, P, Nw);
        for (int e = 0; e < 16; ++e) poses_rm[16 * i + e] = (float)Nw[e];
    }
    return OPB_OK;
}
int ba_check(int n_poses, const float *poses, int n_corr, const int32_t *src_id, const int32_t *tgt_id)
{
    if (n_poses < 0 || n_corr < 0 || (n_poses && !poses) || (n_corr && (!src_id || !tgt_id))) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    for (int k = 0; k < n_corr; ++k)
        if (src_id[k] < 0 || src_id[k] >= n_poses || tgt_id[k] < 1 || tgt_id[k] >= n_poses)
        {
            // the reference indexes block (target_id - 1): a pair whose target is frame 0 writes at row -6 there
            set_error("frame pair %d: source must be in [0, n_poses), target in [1, n_poses)", k);
            return OPB_ERR_INVALID;
        }
    return OPB_OK;
}
} // namespace

int opb_simple_ba_from_sums(int n_poses, float *poses_colmajor, int n_corr, const int32_t *src_id, const int32_t *tgt_id, const double *sums28)
{
    int rc = ba_check(n_poses, poses_colmajor, n_corr, src_id, tgt_id);
    if (rc) return rc;
    if (!sums28 && n_corr) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    if (n_poses < 3) return OPB_OK;
    if (n_corr < n_poses - 1) { set_error("SimpleBA: there are unconnected components"); return OPB_ERR_INVALID; }
    std::vector<float> rm((size_t)16 * n_poses);
    for (int i = 0; i < n_poses; ++i)
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) rm[16 * i + 4 * r + c] = poses_colmajor[16 * i + 4 * c + r];
    rc = ba_solve_and_update(n_poses, rm.data(), n_corr, src_id, tgt_id, sums28);
    if (rc) return rc;
    for (int i = 0; i < n_poses; ++i)
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) poses_colmajor[16 * i + 4 * c + r] = rm[16 * i + 4 * r + c];
    return OPB_OK;
}

int opb_simple_ba(int device, int n_poses, float *poses_colmajor, int n_corr, const int32_t *src_id, const int32_t *tgt_id, const int64_t *offset,
                  const float *a_xyz, const float *b_xyz, int max_iteration)
{
    int rc = ba_check(n_poses, poses_colmajor, n_corr, src_id, tgt_id);
    if (rc) return rc;
    if (n_corr && (!offset || !a_xyz || !b_xyz)) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    if (n_poses < 3) return OPB_OK; // "Too few optimization variables, No need to optimize." (:84-88)
    if (n_corr < n_poses - 1) { set_error("SimpleBA: there are unconnected components"); return OPB_ERR_INVALID; } // (:89-93): the reference returns
    for (int k = 0; k < n_corr; ++k)
        if (offset[k + 1] < offset[k] || offset[0] != 0) { set_error("offsets must start at 0 and not decrease"); return OPB_ERR_INVALID; }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
    {
        set_error("no CUDA device %d (the library has no CPU path)", device);
        return OPB_ERR_CUDA;
    }
    OPB_CUDA(cudaSetDevice(device));
    const size_t n_pts = (size_t)offset[n_corr];
    float *d_a = nullptr, *d_b = nullptr, *d_poses = nullptr;
    int *d_src = nullptr, *d_tgt = nullptr;
    long long *d_off = nullptr;
    double *d_sums = nullptr;
    std::vector<float> rm((size_t)16 * n_poses);
    std::vector<double> sums((size_t)n_corr * kBaSums);
    for (int i = 0; i < n_poses; ++i)
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) rm[16 * i + 4 * r + c] = poses_colmajor[16 * i + 4 * c + r];
    auto release = [&]() { cudaFree(d_a); cudaFree(d_b); cudaFree(d_poses); cudaFree(d_src); cudaFree(d_tgt); cudaFree(d_off); cudaFree(d_sums); };
    rc = OPB_OK;
    do
    {
#define OPB_TRY(expr) if ((expr) != cudaSuccess) { set_error("%s failed: %s", #expr, cudaGetErrorString(cudaGetLastError())); rc = OPB_ERR_CUDA; break; }
        OPB_TRY(cudaMalloc((void **)&d_a, (n_pts + 1) * 3 * sizeof(float)));
        OPB_TRY(cudaMalloc((void **)&d_b, (n_pts + 1) * 3 * sizeof(float)));
        OPB_TRY(cudaMalloc((void **)&d_poses, rm.size() * sizeof(float)));
        OPB_TRY(cudaMalloc((void **)&d_src, (size_t)n_corr * sizeof(int)));
        OPB_TRY(cudaMalloc((void **)&d_tgt, (size_t)n_corr * sizeof(int)));
        OPB_TRY(cudaMalloc((void **)&d_off, (size_t)(n_corr + 1) * sizeof(long long)));
        OPB_TRY(cudaMalloc((void **)&d_sums, sums.size() * sizeof(double)));
        OPB_TRY(cudaMemcpy(d_a, a_xyz, n_pts * 3 * sizeof(float), cudaMemcpyDefault));
        OPB_TRY(cudaMemcpy(d_b, b_xyz, n_pts * 3 * sizeof(float), cudaMemcpyDefault));
        OPB_TRY(cudaMemcpy(d_src, src_id, (size_t)n_corr * sizeof(int), cudaMemcpyHostToDevice));
        OPB_TRY(cudaMemcpy(d_tgt, tgt_id, (size_t)n_corr * sizeof(int), cudaMemcpyHostToDevice));
        OPB_TRY(cudaMemcpy(d_off, offset, (size_t)(n_corr + 1) * sizeof(long long), cudaMemcpyHostToDevice));
        for (int iter = 0; iter < max_iteration && rc == OPB_OK; ++iter)
        {
            OPB_TRY(cudaMemcpy(d_poses, rm.data(), rm.size() * sizeof(float), cudaMemcpyHostToDevice));
            ba_pair_sums_kernel<<<n_corr, kBaThreads>>>(d_poses, d_src, d_tgt, d_off, d_a, d_b, d_sums);
            OPB_TRY(cudaGetLastError());
            OPB_TRY(cudaMemcpy(sums.data(), d_sums, sums.size() * sizeof(double), cudaMemcpyDeviceToHost));
            rc = ba_solve_and_update(n_poses, rm.data(), n_corr, src_id, tgt_id, sums.data());
        }
#undef OPB_TRY
    } while (0);
    release();
    if (rc) return rc;
    for (int i = 0; i < n_poses; ++i)
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) poses_colmajor[16 * i + 4 * c + r] = rm[16 * i + 4 * r + c];
    return OPB_OK;
}
