// K4-K6: dense RGB-D odometry (photometric + geometric, coarse to fine) on the device.
//
// Reference path rebuilt here (file:line relative to the reference tree):
//   Odometry::DenseTracking (both overloads)        src/Odometry/Odometry.cpp:463-523, 526-608
//   Odometry::InitializeRGBDDenseTracking           src/Odometry/Odometry.cpp:609-620
//   Odometry::CreateImagePyramid / XYZ pyramid      src/Odometry/Odometry.cpp:436-461, src/Geometry/Geometry.cpp:72-106
//   Odometry::MultiScaleComputing                   src/Odometry/Odometry.cpp:621-685
//   ConvertDepthTo32FNaN / ConvertColorToIntensity32F   src/Odometry/DenseOdometryFunction.cpp:26-71
//   ComputeCorrespondencePixelWise                  src/Odometry/DenseOdometryFunction.cpp:8-25,72-128
//   NormalizeIntensity                              src/Odometry/DenseOdometryFunction.cpp:129-144
//   ComputeJacobian{Hybrid,Photo,Depth}Term, ComputeJTJandJTr*, DoSingleIteration*   :146-475
//   tool::CreatePyramid / SobelFiltering / GaussianFiltering / Convert2Gray / LinearTransform
//                                                   src/Tool/ImageProcessing.cpp:6-63 (OpenCV calls; OpenCV is not
//                                                   vendored by the reference: the filter arithmetic is the one
//                                                   DESIGN.md section 2 defines and the tests check against cv2)
//
// Design.  A frame's dense cache (geometry::RGBDFrame: gray, depth32f, 6 three-level pyramids) lives in one device
// allocation (3.2 MB at 640x480, L2-resident); the XYZ image is never materialised, back-projection is three flops.
// One solver iteration is two launches:
//   odo_candidates_kernel  thread = source pixel: warp it with d * K R K^-1 (u,v,1) + K t, test the target depth
//                          -> candidate {target pixel, transformed depth}
//   odo_iteration_kernel   thread = source pixel: resolves the reference's order-dependent occlusion filter exactly
//                          (the accept decision of pixel s is NOT(accept(t(s))) along a chain of strictly decreasing
//                          raster indices; the chain is walked to its first unconditional node), evaluates the two
//                          Jacobian rows in float exactly as the reference writes them, accumulates the 29 scalars
//                          (21 J^T J + 6 J^T r + sum r^2 + count) in double: registers -> warp shuffles -> per-CTA
//                          partial in a fixed slot; the last CTA to finish sums the partials in a fixed order
//                          (deterministic), solves the 6x6 system, applies the SE(3) exponential and updates the pose.
// The host never sees a pose until the coarse-to-fine schedule is over; the reference's data-dependent early exit
// of a level (correspondence ratio > 0.9) is a device flag later launches of that level test.
#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

#include "../../include/onepiece_b200.h"
#include "opb_common.cuh"
#include "opb_linalg.h"

namespace opb
{
constexpr int kOdoThreads = 256;
constexpr int kOdoPacket = 32; // doubles per partial (29 used)
constexpr int kMaxLevels = OPB_ODO_MAX_LEVELS;
constexpr int kMaxTrace = OPB_ODO_MAX_TRACE;

struct OdoCam
{
    float fx, fy, cx, cy;
    int w, h;
};

struct OdoState // device-resident solver state
{
    float T[16]; // source -> target, column-major float
    double packet[kOdoPacket];
    unsigned int blocks_done;
    int iteration;        // executed iterations
    int break_level;      // level whose remaining iterations are skipped (-1: none)
    int last_count;       // correspondences of the last executed iteration
    int n_pairs;          // result of the last ordered compaction
    float mean_src, mean_tgt; // NormalizeIntensity
    double rmse_sum;
    unsigned long long tail_ns; // profiling: time the last CTA spent summing partials + solving, accumulated over the call
    unsigned long long phase_ns[4]; // persistent loop, CTA 0: candidates, mid barrier, reduction (+ solve if last), release wait
    int trace_count[kMaxTrace];
    float trace_T[kMaxTrace][16];
};

struct FrameImages // pointers into one frame's allocation; what: 0 gray, 1 depth, 2 gray dx, 3 gray dy, 4 depth dx, 5 depth dy
{
    float *img[6][kMaxLevels];
};

__device__ __forceinline__ int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

// ---------------------------------------------------------------------------------------------------------
// K4: per-frame pre-processing
// ---------------------------------------------------------------------------------------------------------
// ConvertColorToIntensity32F (cvtColor RGB2GRAY on the stored bytes, 15-bit fixed point, then / 255) and
// ConvertDepthTo32FNaN (metres, NaN outside (0.5, 4))
__global__ void odo_convert_kernel(const uint8_t *__restrict__ bgr, const void *__restrict__ depth, int is_u16, float depth_scale,
                                   int n, float *__restrict__ gray, float *__restrict__ d32)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int g8 = (9798 * (int)bgr[3 * i] + 19235 * (int)bgr[3 * i + 1] + 3735 * (int)bgr[3 * i + 2] + 16384) >> 15;
    gray[i] = fdiv((float)g8, 255.0f);
    float out = __int_as_float(0x7fc00000);
    if (is_u16)
    {
        const unsigned short v = ((const unsigned short *)depth)[i];
        // `v > MIN_DEPTH * depth_scale`: the literal is double
        if ((double)v > 0.5 * (double)depth_scale && (double)v < 4.0 * (double)depth_scale) out = fdiv((float)v, depth_scale);
    }
    else
    {
        const float v = ((const float *)depth)[i];
        if (v > 0.5f && v < 4.0f) out = v;
    }
    d32[i] = out;
}

// GaussianBlur 3x3 sigma 0: separable [1/4 1/2 1/4], rows first, BORDER_REFLECT_101; blockIdx.z picks the image
__global__ void odo_blur3_kernel(const float *__restrict__ src0, const float *__restrict__ src1, int w, int h, float *__restrict__ dst0,
                                 float *__restrict__ dst1)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const float *src = blockIdx.z ? src1 : src0;
    float *dst = blockIdx.z ? dst1 : dst0;
    const int xl = reflect101(x - 1, w), xr = reflect101(x + 1, w);
    float t[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        const float *r = src + (size_t)reflect101(y - 1 + k, h) * w;
        t[k] = fadd(fmul(0.5f, r[x]), fmul(0.25f, fadd(r[xl], r[xr])));
    }
    dst[(size_t)y * w + x] = fadd(fmul(0.5f, t[1]), fmul(0.25f, fadd(t[0], t[2])));
}

// pyrDown to (w/2, h/2): separable [1 4 6 4 1], every second sample, / 256 after the column pass
__global__ void odo_pyrdown_kernel(const float *__restrict__ src0, const float *__restrict__ src1, int w, int h, float *__restrict__ dst0,
                                   float *__restrict__ dst1)
{
    const int ow = w / 2, oh = h / 2;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= ow || y >= oh) return;
    const float *src = blockIdx.z ? src1 : src0;
    float *dst = blockIdx.z ? dst1 : dst0;
    int xs[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) xs[k] = reflect101(2 * x - 2 + k, w);
    float t[5];
#pragma unroll
    for (int k = 0; k < 5; ++k)
    {
        const float *r = src + (size_t)reflect101(2 * y - 2 + k, h) * w;
        t[k] = fadd(fadd(fadd(fmul(r[xs[2]], 6.0f), fmul(fadd(r[xs[1]], r[xs[3]]), 4.0f)), r[xs[0]]), r[xs[4]]);
    }
    dst[(size_t)y * ow + x] = fmul(fadd(fadd(fadd(fmul(t[2], 6.0f), fmul(fadd(t[1], t[3]), 4.0f)), t[0]), t[4]), 1.0f / 256.0f);
}

// Sobel 3x3 (CV_32F): blockIdx.z = 4 * level + 2 * image + (0: dx, 1: dy) over all levels of both pyramids
__global__ void odo_sobel_kernel(FrameImages f, int w0, int h0, int levels)
{
    const int level = blockIdx.z >> 2, a = (blockIdx.z >> 1) & 1, dy = blockIdx.z & 1;
    if (level >= levels) return;
    const int w = w0 >> level, h = h0 >> level;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const float *src = f.img[a][level];
    float *dst = f.img[2 + 2 * a + dy][level];
    const int xl = reflect101(x - 1, w), xr = reflect101(x + 1, w);
    float t[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        const float *r = src + (size_t)reflect101(y - 1 + k, h) * w;
        t[k] = dy ? fadd(fadd(r[xl], fmul(2.0f, r[x])), r[xr]) : fsub(r[xr], r[xl]);
    }
    dst[(size_t)y * w + x] = dy ? fsub(t[2], t[0]) : fadd(fadd(t[0], fmul(2.0f, t[1])), t[2]);
}

// ---------------------------------------------------------------------------------------------------------
// K5: correspondences
// ---------------------------------------------------------------------------------------------------------
// K R K^-1 and K t in float, Eigen's evaluation order: 3x3 inverse by cofactors (compute_inverse_size3), lazy 3x3
// products coefficient by coefficient as a0*b0 + (a1*b1 + a2*b2)
__device__ void warp_matrices(const OdoCam &cam, const float *T, float *M, float *Kt)
{
    const float K[9] = {cam.fx, 0, cam.cx, 0, cam.fy, cam.cy, 0, 0, 1};
    const float R[9] = {T[0], T[4], T[8], T[1], T[5], T[9], T[2], T[6], T[10]};
    float cof[9], Kinv[9], KR[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
        {
            const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            cof[i * 3 + j] = fsub(fmul(K[i1 * 3 + j1], K[i2 * 3 + j2]), fmul(K[i1 * 3 + j2], K[i2 * 3 + j1]));
        }
    const float det = fadd(fadd(fmul(cof[0], K[0]), fmul(cof[3], K[3])), fmul(cof[6], K[6]));
    const float invdet = fdiv(1.0f, det);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Kinv[i * 3 + j] = fmul(cof[j * 3 + i], invdet);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) KR[i * 3 + j] = dot3(K[i * 3], K[i * 3 + 1], K[i * 3 + 2], R[j], R[3 + j], R[6 + j]);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[i * 3 + j] = dot3(KR[i * 3], KR[i * 3 + 1], KR[i * 3 + 2], Kinv[j], Kinv[3 + j], Kinv[6 + j]);
    for (int i = 0; i < 3; ++i) Kt[i] = dot3(K[i * 3], K[i * 3 + 1], K[i * 3 + 2], T[12], T[13], T[14]);
}

struct OdoArgs
{
    FrameImages src, tgt;
    OdoCam cam;      // of this level
    int level;
    int term;        // 0 hybrid, 1 photo, 2 depth
    int full_pixels; // width * height of level 0 (the early-exit ratio always uses the full resolution, Odometry.cpp:668)
    int2 *cand;      // per source pixel {target pixel index or -1, transformed depth bits}
    unsigned char *accepted;
    double *partials;
    OdoState *st;
    const float *T_override; // identity for the NormalizeIntensity correspondences, else nullptr (= st->T)
};

// the candidate of source pixel s (its target pixel and transformed depth, or -1): a pure function of the pixel and the pose
__device__ __forceinline__ int2 candidate_of(const float *__restrict__ sd, const float *__restrict__ td, const float *sM, int w, int h, int s)
{
    const int i = s / w, j = s - i * w;
    const float d_s = sd[s];
    int2 c = make_int2(-1, 0);
    if (d_s == d_s)
    {
        float uv[3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
            uv[r] = fadd(fadd(fmul(fmul(d_s, sM[r * 3]), (float)j), fadd(fmul(fmul(d_s, sM[r * 3 + 1]), (float)i), fmul(d_s, sM[r * 3 + 2]))),
                         sM[9 + r]);
        const float tds = uv[2];
        // (int)(x / z + 0.5): float division, double add, truncation (x86 semantics)
        const int u_t = cvtt_x86(__dadd_rn((double)fdiv(uv[0], tds), 0.5));
        const int v_t = cvtt_x86(__dadd_rn((double)fdiv(uv[1], tds), 0.5));
        if (u_t >= 0 && u_t < w && v_t >= 0 && v_t < h)
        {
            const float d_t = td[v_t * w + u_t];
            if (d_t == d_t && (double)fabsf(fsub(d_t, tds)) < 0.05) c = make_int2(v_t * w + u_t, __float_as_int(tds));
        }
    }
    return c;
}
__device__ __forceinline__ void candidates_phase(const OdoArgs &a, float *sM)
{
    __syncthreads(); // the previous user of the shared block is done
    if (threadIdx.x == 0) warp_matrices(a.cam, a.T_override ? a.T_override : a.st->T, sM, sM + 9);
    __syncthreads();
    const int w = a.cam.w, h = a.cam.h, n = w * h;
    const float *__restrict__ sd = a.src.img[1][a.level];
    const float *__restrict__ td = a.tgt.img[1][a.level];
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) a.cand[s] = candidate_of(sd, td, sM, w, h, s);
}
__global__ void __launch_bounds__(kOdoThreads) odo_candidates_kernel(OdoArgs a)
{
    if (a.st->break_level == a.level && !a.T_override) return;
    __shared__ float sM[12];
    candidates_phase(a, sM);
}

// AddElementToCorrespondenceMap, resolved without the raster-order loop.  The reference accepts source pixel s (a valid
// candidate with target t and transformed depth d) iff wraping_depth[t] == -1 or wraping_depth[t] > d at the moment s is
// visited, where wraping_depth[t] was written when SOURCE pixel t was accepted earlier in raster order.  Hence
//   t >= s, t not a valid candidate, or d(t) > d(s)  ->  accept(s) = true
//   otherwise                                         ->  accept(s) = !accept(t)
// and t < s strictly along such a chain, so walking it terminates at an unconditional "true".
// (no __restrict__ on the list: the persistent kernels write it in an earlier phase of the same launch, so its loads must stay
// ordinary coherent loads -- a restrict-qualified const pointer lets the compiler use the non-coherent path)
__device__ __forceinline__ bool resolve_accept(const int2 *cand, int s, int2 c)
{
    if (c.x < 0) return false;
    bool acc = true;
    int cur = s;
    for (;;)
    {
        const int t = c.x;
        if (t >= cur) break;
        const int2 ct = cand[t];
        if (ct.x < 0) break;
        if (__int_as_float(ct.y) > __int_as_float(c.y)) break;
        acc = !acc;
        cur = t;
        c = ct;
    }
    return acc;
}

// The same resolution by asynchronous pointer jumping (second persistent loop form).  accept(s) is the parity of the length of the
// chain s -> parent(s) -> ... -> root, where parent(s) = t(s) when the loop above would step to it.  Every pixel publishes a word
// {ancestor << 1 | parity of the path to that ancestor} (its parent first) and then repeatedly replaces it by its ancestor's word
// composed with its own.  Any word ever stored for a pixel is a true statement about one of its ancestors, words are single 32-bit
// stores written only by their owner, and an ancestor's index is smaller than its descendant's, so readers may see any mixture of
// old and new words: the answer is the same, only the number of steps differs (logarithmic in the chain length when the
// pixels of a chain advance together, instead of the 35-pixel average / 147-pixel maximum walks of the bench frames).
constexpr unsigned int kJumpUnset = 0xFFFFFFFFu;
__device__ __forceinline__ unsigned int chain_word(const int2 *cand, int s, int2 c)
{
    const int t = c.x;
    if (t >= s) return (unsigned int)s << 1;
    const int2 ct = cand[t];
    if (ct.x < 0 || __int_as_float(ct.y) > __int_as_float(c.y)) return (unsigned int)s << 1;
    return ((unsigned int)t << 1) | 1u;
}
__device__ __forceinline__ bool resolve_accept_jumping(const int2 *cand, unsigned int *jump, int s, int2 c)
{
    if (c.x < 0) return false;
    unsigned int w = chain_word(cand, s, c);
    __stcg(&jump[s], w);
    unsigned int a = w >> 1, p = w & 1u;
    if (a == (unsigned int)s) return true;
    for (;;)
    {
        unsigned int wa = __ldcg(&jump[a]);
        if (wa == kJumpUnset) wa = chain_word(cand, (int)a, cand[a]); // its owner has not got there yet
        const unsigned int a2 = wa >> 1;
        if (a2 == a) break; // a is a root
        p ^= wa & 1u;
        a = a2;
        __stcg(&jump[s], (a << 1) | p);
    }
    return p == 0u;
}

// ... without a candidate list at all: a pixel's candidate is a pure function of the pixel and the pose (candidate_of), so the
// first link of a pixel, and the word of an ancestor whose owner has not published yet, are computed on the spot.  Nothing a
// pixel reads then depends on another thread having got anywhere: no grid barrier between candidates and acceptance.
struct CandFn
{
    const float *sd, *td, *M;
    int w, h;
    __device__ __forceinline__ int2 operator()(int s) const { return candidate_of(sd, td, M, w, h, s); }
};
__device__ __forceinline__ unsigned int chain_word_lazy(const CandFn &f, int s, int2 c)
{
    const int t = c.x;
    if (t >= s) return (unsigned int)s << 1;
    const int2 ct = f(t);
    if (ct.x < 0 || __int_as_float(ct.y) > __int_as_float(c.y)) return (unsigned int)s << 1;
    return ((unsigned int)t << 1) | 1u;
}
__device__ __forceinline__ bool resolve_accept_lazy(const CandFn &f, unsigned int *jump, int s, int2 c)
{
    if (c.x < 0) return false;
    unsigned int w = chain_word_lazy(f, s, c);
    __stcg(&jump[s], w);
    unsigned int a = w >> 1, p = w & 1u;
    if (a == (unsigned int)s) return true;
    for (;;)
    {
        unsigned int wa = __ldcg(&jump[a]);
        if (wa == kJumpUnset) wa = chain_word_lazy(f, (int)a, f((int)a)); // its owner has not got there yet
        const unsigned int a2 = wa >> 1;
        if (a2 == a) break; // a is a root
        p ^= wa & 1u;
        a = a2;
        __stcg(&jump[s], (a << 1) | p);
    }
    return p == 0u;
}

__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One Jacobian row pair of a correspondence, float arithmetic in the reference's order (DenseOdometryFunction.cpp:146-296);
// adds its (exact, double) products to the per-thread double accumulators acc[0..20] (upper triangle of J^T J), acc[21..26] (J^T r), acc[27] (r^2)
struct RowCtx // what compute_rows reads of one pyramid level
{
    const float *sgray, *sdepth, *tgray, *tdepth, *tgx, *tgy, *tdx, *tdy;
    float fx, fy, cx, cy;
    int w;
};
__device__ __forceinline__ RowCtx row_ctx(const OdoArgs &a)
{
    const int l = a.level;
    RowCtx c;
    c.sgray = a.src.img[0][l]; c.sdepth = a.src.img[1][l];
    c.tgray = a.tgt.img[0][l]; c.tdepth = a.tgt.img[1][l];
    c.tgx = a.tgt.img[2][l]; c.tgy = a.tgt.img[3][l]; c.tdx = a.tgt.img[4][l]; c.tdy = a.tgt.img[5][l];
    c.fx = a.cam.fx; c.fy = a.cam.fy; c.cx = a.cam.cx; c.cy = a.cam.cy; c.w = a.cam.w;
    return c;
}
template <int TERM>
__device__ __forceinline__ int compute_rows(const RowCtx &a, const float *T, int s, int t, float (*J)[6], float *res)
{
    const int w = a.w;
    const int v_s = s / w, u_s = s - v_s * w;
    const float fx = a.fx, fy = a.fy;
    // source_XYZ[v_s][u_s] (TransformToMatXYZ, Geometry.cpp:72-106)
    const float z = a.sdepth[s];
    float p0 = -1.0f, p1 = -1.0f, p2 = -1.0f;
    if (z > 0)
    {
        p0 = fdiv(fmul(fsub((float)u_s, a.cx), z), fx);
        p1 = fdiv(fmul(fsub((float)v_s, a.cy), z), fy);
        p2 = z;
    }
    // R * p + t: Eigen's fixed-size order m0 + (m1 + m2), then + t
    const float q0 = fadd(fadd(fmul(T[0], p0), fadd(fmul(T[4], p1), fmul(T[8], p2))), T[12]);
    const float q1 = fadd(fadd(fmul(T[1], p0), fadd(fmul(T[5], p1), fmul(T[9], p2))), T[13]);
    const float q2 = fadd(fadd(fmul(T[2], p0), fadd(fmul(T[6], p1), fmul(T[10], p2))), T[14]);
    const float invz = (float)(1.0 / (double)q2);
    const float sq = 0.70710678118654752440f; // (float)sqrt(0.5) == (float)sqrt(1 - 0.5)
    int rows = 0;
    if (TERM == 0 || TERM == 1)
    {
        const float diff = fsub(a.tgray[t], a.sgray[s]);
        const float gx = fmul(0.125f, a.tgx[t]), gy = fmul(0.125f, a.tgy[t]);
        const float c0 = fmul(fmul(gx, fx), invz), c1 = fmul(fmul(gy, fy), invz);
        const float c2 = fmul(-fadd(fmul(c0, q0), fmul(c1, q1)), invz);
        float *j = J[rows];
        j[0] = c0; j[1] = c1; j[2] = c2;
        j[3] = fadd(fmul(-q2, c1), fmul(q1, c2));
        j[4] = fsub(fmul(q2, c0), fmul(q0, c2));
        j[5] = fadd(fmul(-q1, c0), fmul(q0, c1));
        if (TERM == 0)
        {
#pragma unroll
            for (int e = 0; e < 6; ++e) j[e] = fmul(sq, j[e]);
            res[rows] = fmul(sq, diff);
        }
        else res[rows] = diff;
        ++rows;
    }
    if (TERM == 0 || TERM == 2)
    {
        float gx = fmul(0.125f, a.tdx[t]), gy = fmul(0.125f, a.tdy[t]);
        if (gx != gx) gx = 0.0f;
        if (gy != gy) gy = 0.0f;
        const float diff = fsub(a.tdepth[t], q2);
        const float d0 = fmul(fmul(gx, fx), invz), d1 = fmul(fmul(gy, fy), invz);
        const float d2 = fmul(-fadd(fmul(d0, q0), fmul(d1, q1)), invz);
        float *j = J[rows];
        j[0] = d0; j[1] = d1; j[2] = fsub(d2, 1.0f);
        j[3] = fsub(fadd(fmul(-q2, d1), fmul(q1, d2)), q1);
        j[4] = fadd(fsub(fmul(q2, d0), fmul(q0, d2)), q0);
        j[5] = fadd(fmul(-q1, d0), fmul(q0, d1));
        if (TERM == 0)
        {
#pragma unroll
            for (int e = 0; e < 6; ++e) j[e] = fmul(sq, j[e]);
            res[rows] = fmul(sq, diff);
        }
        else res[rows] = diff;
        ++rows;
    }
    return rows;
}
template <int TERM>
__device__ __forceinline__ void accumulate_rows(const OdoArgs &a, const float *T, int s, int t, double *acc)
{
    float J[2][6], res[2];
    const RowCtx ctx = row_ctx(a);
    const int rows = compute_rows<TERM>(ctx, T, s, t, J, res);
#pragma unroll
    for (int r = 0; r < 2; ++r)
    {
        if (r >= rows) break;
        int k = 0;
#pragma unroll
        for (int p = 0; p < 6; ++p)
#pragma unroll
            for (int q = p; q < 6; ++q) acc[k++] += (double)J[r][p] * (double)J[r][q]; // exact product of two floats, double sum
#pragma unroll
        for (int p = 0; p < 6; ++p) acc[21 + p] += (double)J[r][p] * (double)res[r];
        acc[27] += (double)res[r] * (double)res[r];
    }
}

// the pose update of DoSingleIteration (DenseOdometryFunction.cpp:402-411) from the 29-scalar packet: T <- exp(delta) * T.  One thread.
__device__ __noinline__ void odo_solve_core(const double *P, float *T)
{
    double JTJ[36], nJTr[6], x[6], dT[16];
#pragma unroll
    for (int p = 0; p < 6; ++p)
#pragma unroll
        for (int q = p; q < 6; ++q)
        {
            const double v = P[p * 6 - (p * (p - 1)) / 2 + (q - p)]; // position of (p, q) in the row-wise upper triangle
            JTJ[p * 6 + q] = v;
            JTJ[q * 6 + p] = v;
        }
#pragma unroll
    for (int p = 0; p < 6; ++p) nJTr[p] = -P[21 + p];
    linalg::solve_normal_equations6(JTJ, nJTr, x);
#pragma unroll
    for (int p = 0; p < 6; ++p) x[p] = (double)(float)x[p]; // the reference's delta is float32
    linalg::se3_exp(x, dT);
    float dTf[16], Tn[16];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) dTf[c * 4 + r] = (float)dT[r * 4 + c];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int r = 0; r < 4; ++r)
            Tn[c * 4 + r] = fadd(fadd(fadd(fmul(dTf[r], T[c * 4]), fmul(dTf[4 + r], T[c * 4 + 1])), fmul(dTf[8 + r], T[c * 4 + 2])),
                                 fmul(dTf[12 + r], T[c * 4 + 3]));
#pragma unroll
    for (int e = 0; e < 16; ++e) T[e] = Tn[e];
}
// ... on the device-resident state, run by one thread of the last CTA (separate-launch form and first persistent form)
__device__ void solve_and_update(const OdoArgs &a, OdoState *st)
{
    const double *P = st->packet;
    const int n = (int)(P[28] + 0.5);
    float Tn[16];
    for (int e = 0; e < 16; ++e) Tn[e] = st->T[e];
    odo_solve_core(P, Tn);
    for (int e = 0; e < 16; ++e) st->T[e] = Tn[e];
    const int it = st->iteration;
    if (it < kMaxTrace)
    {
        st->trace_count[it] = n;
        for (int e = 0; e < 16; ++e) st->trace_T[it][e] = Tn[e];
    }
    st->iteration = it + 1;
    st->last_count = n;
    // if ((float)correspondences.size() / (height * width) > MAX_INLIER_RATIO_DENSE) break;   (Odometry.cpp:668)
    if ((double)fdiv((float)n, (float)a.full_pixels) > 0.9) st->break_level = a.level;
}

// mode 0: solver iteration; mode 1: only the accept flags and the count (NormalizeIntensity / teacher-forced listing)
struct OdoShared
{
    double part[kOdoThreads / 32][kOdoPacket];
    float T[16];
    float M[12];
    bool last;
};
// returns true in every thread of the CTA that finished last (the one that summed the partials and solved)
template <int TERM, int MODE>
__device__ __forceinline__ bool iteration_phase(const OdoArgs &a, OdoShared &sh)
{
    double (*s_part)[kOdoPacket] = sh.part;
    float *sT = sh.T;
    bool &s_last = sh.last;
    __syncthreads(); // the previous user of the shared block is done
    if (threadIdx.x < 16) sT[threadIdx.x] = a.st->T[threadIdx.x];
    __syncthreads();
    const int n = a.cam.w * a.cam.h;
    double acc[29];
#pragma unroll
    for (int k = 0; k < 29; ++k) acc[k] = 0.0;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x)
    {
        const int2 c = a.cand[s];
        const bool ok = resolve_accept(a.cand, s, c);
        a.accepted[s] = ok;
        if (!ok) continue;
        acc[28] += 1.0;
        if (MODE == 0) accumulate_rows<TERM>(a, sT, s, c.x, acc);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 29; ++k)
    {
        const double v = warp_sum_d(acc[k]);
        if (lane == 0) s_part[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 29)
    {
        double v = 0.0;
        for (int wp = 0; wp < kOdoThreads / 32; ++wp) v += s_part[wp][threadIdx.x];
        a.partials[(size_t)blockIdx.x * kOdoPacket + threadIdx.x] = v;
    }
    // last CTA: fixed-order sum of the partials, then the solve
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&a.st->blocks_done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return false;
    unsigned long long t_tail = 0;
    if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_tail));
    __threadfence();
    {
        const int k = threadIdx.x & 31, chain = threadIdx.x >> 5;
        double v = 0.0;
        if (k < 29)
        {
#pragma unroll 8
            for (unsigned int b = chain; b < gridDim.x; b += kOdoThreads / 32) v += __ldcg(&a.partials[(size_t)b * kOdoPacket + k]);
        }
        s_part[chain][k] = v;
        __syncthreads();
        if (threadIdx.x < 29)
        {
            double tot = 0.0;
            for (int c = 0; c < kOdoThreads / 32; ++c) tot += s_part[c][threadIdx.x];
            a.st->packet[threadIdx.x] = tot;
        }
    }
    __syncthreads();
    if (threadIdx.x != 0) return true;
    a.st->blocks_done = 0;
    if (MODE == 0) solve_and_update(a, a.st);
    else a.st->last_count = (int)(a.st->packet[28] + 0.5);
    unsigned long long t_end;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
    a.st->tail_ns += t_end - t_tail;
    return true;
}
template <int TERM, int MODE>
__global__ void __launch_bounds__(kOdoThreads) odo_iteration_kernel(OdoArgs a)
{
    if (MODE == 0 && a.st->break_level == a.level) return;
    __shared__ OdoShared sh;
    iteration_phase<TERM, MODE>(a, sh);
}

// MultiScaleComputing (Odometry.cpp:621-685) as ONE persistent cooperative launch: all levels, all iterations.  An iteration
// is the two phases of the kernels above separated by a grid-wide barrier (the acceptance test of a pixel reads the
// candidates of other pixels); the last CTA to finish the reduction solves, updates the pose and releases the grid into the
// next iteration.  Coarse levels keep the whole grid alive but idle: their iterations cost two barriers and the solve instead
// of two launches.
__device__ __forceinline__ unsigned long long timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
struct OdoLoopArgs
{
    OdoArgs base;
    OdoCam cams[kMaxLevels];
    int iterations[kMaxLevels];
    int levels;
    unsigned int *sync; // [0] arrivals of the mid-iteration barrier (monotonic), [1] iterations released
};
template <int TERM>
__global__ void __launch_bounds__(kOdoThreads, 2) odo_loop_kernel(OdoLoopArgs L)
{
    __shared__ OdoShared sh;
    unsigned int n_sync = 0; // barriers passed so far, identical in every thread of the grid
    for (int l = L.levels - 1; l >= 0; --l)
    {
        OdoArgs a = L.base;
        a.level = l;
        a.cam = L.cams[l];
        for (int j = 0; j < L.iterations[l]; ++j)
        {
            // written by the previous iteration's solve, which every CTA has waited for: uniform over the grid
            if (*(volatile int *)&a.st->break_level == l) break;
            const bool stamp = blockIdx.x == 0 && threadIdx.x == 0;
            unsigned long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
            if (stamp) t0 = timer_ns();
            candidates_phase(a, sh.M);
            if (stamp) t1 = timer_ns();
            ++n_sync;
            __syncthreads();
            if (threadIdx.x == 0)
            {
                __threadfence();
                atomicAdd(&L.sync[0], 1u);
                while (*(volatile unsigned int *)&L.sync[0] < n_sync * gridDim.x) { }
                __threadfence();
            }
            __syncthreads();
            if (stamp) t2 = timer_ns();
            const bool last = iteration_phase<TERM, 0>(a, sh);
            if (stamp) t3 = timer_ns();
            if (threadIdx.x == 0)
            {
                if (last)
                {
                    __threadfence();
                    atomicExch(&L.sync[1], n_sync); // releases the grid
                }
                else
                    while (*(volatile unsigned int *)&L.sync[1] < n_sync) { }
                __threadfence();
            }
            __syncthreads();
            if (stamp)
            {
                t4 = timer_ns();
                a.st->phase_ns[0] += t1 - t0; a.st->phase_ns[1] += t2 - t1; a.st->phase_ns[2] += t3 - t2; a.st->phase_ns[3] += t4 - t3;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// MultiScaleComputing, second persistent form (the default).  Where the first form's 19 us per iteration go, measured
// (profiles/r02_odo_loop_kernel_ncu_full_before.md, opb_odometry_last_phases): 1.4 us candidates, 1.5 us grid barrier, 6 us in which
// every thread keeps 29 double sums and a warp folds them with 290 shuffles, 10 us in which 295 CTAs wait for one CTA to sum 296
// partials and for one of its threads to solve.  This form removes the per-thread sums and the serial owner:
//   * (measured and dropped on the way: WALKING the chains over recomputed candidates instead of the list -- the chains are 35
//     pixels long on average and up to 150 on the bench frames, and two dependent loads plus the projection per step cost more
//     than the barrier that goes away: 0.58 ms per call against 0.55; pointer jumping needs the recomputation only for first links);
//   * the sums are an 8x8 outer-product accumulation on the FP64 tensor-core op (warp_fold_outer8): a Jacobian row contributes
//     c c^T with c = (J0..J5, r, 1) -- J^T J, J^T r, r^2 and the count are entries of that matrix; the hybrid term's second row
//     goes in with c = (J0..J5, r, 0).  Products of floats are exact in double, which is also what the separate-launch kernels and
//     the oracle sum: the forms differ only in the order of the double additions (~1e-16 relative);
//   * candidates are not listed at all: a pixel's candidate is recomputed where a chain needs it, and the chains are resolved by
//     asynchronous pointer jumping (resolve_accept_lazy), so the grid barrier between candidates and acceptance is gone;
//   * one CTA of 1024 threads per SM publishes its 8x8 partial, announces it on one counter, waits until all have, and then EVERY
//     CTA sums all partials in the same fixed order and solves the same 6x6 system: nobody waits for an owner, the pose lives
//     in shared memory, the result is deterministic and identical on all CTAs (CTA 0 records the trace).
// The candidate / accept lists that the ordered compaction after the loop reads are written by the iterations of `list_level`.
// ---------------------------------------------------------------------------------------------------------
constexpr int kOdo2Threads = 1024;
constexpr int kOdo2Warps = kOdo2Threads / 32;
constexpr int kOdo2Chunks = kOdo2Threads / 64; // groups that each sum every 16th partial

// entries of the row-major 8x8 sum matrix that feed the 29-scalar packet: upper triangle of the leading 6x6, column 6 down to
// (6,6), and (7,7)
__host__ __device__ constexpr unsigned long long odo2_need_mask()
{
    unsigned long long m = 0;
    for (int p = 0; p < 6; ++p)
        for (int q = p; q < 7; ++q) m |= 1ull << (p * 8 + q);
    m |= 1ull << (6 * 8 + 6);
    m |= 1ull << (7 * 8 + 7);
    return m;
}
// ... and where component k of the packet comes from
__host__ __device__ constexpr int odo2_packet_source(int k)
{
    if (k < 21)
    {
        int p = 0, first = 0;
        while (k >= first + (6 - p)) { first += 6 - p; ++p; }
        return p * 8 + p + (k - first);
    }
    if (k < 27) return (k - 21) * 8 + 6;
    return k == 27 ? 6 * 8 + 6 : (k == 28 ? 7 * 8 + 7 : -1);
}
struct Odo2Shared
{
    union
    {
        float stage[kOdo2Warps][32 * 8]; // during the trips: per warp, 32 rows x 8 components
        double wsum[kOdo2Warps][64];     // after the trips: per-warp 8x8 sums
        double chunk[kOdo2Chunks][64];   // partial sums of the cross-CTA reduction
    } u;
    double sum64[64];
    double packet[kOdoPacket];
    float T[16];
    float M[12];
    RowCtx ctx;
    int break_level, iteration, last_count;
    unsigned long long t[6], ph[4]; // profiling stamps of CTA 0
};
struct OdoLoop2Args
{
    OdoArgs base;
    OdoCam cams[kMaxLevels];
    int iterations[kMaxLevels];
    int levels;
    int list_level;       // iterations of this level write a.cand / a.accepted (the compaction after the loop reads them)
    double *partials;     // [2][gridDim.x][64]
    unsigned int *jump;   // [2][jump_stride]: per pixel chain words of resolve_accept_lazy, two generations (this iteration's, the next one's)
    unsigned int jump_stride;
    unsigned int *sync;   // arrivals, monotonic over the launch
};
template <int TERM>
__global__ void __launch_bounds__(kOdo2Threads, 1) odo_loop2_kernel(const __grid_constant__ OdoLoop2Args L)
{
    __shared__ Odo2Shared sh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_cta = gridDim.x;
    const bool stamp = blockIdx.x == 0 && threadIdx.x == 0;
    OdoState *st = L.base.st;
    if (threadIdx.x < 16) sh.T[threadIdx.x] = st->T[threadIdx.x];
    if (threadIdx.x == 32) { sh.break_level = -1; sh.iteration = 0; sh.last_count = 0; }
    if (threadIdx.x < 4) sh.ph[threadIdx.x] = 0;
    __syncthreads();
    unsigned int n_sync = 0; // barriers passed so far, identical in every thread of the grid
    for (int l = L.levels - 1; l >= 0; --l)
    {
        const OdoCam cam = L.cams[l];
        const int w = cam.w, h = cam.h, n = w * h;
        const int n_trips = (n + 31) >> 5;
        // consecutive trips go to different CTAs: the coarse levels keep every SM's load pipes busy instead of a few CTAs'
        const int first_trip = warp * n_cta + blockIdx.x, trip_step = kOdo2Warps * n_cta;
        const float *__restrict__ sd = L.base.src.img[1][l];
        const float *__restrict__ td = L.base.tgt.img[1][l];
        __syncthreads();
        if (threadIdx.x == 0)
        {
            RowCtx c;
            c.sgray = L.base.src.img[0][l]; c.sdepth = L.base.src.img[1][l];
            c.tgray = L.base.tgt.img[0][l]; c.tdepth = L.base.tgt.img[1][l];
            c.tgx = L.base.tgt.img[2][l]; c.tgy = L.base.tgt.img[3][l]; c.tdx = L.base.tgt.img[4][l]; c.tdy = L.base.tgt.img[5][l];
            c.fx = cam.fx; c.fy = cam.fy; c.cx = cam.cx; c.cy = cam.cy; c.w = cam.w;
            sh.ctx = c;
        }
        const bool lists = l == L.list_level;
        int2 *cand = L.base.cand;
        for (int j = 0; j < L.iterations[l]; ++j)
        {
            if (sh.break_level == l) break; // every CTA computed the same value
            if (stamp) sh.t[0] = timer_ns();
            if (threadIdx.x == 0) warp_matrices(cam, sh.T, sh.M, sh.M + 9);
            __syncthreads();
            // ---- one phase: candidate, acceptance (pointer jumping over lazily computed chains), Jacobian rows, fold ----
            unsigned int *jump = L.jump + (size_t)(sh.iteration & 1) * L.jump_stride;
            const CandFn candfn = {sd, td, sh.M, w, h};
            double c0 = 0.0, c1 = 0.0;
            float *stage = sh.u.stage[warp];
            for (int trip = first_trip; trip < n_trips; trip += trip_step)
            {
                const int s = trip * 32 + lane;
                float J[2][6], res[2];
                int rows = 0;
                bool ok = false;
                if (s < n)
                {
                    const int2 c = candfn(s);
                    if (lists) cand[s] = c;
                    ok = resolve_accept_lazy(candfn, jump, s, c); // AddElementToCorrespondenceMap, resolved along the chain
                    if (lists) L.base.accepted[s] = ok;
                    if (ok) rows = compute_rows<TERM>(sh.ctx, sh.T, s, c.x, J, res);
                }
                float comp[8];
#pragma unroll
                for (int e = 0; e < 6; ++e) comp[e] = rows > 0 ? J[0][e] : 0.0f;
                comp[6] = rows > 0 ? res[0] : 0.0f;
                comp[7] = ok ? 1.0f : 0.0f;
                warp_fold_outer8(stage, lane, comp, c0, c1);
                if (TERM == 0)
                {
#pragma unroll
                    for (int e = 0; e < 6; ++e) comp[e] = rows > 1 ? J[1][e] : 0.0f;
                    comp[6] = rows > 1 ? res[1] : 0.0f;
                    comp[7] = 0.0f;
                    warp_fold_outer8(stage, lane, comp, c0, c1);
                }
            }
            {   // the words of the NEXT iteration start out unset (its level may have four times the pixels: clear the whole buffer)
                unsigned int *jump_next = L.jump + (size_t)((sh.iteration + 1) & 1) * L.jump_stride;
                for (unsigned int i = blockIdx.x * kOdo2Threads + threadIdx.x; i < L.jump_stride; i += gridDim.x * kOdo2Threads)
                    __stcg(&jump_next[i], kJumpUnset);
            }
            if (stamp) sh.t[1] = sh.t[2] = timer_ns();
            // ---- CTA partial ----
            __syncthreads(); // every warp is done with its staging area (aliased below)
            *reinterpret_cast<double2 *>(&sh.u.wsum[warp][2 * lane]) = make_double2(c0, c1);
            __syncthreads();
            constexpr unsigned long long need = odo2_need_mask();
            const int e = threadIdx.x & 63, grp = threadIdx.x >> 6;
            const bool needed = (need >> e) & 1ull;
            double v = 0.0;
            if (needed && grp < 8)
                v = (sh.u.wsum[4 * grp][e] + sh.u.wsum[4 * grp + 1][e]) + (sh.u.wsum[4 * grp + 2][e] + sh.u.wsum[4 * grp + 3][e]);
            __syncthreads();
            if (grp < 8) sh.u.chunk[grp][e] = v;
            __syncthreads();
            double *mine = L.partials + ((size_t)(sh.iteration & 1) * n_cta + blockIdx.x) * 64;
            if (threadIdx.x < 64 && needed)
            {
                double t = 0.0;
#pragma unroll
                for (int k = 0; k < 8; ++k) t += sh.u.chunk[k][e];
                __stcg(&mine[e], t);
            }
            __threadfence();
            __syncthreads();
            // ---- arrive, wait for everybody ----
            const double *all = L.partials + (size_t)(sh.iteration & 1) * n_cta * 64;
            ++n_sync;
            if (threadIdx.x == 0)
            {
                atomicAdd(L.sync, 1u);
                const unsigned int want = n_sync * (unsigned int)n_cta;
                while (*(volatile unsigned int *)L.sync < want) { }
                __threadfence();
            }
            __syncthreads();
            if (stamp) sh.t[3] = timer_ns();
            // ---- all partials -> the 8x8 sums, in one fixed order, on every CTA ----
            {
                double t = 0.0;
                if (needed)
                {
                    if (n_cta <= kOdo2Chunks * 10)
                    {
                        double ld[10];
#pragma unroll
                        for (int u = 0; u < 10; ++u)
                        {
                            const int b = grp + u * kOdo2Chunks;
                            ld[u] = b < n_cta ? __ldcg(&all[(size_t)b * 64 + e]) : 0.0;
                        }
#pragma unroll
                        for (int u = 0; u < 10; ++u) t += ld[u];
                    }
                    else
                        for (int b = grp; b < n_cta; b += kOdo2Chunks) t += __ldcg(&all[(size_t)b * 64 + e]);
                }
                sh.u.chunk[grp][e] = t;
                __syncthreads();
                if (threadIdx.x < 64)
                {
                    double tot = 0.0;
                    if (needed)
#pragma unroll
                        for (int k = 0; k < kOdo2Chunks; ++k) tot += sh.u.chunk[k][e];
                    sh.sum64[e] = tot;
                }
                __syncthreads();
                if (threadIdx.x < kOdoPacket)
                {
                    const int src = odo2_packet_source(threadIdx.x);
                    sh.packet[threadIdx.x] = src >= 0 ? sh.sum64[src] : 0.0;
                }
                __syncthreads();
            }
            if (stamp) sh.t[4] = timer_ns();
            // ---- the same solve on every CTA (DoSingleIteration's tail, MultiScaleComputing's early exit) ----
            if (threadIdx.x == 0)
            {
                const int cnt = (int)(sh.packet[28] + 0.5);
                odo_solve_core(sh.packet, sh.T);
                const int it = sh.iteration;
                if (blockIdx.x == 0 && it < kMaxTrace)
                {
                    st->trace_count[it] = cnt;
                    for (int k = 0; k < 16; ++k) st->trace_T[it][k] = sh.T[k];
                }
                sh.iteration = it + 1;
                sh.last_count = cnt;
                if ((double)fdiv((float)cnt, (float)L.base.full_pixels) > 0.9) sh.break_level = l;
            }
            __syncthreads();
            if (stamp)
            {
                sh.t[5] = timer_ns();
                // candidates, acceptance, rows | (unused) | publish + barrier + sum of partials | solve
                sh.ph[0] += sh.t[1] - sh.t[0]; sh.ph[1] += sh.t[2] - sh.t[1]; sh.ph[2] += sh.t[4] - sh.t[2]; sh.ph[3] += sh.t[5] - sh.t[4];
            }
        }
    }
    if (stamp)
    {
        for (int k = 0; k < 16; ++k) st->T[k] = sh.T[k];
        for (int k = 0; k < kOdoPacket; ++k) st->packet[k] = sh.packet[k];
        st->iteration = sh.iteration;
        st->last_count = sh.last_count;
        st->break_level = sh.break_level;
        for (int k = 0; k < 4; ++k) st->phase_ns[k] += sh.ph[k];
        st->tail_ns += sh.t[5] - sh.t[3];
    }
}

// ---------------------------------------------------------------------------------------------------------
// ordered compaction of the accepted pixels (raster order of the source, like the reference's second loop)
// ---------------------------------------------------------------------------------------------------------
constexpr int kCompactTile = 1024;
__device__ __forceinline__ unsigned int block_scan_1024(unsigned int v, unsigned int *warp_sums, unsigned int *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const unsigned int m = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += m;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0)
    {
        unsigned int ws = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const unsigned int m = __shfl_up_sync(0xffffffffu, ws, o);
            if (lane >= o) ws += m;
        }
        warp_sums[lane] = ws;
    }
    __syncthreads();
    *total = warp_sums[31];
    const unsigned int r = inc + (warp ? warp_sums[warp - 1] : 0u);
    __syncthreads();
    return r;
}
__global__ void __launch_bounds__(1024) odo_tile_counts_kernel(const unsigned char *__restrict__ accepted, int n, unsigned int *tile_counts)
{
    __shared__ unsigned int warp_sums[32];
    const int i = blockIdx.x * kCompactTile + threadIdx.x;
    unsigned int total;
    block_scan_1024(i < n ? accepted[i] : 0u, warp_sums, &total);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) odo_tile_scan_kernel(unsigned int *tile_counts, int n_tiles, OdoState *st)
{
    __shared__ unsigned int warp_sums[32];
    unsigned int carry = 0;
    for (int base = 0; base < n_tiles; base += 1024)
    {
        const int i = base + threadIdx.x;
        const unsigned int v = i < n_tiles ? tile_counts[i] : 0u;
        unsigned int total;
        const unsigned int inc = block_scan_1024(v, warp_sums, &total);
        if (i < n_tiles) tile_counts[i] = carry + inc - v;
        carry += total;
    }
    if (threadIdx.x == 0) st->n_pairs = (int)carry;
}
// pairs: (v_s, u_s, v_t, u_t) per accepted pixel
__global__ void __launch_bounds__(1024) odo_compact_kernel(const unsigned char *__restrict__ accepted, const int2 *__restrict__ cand, int n,
                                                           int w, const unsigned int *__restrict__ tile_offsets, uint4 *__restrict__ pairs)
{
    __shared__ unsigned int warp_sums[32];
    const int i = blockIdx.x * kCompactTile + threadIdx.x;
    const unsigned int f = i < n ? accepted[i] : 0u;
    unsigned int total;
    const unsigned int inc = block_scan_1024(f, warp_sums, &total);
    if (!f) return;
    const unsigned int pos = tile_offsets[blockIdx.x] + inc - 1u;
    const int t = cand[i].x;
    pairs[pos] = make_uint4((unsigned int)(i / w), (unsigned int)(i % w), (unsigned int)(t / w), (unsigned int)(t % w));
}

// ---------------------------------------------------------------------------------------------------------
// NormalizeIntensity: the reference's means are SEQUENTIAL float32 sums over the correspondence list
// (DenseOdometryFunction.cpp:131-141); 0.5 / mean then scales the image, so the sum has to be reproduced bit for bit.
//
// One CTA per image.  All warps gather the next chunk of 32 K values (list order) into registers while the current chunk,
// staged in shared memory, is folded into the running sum S, 32 values per step.  A step is exact without 32 dependent additions whenever S
// stays inside its binade [2^e, 2^(e+1)): there fl(S + x) = S + U * rne(x / U) with U = ulp(S) = 2^(e-23), because S is
// a multiple of U -- so the 32 roundings are independent, their integer sum is a shuffle reduction, and S advances by
// one exact integer add.  Steps that cross a binade, hit a rounding tie (whose direction depends on the parity of the
// running sum) or see a negative / non-finite value take the literal path: 32 dependent float additions in list order.
// ---------------------------------------------------------------------------------------------------------

__device__ __forceinline__ float sequential_add32(float sum, float v, int count)
{
    if (count == 32)
    {
#pragma unroll
        for (int k = 0; k < 32; ++k) sum = fadd(sum, __shfl_sync(0xffffffffu, v, k));
    }
    else
        for (int k = 0; k < count; ++k) sum = fadd(sum, __shfl_sync(0xffffffffu, v, k));
    return sum;
}
// the binade of a running sum as its biased exponent, or -1 when the shortcut does not apply (zero, tiny, huge, negative, NaN)
__device__ __forceinline__ int binade_of(float sum)
{
    const unsigned int sb = __float_as_uint(sum);
    const int e = (int)(sb >> 23) & 0xff;
    return (e >= 64 && e <= 200 && !(sb >> 31)) ? e : -1;
}
// 32 values against the binade e: the integer number of ulps they add (warp-uniform), or -1 if any lane needs the literal path
__device__ __forceinline__ int ulps_of_step(float x, int e)
{
    const float inv_u = __uint_as_float((unsigned int)(277 - e) << 23); // 2^(23 - (e - 127))
    const float mq = x * inv_u;                                         // exact power-of-two scaling
    const float r = rintf(mq);
    const bool lane_ok = e >= 0 && x >= 0.0f && mq < 16777216.0f && fabsf(mq - r) != 0.5f;
    int t = lane_ok ? (int)r : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o); // < 2^29
    return __all_sync(0xffffffffu, lane_ok) ? t : -1;
}
// sum advanced by `ulps` units of its binade e (exact); false if that would leave the binade
__device__ __forceinline__ bool advance_in_binade(float &sum, int e, int ulps)
{
    const int si = (int)(sum * __uint_as_float((unsigned int)(277 - e) << 23)); // exact: in [2^23, 2^24)
    const int sn = si + ulps;
    if (ulps < 0 || sn > (1 << 24)) return false;
    sum = (float)sn * __uint_as_float((unsigned int)(e - 23) << 23); // exact
    return true;
}

// the two intensity lists of NormalizeIntensity (source image at (v_s, u_s), target image at (v_t, u_t)), gathered by the whole
// grid into contiguous arrays: the sequential kernel below runs on two SMs only and must not do 2 x 300 k scattered loads
__global__ void __launch_bounds__(256) odo_gather_gray_kernel(const uint4 *__restrict__ pairs, const OdoState *st, const float *__restrict__ sgray,
                                                              const float *__restrict__ tgray, int w, float *__restrict__ vals, int stride)
{
    const int n = st->n_pairs;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint4 p = pairs[i];
        vals[i] = sgray[p.x * w + p.y];
        vals[stride + i] = tgray[p.z * w + p.w];
    }
}

// Step totals for every binade the running sum can visit: totals[which][b][step] = integer number of ulps the 32 values of
// `step` add to a sum in binade kBinLo + b, or -1 when a value needs the literal path there.  One warp per (list, step), the
// whole grid: this is the part of the work that does not depend on the running sum.
constexpr int kBinLo = 121, kBinN = 32; // biased exponents 121..152: sums from 2^-6 up to 2^26
__global__ void __launch_bounds__(256) odo_mean_totals_kernel(const float *__restrict__ vals, int stride, const OdoState *st, int *__restrict__ totals,
                                                              int max_steps)
{
    const int n = st->n_pairs, n_steps = (n + 31) / 32;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int ws = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ws < 2 * n_steps; ws += warps)
    {
        const int which = ws >= n_steps, step = ws - which * n_steps;
        const int idx = step * 32 + lane;
        const float x = idx < n ? vals[(size_t)which * stride + idx] : 0.0f;
        int mine = -1;
#pragma unroll 8
        for (int b = 0; b < kBinN; ++b)
        {
            const int t = ulps_of_step(x, kBinLo + b);
            if (lane == b) mine = t;
        }
        totals[((size_t)which * kBinN + lane) * max_steps + step] = mine;
    }
}

// The sequential part, one warp per list: the running sum S walks the steps; while it stays in a binade the totals of that
// binade are folded 32 steps at a time (prefix scan, one exact integer add), a step that would leave the binade or that holds
// a rounding tie is taken on its own -- its own binade, literal 32 additions when the shortcut fails -- and so are the steps
// after it until 16 in a row took the shortcut (while the sum is small its ulp is close to the values' own and ties abound).
__global__ void __launch_bounds__(32) odo_sequential_mean_kernel(const float *__restrict__ vals, int stride, const int *__restrict__ totals,
                                                                 int max_steps, OdoState *st)
{
    const int which = blockIdx.x; // 0: source image at (v_s, u_s), 1: target image at (v_t, u_t)
    const float *__restrict__ list = vals + (size_t)which * stride;
    const int n = st->n_pairs, n_steps = (n + 31) / 32;
    const int lane = threadIdx.x;
    float sum = 0.0f;
    int pos = 0;
    while (pos < n_steps)
    {
        const int e_spec = binade_of(sum);
        const int b = e_spec - kBinLo;
        int f = pos;
        if (e_spec >= 0 && b >= 0 && b < kBinN)
        {
            const int *__restrict__ tot = totals + ((size_t)which * kBinN + b) * max_steps;
            const int si0 = (int)(sum * __uint_as_float((unsigned int)(277 - e_spec) << 23));
            int carry = 0, applied = 0;
            bool failed = false;
            f = n_steps;
            // rows of 32 steps, sixteen rows of totals in flight at a time (a row costs one global-memory latency otherwise)
            for (int r0 = pos >> 5; r0 * 32 < n_steps && !failed; r0 += 16)
            {
                int tq[16];
#pragma unroll
                for (int q = 0; q < 16; ++q)
                {
                    const int step = (r0 + q) * 32 + lane;
                    tq[q] = step >= pos && step < n_steps ? tot[step] : 0;
                }
#pragma unroll
                for (int q = 0; q < 16; ++q)
                {
                    if (failed || (r0 + q) * 32 >= n_steps) break;
                    const int r = r0 + q;
                    const int step = r * 32 + lane;
                    const bool valid = step >= pos && step < n_steps;
                    const int t = tq[q];
                    const bool bad = valid && t < 0;
                    const int own = bad ? 0 : t;
                    int inc = own;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1)
                    {
                        const int v = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= o) inc += v;
                    }
                    const bool over = valid && (long long)si0 + carry + inc > (1 << 24); // this step would leave the binade
                    const unsigned int mask = __ballot_sync(0xffffffffu, bad || over);
                    if (mask)
                    {
                        const int fl = __ffs(mask) - 1;
                        f = r * 32 + fl;
                        applied = carry + __shfl_sync(0xffffffffu, inc - own, fl); // ulps of the steps before f
                        failed = true;
                    }
                    else
                        carry += __shfl_sync(0xffffffffu, inc, 31);
                }
            }
            if (!failed) applied = carry;
            if (applied > 0) advance_in_binade(sum, e_spec, applied);
        }
        int calm = 0;
        while (f < n_steps && calm < 16)
        {
            // the values of the next 16 steps in flight together
            float xq[16];
#pragma unroll
            for (int q = 0; q < 16; ++q)
            {
                const int idx = (f + q) * 32 + lane;
                xq[q] = idx < n ? list[idx] : 0.0f;
            }
#pragma unroll
            for (int q = 0; q < 16; ++q)
            {
                if (f >= n_steps || calm >= 16) break;
                const float x = xq[q];
                const int e = binade_of(sum);
                const int t = ulps_of_step(x, e);
                if (e >= 0 && t >= 0 && advance_in_binade(sum, e, t)) ++calm;
                else
                {
                    sum = sequential_add32(sum, x, min(32, n - f * 32));
                    calm = 0;
                }
                ++f;
            }
        }
        pos = f;
    }
    if (lane == 0)
    {
        const float mean = fdiv(sum, (float)n); // mean /= (float)correspondence.size()
        if (which) st->mean_tgt = mean; else st->mean_src = mean;
    }
}
// tool::LinearTransform(image, 0.5 / mean, 0.0): the scale is computed in double and passed as float
__global__ void odo_scale_kernel(float *__restrict__ sgray, float *__restrict__ tgray, int n, const OdoState *st)
{
    const float ss = (float)(0.5 / (double)st->mean_src), ts = (float)(0.5 / (double)st->mean_tgt);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        sgray[i] = fadd(fmul(sgray[i], ss), 0.0f);
        if (tgray != sgray) tgray[i] = fadd(fmul(tgray[i], ts), 0.0f);
    }
}

// ---------------------------------------------------------------------------------------------------------
// result assembly: correspondence_set pairs xyz_s[v_s][u_s] with xyz_t[v_s][u_s] (the SAME pixel on the target's
// XYZ image -- reference quirk, Odometry.cpp:672-682) and rmse = ComputeReprojectionError3D (Geometry.cpp:45-59)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void backproject(const OdoCam &cam, float z, int u, int v, float *p)
{
    p[0] = p[1] = p[2] = -1.0f;
    if (z > 0)
    {
        p[0] = fdiv(fmul(fsub((float)u, cam.cx), z), cam.fx);
        p[1] = fdiv(fmul(fsub((float)v, cam.cy), z), cam.fy);
        p[2] = z;
    }
}
__global__ void __launch_bounds__(kOdoThreads) odo_rmse_kernel(const uint4 *__restrict__ pairs, OdoState *st, const float *__restrict__ sdepth,
                                                               const float *__restrict__ tdepth, OdoCam cam, float *__restrict__ corr_xyz)
{
    const int n = st->n_pairs;
    const float *T = st->T;
    double sum = 0.0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
    {
        const uint4 p = pairs[k];
        const int s = p.x * cam.w + p.y;
        float a[3], b[3];
        backproject(cam, sdepth[s], (int)p.y, (int)p.x, a);
        backproject(cam, tdepth[s], (int)p.y, (int)p.x, b);
        if (corr_xyz)
        {
            float *o = corr_xyz + 6 * (size_t)k;
            o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; o[3] = b[0]; o[4] = b[1]; o[5] = b[2];
        }
        // TransformPoint: T * (x, y, z, 1), head<3>() / w
        const float wv = row_xyz1(T[3], T[7], T[11], T[15], a[0], a[1], a[2]);
        const float e0 = fsub(fdiv(row_xyz1(T[0], T[4], T[8], T[12], a[0], a[1], a[2]), wv), b[0]);
        const float e1 = fsub(fdiv(row_xyz1(T[1], T[5], T[9], T[13], a[0], a[1], a[2]), wv), b[1]);
        const float e2 = fsub(fdiv(row_xyz1(T[2], T[6], T[10], T[14], a[0], a[1], a[2]), wv), b[2]);
        sum += (double)fadd(fmul(e0, e0), fadd(fmul(e1, e1), fmul(e2, e2)));
    }
    sum = warp_sum_d(sum);
    if ((threadIdx.x & 31) == 0 && sum != 0.0) atomicAdd(&st->rmse_sum, sum);
}

} // namespace opb

using namespace opb;

// Device buffers of released frames, kept for the next frame of the same size: a tracking loop creates and releases one
// RGBDFrame per image, and cudaMalloc / cudaFree (which synchronises the device) cost more than the tracking itself.
// Shared between the odometry object and its frames so that either may be destroyed first.
struct FrameBuffers
{
    uint8_t *bgr = nullptr;
    void *depth = nullptr;
    float *images = nullptr;
};
struct FramePool
{
    int device = 0;
    std::mutex lock;
    std::vector<FrameBuffers> free_sets;
    ~FramePool()
    {
        cudaSetDevice(device);
        for (FrameBuffers &b : free_sets) { cudaFree(b.bgr); cudaFree(b.depth); cudaFree(b.images); }
        cudaGetLastError();
    }
};
constexpr size_t kFramePoolCap = 8;

struct opb_odometry
{
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    opb_odometry_desc desc;
    OdoCam cams[kMaxLevels];
    int2 *d_cand = nullptr;
    unsigned char *d_accepted = nullptr;
    double *d_partials = nullptr;
    unsigned int *d_tiles = nullptr;
    uint4 *d_pairs = nullptr;
    float *d_corr_xyz = nullptr;
    float *d_identity = nullptr;
    float *d_tmp = nullptr; // 2 images: un-blurred gray and depth
    OdoState *d_state = nullptr;
    OdoState *h_state = nullptr; // pinned
    int max_blocks = 0;
    std::shared_ptr<FramePool> frame_pool;
    float *d_vals = nullptr;        // NormalizeIntensity: the two gathered intensity lists
    int *d_mean_totals = nullptr;   // ... and their step totals per binade
    unsigned int *d_sync = nullptr; // barrier words of the persistent loop kernel
    int coop_ctas_per_sm = 0;       // resident CTAs per SM of odo_loop_kernel; 0: cooperative launch unavailable
    int loop_form = -1;             // opb_odometry_set_loop_form: -1 default (OPB_ODO_PERSISTENT, else 1)
    bool loop2_ok = false;          // odo_loop2_kernel (one CTA of 1024 threads per SM) can be launched cooperatively
    double *d_partials2 = nullptr;  // its per-CTA 8x8 partials, two generations
    unsigned int *d_jump = nullptr; // ... and its per-pixel chain words
    bool profiling = false;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    float last_ms = 0;
};

struct opb_frame
{
    std::shared_ptr<FramePool> pool;
    opb_odometry *owner = nullptr; // identity check only, never dereferenced on destruction
    int device = 0;
    int w = 0, h = 0, depth_type = 0;
    uint8_t *d_bgr = nullptr;
    void *d_depth = nullptr;
    float *d_images = nullptr;
    FrameImages im;
    bool initialized = false; // gray + depth32f
    bool preprocessed = false; // + pyramids (IsPreprocessedDense)
};

static size_t level_pixels(const opb_odometry *o, int l) { return (size_t)(o->desc.width >> l) * (size_t)(o->desc.height >> l); }

static void setup_cameras(opb_odometry *o)
{
    // Odometry::CreatePyramidCameras (Odometry.h:110-120) / PinholeCamera::GenerateNextPyramid (Camera.h:38-42)
    o->cams[0] = {o->desc.fx, o->desc.fy, o->desc.cx, o->desc.cy, o->desc.width, o->desc.height};
    for (int l = 1; l < kMaxLevels; ++l)
        o->cams[l] = {o->cams[l - 1].fx / 2, o->cams[l - 1].fy / 2, o->cams[l - 1].cx / 2, o->cams[l - 1].cy / 2, o->cams[l - 1].w / 2,
                      o->cams[l - 1].h / 2};
}

static int check_desc(const opb_odometry_desc *d)
{
    if (d->width <= 0 || d->height <= 0 || (long long)d->width * d->height > (1 << 26)) { set_error("bad image size %dx%d", d->width, d->height); return OPB_ERR_INVALID; }
    if (d->levels < 1 || d->levels > kMaxLevels) { set_error("levels must be 1..%d", kMaxLevels); return OPB_ERR_INVALID; }
    if ((d->width >> (d->levels - 1)) < 2 || (d->height >> (d->levels - 1)) < 2) { set_error("image too small for %d levels", d->levels); return OPB_ERR_INVALID; }
    int total = 0;
    for (int l = 0; l < d->levels; ++l)
    {
        if (d->iterations[l] < 0) { set_error("negative iteration count"); return OPB_ERR_INVALID; }
        total += d->iterations[l];
    }
    if (!(d->fx != 0.0f && d->fy != 0.0f)) { set_error("zero focal length"); return OPB_ERR_INVALID; }
    (void)total;
    return OPB_OK;
}

extern "C"
{
void opb_odometry_desc_default(opb_odometry_desc *d)
{
    if (!d) return;
    memset(d, 0, sizeof(*d));
    // camera::PinholeCamera() (Camera.h:94-105)
    d->fx = 514.817f; d->fy = 515.375f; d->cx = 318.771f; d->cy = 238.447f;
    d->width = 640; d->height = 480; d->depth_scale = 1000.0f;
    d->levels = 3;       // Odometry.h:168
    d->iterations[0] = 4; d->iterations[1] = 8; d->iterations[2] = 16; // Odometry.h:170, indexed by level
}

void opb_odometry_destroy(opb_odometry *o)
{
    if (!o) return;
    cudaSetDevice(o->device);
    if (o->stream) cudaStreamSynchronize(o->stream);
    cudaFree(o->d_cand); cudaFree(o->d_accepted); cudaFree(o->d_partials); cudaFree(o->d_tiles); cudaFree(o->d_pairs);
    cudaFree(o->d_corr_xyz); cudaFree(o->d_identity); cudaFree(o->d_tmp); cudaFree(o->d_state); cudaFree(o->d_sync); cudaFree(o->d_partials2); cudaFree(o->d_jump); cudaFree(o->d_vals); cudaFree(o->d_mean_totals);
    if (o->h_state) cudaFreeHost(o->h_state);
    for (int i = 0; i < 2; ++i) if (o->ev[i]) cudaEventDestroy(o->ev[i]);
    if (o->own_stream && o->stream) cudaStreamDestroy(o->stream);
    cudaGetLastError();
    delete o;
}

int opb_odometry_create(const opb_odometry_desc *desc, opb_odometry **out)
{
    if (!desc || !out) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    *out = nullptr;
    int rc = check_desc(desc);
    if (rc) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
        cudaGetLastError();
        set_error("no CUDA device: onepiece_b200 has no CPU path");
        return OPB_ERR_CUDA;
    }
    if (desc->device < 0 || desc->device >= ndev) { set_error("device %d out of range (%d devices)", desc->device, ndev); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(desc->device));
    opb_odometry *o = new opb_odometry();
    o->device = desc->device;
    o->desc = *desc;
    setup_cameras(o);
    o->frame_pool = std::make_shared<FramePool>();
    o->frame_pool->device = desc->device;
    cudaDeviceProp prop;
    OPB_CUDA(cudaGetDeviceProperties(&prop, desc->device));
    o->sm_count = prop.multiProcessorCount;
    if (desc->stream) o->stream = (cudaStream_t)desc->stream;
    else { OPB_CUDA(cudaStreamCreateWithFlags(&o->stream, cudaStreamNonBlocking)); o->own_stream = true; }
    const size_t n = level_pixels(o, 0);
    o->max_blocks = o->sm_count * 4;
    const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    cudaError_t e = cudaMalloc(&o->d_cand, n * sizeof(int2));
    if (e == cudaSuccess) e = cudaMalloc(&o->d_accepted, n);
    if (e == cudaSuccess) e = cudaMalloc(&o->d_partials, (size_t)o->max_blocks * kOdoPacket * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&o->d_tiles, (n / kCompactTile + 2) * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&o->d_pairs, n * sizeof(uint4));
    if (e == cudaSuccess) e = cudaMalloc(&o->d_corr_xyz, n * 6 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&o->d_identity, sizeof(I));
    if (e == cudaSuccess) e = cudaMalloc(&o->d_tmp, 2 * n * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&o->d_state, sizeof(OdoState));
    if (e == cudaSuccess) e = cudaHostAlloc(&o->h_state, sizeof(OdoState), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaMemcpy(o->d_identity, I, sizeof(I), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(o->d_state, 0, sizeof(OdoState));
    if (e == cudaSuccess) e = cudaMalloc(&o->d_sync, 4 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&o->d_vals, 2 * n * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&o->d_mean_totals, 2 * (size_t)kBinN * ((n + 31) / 32) * sizeof(int));
    if (e == cudaSuccess)
    {
        int coop = 0, occ[3] = {0, 0, 0};
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, desc->device);
        if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], odo_loop_kernel<0>, kOdoThreads, 0) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], odo_loop_kernel<1>, kOdoThreads, 0) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[2], odo_loop_kernel<2>, kOdoThreads, 0) == cudaSuccess)
            o->coop_ctas_per_sm = occ[0] < occ[1] ? (occ[0] < occ[2] ? occ[0] : occ[2]) : (occ[1] < occ[2] ? occ[1] : occ[2]);
        cudaGetLastError();
        if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], odo_loop2_kernel<0>, kOdo2Threads, 0) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], odo_loop2_kernel<1>, kOdo2Threads, 0) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[2], odo_loop2_kernel<2>, kOdo2Threads, 0) == cudaSuccess &&
            occ[0] > 0 && occ[1] > 0 && occ[2] > 0 &&
            cudaMalloc(&o->d_partials2, (size_t)2 * o->sm_count * 64 * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&o->d_jump, 2 * n * sizeof(unsigned int)) == cudaSuccess)
            o->loop2_ok = true;
        cudaGetLastError();
    }
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreate(&o->ev[i]);
    if (e != cudaSuccess)
    {
        set_error("odometry workspace allocation failed: %s", cudaGetErrorString(e));
        opb_odometry_destroy(o);
        return OPB_ERR_CUDA;
    }
    *out = o;
    return OPB_OK;
}

int opb_odometry_set_profiling(opb_odometry *o, int on)
{
    if (!o) { set_error("odometry is NULL"); return OPB_ERR_INVALID; }
    o->profiling = on != 0;
    return OPB_OK;
}
int opb_odometry_set_loop_form(opb_odometry *o, int form)
{
    if (!o) { set_error("odometry is NULL"); return OPB_ERR_INVALID; }
    if (form < -1 || form > 2) { set_error("loop form %d: -1 default, 0 separate launches, 1 second persistent form, 2 first persistent form", form); return OPB_ERR_INVALID; }
    o->loop_form = form;
    return OPB_OK;
}
int opb_odometry_last_phases(opb_odometry *o, uint64_t phase_ns[4])
{
    if (!o || !phase_ns) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    for (int i = 0; i < 4; ++i) phase_ns[i] = o->h_state->phase_ns[i];
    return OPB_OK;
}
int opb_odometry_last_timing(opb_odometry *o, float *tracking_ms, float *solve_tail_us)
{
    if (!o) { set_error("odometry is NULL"); return OPB_ERR_INVALID; }
    if (tracking_ms) *tracking_ms = o->last_ms;
    if (solve_tail_us) *solve_tail_us = o->h_state->iteration > 0 ? (float)o->h_state->tail_ns * 1e-3f / (float)(o->h_state->iteration + 1) : 0.0f;
    return OPB_OK;
}

void opb_frame_destroy(opb_frame *f)
{
    if (!f) return;
    bool kept = false;
    if (f->pool && f->d_bgr && f->d_depth && f->d_images)
    {
        // the next user of these buffers uploads on the odometry's stream, i.e. after every kernel that still reads them
        std::lock_guard<std::mutex> g(f->pool->lock);
        if (f->pool->free_sets.size() < kFramePoolCap)
        {
            FrameBuffers b;
            b.bgr = f->d_bgr; b.depth = f->d_depth; b.images = f->d_images;
            f->pool->free_sets.push_back(b);
            kept = true;
        }
    }
    if (!kept)
    {
        cudaSetDevice(f->device); // cudaFree synchronises the device: no kernel can still be reading the frame
        cudaFree(f->d_bgr); cudaFree(f->d_depth); cudaFree(f->d_images);
        cudaGetLastError();
    }
    delete f;
}

// geometry::RGBDFrame(rgb, depth) (RGBDFrame.h:14-19): keeps the raw images; the dense cache is filled on first use
int opb_frame_create(opb_odometry *o, const uint8_t *bgr, const void *depth, int depth_type, opb_frame **out)
{
    if (!o || !bgr || !depth || !out) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    *out = nullptr;
    if (depth_type != OPB_DEPTH_F32 && depth_type != OPB_DEPTH_U16)
    {
        // ConvertDepthTo32FNaN: "Unknown depth image type" + exit(1) in the reference (DenseOdometryFunction.cpp:51-55)
        set_error("[ImageProcessing]::[ERROR]::Unknown depth image type: %d", depth_type);
        return OPB_ERR_UNSUPPORTED;
    }
    OPB_CUDA(cudaSetDevice(o->device));
    opb_frame *f = new opb_frame();
    f->owner = o; f->device = o->device; f->w = o->desc.width; f->h = o->desc.height; f->depth_type = depth_type;
    const size_t n = level_pixels(o, 0);
    size_t total = 0;
    for (int l = 0; l < o->desc.levels; ++l) total += 6 * level_pixels(o, l);
    f->pool = o->frame_pool;
    cudaError_t e = cudaSuccess;
    {
        std::lock_guard<std::mutex> g(f->pool->lock);
        if (!f->pool->free_sets.empty())
        {
            const FrameBuffers b = f->pool->free_sets.back();
            f->pool->free_sets.pop_back();
            f->d_bgr = b.bgr; f->d_depth = b.depth; f->d_images = b.images;
        }
    }
    if (!f->d_images)
    {
        e = cudaMalloc(&f->d_bgr, n * 3);
        if (e == cudaSuccess) e = cudaMalloc(&f->d_depth, n * 4); // sized for float depth, so that any frame can reuse it
        if (e == cudaSuccess) e = cudaMalloc(&f->d_images, total * sizeof(float));
    }
    if (e != cudaSuccess)
    {
        set_error("frame allocation failed: %s", cudaGetErrorString(e));
        opb_frame_destroy(f);
        return OPB_ERR_CUDA;
    }
    float *p = f->d_images;
    memset(&f->im, 0, sizeof(f->im));
    for (int a = 0; a < 6; ++a)
        for (int l = 0; l < o->desc.levels; ++l) { f->im.img[a][l] = p; p += level_pixels(o, l); }
    e = cudaMemcpyAsync(f->d_bgr, bgr, n * 3, cudaMemcpyDefault, o->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(f->d_depth, depth, n * (depth_type == OPB_DEPTH_U16 ? 2 : 4), cudaMemcpyDefault, o->stream);
    if (e != cudaSuccess)
    {
        set_error("frame upload failed: %s", cudaGetErrorString(e));
        opb_frame_destroy(f);
        return OPB_ERR_CUDA;
    }
    if (cudaStreamSynchronize(o->stream) != cudaSuccess)
    {
        set_error("frame upload failed");
        opb_frame_destroy(f);
        return OPB_ERR_CUDA;
    }
    *out = f;
    return OPB_OK;
}
} // extern "C"

static const dim3 kImgBlock(32, 8);
static dim3 img_grid(int w, int h, int z) { return dim3((w + 31) / 32, (h + 7) / 8, z); }

// InitializeRGBDDenseTracking (Odometry.cpp:609-620)
static int frame_initialize(opb_odometry *o, opb_frame *f)
{
    const int n = (int)level_pixels(o, 0);
    cudaStream_t s = o->stream;
    odo_convert_kernel<<<(n + 255) / 256, 256, 0, s>>>(f->d_bgr, f->d_depth, f->depth_type == OPB_DEPTH_U16, o->desc.depth_scale, n, o->d_tmp,
                                                       o->d_tmp + n);
    odo_blur3_kernel<<<img_grid(f->w, f->h, 2), kImgBlock, 0, s>>>(o->d_tmp, o->d_tmp + n, f->w, f->h, f->im.img[0][0], f->im.img[1][0]);
    OPB_CUDA(cudaGetLastError());
    f->initialized = true;
    return OPB_OK;
}
// CreateImagePyramid (Odometry.cpp:436-449)
static int frame_pyramids(opb_odometry *o, opb_frame *f)
{
    cudaStream_t s = o->stream;
    for (int l = 1; l < o->desc.levels; ++l)
        odo_pyrdown_kernel<<<img_grid(f->w >> l, f->h >> l, 2), kImgBlock, 0, s>>>(f->im.img[0][l - 1], f->im.img[1][l - 1], f->w >> (l - 1),
                                                                                  f->h >> (l - 1), f->im.img[0][l], f->im.img[1][l]);
    odo_sobel_kernel<<<img_grid(f->w, f->h, 4 * o->desc.levels), kImgBlock, 0, s>>>(f->im, f->w, f->h, o->desc.levels);
    OPB_CUDA(cudaGetLastError());
    f->preprocessed = true;
    return OPB_OK;
}

static int grid_for(const opb_odometry *o, size_t n)
{
    const size_t need = (n + kOdoThreads - 1) / kOdoThreads;
    return (int)(need < (size_t)o->max_blocks ? need : (size_t)o->max_blocks);
}

static OdoArgs make_args(opb_odometry *o, opb_frame *S, opb_frame *T, int level, int term)
{
    OdoArgs a;
    a.src = S->im; a.tgt = T->im; a.cam = o->cams[level]; a.level = level; a.term = term;
    a.full_pixels = o->desc.width * o->desc.height;
    a.cand = o->d_cand; a.accepted = o->d_accepted; a.partials = o->d_partials; a.st = o->d_state; a.T_override = nullptr;
    return a;
}

static void launch_iteration(opb_odometry *o, const OdoArgs &a, bool count_only)
{
    const size_t n = (size_t)a.cam.w * a.cam.h;
    const int nb = grid_for(o, n);
    cudaStream_t s = o->stream;
    odo_candidates_kernel<<<nb, kOdoThreads, 0, s>>>(a);
    if (count_only) odo_iteration_kernel<0, 1><<<nb, kOdoThreads, 0, s>>>(a);
    else if (a.term == 0) odo_iteration_kernel<0, 0><<<nb, kOdoThreads, 0, s>>>(a);
    else if (a.term == 1) odo_iteration_kernel<1, 0><<<nb, kOdoThreads, 0, s>>>(a);
    else odo_iteration_kernel<2, 0><<<nb, kOdoThreads, 0, s>>>(a);
}

// accepted flags + candidates of level `level` -> ordered pair list in o->d_pairs, count in state.n_pairs
static void launch_compaction(opb_odometry *o, int level)
{
    const int n = (int)level_pixels(o, level);
    const int tiles = (n + kCompactTile - 1) / kCompactTile;
    cudaStream_t s = o->stream;
    odo_tile_counts_kernel<<<tiles, 1024, 0, s>>>(o->d_accepted, n, o->d_tiles);
    odo_tile_scan_kernel<<<1, 1024, 0, s>>>(o->d_tiles, tiles, o->d_state);
    odo_compact_kernel<<<tiles, 1024, 0, s>>>(o->d_accepted, o->d_cand, n, o->cams[level].w, o->d_tiles, o->d_pairs);
}

// the identity-pose correspondences + NormalizeIntensity on both level-0 gray images, in place (Odometry.cpp:588-595)
static void launch_normalize(opb_odometry *o, opb_frame *S, opb_frame *T)
{
    OdoArgs a = make_args(o, S, T, 0, 0);
    a.T_override = o->d_identity;
    launch_iteration(o, a, true);
    launch_compaction(o, 0);
    cudaStream_t s = o->stream;
    const int n0 = (int)level_pixels(o, 0);
    odo_gather_gray_kernel<<<grid_for(o, n0), kOdoThreads, 0, s>>>(o->d_pairs, o->d_state, S->im.img[0][0], T->im.img[0][0], o->cams[0].w,
                                                                  o->d_vals, n0);
    const int max_steps = (n0 + 31) / 32;
    odo_mean_totals_kernel<<<grid_for(o, (size_t)n0 * 2), kOdoThreads, 0, s>>>(o->d_vals, n0, o->d_state, o->d_mean_totals, max_steps);
    odo_sequential_mean_kernel<<<2, 32, 0, s>>>(o->d_vals, n0, o->d_mean_totals, max_steps, o->d_state);
    const int n = n0;
    odo_scale_kernel<<<grid_for(o, n), kOdoThreads, 0, s>>>(S->im.img[0][0], T->im.img[0][0], n, o->d_state);
}

static int reset_state(opb_odometry *o, const float *init_T)
{
    OdoState *h = o->h_state;
    memset(h, 0, sizeof(OdoState));
    memcpy(h->T, init_T, 16 * sizeof(float));
    h->break_level = -1;
    OPB_CUDA(cudaMemcpyAsync(o->d_state, h, sizeof(OdoState), cudaMemcpyHostToDevice, o->stream));
    return OPB_OK;
}

// MultiScaleComputing + result assembly; both frames pre-processed
static int run_tracking(opb_odometry *o, opb_frame *S, opb_frame *T, int term, opb_tracking_result *res, uint32_t *pixel_pairs,
                        size_t pairs_cap, float *corr_xyz)
{
    cudaStream_t s = o->stream;
    static const int k_env = getenv("OPB_ODO_PERSISTENT") ? atoi(getenv("OPB_ODO_PERSISTENT")) : 1;
    const int k_persistent = o->loop_form >= 0 ? o->loop_form : k_env;
    int last_level = -1;
    for (int l = 0; l < o->desc.levels && last_level < 0; ++l)
        if (o->desc.iterations[l] > 0) last_level = l;
    if (k_persistent == 1 && o->loop2_ok)
    {
        // second persistent form: one CTA of 1024 threads per SM, one grid barrier per iteration, every CTA sums and solves
        OdoLoop2Args L;
        L.base = make_args(o, S, T, 0, term);
        for (int l = 0; l < kMaxLevels; ++l) { L.cams[l] = o->cams[l]; L.iterations[l] = l < o->desc.levels ? o->desc.iterations[l] : 0; }
        L.levels = o->desc.levels;
        L.list_level = last_level;
        L.partials = o->d_partials2;
        L.jump = o->d_jump;
        L.jump_stride = (unsigned int)level_pixels(o, 0);
        OPB_CUDA(cudaMemsetAsync(o->d_jump, 0xFF, (size_t)L.jump_stride * sizeof(unsigned int), s)); // generation 0 starts out unset
        L.sync = o->d_sync;
        OPB_CUDA(cudaMemsetAsync(o->d_sync, 0, 4 * sizeof(unsigned int), s));
        void *kargs[] = {(void *)&L};
        const void *fn = term == 0 ? (const void *)odo_loop2_kernel<0> : (term == 1 ? (const void *)odo_loop2_kernel<1> : (const void *)odo_loop2_kernel<2>);
        OPB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(o->sm_count), dim3(kOdo2Threads), kargs, 0, s));
    }
    else if (k_persistent && o->coop_ctas_per_sm > 0)
    {
        OdoLoopArgs L;
        L.base = make_args(o, S, T, 0, term);
        for (int l = 0; l < kMaxLevels; ++l) { L.cams[l] = o->cams[l]; L.iterations[l] = l < o->desc.levels ? o->desc.iterations[l] : 0; }
        L.levels = o->desc.levels;
        L.sync = o->d_sync;
        OPB_CUDA(cudaMemsetAsync(o->d_sync, 0, 4 * sizeof(unsigned int), s));
        const int per_sm = o->coop_ctas_per_sm < 2 ? o->coop_ctas_per_sm : 2;
        int nb = grid_for(o, level_pixels(o, 0));
        if (nb > o->sm_count * per_sm) nb = o->sm_count * per_sm;
        void *kargs[] = {(void *)&L};
        const void *fn = term == 0 ? (const void *)odo_loop_kernel<0> : (term == 1 ? (const void *)odo_loop_kernel<1> : (const void *)odo_loop_kernel<2>);
        OPB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(nb), dim3(kOdoThreads), kargs, 0, s));
    }
    else
        for (int l = o->desc.levels - 1; l >= 0; --l)
        {
            OdoArgs a = make_args(o, S, T, l, term);
            for (int j = 0; j < o->desc.iterations[l]; ++j) launch_iteration(o, a, false);
        }
    // the correspondences of the last executed iteration, in raster order (level 0 unless its iteration count is 0;
    // the reference then indexes the level-0 XYZ images with the coarser level's pixel coordinates, and so does this)
    if (last_level >= 0) launch_compaction(o, last_level);
    const int n0 = (int)level_pixels(o, 0);
    odo_rmse_kernel<<<grid_for(o, n0), kOdoThreads, 0, s>>>(o->d_pairs, o->d_state, S->im.img[1][0], T->im.img[1][0], o->cams[0],
                                                            corr_xyz ? o->d_corr_xyz : nullptr);
    OPB_CUDA(cudaGetLastError());
    OPB_CUDA(cudaMemcpyAsync(o->h_state, o->d_state, sizeof(OdoState), cudaMemcpyDeviceToHost, s));
    if (o->profiling) OPB_CUDA(cudaEventRecord(o->ev[1], s));
    OPB_CUDA(cudaStreamSynchronize(s));
    if (o->profiling) cudaEventElapsedTime(&o->last_ms, o->ev[0], o->ev[1]);
    const OdoState *h = o->h_state;
    memset(res, 0, sizeof(*res));
    memcpy(res->T, h->T, sizeof(res->T));
    res->n_correspondences = (size_t)h->n_pairs;
    res->iterations = h->iteration;
    res->rmse = sqrt(h->rmse_sum / (double)h->n_pairs);
    // return (float)correspondences.size() / (height * width) >= MIN_INLIER_RATIO_DENSE;   (Odometry.cpp:684)
    res->tracking_success = (double)((float)h->n_pairs / (float)(o->desc.width * o->desc.height)) >= 0.3;
    const int nt = h->iteration < kMaxTrace ? h->iteration : kMaxTrace;
    for (int i = 0; i < nt; ++i)
    {
        res->corr_per_iteration[i] = h->trace_count[i];
        memcpy(res->T_per_iteration[i], h->trace_T[i], 16 * sizeof(float));
    }
    const size_t np = (size_t)h->n_pairs;
    if (pixel_pairs && pairs_cap)
        OPB_CUDA(cudaMemcpy(pixel_pairs, o->d_pairs, (np < pairs_cap ? np : pairs_cap) * sizeof(uint4), cudaMemcpyDeviceToHost));
    if (corr_xyz && pairs_cap)
        OPB_CUDA(cudaMemcpy(corr_xyz, o->d_corr_xyz, (np < pairs_cap ? np : pairs_cap) * 6 * sizeof(float), cudaMemcpyDeviceToHost));
    res->status = OPB_OK;
    return OPB_OK;
}

static int check_pair(opb_odometry *o, opb_frame *S, opb_frame *T, int term)
{
    if (!o || !S || !T) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    if (S->owner != o || T->owner != o) { set_error("frame belongs to another odometry object"); return OPB_ERR_INVALID; }
    if (term < 0 || term > 2) { set_error("term_type must be 0 (hybrid), 1 (photo) or 2 (geometry)"); return OPB_ERR_INVALID; }
    return OPB_OK;
}

extern "C"
{
int opb_frame_preprocess(opb_odometry *o, opb_frame *f)
{
    if (!o || !f || f->owner != o) { set_error("bad frame"); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(o->device));
    if (f->preprocessed) return OPB_OK;
    int rc = frame_initialize(o, f);
    if (rc == OPB_OK) rc = frame_pyramids(o, f);
    return rc;
}

int opb_frame_is_preprocessed(const opb_frame *f) { return f && f->preprocessed; }

int opb_frame_image(opb_odometry *o, opb_frame *f, int what, int level, float *out)
{
    if (!o || !f || !out || f->owner != o) { set_error("bad argument"); return OPB_ERR_INVALID; }
    if (what < 0 || what >= 6 || level < 0 || level >= o->desc.levels) { set_error("no such image"); return OPB_ERR_INVALID; }
    if (!f->preprocessed) { set_error("frame is not pre-processed"); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(o->device));
    OPB_CUDA(cudaMemcpyAsync(out, f->im.img[what][level], level_pixels(o, level) * sizeof(float), cudaMemcpyDeviceToHost, o->stream));
    OPB_CUDA(cudaStreamSynchronize(o->stream));
    return OPB_OK;
}

int opb_odometry_dense_tracking_frames(opb_odometry *o, opb_frame *source, opb_frame *target, const float init_T[16], int term_type,
                                       opb_tracking_result *result, uint32_t *pixel_pairs, size_t pairs_cap, float *corr_xyz)
{
    int rc = check_pair(o, source, target, term_type);
    if (rc) return rc;
    if (!init_T || !result) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(o->device));
    if (o->profiling) OPB_CUDA(cudaEventRecord(o->ev[0], o->stream));
    // Odometry.cpp:571-587: pre-process once per frame
    if ((rc = opb_frame_preprocess(o, source))) return rc;
    if ((rc = opb_frame_preprocess(o, target))) return rc;
    if ((rc = reset_state(o, init_T))) return rc;
    launch_normalize(o, source, target);
    return run_tracking(o, source, target, term_type, result, pixel_pairs, pairs_cap, corr_xyz);
}

int opb_odometry_dense_tracking(opb_odometry *o, const uint8_t *src_bgr, const uint8_t *tgt_bgr, const void *src_depth,
                                const void *tgt_depth, int depth_type, const float init_T[16], int term_type,
                                opb_tracking_result *result, uint32_t *pixel_pairs, size_t pairs_cap, float *corr_xyz)
{
    if (!o || !init_T || !result) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    if (term_type < 0 || term_type > 2) { set_error("term_type must be 0 (hybrid), 1 (photo) or 2 (geometry)"); return OPB_ERR_INVALID; }
    opb_frame *S = nullptr, *T = nullptr;
    OPB_CUDA(cudaSetDevice(o->device));
    if (o->profiling) OPB_CUDA(cudaEventRecord(o->ev[0], o->stream));
    int rc = opb_frame_create(o, src_bgr, src_depth, depth_type, &S);
    if (rc == OPB_OK) rc = opb_frame_create(o, tgt_bgr, tgt_depth, depth_type, &T);
    // Odometry.cpp:482-509: initialise both, normalise the full-resolution gray images, THEN build the pyramids
    if (rc == OPB_OK) rc = frame_initialize(o, S);
    if (rc == OPB_OK) rc = frame_initialize(o, T);
    if (rc == OPB_OK) rc = reset_state(o, init_T);
    if (rc == OPB_OK)
    {
        launch_normalize(o, S, T);
        rc = frame_pyramids(o, S);
    }
    if (rc == OPB_OK) rc = frame_pyramids(o, T);
    if (rc == OPB_OK) rc = run_tracking(o, S, T, term_type, result, pixel_pairs, pairs_cap, corr_xyz);
    opb_frame_destroy(S);
    opb_frame_destroy(T);
    return rc;
}

int opb_odometry_single_iteration(opb_odometry *o, opb_frame *source, opb_frame *target, int level, float T_inout[16], int term_type,
                                  double sums43[43], uint32_t *pixel_pairs, size_t pairs_cap, size_t *n_pairs)
{
    int rc = check_pair(o, source, target, term_type);
    if (rc) return rc;
    if (!T_inout || level < 0 || level >= o->desc.levels) { set_error("bad argument"); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(o->device));
    if ((rc = opb_frame_preprocess(o, source))) return rc;
    if ((rc = opb_frame_preprocess(o, target))) return rc;
    if ((rc = reset_state(o, T_inout))) return rc;
    OdoArgs a = make_args(o, source, target, level, term_type);
    launch_iteration(o, a, false);
    launch_compaction(o, level);
    OPB_CUDA(cudaGetLastError());
    OPB_CUDA(cudaMemcpyAsync(o->h_state, o->d_state, sizeof(OdoState), cudaMemcpyDeviceToHost, o->stream));
    OPB_CUDA(cudaStreamSynchronize(o->stream));
    const OdoState *h = o->h_state;
    memcpy(T_inout, h->T, 16 * sizeof(float));
    if (sums43)
    {
        int k = 0;
        for (int p = 0; p < 6; ++p)
            for (int q = p; q < 6; ++q) { sums43[p * 6 + q] = h->packet[k]; sums43[q * 6 + p] = h->packet[k]; ++k; }
        for (int p = 0; p < 6; ++p) sums43[36 + p] = h->packet[21 + p];
        sums43[42] = h->packet[27];
    }
    if (n_pairs) *n_pairs = (size_t)h->n_pairs;
    const size_t np = (size_t)h->n_pairs;
    if (pixel_pairs && pairs_cap)
        OPB_CUDA(cudaMemcpy(pixel_pairs, o->d_pairs, (np < pairs_cap ? np : pairs_cap) * sizeof(uint4), cudaMemcpyDeviceToHost));
    return OPB_OK;
}
} // extern "C"
