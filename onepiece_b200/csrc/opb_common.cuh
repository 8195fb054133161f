// Shared helpers of the sm_100a kernels: error plumbing and the "reference arithmetic" primitives.
//
// The reference is compiled with -O3 -msse4.2 (CMakeLists.txt:149-150): scalar SSE float math, no FMA
// contraction, IEEE division.  Discrete decisions on the path (pixel selection by truncation, |sdf| < trunc,
// sdf > 0 in Marching Cubes) flip on 1-ulp differences, so every float operation that feeds one is written
// with the non-contracting intrinsics below, in the reference's operation order (SURVEY.md §7 hard part 1).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

namespace opb
{
void set_error(const char *fmt, ...);

#define OPB_CUDA(expr)                                                                          \
    do                                                                                          \
    {                                                                                           \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
        {                                                                                       \
            opb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return OPB_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

// ---- float ops that ptxas must not fuse ----
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// Correctly rounded a/b for several numerators sharing one divisor: y = RN(1/b) once (rcp_rn), then per
// numerator q = RN(a*y), r = a - b*q (exact in one FMA), RN(q + r*y).  With y the correctly rounded reciprocal
// this last step returns RN(a/b) (Markstein's quotient-refinement theorem); tests/test_division.py checks it
// against IEEE division on 4.6e8 operand pairs, including every integer divisor up to 70,000 (the weights W of
// TSDFVoxel::operator+) and all-ones significands.  Operands outside the guarded range take __fdiv_rn.
__device__ __forceinline__ bool rcp_safe(float b)
{
    const float ab = fabsf(b);
    return ab > 1e-18f && ab < 1e18f; // also false for NaN, 0, inf
}
__device__ __forceinline__ float div_by(float a, float b, float y, bool safe)
{
    if (!safe) return __fdiv_rn(a, b);
    const float q = __fmul_rn(a, y);
    const float r = __fmaf_rn(-b, q, a);
    return __fmaf_rn(r, y, q);
}

// x86 cvttsd2si semantics: C truncation toward zero, INT_MIN ("integer indefinite") for NaN and for values
// that do not fit -- CUDA's own conversion saturates and maps NaN to 0, which would select pixel 0.
__device__ __forceinline__ int cvtt_x86(double d)
{
    if (!(d > -2147483649.0 && d < 2147483648.0)) return INT32_MIN;
    return __double2int_rz(d);
}
// float -> int as g++ emits it (cvttss2si)
__device__ __forceinline__ int cvtt_x86(float f)
{
    if (!(f >= -2147483648.0f && f < 2147483648.0f)) return INT32_MIN;
    return __float2int_rz(f);
}

// int u = a + 0.5 + c with a, c float:  the 0.5 literal is double, so both adds are done in double
// (Integrator.cpp:19-20,61-62)
__device__ __forceinline__ int pixel_index(float a, float c)
{
    return cvtt_x86(__dadd_rn(__dadd_rn((double)a, 0.5), (double)c));
}

// Eigen 3.3.7 Matrix4f * Vector4f(x, y, z, 1) row, SSE gemv order: ((m0*x + m1*y) + m2*z) + m3*1
__device__ __forceinline__ float row_xyz1(float m0, float m1, float m2, float m3, float x, float y, float z)
{
    return fadd(fadd(fadd(fmul(m0, x), fmul(m1, y)), fmul(m2, z)), m3);
}
// Eigen 3.3.7 fixed-size-3 redux order: a0*b0 + (a1*b1 + a2*b2)
__device__ __forceinline__ float dot3(float a0, float a1, float a2, float b0, float b1, float b2)
{
    return fadd(fmul(a0, b0), fadd(fmul(a1, b1), fmul(a2, b2)));
}

// monotone float <-> uint mapping for atomicMin/atomicMax on floats
__device__ __forceinline__ unsigned int float_to_ordered(float f)
{
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(unsigned int u)
{
    unsigned int b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// floor division / modulo of ints by a positive int
__host__ __device__ __forceinline__ int floor_div(int a, int b)
{
    int q = a / b;
    return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}
__host__ __device__ __forceinline__ int floor_mod(int a, int b) { return a - floor_div(a, b) * b; }

// ---------------------------------------------------------------------------------------------------------
// FP64 tensor-core op as a reduction primitive (the persistent solver loops of opb_icp.cu and opb_odometry.cu): a warp folds the
// outer products c c^T of its 32 eight-component float vectors into an 8x8 double matrix held as two doubles per lane (lane L:
// entries 2L, 2L+1 of the row-major matrix).  Products of floats are exact in double; accumulation is double.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// the warp's 32 vectors -> shared memory -> eight 8x8x4 outer-product accumulations (A = B^T = 8 components x 4 points).  The
// eight products go to four independent accumulators (a dependent DMMA chain costs ~140 cycles per link, independent ones issue
// every 16: scripts/micro/dmma_rate.cu) and are folded with plain additions.
__device__ __forceinline__ void warp_fold_outer8(float *stage, int lane, const float *comp, double &c0, double &c1)
{
    *reinterpret_cast<float4 *>(stage + lane * 8) = make_float4(comp[0], comp[1], comp[2], comp[3]);
    *reinterpret_cast<float4 *>(stage + lane * 8 + 4) = make_float4(comp[4], comp[5], comp[6], comp[7]);
    __syncwarp();
    double t0[4] = {0.0, 0.0, 0.0, 0.0}, t1[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < 8; ++j)
    {
        const double v = (double)stage[(4 * j + (lane & 3)) * 8 + (lane >> 2)];
        dmma_8x8x4(t0[j & 3], t1[j & 3], v, v);
    }
    __syncwarp();
    c0 += (t0[0] + t0[1]) + (t0[2] + t0[3]);
    c1 += (t1[0] + t1[1]) + (t1[2] + t1[3]);
}

} // namespace opb
