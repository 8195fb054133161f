// TEST INFRASTRUCTURE ONLY.  A minimal CUDA execution model on host threads, just enough to run the device code of
// onepiece_b200/csrc/opb_kdtree.cu (block-wide barriers, warp shuffles, shared memory, atomics) in the CPU container:
// one std::thread per CUDA thread of a block, blocks one after the other, __syncthreads = a pthread barrier, shuffles = a
// per-warp exchange buffer between two warp barriers.  The arithmetic intrinsics map to plain IEEE float operations (the
// harness is compiled with -ffp-contract=off), so results are bit-comparable with the oracle.
#pragma once
#include <pthread.h>

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)

struct Dim3 { int x = 1, y = 1, z = 1; };
inline thread_local Dim3 threadIdx, blockIdx;
inline Dim3 blockDim, gridDim;

namespace emu
{
inline pthread_barrier_t block_bar;
inline pthread_barrier_t warp_bar[32];
inline uint32_t warp_slot[32][32];
inline std::mutex atomic_mutex;

template <class F>
void launch(int grid, int block, F f)
{
    gridDim.x = grid;
    blockDim.x = block;
    pthread_barrier_init(&block_bar, nullptr, (unsigned)block);
    const int warps = (block + 31) / 32;
    for (int w = 0; w < warps; ++w) pthread_barrier_init(&warp_bar[w], nullptr, (unsigned)std::min(32, block - 32 * w));
    std::vector<std::thread> ts;
    for (int t = 0; t < block; ++t)
        ts.emplace_back([&, t] {
            threadIdx.x = t;
            for (int b = 0; b < grid; ++b)
            {
                blockIdx.x = b;
                f();
                pthread_barrier_wait(&block_bar);
            }
        });
    for (auto &t : ts) t.join();
    pthread_barrier_destroy(&block_bar);
    for (int w = 0; w < warps; ++w) pthread_barrier_destroy(&warp_bar[w]);
}
inline uint64_t warp_slot64[32][32];
template <class T>
T exchange(T v, int src_lane)
{
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "32- and 64-bit shuffles only");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    warp_slot64[warp][lane] = bits;
    pthread_barrier_wait(&warp_bar[warp]);
    uint64_t r = src_lane >= 0 && src_lane < 32 ? warp_slot64[warp][src_lane] : bits;
    pthread_barrier_wait(&warp_bar[warp]);
    T out;
    memcpy(&out, &r, sizeof(T));
    return out;
}
} // namespace emu

inline void __syncthreads() { pthread_barrier_wait(&emu::block_bar); }
template <class T> T __shfl_xor_sync(unsigned, T v, int o) { return emu::exchange(v, (int)(threadIdx.x & 31) ^ o); }
template <class T> T __shfl_up_sync(unsigned, T v, int o) { const int lane = threadIdx.x & 31; return emu::exchange(v, lane >= o ? lane - o : lane); }
inline int atomicAdd(int *p, int v) { std::lock_guard<std::mutex> g(emu::atomic_mutex); const int old = *p; *p = old + v; return old; }
inline int atomicMax(int *p, int v) { std::lock_guard<std::mutex> g(emu::atomic_mutex); const int old = *p; if (v > old) *p = v; return old; }
inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) { std::lock_guard<std::mutex> g(emu::atomic_mutex); const unsigned long long old = *p; if (v > old) *p = v; return old; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fsqrt_rn(float a) { return sqrtf(a); }
struct float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
using std::isfinite;
using std::max;
using std::min;
