// Device-side layout of the TSDF volume (replaces the reference's
// std::unordered_map<CubeID, VoxelCube> of 20-byte AoS voxels, CubeHandler.h:22, VoxelCube.h:95-197).
//
//   block pool   max_cubes slots; a slot holds one 8^3 cube as five float planes of 512 values
//                [sdf | weight | c0 | c1 | c2]  (10,240 B, the same 20 B/voxel as the reference, but
//                plane-major so that a warp touching 32 consecutive voxels moves whole 128 B lines)
//   slot_ids     3 x int32 per slot: the CubeID stored in it
//   hash table   open addressing, 64-bit packed CubeID -> slot, capacity = pow2 >= 2*max_cubes
//   frame list   slots selected for the frame being integrated (CubeHandler::PrepareCubes' cube_id_list)
//   counters     allocation bump pointer, frame counters, bounding box (ordered-uint encoding)
#pragma once
#include "opb_common.cuh"

namespace opb
{
constexpr int kCube = 8;                 // CUBE_SIZE (VoxelCube.h:4)
constexpr int kCubeVoxels = 512;         // 8^3
constexpr int kPlanes = 5;               // sdf, weight, c0, c1, c2
constexpr int kSlotFloats = kCubeVoxels * kPlanes;
constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFull;
constexpr int kIdBias = 1 << 20;         // cube ids are packed 21 bits per axis

struct FrameCounters
{
    unsigned int bbox_min[3]; // ordered-uint floats
    unsigned int bbox_max[3];
    int candidate_cubes;
    int frame_cubes;
    int overflow;   // kOverflowPool: pool or table full, cubes were skipped; kOverflowRange: cube ids beyond 21 bits per axis
    int wild_frame; // a depth value outside [1e-6, 1e6] m was seen: the update kernel takes its IEEE-division path
    unsigned long long updated_voxels;
};

struct VolumeDev
{
    float *pool;              // max_cubes * kSlotFloats (OPB_STORAGE_PACKED16: the float mirror, materialised on demand, else NULL)
    uint2 *pool16;            // OPB_STORAGE_PACKED16: max_cubes * 512 voxels of 8 bytes {half sdf | half weight << 16, b | g<<8 | r<<16}
    int *slot_ids;            // max_cubes * 3
    unsigned long long *keys; // table_cap
    int *vals;                // table_cap
    int4 *frame_list;         // max_cubes entries {slot, i, j, k}: one 16-byte load gives the update kernel all it needs
    int *n_alloc;             // cubes allocated so far
    int *tainted;             // != 0 after an upload of values outside the range the fast quotient is exact for
    float2 *texels;           // per-frame W*H texels {depth in metres (f32), b | g<<8 | r<<16 as raw bits}
    FrameCounters *fc;        // counters of the frame in flight
    volatile int *host_flags; // mapped pinned host memory, sticky until the host clears it: [0] pool full, [1] ids out of range
    int max_cubes;
    unsigned int table_mask;
};

// per-frame constants, passed to the kernels by value
struct FrameParams
{
    float fx, fy, cx, cy;
    int width, height;
    float depth_scale;
    int depth_u16;
    float res;       // VoxelResolution
    float cube_res;  // VoxelResolution * CUBE_SIZE
    float half_res;  // VoxelResolution / 2
    float trunc;
    float pose[16];  // camera-to-world, column-major
    float pinv[16];  // Eigen-order inverse of pose, column-major
    float planes[24];
    double cx_d, cy_d;       // (double)cx, (double)cy: second operand of the reference's double adds
    double width_d, height_d;
    int shard_rank, shard_world, shard_axis, shard_slab;
    int exact_division; // host decision: pose / intrinsics outside the tame range -> IEEE-division path
    int min_new_slot;   // > 0: list only cubes allocated by this very pass (slot >= min_new_slot) -- the re-run after the pool grew
};
constexpr int kOverflowPool = 1, kOverflowRange = 2;
__device__ __forceinline__ void raise_overflow(const VolumeDev &v, int bit)
{
    atomicOr(&v.fc->overflow, bit);
    if (v.host_flags) v.host_flags[bit == kOverflowPool ? 0 : 1] = 1;
}

__host__ __device__ __forceinline__ bool pack_id(int i, int j, int k, unsigned long long &key)
{
    const unsigned int a = (unsigned int)(i + kIdBias), b = (unsigned int)(j + kIdBias), c = (unsigned int)(k + kIdBias);
    if ((a | b | c) >> 21) return false;
    key = ((unsigned long long)a << 42) | ((unsigned long long)b << 21) | (unsigned long long)c;
    return true;
}
__host__ __device__ __forceinline__ unsigned int hash_key(unsigned long long k)
{
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (unsigned int)k;
}
// slot of a cube or -1
__device__ __forceinline__ int table_find(const VolumeDev &v, int i, int j, int k)
{
    unsigned long long key;
    if (!pack_id(i, j, k, key)) return -1;
    unsigned int h = hash_key(key) & v.table_mask;
    for (;;)
    {
        const unsigned long long cur = v.keys[h];
        if (cur == key) return v.vals[h];
        if (cur == kEmptyKey) return -1;
        h = (h + 1) & v.table_mask;
    }
}

// slot of a cube, allocating it from the pool (bump pointer) on first sight; -1 and fc->overflow on a full pool / table
__device__ __forceinline__ int table_find_or_insert(const VolumeDev &v, int i, int j, int k)
{
    unsigned long long key;
    if (!pack_id(i, j, k, key)) { raise_overflow(v, kOverflowRange); return -1; }
    unsigned int h = hash_key(key) & v.table_mask;
    for (unsigned int probe = 0; probe <= v.table_mask; ++probe)
    {
        const unsigned long long prev = atomicCAS(&v.keys[h], kEmptyKey, key);
        if (prev == kEmptyKey)
        {
            const int slot = atomicAdd(v.n_alloc, 1);
            if (slot >= v.max_cubes)
            {
                raise_overflow(v, kOverflowPool);
                v.vals[h] = -1;
                return -1;
            }
            v.vals[h] = slot;
            v.slot_ids[3 * slot] = i;
            v.slot_ids[3 * slot + 1] = j;
            v.slot_ids[3 * slot + 2] = k;
            return slot;
        }
        if (prev == key) return v.vals[h]; // inserted by an earlier frame (a cube is tested once per frame)
        h = (h + 1) & v.table_mask;
    }
    raise_overflow(v, kOverflowPool);
    return -1;
}

// VoxelCentroidOffSet component (VoxelCube.h:48-61): x*VoxelResolution + half_resolution
__device__ __forceinline__ float centroid_offset(int x, float res, float half_res)
{
    return fadd(fmul((float)x, res), half_res);
}
// cube origin component: CubeID * CUBE_SIZE * VoxelResolution == CubeID * (VoxelResolution*CUBE_SIZE)
// (GetGlobalPoint VoxelCube.h:75-80, GetOrigin :143-147, PrepareCubes CubeHandler.cpp:163,175 -- the factor 8 is
// a power of two, so both association orders round identically)
__device__ __forceinline__ float cube_origin(int id, float cube_res) { return fmul((float)id, cube_res); }

} // namespace opb
