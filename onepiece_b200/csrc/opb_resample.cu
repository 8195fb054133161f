// Volume resampling and merging on the device.
//
// Reference path rebuilt here (file:line relative to the reference tree):
//   CubeHandler::Transform            src/Integration/CubeHandler.h:242-298  (AddTransformedCube :199-225)
//   CubeHandler::TransformNearest     src/Integration/CubeHandler.h:299-338  (AddTransformedCubeNearest :226-241)
//   ReadVoxelInterpolate              src/Integration/VoxelCube.cpp:6-50
//   TSDFVoxel operator+ * / add       src/Integration/TSDFVoxel.h:24-74
//   CubeHandler::Merge                src/Integration/CubeHandler.h:145-177
//   CubePara::GetGlobalPoint/GetCubeID/GetVoxelID   src/Integration/VoxelCube.h:63-86
// Callers: example/ImageSequenceIntegration.cpp:48 (TransformNearest), example/MergeMultipleSubmaps.cpp:40-41.
//
// Both transforms are two gathers.  Pass 1 walks every voxel of the source, pushes its centre through `trans` and
// registers the cube(s) it lands in with the result volume (hash insert; a warp first agrees on its distinct target cubes so
// one lane per cube does the atomic).  Pass 2 walks every voxel of the result (one CTA per cube), pulls its centre back
// through trans^-1 and reads the source: one voxel (nearest) or the 8 surrounding ones blended by ReadVoxelInterpolate,
// in the reference's float operation order -- the result is bit-identical to the reference's.
//
// Reference quirk kept on purpose: TransformNearest does not copy c_para into its result, so the allocation pass (which runs
// on the result handler) uses CubePara's default VoxelResolution 0.01 whatever the source resolution is, while the sampling
// pass (a member of the source handler) uses the source resolution.  The C-ABI takes the result resolution as an argument;
// the drop-in class passes what the reference would use.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../include/onepiece_b200.h"
#include "opb_host_math.h"
#include "opb_volume.cuh"
#include "opb_volume_host.h"

namespace opb
{
struct Voxel
{
    float sdf, weight, c0, c1, c2;
};
__device__ __forceinline__ Voxel default_voxel() { return Voxel{999.0f, 0.0f, -1.0f, -1.0f, -1.0f}; }

// TSDFVoxel::operator+ (TSDFVoxel.h:24-39)
__device__ __forceinline__ Voxel vox_plus(const Voxel &a, const Voxel &b)
{
    if (a.weight == 0) return b;
    if (b.weight == 0) return a;
    Voxel r = default_voxel();
    r.weight = fadd(a.weight, b.weight);
    if (r.weight != 0)
    {
        r.sdf = fdiv(fadd(fmul(a.weight, a.sdf), fmul(b.weight, b.sdf)), r.weight);
        r.c0 = fdiv(fadd(fmul(a.weight, a.c0), fmul(b.weight, b.c0)), r.weight);
        r.c1 = fdiv(fadd(fmul(a.weight, a.c1), fmul(b.weight, b.c1)), r.weight);
        r.c2 = fdiv(fadd(fmul(a.weight, a.c2), fmul(b.weight, b.c2)), r.weight);
    }
    return r;
}
// TSDFVoxel::add (TSDFVoxel.h:40-53)
__device__ __forceinline__ Voxel vox_add(const Voxel &a, const Voxel &b)
{
    if (a.weight == 0) return b;
    if (b.weight == 0) return a;
    return Voxel{fadd(a.sdf, b.sdf), fadd(a.weight, b.weight), fadd(a.c0, b.c0), fadd(a.c1, b.c1), fadd(a.c2, b.c2)};
}
// TSDFVoxel::operator* (TSDFVoxel.h:58-70)
__device__ __forceinline__ Voxel vox_mul(const Voxel &a, float w)
{
    if (w == 0 || a.weight == 0) return default_voxel();
    return Voxel{fmul(a.sdf, w), fmul(a.weight, w), fmul(a.c0, w), fmul(a.c1, w), fmul(a.c2, w)};
}
// one level of ReadVoxelInterpolate (VoxelCube.cpp:16-47); operator/ multiplies by the reciprocal (TSDFVoxel.h:71-74)
__device__ __forceinline__ Voxel vox_lerp(const Voxel &a, const Voxel &b, float t)
{
    if (!(a.weight != 0 || b.weight != 0)) return default_voxel();
    const float s = fsub(1.0f, t);
    const float den = fadd(fmul(s, a.weight != 0 ? 1.0f : 0.0f), fmul(t, b.weight != 0 ? 1.0f : 0.0f));
    return vox_mul(vox_add(vox_mul(a, s), vox_mul(b, t)), fdiv(1.0f, den));
}

struct ResampleParams
{
    float fwd[16];   // trans, column-major
    float inv[16];   // Eigen-order inverse of trans
    float alloc_res; // VoxelResolution of the result handler (allocation pass)
    float res;       // VoxelResolution of the source handler (sampling pass)
    int nearest;
};

// CubePara::GetGlobalPoint (VoxelCube.h:75-80) for voxel n of cube id at resolution res
__device__ __forceinline__ void global_point(const int *id, int n, float res, float &x, float &y, float &z)
{
    const float half = fdiv(res, 2.0f);
    x = fadd(fmul(fmul((float)id[0], (float)kCube), res), fadd(fmul((float)(n & 7), res), half));
    y = fadd(fmul(fmul((float)id[1], (float)kCube), res), fadd(fmul((float)((n >> 3) & 7), res), half));
    z = fadd(fmul(fmul((float)id[2], (float)kCube), res), fadd(fmul((float)(n >> 6), res), half));
}
// trans * Vector4(x, y, z, 1), head<3>() / w (CubeHandler.h:205-208)
__device__ __forceinline__ void transform_h(const float *m, float x, float y, float z, float &ox, float &oy, float &oz)
{
    const float w = row_xyz1(m[3], m[7], m[11], m[15], x, y, z);
    ox = fdiv(row_xyz1(m[0], m[4], m[8], m[12], x, y, z), w);
    oy = fdiv(row_xyz1(m[1], m[5], m[9], m[13], x, y, z), w);
    oz = fdiv(row_xyz1(m[2], m[6], m[10], m[14], x, y, z), w);
}
__device__ __forceinline__ int voxel_coord(float p, float res) { return cvtt_x86(floorf(fdiv(p, res))); }

// pass 1: thread = source voxel
__global__ void __launch_bounds__(256) resample_alloc_kernel(VolumeDev src, int n_src, VolumeDev dst, const __grid_constant__ ResampleParams p)
{
    const long long total = (long long)n_src * kCubeVoxels;
    for (long long base = (long long)blockIdx.x * blockDim.x; base < total; base += (long long)gridDim.x * blockDim.x)
    {
        const long long t = base + threadIdx.x; // a warp stays inside one source cube (512 = 16 warps)
        const bool live = t < total;
        int n0[3] = {0, 0, 0};
        if (live)
        {
            const int slot = (int)(t >> 9), n = (int)(t & 511);
            float gx, gy, gz, px, py, pz;
            global_point(&src.slot_ids[3 * slot], n, p.alloc_res, gx, gy, gz);
            transform_h(p.fwd, gx, gy, gz, px, py, pz);
            if (!p.nearest)
            {
                const float half = fdiv(p.alloc_res, 2.0f);
                px = fsub(px, half); py = fsub(py, half); pz = fsub(pz, half);
            }
            n0[0] = voxel_coord(px, p.alloc_res); n0[1] = voxel_coord(py, p.alloc_res); n0[2] = voxel_coord(pz, p.alloc_res);
        }
        const int corners = p.nearest ? 1 : 8;
        for (int i = 0; i < corners; ++i)
        {
            const int ci = (n0[0] + (i & 1)) >> 3, cj = (n0[1] + ((i >> 1) & 1)) >> 3, ck = (n0[2] + (i >> 2)) >> 3; // floor(v / 8)
            unsigned long long key = kEmptyKey;
            const bool ok = live && pack_id(ci, cj, ck, key);
            if (live && !ok) raise_overflow(dst, kOverflowRange);
            // one lane per distinct cube of the warp performs the insert
            const unsigned int peers = __match_any_sync(0xffffffffu, ok ? key : kEmptyKey);
            if (ok && (__ffs(peers) - 1) == (int)(threadIdx.x & 31)) table_find_or_insert(dst, ci, cj, ck);
        }
    }
}

// The order in which the reference would have CREATED the cubes of the result (AddTransformedCube / AddTransformedCubeNearest,
// CubeHandler.h:198-241: source cubes in the iteration order of the source's map, voxels 0..511, the eight neighbours 0..7, a cube is
// inserted at its first touch).  rank[slot] = position of a source cube in that iteration; first_touch[result slot] = smallest
// ((rank * 512 + voxel) * 8 + neighbour) that maps into the cube.
__global__ void slot_ranks_kernel(VolumeDev src, const int *ids_in_order, int n, unsigned int *rank, int *missing)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const int s = table_find(src, ids_in_order[3 * i], ids_in_order[3 * i + 1], ids_in_order[3 * i + 2]);
        if (s < 0 || s >= n) atomicAdd(missing, 1);
        else rank[s] = (unsigned int)i;
    }
}
__global__ void __launch_bounds__(256) resample_first_touch_kernel(VolumeDev src, int n_src, const unsigned int *rank, VolumeDev dst,
                                                                   const __grid_constant__ ResampleParams p, unsigned long long *first_touch)
{
    const long long total = (long long)n_src * kCubeVoxels;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x)
    {
        const int slot = (int)(t >> 9), n = (int)(t & 511);
        float gx, gy, gz, px, py, pz;
        global_point(&src.slot_ids[3 * slot], n, p.alloc_res, gx, gy, gz);
        transform_h(p.fwd, gx, gy, gz, px, py, pz);
        if (!p.nearest)
        {
            const float half = fdiv(p.alloc_res, 2.0f);
            px = fsub(px, half); py = fsub(py, half); pz = fsub(pz, half);
        }
        const int n0[3] = {voxel_coord(px, p.alloc_res), voxel_coord(py, p.alloc_res), voxel_coord(pz, p.alloc_res)};
        const unsigned long long base_key = ((unsigned long long)rank[slot] * kCubeVoxels + (unsigned long long)n) * 8ull;
        const int corners = p.nearest ? 1 : 8;
        int last = -1;
        for (int i = 0; i < corners; ++i)
        {
            const int ci = (n0[0] + (i & 1)) >> 3, cj = (n0[1] + ((i >> 1) & 1)) >> 3, ck = (n0[2] + (i >> 2)) >> 3;
            const int d = table_find(dst, ci, cj, ck);
            if (d < 0 || d == last) continue; // (the same cube as the neighbour before: its key is already smaller)
            last = d;
            if (base_key + i < first_touch[d]) atomicMin(&first_touch[d], base_key + (unsigned long long)i);
        }
    }
}

__device__ __forceinline__ Voxel read_voxel(const VolumeDev &v, int x, int y, int z)
{
    const int ci = x >> 3, cj = y >> 3, ck = z >> 3;
    const int slot = table_find(v, ci, cj, ck);
    if (slot < 0) return default_voxel();
    const int vid = (x - (ci << 3)) + ((y - (cj << 3)) << 3) + ((z - (ck << 3)) << 6);
    const float *b = v.pool + (size_t)slot * kSlotFloats + vid;
    return Voxel{b[0], b[kCubeVoxels], b[2 * kCubeVoxels], b[3 * kCubeVoxels], b[4 * kCubeVoxels]};
}

// pass 2: CTA = result cube, thread = voxel
__global__ void __launch_bounds__(kCubeVoxels) resample_fill_kernel(VolumeDev src, VolumeDev dst, int n_dst, const __grid_constant__ ResampleParams p)
{
    const int slot = blockIdx.x;
    if (slot >= n_dst) return;
    const int n = threadIdx.x;
    float gx, gy, gz, px, py, pz;
    global_point(&dst.slot_ids[3 * slot], n, p.res, gx, gy, gz);
    transform_h(p.inv, gx, gy, gz, px, py, pz);
    Voxel got;
    if (p.nearest)
        got = read_voxel(src, voxel_coord(px, p.res), voxel_coord(py, p.res), voxel_coord(pz, p.res));
    else
    {
        const float half = fdiv(p.res, 2.0f);
        px = fsub(px, half); py = fsub(py, half); pz = fsub(pz, half);
        const int x = voxel_coord(px, p.res), y = voxel_coord(py, p.res), z = voxel_coord(pz, p.res);
        Voxel v8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v8[i] = read_voxel(src, x + (i & 1), y + ((i >> 1) & 1), z + (i >> 2));
        // ReadVoxelInterpolate (VoxelCube.cpp:11-13): weight = (position - neighbour0 * res) / res
        const float xw = fdiv(fsub(px, fmul((float)x, p.res)), p.res), yw = fdiv(fsub(py, fmul((float)y, p.res)), p.res),
                    zw = fdiv(fsub(pz, fmul((float)z, p.res)), p.res);
        const Voxel z1 = vox_lerp(vox_lerp(v8[0], v8[1], xw), vox_lerp(v8[2], v8[3], xw), yw);
        const Voxel z2 = vox_lerp(vox_lerp(v8[4], v8[5], xw), vox_lerp(v8[6], v8[7], xw), yw);
        got = vox_lerp(z1, z2, zw);
    }
    // v_cube.voxels[voxel_id] += result on a freshly constructed voxel (weight 0): operator+ returns the right operand
    float *o = dst.pool + (size_t)slot * kSlotFloats + n;
    o[0] = got.sdf; o[kCubeVoxels] = got.weight; o[2 * kCubeVoxels] = got.c0; o[3 * kCubeVoxels] = got.c1; o[4 * kCubeVoxels] = got.c2;
}

// Merge: CTA = cube of `other`
__global__ void __launch_bounds__(kCubeVoxels) merge_count_kernel(VolumeDev dst, VolumeDev other, int n_other, int *n_new)
{
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_other; s += gridDim.x * blockDim.x)
        if (table_find(dst, other.slot_ids[3 * s], other.slot_ids[3 * s + 1], other.slot_ids[3 * s + 2]) < 0) atomicAdd(n_new, 1);
}
__global__ void __launch_bounds__(kCubeVoxels) merge_kernel(VolumeDev dst, VolumeDev other, int n_other)
{
    const int s = blockIdx.x;
    if (s >= n_other) return;
    __shared__ int s_slot;
    if (threadIdx.x == 0) s_slot = table_find_or_insert(dst, other.slot_ids[3 * s], other.slot_ids[3 * s + 1], other.slot_ids[3 * s + 2]);
    __syncthreads();
    if (s_slot < 0) return;
    const int n = threadIdx.x;
    const float *b = other.pool + (size_t)s * kSlotFloats + n;
    float *a = dst.pool + (size_t)s_slot * kSlotFloats + n;
    const Voxel vb{b[0], b[kCubeVoxels], b[2 * kCubeVoxels], b[3 * kCubeVoxels], b[4 * kCubeVoxels]};
    const Voxel va{a[0], a[kCubeVoxels], a[2 * kCubeVoxels], a[3 * kCubeVoxels], a[4 * kCubeVoxels]};
    // a cube missing from dst is copied (CubeHandler.h:155-159): its slot holds default voxels, whose weight 0 makes
    // operator+ return the other voxel -- the same thing
    const Voxel r = vox_plus(va, vb);
    a[0] = r.sdf; a[kCubeVoxels] = r.weight; a[2 * kCubeVoxels] = r.c0; a[3 * kCubeVoxels] = r.c1; a[4 * kCubeVoxels] = r.c2;
}
} // namespace opb

using namespace opb;

extern "C"
{
int opb_volume_get_desc(opb_volume *v, opb_volume_desc *out)
{
    if (!v || !out) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    *out = v->desc;
    return OPB_OK;
}

int opb_volume_transform(opb_volume *src, const float trans_cm[16], int nearest, float result_voxel_resolution, int32_t result_max_cubes,
                         opb_volume **out)
{
    if (!src || !trans_cm || !out) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    *out = nullptr;
    OPB_CUDA(cudaSetDevice(src->desc.device));
    size_t n_src = 0;
    int rc = volume_require_float(src, "opb_volume_transform");
    if (rc == OPB_OK) rc = opb_volume_num_cubes(src, &n_src); // synchronizes
    if (rc) return rc;
    ResampleParams p;
    memcpy(p.fwd, trans_cm, sizeof(p.fwd));
    hostmath::mat4_inverse_colmajor(trans_cm, p.inv); // trans.inverse() (CubeHandler.h:266,318)
    p.res = src->desc.voxel_resolution;
    p.alloc_res = result_voxel_resolution > 0 ? result_voxel_resolution : p.res;
    p.nearest = nearest ? 1 : 0;
    opb_volume_desc d = src->desc;
    d.voxel_resolution = p.alloc_res;
    d.shard_world = 1; d.shard_rank = 0; // the result is a plain volume; resample every shard and merge to keep a partition
    d.stream = nullptr;
    long long cap = result_max_cubes > 0 ? result_max_cubes : (long long)(2 * n_src + 1024);
    for (int attempt = 0;; ++attempt)
    {
        if (cap > 0x7FFFFFFF) cap = 0x7FFFFFFF;
        d.max_cubes = (int32_t)cap;
        opb_volume *r = nullptr;
        rc = opb_volume_create(&d, &r);
        if (rc) return rc;
        cudaStream_t s = r->stream;
        if (n_src)
        {
            const long long total = (long long)n_src * kCubeVoxels;
            const int blocks = (int)((total + 255) / 256 < (long long)r->sm_count * 16 ? (total + 255) / 256 : (long long)r->sm_count * 16);
            resample_alloc_kernel<<<blocks, 256, 0, s>>>(src->dev, (int)n_src, r->dev, p);
        }
        FrameCounters fc;
        int n_dst = 0;
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(&fc, r->dev.fc, sizeof(fc), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&n_dst, r->dev.n_alloc, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { opb_volume_destroy(r); set_error("volume transform failed: %s", cudaGetErrorString(e)); return OPB_ERR_CUDA; }
        if (fc.overflow)
        {
            opb_volume_destroy(r);
            if (result_max_cubes > 0 || attempt >= 4 || cap >= 0x7FFFFFFF)
            {
                set_error("transformed volume needs more than %lld cubes", cap);
                return OPB_ERR_CAPACITY;
            }
            cap *= 2; // a rotated thin shell touches more cubes than the axis-aligned one; retry with room
            continue;
        }
        if (n_dst > 0) resample_fill_kernel<<<n_dst, kCubeVoxels, 0, s>>>(src->dev, r->dev, n_dst, p);
        // interpolated weights are no longer small integers: later integrations into the result use IEEE division
        int tainted = 1;
        if (nearest) e = cudaMemcpyAsync(r->dev.tainted, src->dev.tainted, sizeof(int), cudaMemcpyDeviceToDevice, s);
        else e = cudaMemcpyAsync(r->dev.tainted, &tainted, sizeof(int), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { opb_volume_destroy(r); set_error("volume transform failed: %s", cudaGetErrorString(e)); return OPB_ERR_CUDA; }
        *out = r;
        return OPB_OK;
    }
}

int opb_volume_transform_ordered(opb_volume *src, const float trans_cm[16], int nearest, float result_voxel_resolution, int32_t result_max_cubes,
                                 const int32_t *src_ids_in_order, size_t n_src_ids, opb_volume **out, int32_t **result_ids_in_order,
                                 size_t *n_result)
{
    if (!src || !out || !result_ids_in_order || !n_result || (n_src_ids && !src_ids_in_order)) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    *result_ids_in_order = nullptr;
    *n_result = 0;
    size_t n_src = 0;
    int rc = opb_volume_num_cubes(src, &n_src);
    if (rc) return rc;
    if (n_src_ids != n_src) { set_error("the cube sequence has %zu entries, the source volume %zu cubes", n_src_ids, n_src); return OPB_ERR_INVALID; }
    rc = opb_volume_transform(src, trans_cm, nearest, result_voxel_resolution, result_max_cubes, out);
    if (rc) return rc;
    opb_volume *r = *out;
    size_t n_dst = 0;
    rc = opb_volume_num_cubes(r, &n_dst);
    if (rc == OPB_OK && n_dst == 0) return OPB_OK;
    auto fail = [&](int code) { opb_volume_destroy(r); *out = nullptr; return code; };
    if (rc) return fail(rc);
    ResampleParams p;
    memcpy(p.fwd, trans_cm, sizeof(p.fwd));
    hostmath::mat4_inverse_colmajor(trans_cm, p.inv);
    p.res = src->desc.voxel_resolution;
    p.alloc_res = r->desc.voxel_resolution;
    p.nearest = nearest ? 1 : 0;
    cudaStream_t s = r->stream;
    int *d_ids = nullptr;
    unsigned int *d_rank = nullptr;
    unsigned long long *d_touch = nullptr;
    int *d_missing = nullptr;
    cudaError_t e = cudaMalloc(&d_ids, n_src * 3 * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&d_rank, n_src * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&d_touch, n_dst * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&d_missing, sizeof(int));
    std::vector<unsigned long long> touch(n_dst);
    std::vector<int> ids(n_dst * 3);
    int missing = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_ids, src_ids_in_order, n_src * 3 * sizeof(int), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_missing, 0, sizeof(int), s);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_touch, 0xFF, n_dst * sizeof(unsigned long long), s);
    if (e == cudaSuccess)
    {
        slot_ranks_kernel<<<r->sm_count * 4, 256, 0, s>>>(src->dev, d_ids, (int)n_src, d_rank, d_missing);
        const long long total = (long long)n_src * kCubeVoxels;
        const int blocks = (int)((total + 255) / 256 < (long long)r->sm_count * 16 ? (total + 255) / 256 : (long long)r->sm_count * 16);
        resample_first_touch_kernel<<<blocks, 256, 0, s>>>(src->dev, (int)n_src, d_rank, r->dev, p, d_touch);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&missing, d_missing, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(touch.data(), d_touch, n_dst * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ids.data(), r->dev.slot_ids, n_dst * 3 * sizeof(int), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d_ids); cudaFree(d_rank); cudaFree(d_touch); cudaFree(d_missing);
    if (e != cudaSuccess) { set_error("ordered transform failed: %s", cudaGetErrorString(e)); return fail(OPB_ERR_CUDA); }
    if (missing) { set_error("%d cubes of the sequence are not in the source volume", missing); return fail(OPB_ERR_INVALID); }
    std::vector<size_t> perm(n_dst);
    for (size_t i = 0; i < n_dst; ++i) perm[i] = i;
    std::sort(perm.begin(), perm.end(), [&](size_t a, size_t b) { return touch[a] < touch[b]; });
    int32_t *o = (int32_t *)malloc(n_dst * 3 * sizeof(int32_t));
    if (!o) { set_error("host allocation failed"); return fail(OPB_ERR_CAPACITY); }
    for (size_t i = 0; i < n_dst; ++i)
        for (int a = 0; a < 3; ++a) o[3 * i + a] = ids[3 * perm[i] + a];
    *result_ids_in_order = o;
    *n_result = n_dst;
    return OPB_OK;
}

int opb_volume_merge(opb_volume *dst, opb_volume *other)
{
    if (!dst || !other) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    if (dst == other) { set_error("cannot merge a volume into itself"); return OPB_ERR_INVALID; }
    if (dst->desc.device != other->desc.device) { set_error("volumes live on different devices"); return OPB_ERR_INVALID; }
    if (dst->desc.voxel_resolution != other->desc.voxel_resolution)
    {
        // the reference prints this warning and returns without merging (CubeHandler.h:147-151)
        set_error("[Warning]::[MergeVoxelHash]::Voxel resolution is not identical.");
        return OPB_ERR_INVALID;
    }
    OPB_CUDA(cudaSetDevice(dst->desc.device));
    size_t n_dst = 0, n_other = 0;
    int rc = volume_require_float(dst, "opb_volume_merge");
    if (rc == OPB_OK) rc = volume_require_float(other, "opb_volume_merge");
    if (rc) return rc;
    rc = opb_volume_num_cubes(dst, &n_dst);
    if (rc == OPB_OK) rc = opb_volume_num_cubes(other, &n_other);
    if (rc || n_other == 0) return rc;
    if (dst->n_ghost) { rc = halo_drop_ghosts(dst); if (rc) return rc; }
    cudaStream_t s = dst->stream;
    int *d_new = nullptr, n_new = 0;
    OPB_CUDA(cudaMalloc(&d_new, sizeof(int)));
    cudaError_t e = cudaMemsetAsync(d_new, 0, sizeof(int), s);
    merge_count_kernel<<<dst->sm_count, kCubeVoxels, 0, s>>>(dst->dev, other->dev, (int)n_other, d_new);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&n_new, d_new, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d_new);
    if (e != cudaSuccess) { set_error("volume merge failed: %s", cudaGetErrorString(e)); return OPB_ERR_CUDA; }
    if (n_dst + (size_t)n_new > (size_t)dst->dev.max_cubes)
    {   // the reference's map is unbounded (CubeHandler.h:145-177): make room
        rc = volume_grow(dst, (long long)(n_dst + (size_t)n_new));
        if (rc) return rc;
    }
    merge_kernel<<<(unsigned int)n_other, kCubeVoxels, 0, s>>>(dst->dev, other->dev, (int)n_other);
    OPB_CUDA(cudaGetLastError());
    int t_other = 0;
    OPB_CUDA(cudaMemcpyAsync(&t_other, other->dev.tainted, sizeof(int), cudaMemcpyDeviceToHost, s));
    OPB_CUDA(cudaStreamSynchronize(s));
    if (t_other)
    {
        OPB_CUDA(cudaMemcpyAsync(dst->dev.tainted, &t_other, sizeof(int), cudaMemcpyHostToDevice, s));
        OPB_CUDA(cudaStreamSynchronize(s));
    }
    return OPB_OK;
}
} // extern "C"
