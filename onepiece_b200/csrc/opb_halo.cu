// Boundary-cube ("halo") exchange for Marching Cubes over a volume partitioned across GPUs.
//
// Reference behaviour being preserved (file:line relative to the reference tree): a cell on the x/y/z = 7 face of a
// cube takes its +1 corners from up to seven neighbour cubes (CubeHandler::GenerateMeshByCube,
// src/Integration/CubeHandler.cpp:83-99) and is skipped when one of them is absent.  With sub-volume ownership by
// slabs along `shard_axis` (opb_volume_desc), the +1 neighbours of a cube in the LAST layer of a slab belong to the
// owner of the next slab.  MC reads only one voxel layer of those cubes: the layer with axis coordinate 0.
//
//   export   every owned cube whose axis id is the FIRST of its slab -> (cube id, 64 voxels x 5 planes of that layer);
//            all of them are needed by exactly one peer, the owner of the previous slab: rank-1 (mod world)
//   import   received cubes become GHOST slots at the top of the block pool, registered in the hash table so the
//            Marching-Cubes kernels find them as neighbours; they are never listed for integration, never downloaded
//            and never emit cells (the mesh kernels walk slots [0, n_alloc) only)
//   clear    drops the ghosts (done implicitly by the next integrate / upload / clear)
// The transport between the two calls (NCCL send/recv of the two device buffers) belongs to the host program: see
// onepiece_b200/fusion.py.  The data volume is 1,292 B per boundary cube instead of the 10,240 B of a whole cube.
#include <vector>

#include "../../include/onepiece_b200.h"
#include "opb_volume.cuh"
#include "opb_volume_host.h"

namespace opb
{
constexpr int kLayerVoxels = kCube * kCube;            // 64
constexpr int kLayerFloats = kLayerVoxels * kPlanes;   // 320

// voxel index (x + 8y + 64z, VoxelCube.h:56) of element e = u + 8w of the layer `axis coordinate == 0`
__device__ __forceinline__ int layer_voxel(int axis, int e)
{
    const int u = e & 7, w = e >> 3;
    return axis == 0 ? (u << 3) + (w << 6) : (axis == 1 ? u + (w << 6) : u + (w << 3));
}

__global__ void halo_mark_kernel(VolumeDev v, int n_slots, int axis, int slab, int *list, int *count)
{
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += gridDim.x * blockDim.x)
        if (floor_mod(v.slot_ids[3 * s + axis], slab) == 0) list[atomicAdd(count, 1)] = s;
}

// one CTA of 64 threads per listed cube
__global__ void __launch_bounds__(kLayerVoxels) halo_pack_kernel(VolumeDev v, const int *list, int n, int axis, int *ids, float *layers)
{
    const int c = blockIdx.x;
    if (c >= n) return;
    const int slot = list[c], e = threadIdx.x;
    if (e < 3) ids[3 * c + e] = v.slot_ids[3 * slot + e];
    const float *base = v.pool + (size_t)slot * kSlotFloats + layer_voxel(axis, e);
    float *out = layers + (size_t)c * kLayerFloats + e;
#pragma unroll
    for (int p = 0; p < kPlanes; ++p) out[p * kLayerVoxels] = base[p * kCubeVoxels];
}

__global__ void __launch_bounds__(kLayerVoxels) halo_unpack_kernel(VolumeDev v, int first_slot, int n, int axis, const int *ids,
                                                                   const float *layers, int *rejected)
{
    const int c = blockIdx.x;
    if (c >= n) return;
    const int slot = first_slot + c, e = threadIdx.x;
    __shared__ int s_ok;
    if (e == 0)
    {
        const int i = ids[3 * c], j = ids[3 * c + 1], k = ids[3 * c + 2];
        v.slot_ids[3 * slot] = i; v.slot_ids[3 * slot + 1] = j; v.slot_ids[3 * slot + 2] = k;
        unsigned long long key;
        int ok = pack_id(i, j, k, key) ? 1 : 0;
        if (ok)
        {
            unsigned int h = hash_key(key) & v.table_mask;
            for (;;)
            {
                const unsigned long long prev = atomicCAS(&v.keys[h], kEmptyKey, key);
                if (prev == kEmptyKey) { v.vals[h] = slot; break; }
                if (prev == key) { ok = 0; break; } // the cube already lives here (owned, or sent twice): keep the first
                h = (h + 1) & v.table_mask;
            }
        }
        if (!ok) atomicAdd(rejected, 1);
        s_ok = ok;
    }
    __syncthreads();
    if (!s_ok) return;
    float *base = v.pool + (size_t)slot * kSlotFloats + layer_voxel(axis, e);
    const float *in = layers + (size_t)c * kLayerFloats + e;
#pragma unroll
    for (int p = 0; p < kPlanes; ++p) base[p * kCubeVoxels] = in[p * kLayerVoxels];
}

static int halo_scratch(opb_volume *v, size_t bytes)
{
    if (v->halo_scratch_bytes >= bytes) return OPB_OK;
    cudaFree(v->halo_scratch);
    v->halo_scratch = nullptr;
    v->halo_scratch_bytes = 0;
    OPB_CUDA(cudaMalloc(&v->halo_scratch, bytes));
    v->halo_scratch_bytes = bytes;
    return OPB_OK;
}

int halo_drop_ghosts(opb_volume *v)
{
    if (v->n_ghost == 0) return OPB_OK;
    cudaStream_t s = v->stream;
    int n_alloc = 0;
    OPB_CUDA(cudaMemcpyAsync(&n_alloc, v->dev.n_alloc, sizeof(int), cudaMemcpyDeviceToHost, s));
    OPB_CUDA(cudaStreamSynchronize(s));
    if (n_alloc > v->dev.max_cubes) n_alloc = v->dev.max_cubes;
    int rc = volume_reinit_slots(v, (size_t)(v->dev.max_cubes - v->n_ghost), (size_t)v->n_ghost);
    if (rc == OPB_OK) rc = volume_rebuild_table(v, n_alloc);
    if (rc) return rc;
    v->n_ghost = 0;
    return OPB_OK;
}
} // namespace opb

using namespace opb;

extern "C"
{
int opb_volume_halo_export(opb_volume *v, int32_t *ids, float *layers, size_t cap, size_t *n_out)
{
    if (!v || !n_out) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    *n_out = 0;
    if (v->desc.shard_world <= 1) return OPB_OK; // nothing is owned elsewhere
    OPB_CUDA(cudaSetDevice(v->desc.device));
    size_t n_cubes = 0;
    int rc = opb_volume_num_cubes(v, &n_cubes); // synchronizes
    if (rc || n_cubes == 0) return rc;
    const int axis = v->desc.shard_axis, slab = v->desc.shard_slab_cubes > 0 ? v->desc.shard_slab_cubes : 1;
    if (axis < 0 || axis > 2) { set_error("shard_axis %d out of range", axis); return OPB_ERR_INVALID; }
    // scratch: [count | list(n_cubes) | ids(3 n) | layers(320 n)], sized for the worst case lazily
    rc = halo_scratch(v, 16 + n_cubes * sizeof(int));
    if (rc) return rc;
    cudaStream_t s = v->stream;
    int *d_count = (int *)v->halo_scratch, *d_list = (int *)((char *)v->halo_scratch + 16);
    OPB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), s));
    halo_mark_kernel<<<v->sm_count * 4, 256, 0, s>>>(v->dev, (int)n_cubes, axis, slab, d_list, d_count);
    int n = 0;
    OPB_CUDA(cudaMemcpyAsync(&n, d_count, sizeof(int), cudaMemcpyDeviceToHost, s));
    OPB_CUDA(cudaStreamSynchronize(s));
    *n_out = (size_t)n;
    if (!ids && !layers) return OPB_OK; // count only
    if (!ids || !layers) { set_error("ids and layers must both be given"); return OPB_ERR_INVALID; }
    if ((size_t)n > cap) { set_error("halo of %d cubes exceeds the capacity %zu of the output buffers", n, cap); return OPB_ERR_CAPACITY; }
    if (n == 0) return OPB_OK;
    // pack into a second scratch region, then copy to wherever the caller's buffers live (host or device)
    std::vector<int> list((size_t)n);
    OPB_CUDA(cudaMemcpy(list.data(), d_list, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    const size_t off_list = 16, off_ids = off_list + (size_t)n * sizeof(int), off_layers = (off_ids + (size_t)n * 3 * sizeof(int) + 15) & ~(size_t)15;
    rc = halo_scratch(v, off_layers + (size_t)n * kLayerFloats * sizeof(float));
    if (rc) return rc;
    char *base = (char *)v->halo_scratch;
    OPB_CUDA(cudaMemcpyAsync(base + off_list, list.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, s));
    halo_pack_kernel<<<n, kLayerVoxels, 0, s>>>(v->dev, (const int *)(base + off_list), n, axis, (int *)(base + off_ids),
                                                (float *)(base + off_layers));
    OPB_CUDA(cudaGetLastError());
    OPB_CUDA(cudaMemcpyAsync(ids, base + off_ids, (size_t)n * 3 * sizeof(int), cudaMemcpyDefault, s));
    OPB_CUDA(cudaMemcpyAsync(layers, base + off_layers, (size_t)n * kLayerFloats * sizeof(float), cudaMemcpyDefault, s));
    OPB_CUDA(cudaStreamSynchronize(s));
    return OPB_OK;
}

int opb_volume_halo_import(opb_volume *v, const int32_t *ids, const float *layers, size_t n)
{
    if (!v || (n && (!ids || !layers))) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    if (n == 0) return OPB_OK;
    OPB_CUDA(cudaSetDevice(v->desc.device));
    size_t n_cubes = 0;
    int rc = opb_volume_num_cubes(v, &n_cubes);
    if (rc) return rc;
    if (n_cubes + (size_t)v->n_ghost + n > (size_t)v->dev.max_cubes)
    {
        set_error("%zu ghost cubes do not fit: %zu owned + %d ghosts of max_cubes=%d", n, n_cubes, v->n_ghost, v->dev.max_cubes);
        return OPB_ERR_CAPACITY;
    }
    const int axis = v->desc.shard_axis;
    if (axis < 0 || axis > 2) { set_error("shard_axis %d out of range", axis); return OPB_ERR_INVALID; }
    const size_t off_ids = 16, off_layers = (off_ids + n * 3 * sizeof(int) + 15) & ~(size_t)15;
    rc = halo_scratch(v, off_layers + n * kLayerFloats * sizeof(float));
    if (rc) return rc;
    cudaStream_t s = v->stream;
    char *base = (char *)v->halo_scratch;
    OPB_CUDA(cudaMemsetAsync(base, 0, sizeof(int), s));
    OPB_CUDA(cudaMemcpyAsync(base + off_ids, ids, n * 3 * sizeof(int), cudaMemcpyDefault, s));
    OPB_CUDA(cudaMemcpyAsync(base + off_layers, layers, n * kLayerFloats * sizeof(float), cudaMemcpyDefault, s));
    const int first = v->dev.max_cubes - v->n_ghost - (int)n;
    halo_unpack_kernel<<<(unsigned int)n, kLayerVoxels, 0, s>>>(v->dev, first, (int)n, axis, (const int *)(base + off_ids),
                                                               (const float *)(base + off_layers), (int *)base);
    OPB_CUDA(cudaGetLastError());
    OPB_CUDA(cudaStreamSynchronize(s));
    v->n_ghost += (int)n;
    return OPB_OK;
}

int opb_volume_halo_clear(opb_volume *v)
{
    if (!v) { set_error("volume is NULL"); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(v->desc.device));
    int rc = halo_drop_ghosts(v);
    if (rc) return rc;
    OPB_CUDA(cudaStreamSynchronize(v->stream));
    return OPB_OK;
}

int opb_volume_num_ghost_cubes(opb_volume *v, size_t *n)
{
    if (!v || !n) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    *n = (size_t)v->n_ghost;
    return OPB_OK;
}
} // extern "C"
