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
#include <cstddef>
#include <cstring>
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

// ---------------------------------------------------------------------------------------------------------
// The same exchange without a host in the transport: the export kernel writes ids and layers straight into the
// destination rank's receive box (cudaIpc-mapped peer memory, plain stores over NVLink), raises a flag there, and the import
// kernel of that rank -- already queued on its stream -- waits for the flag in its own memory, registers the ghosts and
// acknowledges into the sender's box (so the sender's NEXT export cannot overwrite a box still being read).  One launch on
// either side, no NCCL call, no host round trip between them; the host only reads three counters back at the end.
// ---------------------------------------------------------------------------------------------------------
struct HaloBox
{
    unsigned long long flag;  // written by the rank that exports into this box: number of its completed exports
    unsigned long long ack;   // written by the rank this box's owner exports to: number of imports it has completed
    unsigned int count;       // cubes of the export `flag` announces (may exceed the capacity: then the import fails loudly)
    unsigned int pad[3];
    // int32 ids[3 * cap], then float layers[320 * cap] (16-byte aligned)
};
struct HaloLocal // per-volume device words of the exchange
{
    unsigned int n_sent, tickets_out, tickets_in, n_imported, rejected;
    int error; // 1 flag wait timed out, 2 the received cubes do not fit into the pool, 3 more boundary cubes than the box holds
};
__host__ __device__ inline size_t halo_box_ids_offset() { return sizeof(HaloBox); }
__host__ __device__ inline size_t halo_box_layers_offset(size_t cap) { return (sizeof(HaloBox) + cap * 3 * sizeof(int) + 15) & ~(size_t)15; }
__host__ __device__ inline size_t halo_box_bytes(size_t cap) { return halo_box_layers_offset(cap) + cap * kLayerFloats * sizeof(float); }

__device__ __forceinline__ bool halo_wait(volatile unsigned long long *word, unsigned long long want, int *error)
{
    const unsigned long long t0 = global_timer_ns();
    while (*word < want)
    {
        if (*(volatile int *)error) return false; // the exchange already failed (time limit or capacity): keep the first reason
        if (global_timer_ns() - t0 > 4000000000ull) { atomicCAS(error, 0, 1); return false; }
    }
    return true;
}

constexpr int kHaloThreads = 256;
// one warp per 32 slots: every lane tests one slot, then the warp copies the layers of the marked ones (lane e and e + 32)
__global__ void __launch_bounds__(kHaloThreads) halo_export_peer_kernel(VolumeDev v, int axis, int slab, HaloBox *own, HaloBox *dst, unsigned int dst_cap,
                                                                        unsigned long long epoch, HaloLocal *loc)
{
    __shared__ int s_go;
    if (threadIdx.x == 0) s_go = halo_wait(&own->ack, epoch - 1, &loc->error) ? 1 : 0; // the destination has read our previous export
    __syncthreads();
    const int lane = threadIdx.x & 31;
    if (s_go)
    {
        int *ids = (int *)((char *)dst + halo_box_ids_offset());
        float *layers = (float *)((char *)dst + halo_box_layers_offset(dst_cap));
        int n_alloc = *v.n_alloc;
        if (n_alloc > v.max_cubes) n_alloc = v.max_cubes;
        const int warps = (gridDim.x * kHaloThreads) >> 5, wid = (blockIdx.x * kHaloThreads + threadIdx.x) >> 5;
        for (int base = wid * 32; base < n_alloc; base += warps * 32)
        {
            const int s = base + lane;
            const bool mine = s < n_alloc && floor_mod(v.slot_ids[3 * s + axis], slab) == 0;
            unsigned int m = __ballot_sync(0xffffffffu, mine);
            unsigned int first = 0;
            if (lane == 0 && m) first = atomicAdd(&loc->n_sent, (unsigned int)__popc(m));
            first = __shfl_sync(0xffffffffu, first, 0);
            while (m)
            {
                const int b = __ffs(m) - 1;
                m &= m - 1u;
                const unsigned int c = first++;
                if (c >= dst_cap) continue;
                const int slot = base + b;
                if (lane < 3) ids[3 * c + lane] = v.slot_ids[3 * slot + lane];
#pragma unroll
                for (int half = 0; half < 2; ++half)
                {
                    const int e = lane + 32 * half;
                    const float *src = v.pool + (size_t)slot * kSlotFloats + layer_voxel(axis, e);
                    float *out = layers + (size_t)c * kLayerFloats + e;
#pragma unroll
                    for (int p = 0; p < kPlanes; ++p) out[p * kLayerVoxels] = src[p * kCubeVoxels];
                }
            }
        }
    }
    // the CTA that finishes last publishes the count, then the flag
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(&loc->tickets_out, 1u) == gridDim.x - 1)
    {
        __threadfence();
        const unsigned int n = *(volatile unsigned int *)&loc->n_sent;
        if (n > dst_cap) atomicCAS(&loc->error, 0, 3);
        *(volatile unsigned int *)&dst->count = n;
        __threadfence_system();
        *(volatile unsigned long long *)&dst->flag = epoch;
        loc->tickets_out = 0;
    }
}

__global__ void __launch_bounds__(kHaloThreads) halo_import_peer_kernel(VolumeDev v, int axis, int n_ghost, HaloBox *own, unsigned int own_cap, HaloBox *src,
                                                                        unsigned long long epoch, HaloLocal *loc)
{
    __shared__ int s_n;
    if (threadIdx.x == 0)
    {
        int n = -1;
        if (halo_wait(&own->flag, epoch, &loc->error))
        {
            __threadfence_system();
            n = (int)*(volatile unsigned int *)&own->count;
            int n_alloc = *v.n_alloc;
            if (n_alloc > v.max_cubes) n_alloc = v.max_cubes;
            if ((unsigned int)n > own_cap) { atomicCAS(&loc->error, 0, 3); n = -1; }
            else if ((long long)n_alloc + n_ghost + n > (long long)v.max_cubes) { atomicCAS(&loc->error, 0, 2); n = -1; }
        }
        s_n = n;
    }
    __syncthreads();
    const int n = s_n, lane = threadIdx.x & 31;
    if (n > 0)
    {
        const int *ids = (const int *)((const char *)own + halo_box_ids_offset());
        const float *layers = (const float *)((const char *)own + halo_box_layers_offset(own_cap));
        const int first_slot = v.max_cubes - n_ghost - n;
        const int warps = (gridDim.x * kHaloThreads) >> 5, wid = (blockIdx.x * kHaloThreads + threadIdx.x) >> 5;
        for (int c = wid; c < n; c += warps)
        {
            const int slot = first_slot + c;
            int ok = 0;
            if (lane == 0)
            {
                const int i = __ldcv(&ids[3 * c]), j = __ldcv(&ids[3 * c + 1]), k = __ldcv(&ids[3 * c + 2]);
                v.slot_ids[3 * slot] = i; v.slot_ids[3 * slot + 1] = j; v.slot_ids[3 * slot + 2] = k;
                unsigned long long key;
                ok = pack_id(i, j, k, key) ? 1 : 0;
                if (ok)
                {
                    unsigned int h = hash_key(key) & v.table_mask;
                    for (;;)
                    {
                        const unsigned long long prev = atomicCAS(&v.keys[h], kEmptyKey, key);
                        if (prev == kEmptyKey) { v.vals[h] = slot; break; }
                        if (prev == key) { ok = 0; break; } // the cube already lives here: keep the first
                        h = (h + 1) & v.table_mask;
                    }
                }
                if (!ok) atomicAdd(&loc->rejected, 1u);
            }
            ok = __shfl_sync(0xffffffffu, ok, 0);
            if (!ok) continue;
#pragma unroll
            for (int half = 0; half < 2; ++half)
            {
                const int e = lane + 32 * half;
                float *base = v.pool + (size_t)slot * kSlotFloats + layer_voxel(axis, e);
                const float *in = layers + (size_t)c * kLayerFloats + e;
#pragma unroll
                for (int p = 0; p < kPlanes; ++p) base[p * kCubeVoxels] = __ldcv(&in[p * kLayerVoxels]);
            }
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(&loc->tickets_in, 1u) == gridDim.x - 1)
    {
        loc->n_imported = n > 0 ? (unsigned int)n : 0u;
        loc->tickets_in = 0;
        // the box may be overwritten by the sender's next export from here on
        __threadfence_system();
        if (n >= 0 || loc->error >= 2) *(volatile unsigned long long *)&src->ack = epoch;
    }
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
    if (v->desc.shard_world <= 1) return OPB_OK; // nothing is owned elsewhere (packed volumes are never sharded)
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
    int rc = volume_require_float(v, "opb_volume_halo_import");
    if (rc == OPB_OK) rc = opb_volume_num_cubes(v, &n_cubes);
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
int opb_volume_halo_peer_buffer(opb_volume *v, size_t cap_cubes, void **d_buffer, unsigned char ipc_handle[OPB_IPC_HANDLE_BYTES])
{
    if (!v || !d_buffer) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    if (cap_cubes == 0 || cap_cubes > 0x7fffffffu) { set_error("halo box capacity %zu out of range", cap_cubes); return OPB_ERR_INVALID; }
    if (volume_require_float(v, "the boundary-cube exchange")) return OPB_ERR_UNSUPPORTED;
    OPB_CUDA(cudaSetDevice(v->desc.device));
    if (v->halo_box && v->halo_box_cap != cap_cubes)
    {
        set_error("the halo box of this volume already exists with capacity %zu (peers may have it mapped)", v->halo_box_cap);
        return OPB_ERR_INVALID;
    }
    if (!v->halo_box)
    {
        OPB_CUDA(cudaMalloc(&v->halo_box, halo_box_bytes(cap_cubes)));
        OPB_CUDA(cudaMemset(v->halo_box, 0, sizeof(HaloBox)));
        OPB_CUDA(cudaMalloc(&v->halo_local, sizeof(HaloLocal)));
        OPB_CUDA(cudaMemset(v->halo_local, 0, sizeof(HaloLocal)));
        v->halo_box_cap = cap_cubes;
        cudaFuncAttributes fa; // load both kernels now (a first launch may wait for the device to drain: see opb_icp_create)
        OPB_CUDA(cudaFuncGetAttributes(&fa, halo_export_peer_kernel));
        OPB_CUDA(cudaFuncGetAttributes(&fa, halo_import_peer_kernel));
    }
    *d_buffer = v->halo_box;
    if (ipc_handle)
    {
        cudaIpcMemHandle_t h;
        OPB_CUDA(cudaIpcGetMemHandle(&h, v->halo_box));
        memcpy(ipc_handle, &h, sizeof(h));
    }
    return OPB_OK;
}

int opb_volume_halo_peer_attach(opb_volume *v, void *dst_buffer, size_t dst_cap_cubes, void *src_buffer)
{
    if (!v) { set_error("volume is NULL"); return OPB_ERR_INVALID; }
    if (!v->halo_box) { set_error("call opb_volume_halo_peer_buffer first"); return OPB_ERR_INVALID; }
    if ((dst_buffer == nullptr) != (src_buffer == nullptr)) { set_error("dst_buffer and src_buffer must both be given (or both NULL to detach)"); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(v->desc.device));
    OPB_CUDA(cudaStreamSynchronize(v->stream));
    v->halo_dst = dst_buffer;
    v->halo_src = src_buffer;
    v->halo_dst_cap = dst_cap_cubes;
    return OPB_OK;
}

int opb_volume_halo_exchange_begin(opb_volume *v)
{
    if (!v) { set_error("volume is NULL"); return OPB_ERR_INVALID; }
    if (!v->halo_box || !v->halo_dst || !v->halo_src) { set_error("no peer boxes attached (opb_volume_halo_peer_attach)"); return OPB_ERR_INVALID; }
    if (v->desc.shard_world <= 1) { set_error("the volume is not sharded"); return OPB_ERR_INVALID; }
    const int axis = v->desc.shard_axis, slab = v->desc.shard_slab_cubes > 0 ? v->desc.shard_slab_cubes : 1;
    if (axis < 0 || axis > 2) { set_error("shard_axis %d out of range", axis); return OPB_ERR_INVALID; }
    if (v->halo_pending) { set_error("an exchange is already in flight (opb_volume_halo_exchange_end)"); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(v->desc.device));
    cudaStream_t s = v->stream;
    HaloLocal *loc = (HaloLocal *)v->halo_local;
    const unsigned long long epoch = ++v->halo_epoch;
    OPB_CUDA(cudaMemsetAsync(loc, 0, offsetof(HaloLocal, error), s)); // the error word is sticky
    const int nb = v->sm_count * 4;
    halo_export_peer_kernel<<<nb, kHaloThreads, 0, s>>>(v->dev, axis, slab, (HaloBox *)v->halo_box, (HaloBox *)v->halo_dst, (unsigned int)v->halo_dst_cap,
                                                        epoch, loc);
    halo_import_peer_kernel<<<nb, kHaloThreads, 0, s>>>(v->dev, axis, v->n_ghost, (HaloBox *)v->halo_box, (unsigned int)v->halo_box_cap,
                                                        (HaloBox *)v->halo_src, epoch, loc);
    OPB_CUDA(cudaGetLastError());
    v->halo_pending = true;
    return OPB_OK;
}

int opb_volume_halo_exchange_end(opb_volume *v, size_t *n_sent, size_t *n_imported)
{
    if (!v) { set_error("volume is NULL"); return OPB_ERR_INVALID; }
    if (!v->halo_pending) { set_error("no exchange in flight"); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(v->desc.device));
    HaloLocal h;
    OPB_CUDA(cudaMemcpyAsync(&h, v->halo_local, sizeof(h), cudaMemcpyDeviceToHost, v->stream));
    OPB_CUDA(cudaStreamSynchronize(v->stream));
    v->halo_pending = false;
    if (n_sent) *n_sent = h.n_sent;
    if (n_imported) *n_imported = 0;
    if (h.error == 1) { set_error("halo exchange timed out: a peer rank did not make the matching call"); return OPB_ERR_CUDA; }
    if (h.error == 2) { set_error("the received boundary cubes do not fit into the block pool (max_cubes=%d)", v->dev.max_cubes); return OPB_ERR_CAPACITY; }
    if (h.error == 3) { set_error("more boundary cubes (%u) than a peer's halo box holds", h.n_sent); return OPB_ERR_CAPACITY; }
    v->n_ghost += (int)h.n_imported;
    if (n_imported) *n_imported = h.n_imported;
    return OPB_OK;
}

int opb_volume_halo_exchange_peer(opb_volume *v, size_t *n_sent, size_t *n_imported)
{
    const int rc = opb_volume_halo_exchange_begin(v);
    return rc ? rc : opb_volume_halo_exchange_end(v, n_sent, n_imported);
}
} // extern "C"
