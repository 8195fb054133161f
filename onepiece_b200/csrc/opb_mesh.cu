// K3: Marching Cubes over the device-resident volume.
//
// Reference path rebuilt here (file:line relative to the reference tree):
//   CubeHandler::ExtractTriangleMesh  src/Integration/CubeHandler.cpp:9-44
//   CubeHandler::GenerateMeshByCube   src/Integration/CubeHandler.cpp:70-114   (8-corner gather, neighbour cubes)
//   integration::MarchingCube         src/Integration/MarchingCube.cpp:31-74   (3 fresh vertices per triangle)
//   DetermineCase / InterpolateEdgeVetex  MarchingCube.cpp:9-29
//   TriangleMesh::LoadFromMeshes      src/Geometry/TriangleMesh.cpp:73-94      (concatenation)
// The reference walks cubes in unordered_map order and cells x-outermost; the output here is ordered by pool
// slot and voxel index.  The triangle multiset is identical (tests compare it sorted, bit for bit).
//
// classify -> scan -> emit: one CTA per cube stages the 9x9x9 sdf lattice (own voxels plus the first layer of
// up to 7 neighbour cubes, found through the hash table) in shared memory, each thread classifies one cell,
// a block scan turns triangle counts into offsets; a device-wide scan over the per-cube totals places cubes.
#include <cstdlib>
#include <cstring>

#include "../../include/onepiece_b200.h"
#include "opb_mc_table.h"
#include "opb_volume.cuh"
#include "opb_volume_host.h"

namespace opb
{
constexpr int kLat = kCube + 1; // 9

struct McShared
{
    float sdf[kLat * kLat * kLat];
    unsigned char ok[kLat * kLat * kLat];
    int nb[8];
    int warp_sums[16];
};

__device__ __forceinline__ int lat_index(int x, int y, int z) { return x + kLat * (y + kLat * z); }

// stages the lattice; returns through sm.  All 512 threads participate.
__device__ __forceinline__ void mc_stage(const VolumeDev &vol, int slot, McShared &sm)
{
    const int t = threadIdx.x;
    if (t < 8)
    {
        const int i = vol.slot_ids[3 * slot], j = vol.slot_ids[3 * slot + 1], k = vol.slot_ids[3 * slot + 2];
        sm.nb[t] = t == 0 ? slot : table_find(vol, i + (t & 1), j + ((t >> 1) & 1), k + ((t >> 2) & 1));
    }
    __syncthreads();
    for (int e = t; e < kLat * kLat * kLat; e += blockDim.x)
    {
        const int x = e % kLat, y = (e / kLat) % kLat, z = e / (kLat * kLat);
        const int o = (x >> 3) | ((y >> 3) << 1) | ((z >> 3) << 2);
        const int s = sm.nb[o];
        float sdf = 999.0f;
        bool ok = false;
        if (s >= 0)
        {
            const int vid = (x & 7) + ((y & 7) << 3) + ((z & 7) << 6);
            const float *base = vol.pool + (size_t)s * kSlotFloats;
            sdf = base[vid];
            const float w = base[kCubeVoxels + vid];
            ok = !(sdf >= 1 || w <= 0); // TSDFVoxel::IsValid
        }
        sm.sdf[e] = sdf;
        sm.ok[e] = ok;
    }
    __syncthreads();
}

// case index of the cell owned by this thread, or -1 if any corner is missing/invalid
__device__ __forceinline__ int mc_case(const McShared &sm, int x, int y, int z)
{
    int cs = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
        const int e = lat_index(x + kMcCornerOffset[i][0], y + kMcCornerOffset[i][1], z + kMcCornerOffset[i][2]);
        if (!sm.ok[e]) return -1;
        if (sm.sdf[e] > 0) cs |= 1 << i; // DetermineCase
    }
    return cs;
}

// exclusive block scan over 512 threads; returns this thread's offset, total in *total
__device__ __forceinline__ int block_scan_512(int v, McShared &sm, int *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) sm.warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0)
    {
        int w = lane < 16 ? sm.warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1)
        {
            const int n = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += n;
        }
        if (lane < 16) sm.warp_sums[lane] = w; // inclusive
    }
    __syncthreads();
    const int base = warp == 0 ? 0 : sm.warp_sums[warp - 1];
    *total = sm.warp_sums[15];
    return base + inc - v;
}

// Emission order.  CTA b handles the cube order[b] (slot b without an order): the triangles of the mesh come cube by cube in that
// sequence, and inside a cube cell by cell with x outermost and z innermost -- the nesting of GenerateMeshByCube's loops
// (CubeHandler.cpp:75-79) -- so that a caller who supplies the iteration order of the reference's cube map gets the reference's mesh
// triangle for triangle (opb_volume_extract_mesh_ordered).
__global__ void __launch_bounds__(512) mc_count_kernel(VolumeDev vol, int n_slots, const int *order, unsigned int *cube_tris)
{
    __shared__ McShared sm;
    if ((int)blockIdx.x >= n_slots) return;
    const int slot = order ? order[blockIdx.x] : (int)blockIdx.x;
    mc_stage(vol, slot, sm);
    const int t = threadIdx.x;
    const int cs = mc_case(sm, t >> 6, (t >> 3) & 7, t & 7);
    const int ntri = cs < 0 ? 0 : kMcTriCount[cs];
    int total;
    block_scan_512(ntri, sm, &total);
    if (t == 0) cube_tris[blockIdx.x] = (unsigned int)total;
}

// in-place exclusive scan of n counts by one CTA of 1024 threads; writes the grand total to *total
__global__ void __launch_bounds__(1024) scan_counts_kernel(unsigned int *counts, int n, unsigned long long *total)
{
    __shared__ unsigned long long warp_sums[32];
    __shared__ unsigned long long carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024)
    {
        const int i = base + threadIdx.x;
        const unsigned long long v = i < n ? counts[i] : 0;
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const unsigned long long m = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += m;
        }
        if (lane == 31) warp_sums[warp] = inc;
        __syncthreads();
        if (warp == 0)
        {
            unsigned long long w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const unsigned long long m = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += m;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const unsigned long long excl = carry + (warp ? warp_sums[warp - 1] : 0) + inc - v;
        // triangle offsets are stored as 32-bit: 4 G triangles would be 144 GB of vertices anyway
        if (i < n) counts[i] = (unsigned int)excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(512) mc_emit_kernel(VolumeDev vol, int n_slots, const int *order, const unsigned int *cube_offsets, float res,
                                                      float cube_res, float half_res, float *xyz, float *rgb)
{
    __shared__ McShared sm;
    if ((int)blockIdx.x >= n_slots) return;
    const int slot = order ? order[blockIdx.x] : (int)blockIdx.x;
    mc_stage(vol, slot, sm);
    const int t = threadIdx.x;
    const int x = t >> 6, y = (t >> 3) & 7, z = t & 7;
    const int cs = mc_case(sm, x, y, z);
    const int ntri = cs < 0 ? 0 : kMcTriCount[cs];
    int total;
    const int off = block_scan_512(ntri, sm, &total);
    if (ntri == 0) return;
    size_t out = ((size_t)cube_offsets[blockIdx.x] + off) * 9; // floats: 3 vertices x 3
    const unsigned long long row = kMcCases[cs];
    for (int k = 0; k < 3 * ntri; ++k)
    {
        const int e = (int)((row >> (4 * k)) & 0xF);
        const int ca = kMcEdgeCorners[e][0], cb = kMcEdgeCorners[e][1];
        float pa[3], pb[3], cola[3], colb[3], sa, sb;
#pragma unroll
        for (int side = 0; side < 2; ++side)
        {
            const int c = side ? cb : ca;
            const int lx = x + kMcCornerOffset[c][0], ly = y + kMcCornerOffset[c][1], lz = z + kMcCornerOffset[c][2];
            const int s = sm.nb[(lx >> 3) | ((ly >> 3) << 1) | ((lz >> 3) << 2)];
            const int vx = lx & 7, vy = ly & 7, vz = lz & 7;
            const int vid = vx + (vy << 3) + (vz << 6);
            const float *base = vol.pool + (size_t)s * kSlotFloats;
            float *p = side ? pb : pa, *col = side ? colb : cola;
            // corner position: VoxelCube::GetOrigin + VoxelCentroidOffSet (CubeHandler.cpp:98)
            p[0] = fadd(cube_origin(vol.slot_ids[3 * s], cube_res), centroid_offset(vx, res, half_res));
            p[1] = fadd(cube_origin(vol.slot_ids[3 * s + 1], cube_res), centroid_offset(vy, res, half_res));
            p[2] = fadd(cube_origin(vol.slot_ids[3 * s + 2], cube_res), centroid_offset(vz, res, half_res));
            col[0] = base[2 * kCubeVoxels + vid];
            col[1] = base[3 * kCubeVoxels + vid];
            col[2] = base[4 * kCubeVoxels + vid];
            (side ? sb : sa) = sm.sdf[lat_index(lx, ly, lz)];
        }
        // InterpolateEdgeVetex (MarchingCube.cpp:9-16): corner1 - (sdf1/(sdf2-sdf1)) * (corner2 - corner1)
        const float tt = fdiv(sa, fsub(sb, sa));
#pragma unroll
        for (int c = 0; c < 3; ++c)
        {
            xyz[out + c] = fsub(pa[c], fmul(tt, fsub(pb[c], pa[c])));
            rgb[out + c] = fdiv(fadd(cola[c], colb[c]), 2.0f); // MarchingCube.cpp:59
        }
        out += 3;
    }
}

static int mesh_count(opb_volume *v, int *n_slots_out, unsigned int **d_counts_out, unsigned long long *n_tris, const int *d_order = nullptr)
{
    size_t n = 0;
    int rc = opb_volume_num_cubes(v, &n);
    if (rc) return rc;
    *n_slots_out = (int)n;
    *n_tris = 0;
    *d_counts_out = nullptr;
    if (n == 0) return OPB_OK;
    rc = volume_materialize(v); // packed volumes are meshed from their float mirror
    if (rc) return rc;
    const size_t need = n * sizeof(unsigned int) + 16;
    if (v->mesh_scratch_bytes < need)
    {
        cudaFree(v->mesh_scratch);
        v->mesh_scratch = nullptr;
        v->mesh_scratch_bytes = 0;
        OPB_CUDA(cudaMalloc(&v->mesh_scratch, need));
        v->mesh_scratch_bytes = need;
    }
    unsigned long long *d_total = (unsigned long long *)v->mesh_scratch;
    unsigned int *d_counts = (unsigned int *)((char *)v->mesh_scratch + 16);
    mc_count_kernel<<<(unsigned int)n, 512, 0, v->stream>>>(v->dev, (int)n, d_order, d_counts);
    scan_counts_kernel<<<1, 1024, 0, v->stream>>>(d_counts, (int)n, d_total);
    OPB_CUDA(cudaGetLastError());
    OPB_CUDA(cudaMemcpyAsync(n_tris, d_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, v->stream));
    OPB_CUDA(cudaStreamSynchronize(v->stream));
    *d_counts_out = d_counts;
    return OPB_OK;
}
__global__ void iota_kernel(unsigned int *a, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a[i] = (unsigned int)i;
}
} // namespace opb

using namespace opb;

extern "C"
{
// ExtractTriangleMesh followed by TriangleMesh::ClusteringSimplify(grid_len), the pair every fusion main runs before writing
// its PLY (example/DenseFusion/DenseFusion.cpp:99-105): the raw mesh (72 B per triangle) never leaves the device.
int opb_volume_extract_mesh_clustered(opb_volume *v, float grid_len, float **xyz, float **rgb, uint32_t **tri, size_t *nv, size_t *nt)
{
    if (!v || !xyz || !rgb || !tri || !nv || !nt) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    *xyz = *rgb = nullptr; *tri = nullptr; *nv = *nt = 0;
    if (!(grid_len > 0)) { set_error("[ClusteringMeshSimplification]::[ERROR]::Grid length cannot be less than 0."); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(v->desc.device));
    int n_slots;
    unsigned int *d_counts;
    unsigned long long tris;
    int rc = mesh_count(v, &n_slots, &d_counts, &tris);
    if (rc || tris == 0) return rc;
    if (tris > 0x2AAAAAAAull) { set_error("mesh of %llu triangles exceeds 32-bit vertex indices", tris); return OPB_ERR_CAPACITY; }
    const size_t nfl = (size_t)tris * 9;
    float *d_xyz = nullptr, *d_rgb = nullptr;
    unsigned int *d_tri = nullptr;
    cudaError_t e = cudaMalloc(&d_xyz, nfl * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_rgb, nfl * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_tri, (size_t)tris * 3 * sizeof(unsigned int));
    if (e != cudaSuccess)
    {
        cudaFree(d_xyz); cudaFree(d_rgb); cudaFree(d_tri);
        set_error("cudaMalloc of the mesh buffers failed: %s", cudaGetErrorString(e));
        return OPB_ERR_CUDA;
    }
    const float res = v->desc.voxel_resolution;
    mc_emit_kernel<<<n_slots, 512, 0, v->stream>>>(v->dev, n_slots, nullptr, d_counts, res, (float)kCube * res, res / 2, d_xyz, d_rgb);
    iota_kernel<<<v->sm_count * 8, 256, 0, v->stream>>>(d_tri, (size_t)tris * 3); // three fresh vertices per triangle
    float *r_p = nullptr, *r_c = nullptr;
    unsigned int *r_t = nullptr;
    size_t r_nv = 0, r_nt = 0;
    rc = cudaGetLastError() == cudaSuccess ? OPB_OK : OPB_ERR_CUDA;
    if (rc == OPB_OK)
        rc = clustering_simplify_device(v->sm_count, v->stream, d_xyz, d_rgb, (size_t)tris * 3, d_tri, (size_t)tris, grid_len, &r_p, &r_c, &r_t, &r_nv, &r_nt);
    else set_error("mesh extraction failed");
    cudaFree(d_xyz); cudaFree(d_rgb); cudaFree(d_tri);
    if (rc) return rc;
    rc = mesh_result_to_host(v->stream, r_p, r_c, r_t, r_nv, r_nt, xyz, rgb, tri);
    if (rc) return rc;
    *nv = r_nv; *nt = r_nt;
    return OPB_OK;
}

int opb_volume_count_mesh(opb_volume *v, size_t *nv, size_t *nt)
{
    if (!v || !nv || !nt) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(v->desc.device));
    int n_slots;
    unsigned int *d_counts;
    unsigned long long tris;
    int rc = mesh_count(v, &n_slots, &d_counts, &tris);
    if (rc) return rc;
    *nt = (size_t)tris;
    *nv = (size_t)tris * 3;
    return OPB_OK;
}

static int extract_mesh_impl(opb_volume *v, const int *d_order, float **xyz, float **rgb, uint32_t **tri, size_t *nv, size_t *nt)
{
    if (!v || !xyz || !rgb || !tri || !nv || !nt) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    *xyz = *rgb = nullptr; *tri = nullptr; *nv = *nt = 0;
    OPB_CUDA(cudaSetDevice(v->desc.device));
    int n_slots;
    unsigned int *d_counts;
    unsigned long long tris;
    int rc = mesh_count(v, &n_slots, &d_counts, &tris, d_order);
    if (rc) return rc;
    if (tris == 0) return OPB_OK;
    if (tris > 0xFFFFFFFFull / 3) { set_error("mesh of %llu triangles exceeds 32-bit vertex indices", tris); return OPB_ERR_CAPACITY; }
    const size_t nfl = (size_t)tris * 9;
    float *d_xyz = nullptr, *d_rgb = nullptr;
    OPB_CUDA(cudaMalloc(&d_xyz, nfl * sizeof(float)));
    if (cudaMalloc(&d_rgb, nfl * sizeof(float)) != cudaSuccess)
    {
        cudaFree(d_xyz);
        set_error("cudaMalloc of the mesh colour buffer failed: %s", cudaGetErrorString(cudaGetLastError()));
        return OPB_ERR_CUDA;
    }
    const float res = v->desc.voxel_resolution;
    mc_emit_kernel<<<n_slots, 512, 0, v->stream>>>(v->dev, n_slots, d_order, d_counts, res, (float)kCube * res, res / 2, d_xyz, d_rgb);
    float *h_xyz = (float *)malloc(nfl * sizeof(float)), *h_rgb = (float *)malloc(nfl * sizeof(float));
    uint32_t *h_tri = (uint32_t *)malloc((size_t)tris * 3 * sizeof(uint32_t));
    cudaError_t e = cudaGetLastError();
    if (h_xyz && h_rgb && h_tri && e == cudaSuccess)
    {
        e = cudaMemcpyAsync(h_xyz, d_xyz, nfl * sizeof(float), cudaMemcpyDeviceToHost, v->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_rgb, d_rgb, nfl * sizeof(float), cudaMemcpyDeviceToHost, v->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(v->stream);
    }
    cudaFree(d_xyz);
    cudaFree(d_rgb);
    if (!h_xyz || !h_rgb || !h_tri || e != cudaSuccess)
    {
        free(h_xyz); free(h_rgb); free(h_tri);
        set_error("mesh extraction failed: %s", e != cudaSuccess ? cudaGetErrorString(e) : "host allocation");
        return e != cudaSuccess ? OPB_ERR_CUDA : OPB_ERR_CAPACITY;
    }
    // every triangle owns three fresh vertices (MarchingCube.cpp:65-71)
    for (size_t i = 0; i < (size_t)tris * 3; ++i) h_tri[i] = (uint32_t)i;
    *xyz = h_xyz; *rgb = h_rgb; *tri = h_tri;
    *nt = (size_t)tris;
    *nv = (size_t)tris * 3;
    return OPB_OK;
}
int opb_volume_extract_mesh(opb_volume *v, float **xyz, float **rgb, uint32_t **tri, size_t *nv, size_t *nt)
{
    return extract_mesh_impl(v, nullptr, xyz, rgb, tri, nv, nt);
}

// the cubes of the volume in the caller's sequence -> their slots (-1: no such cube)
__global__ void ids_to_slots_kernel(VolumeDev vol, const int *ids, int n, int *slots, int *missing)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const int s = opb::table_find(vol, ids[3 * i], ids[3 * i + 1], ids[3 * i + 2]);
        slots[i] = s;
        if (s < 0 || s >= *vol.n_alloc) atomicAdd(missing, 1);
    }
}

int opb_volume_extract_mesh_ordered(opb_volume *v, const int32_t *cube_ids, size_t n_ids, float **xyz, float **rgb, uint32_t **tri, size_t *nv,
                                    size_t *nt)
{
    if (!v || !xyz || !rgb || !tri || !nv || !nt || (n_ids && !cube_ids)) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    *xyz = *rgb = nullptr; *tri = nullptr; *nv = *nt = 0;
    OPB_CUDA(cudaSetDevice(v->desc.device));
    size_t n = 0;
    int rc = opb_volume_num_cubes(v, &n);
    if (rc) return rc;
    if (n_ids != n) { set_error("the cube sequence has %zu entries, the volume %zu cubes", n_ids, n); return OPB_ERR_INVALID; }
    if (n == 0) return OPB_OK;
    int *d_ids = nullptr;
    OPB_CUDA(cudaMalloc(&d_ids, (n * 4 + 1) * sizeof(int))); // ids (3n), slots (n), missing (1)
    int *d_slots = d_ids + 3 * n, *d_missing = d_slots + n;
    cudaError_t e = cudaMemcpyAsync(d_ids, cube_ids, n * 3 * sizeof(int), cudaMemcpyHostToDevice, v->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_missing, 0, sizeof(int), v->stream);
    int missing = 0;
    if (e == cudaSuccess)
    {
        ids_to_slots_kernel<<<v->sm_count * 4, 256, 0, v->stream>>>(v->dev, d_ids, (int)n, d_slots, d_missing);
        e = cudaMemcpyAsync(&missing, d_missing, sizeof(int), cudaMemcpyDeviceToHost, v->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(v->stream);
    if (e != cudaSuccess || missing)
    {
        cudaFree(d_ids);
        if (e != cudaSuccess) { set_error("mesh extraction failed: %s", cudaGetErrorString(e)); return OPB_ERR_CUDA; }
        set_error("%d cubes of the sequence are not in the volume", missing);
        return OPB_ERR_INVALID;
    }
    rc = extract_mesh_impl(v, d_slots, xyz, rgb, tri, nv, nt);
    cudaFree(d_ids);
    return rc;
}
} // extern "C"
