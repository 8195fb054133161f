// Mesh post-processing on the device: what the fusion mains run on the Marching-Cubes output before
// they write the PLY (example/DenseFusion/DenseFusion.cpp:105, example/ImageSequenceIntegration.cpp:56,
// example/MergeMultipleSubmaps.cpp:45-46).
//
// Reference path rebuilt here (file:line relative to the reference tree):
//   TriangleMesh::ClusteringSimplify   src/Geometry/TriangleMesh.cpp:53-58
//   ClusteringSimplification           src/Geometry/MeshSimplification.cpp:579-657   (vertex clustering on a hashed grid)
//   UpdateMesh / CompactMesh           src/Geometry/MeshSimplification.cpp:114-139,314-343
//   TriangleMesh::ComputeNormals       src/Geometry/TriangleMesh.cpp:95-127
//
// The reference is one sequential loop over the triangles with an unordered_map of grid cells: the first vertex that falls
// into a cell (in triangle order) becomes the cell's representative, every vertex reference is added to the cell's float
// sum in that same order, and the representative finally moves to sum / count.  The float sum makes the ORDER part of the
// result, so the device version keeps it: vertex references s = 3*triangle + corner are grouped per cell (hash insert ->
// dense cell index -> counting sort), each cell's members are put in ascending s and summed sequentially by one thread.
// Everything else (representative lookup, degenerate-triangle removal, ordered compaction of triangles and of the vertices
// still referenced) is order-free or an ordered scan, so points, colours and triangle indices come out bit-identical and in
// the reference's order.  ComputeNormals is the same pattern with the vertex as the group and the face normal as the value.
#include <cstdlib>
#include <cstring>

#include "../../include/onepiece_b200.h"
#include "opb_common.cuh"

namespace opb
{
constexpr unsigned long long kNoCell = 0xFFFFFFFFFFFFFFFFull;
constexpr int kCellBias = 1 << 20;

// in-place exclusive scan of n counts by one CTA of 1024 threads; the grand total goes to *total
__global__ void __launch_bounds__(1024) post_scan_kernel(unsigned int *counts, int n, unsigned int *total)
{
    __shared__ unsigned int warp_sums[32];
    __shared__ unsigned int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024)
    {
        const int i = base + threadIdx.x;
        const unsigned int v = i < n ? counts[i] : 0;
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
            unsigned int w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const unsigned int m = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += m;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const unsigned int excl = carry + (warp ? warp_sums[warp - 1] : 0) + inc - v;
        if (i < n) counts[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

struct ClusterDev
{
    const float *points;      // nv x 3
    const unsigned int *tri;  // nt x 3 vertex references; nullptr: reference s is point s itself (point-cloud down-sampling)
    int nv, nt;
    int nref;                 // 3 * nt, or the number of points
    float grid_len;
    unsigned long long *keys; // hash table of cells (cap entries)
    unsigned int *vals;       // dense cell index per table entry
    unsigned int cap_mask;
    unsigned int *n_cells;
    int *bad;                 // a vertex outside the 21-bit cell range or a triangle naming a missing vertex
    unsigned int *ref_cell;   // 3 nt: dense cell of every vertex reference
    unsigned int *cell_off;   // n_cells (+1): counts, then exclusive offsets
    unsigned int *cell_fill;  // n_cells
    unsigned int *members;    // 3 nt: references grouped per cell
    unsigned int *cell_rep;   // n_cells: representative vertex
    float *cell_mean;         // n_cells x 3
};

__device__ __forceinline__ unsigned long long cluster_hash(unsigned long long k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}
// GetGridIndex (MeshSimplification.cpp:575-578): floor(p / grid_size) per axis, packed 21 bits each
__device__ __forceinline__ bool cell_key(const ClusterDev &d, unsigned int v, unsigned long long &key)
{
    if (v >= (unsigned int)d.nv) return false;
    const float *p = d.points + 3 * (size_t)v;
    const int a = cvtt_x86(floorf(fdiv(p[0], d.grid_len))), b = cvtt_x86(floorf(fdiv(p[1], d.grid_len))), c = cvtt_x86(floorf(fdiv(p[2], d.grid_len)));
    const unsigned int ua = (unsigned int)(a + kCellBias), ub = (unsigned int)(b + kCellBias), uc = (unsigned int)(c + kCellBias);
    if ((ua | ub | uc) >> 21) return false;
    key = ((unsigned long long)ua << 42) | ((unsigned long long)ub << 21) | uc;
    return true;
}

__device__ __forceinline__ unsigned int ref_vertex(const ClusterDev &d, int s) { return d.tri ? d.tri[s] : (unsigned int)s; }

// A1: register the cell of every vertex reference; the inserting thread draws the dense cell index
__global__ void __launch_bounds__(256) cluster_insert_kernel(ClusterDev d)
{
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < d.nref; s += gridDim.x * blockDim.x)
    {
        unsigned long long key;
        if (!cell_key(d, ref_vertex(d, s), key)) { *d.bad = 1; continue; }
        unsigned int h = (unsigned int)cluster_hash(key) & d.cap_mask;
        for (;;)
        {
            const unsigned long long prev = atomicCAS(&d.keys[h], kNoCell, key);
            if (prev == kNoCell) { d.vals[h] = atomicAdd(d.n_cells, 1u); break; }
            if (prev == key) break;
            h = (h + 1) & d.cap_mask;
        }
    }
}
// A2: dense cell of every reference + population count per cell
__global__ void __launch_bounds__(256) cluster_lookup_kernel(ClusterDev d)
{
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < d.nref; s += gridDim.x * blockDim.x)
    {
        unsigned long long key;
        if (!cell_key(d, ref_vertex(d, s), key)) continue;
        unsigned int h = (unsigned int)cluster_hash(key) & d.cap_mask;
        while (d.keys[h] != key) h = (h + 1) & d.cap_mask;
        const unsigned int c = d.vals[h];
        d.ref_cell[s] = c;
        atomicAdd(&d.cell_off[c], 1u);
    }
}
// C: group the references per cell (order inside a cell is fixed by the next kernel)
__global__ void __launch_bounds__(256) cluster_scatter_kernel(ClusterDev d)
{
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < d.nref; s += gridDim.x * blockDim.x)
    {
        const unsigned int c = d.ref_cell[s];
        d.members[d.cell_off[c] + atomicAdd(&d.cell_fill[c], 1u)] = (unsigned int)s;
    }
}
// Segments too long for a per-thread insertion sort (a coarse grid puts thousands of references into one cell) are listed ...
constexpr int kSmallSegment = 64;
__global__ void __launch_bounds__(256) find_big_segments_kernel(const unsigned int *fill, int n_segments, unsigned int *big_list, unsigned int *n_big)
{
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_segments; c += gridDim.x * blockDim.x)
        if (fill[c] > (unsigned int)kSmallSegment) big_list[atomicAdd(n_big, 1u)] = (unsigned int)c;
}
// ... and sorted by one CTA each: bottom-up merge sort, every element finds its place in the merged run by binary search in
// the sibling run (the keys are distinct reference numbers)
__global__ void __launch_bounds__(1024) sort_big_segments_kernel(unsigned int *members, unsigned int *scratch, const unsigned int *off,
                                                                 const unsigned int *fill, const unsigned int *big_list, const unsigned int *n_big)
{
    for (unsigned int b = blockIdx.x; b < *n_big; b += gridDim.x)
    {
        const unsigned int c = big_list[b], n = fill[c];
        unsigned int *src = members + off[c], *dst = scratch + off[c];
        for (unsigned int w = 1; w < n; w <<= 1)
        {
            for (unsigned int i = threadIdx.x; i < n; i += blockDim.x)
            {
                const unsigned int base = i / (2 * w) * (2 * w);
                const unsigned int mid = min(base + w, n), end = min(base + 2 * w, n);
                const unsigned int key = src[i];
                unsigned int lo = i < mid ? mid : base, hi = i < mid ? end : mid; // the sibling run
                while (lo < hi)
                {
                    const unsigned int m = (lo + hi) >> 1;
                    if (src[m] < key) lo = m + 1;
                    else hi = m;
                }
                const unsigned int pos = i < mid ? i + (lo - mid) : (lo - base) + base + (i - mid);
                dst[pos] = key;
            }
            __syncthreads();
            unsigned int *t = src; src = dst; dst = t;
        }
        if (src != members + off[c])
            for (unsigned int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
        __syncthreads();
    }
}

// D: thread = cell.  Members into ascending reference order (the order of the reference's loop), then the sequential float
// sum grid_to_point[cell] += p (MeshSimplification.cpp:611-613) and points[rep] = sum / count (:135).
__global__ void __launch_bounds__(128) cluster_reduce_kernel(ClusterDev d, int n_cells)
{
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += gridDim.x * blockDim.x)
    {
        unsigned int *m = d.members + d.cell_off[c];
        const int n = (int)d.cell_fill[c];
        for (int i = 1; i < n && n <= kSmallSegment; ++i) // longer segments were sorted by sort_big_segments_kernel
        {
            const unsigned int v = m[i];
            int j = i - 1;
            while (j >= 0 && m[j] > v) { m[j + 1] = m[j]; --j; }
            m[j + 1] = v;
        }
        const unsigned int rep = ref_vertex(d, (int)m[0]);
        const float *p0 = d.points + 3 * (size_t)rep;
        float sx = p0[0], sy = p0[1], sz = p0[2];
        for (int i = 1; i < n; ++i)
        {
            const float *p = d.points + 3 * (size_t)ref_vertex(d, (int)m[i]);
            sx = fadd(sx, p[0]); sy = fadd(sy, p[1]); sz = fadd(sz, p[2]);
        }
        const float cnt = (float)n;
        d.cell_rep[c] = rep;
        d.cell_mean[3 * c] = fdiv(sx, cnt); d.cell_mean[3 * c + 1] = fdiv(sy, cnt); d.cell_mean[3 * c + 2] = fdiv(sz, cnt);
    }
}
// E: triangles onto the representatives; a triangle with two corners in one cell is deleted (MeshSimplification.cpp:641-650)
__global__ void __launch_bounds__(256) cluster_triangles_kernel(ClusterDev d, unsigned int *tri_new, unsigned int *keep)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.nt; i += gridDim.x * blockDim.x)
    {
        const unsigned int v1 = d.cell_rep[d.ref_cell[3 * i]], v2 = d.cell_rep[d.ref_cell[3 * i + 1]], v3 = d.cell_rep[d.ref_cell[3 * i + 2]];
        tri_new[3 * i] = v1; tri_new[3 * i + 1] = v2; tri_new[3 * i + 2] = v3;
        keep[i] = (v1 == v2 || v1 == v3 || v2 == v3) ? 0u : 1u;
    }
}
// G: ordered compaction of the kept triangles + marking of the vertices they reference
__global__ void __launch_bounds__(256) cluster_compact_triangles_kernel(const unsigned int *tri_new, const unsigned int *keep_off, int nt, unsigned int n_keep,
                                                                        unsigned int *out_tri, unsigned int *vertex_used)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nt; i += gridDim.x * blockDim.x)
    {
        const unsigned int o = keep_off[i], next = i + 1 < nt ? keep_off[i + 1] : n_keep;
        if (next == o) continue; // deleted
        for (int k = 0; k < 3; ++k)
        {
            const unsigned int v = tri_new[3 * i + k];
            out_tri[3 * (size_t)o + k] = v;
            vertex_used[v] = 1u;
        }
    }
}
// H: every representative moves to the mean of its cell (UpdateMesh, MeshSimplification.cpp:130-136)
__global__ void __launch_bounds__(256) cluster_move_kernel(ClusterDev d, int n_cells, float *points_out)
{
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += gridDim.x * blockDim.x)
    {
        const size_t v = d.cell_rep[c];
        points_out[3 * v] = d.cell_mean[3 * c]; points_out[3 * v + 1] = d.cell_mean[3 * c + 1]; points_out[3 * v + 2] = d.cell_mean[3 * c + 2];
    }
}
// J: CompactMesh (MeshSimplification.cpp:321-337): referenced vertices keep their relative order
__global__ void __launch_bounds__(256) cluster_gather_kernel(const float *points, const float *colors, const unsigned int *vertex_off, int nv,
                                                             unsigned int n_used, float *out_points, float *out_colors)
{
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x)
    {
        const unsigned int o = vertex_off[v], next = v + 1 < nv ? vertex_off[v + 1] : n_used;
        if (next == o) continue;
        for (int k = 0; k < 3; ++k)
        {
            out_points[3 * (size_t)o + k] = points[3 * (size_t)v + k];
            if (colors) out_colors[3 * (size_t)o + k] = colors[3 * (size_t)v + k];
        }
    }
}
__global__ void __launch_bounds__(256) cluster_remap_kernel(unsigned int *out_tri, size_t n, const unsigned int *vertex_off)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out_tri[i] = vertex_off[out_tri[i]];
}

// ---- PointCloud::DownSample (src/Geometry/PointCloud.cpp:145-189) -----------------------------------------------------------
// The same grouping with the points themselves as references: the first point of a cell (input order) opens the output slot,
// the others are added in input order, the slot is divided by the population.  Colours and normals ride along: an attribute
// is summed over the members the cluster kernels left in ascending order.
__global__ void __launch_bounds__(128) downsample_attribute_kernel(ClusterDev d, int n_cells, const float *attr, float *cell_attr)
{
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += gridDim.x * blockDim.x)
    {
        const unsigned int *m = d.members + d.cell_off[c];
        const int n = (int)d.cell_fill[c];
        const float *a0 = attr + 3 * (size_t)m[0];
        float sx = a0[0], sy = a0[1], sz = a0[2];
        for (int i = 1; i < n; ++i)
        {
            const float *a = attr + 3 * (size_t)m[i];
            sx = fadd(sx, a[0]); sy = fadd(sy, a[1]); sz = fadd(sz, a[2]);
        }
        const float cnt = (float)n;
        cell_attr[3 * c] = fdiv(sx, cnt); cell_attr[3 * c + 1] = fdiv(sy, cnt); cell_attr[3 * c + 2] = fdiv(sz, cnt);
    }
}
__global__ void __launch_bounds__(256) downsample_mark_kernel(ClusterDev d, int n_cells, unsigned int *first_flag)
{
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += gridDim.x * blockDim.x) first_flag[d.cell_rep[c]] = 1u;
}
// output slot of a cell = rank of its first point among all first points (the order in which the reference opens slots)
__global__ void __launch_bounds__(256) downsample_place_kernel(ClusterDev d, int n_cells, const unsigned int *slot_of_point, const float *cell_attr,
                                                               float *out)
{
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += gridDim.x * blockDim.x)
    {
        const size_t o = slot_of_point[d.cell_rep[c]];
        out[3 * o] = cell_attr[3 * c]; out[3 * o + 1] = cell_attr[3 * c + 1]; out[3 * o + 2] = cell_attr[3 * c + 2];
    }
}

// ---- ComputeNormals -------------------------------------------------------------------------------------------------
// Eigen Vector3f::normalize(): z = squaredNorm() in Eigen's 3-term order a0 + (a1 + a2); if (z > 0) v /= sqrt(z)
__device__ __forceinline__ void normalize3(float &x, float &y, float &z)
{
    const float n2 = fadd(fmul(x, x), fadd(fmul(y, y), fmul(z, z)));
    if (n2 > 0)
    {
        const float n = __fsqrt_rn(n2);
        x = fdiv(x, n); y = fdiv(y, n); z = fdiv(z, n);
    }
}
__global__ void __launch_bounds__(256) normals_face_kernel(const float *points, const unsigned int *tri, int nt, int nv, float *face_n,
                                                           unsigned int *vertex_cnt, int *bad)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nt; i += gridDim.x * blockDim.x)
    {
        const unsigned int a = tri[3 * i], b = tri[3 * i + 1], c = tri[3 * i + 2];
        if (a >= (unsigned int)nv || b >= (unsigned int)nv || c >= (unsigned int)nv) { *bad = 1; continue; }
        const float *p1 = points + 3 * (size_t)a, *p2 = points + 3 * (size_t)b, *p3 = points + 3 * (size_t)c;
        const float ax = fsub(p2[0], p1[0]), ay = fsub(p2[1], p1[1]), az = fsub(p2[2], p1[2]);
        const float bx = fsub(p3[0], p1[0]), by = fsub(p3[1], p1[1]), bz = fsub(p3[2], p1[2]);
        float nx = fsub(fmul(ay, bz), fmul(az, by)), ny = fsub(fmul(az, bx), fmul(ax, bz)), nz = fsub(fmul(ax, by), fmul(ay, bx));
        normalize3(nx, ny, nz);
        face_n[3 * i] = nx; face_n[3 * i + 1] = ny; face_n[3 * i + 2] = nz;
        atomicAdd(&vertex_cnt[a], 1u); atomicAdd(&vertex_cnt[b], 1u); atomicAdd(&vertex_cnt[c], 1u);
    }
}
__global__ void __launch_bounds__(256) normals_scatter_kernel(const unsigned int *tri, int nt, const unsigned int *vertex_off, unsigned int *vertex_fill,
                                                              unsigned int *members)
{
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < 3 * nt; s += gridDim.x * blockDim.x)
    {
        const unsigned int v = tri[s];
        members[vertex_off[v] + atomicAdd(&vertex_fill[v], 1u)] = (unsigned int)s;
    }
}
// thread = vertex: faces in reference order (UpdateReferences, MeshSimplification.cpp:531-541), sequential sum, normalise
__global__ void __launch_bounds__(128) normals_vertex_kernel(const float *face_n, const unsigned int *vertex_off, const unsigned int *vertex_fill,
                                                             unsigned int *members, int nv, float *normals)
{
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x)
    {
        unsigned int *m = members + vertex_off[v];
        const int n = (int)vertex_fill[v];
        for (int i = 1; i < n && n <= kSmallSegment; ++i) // longer segments were sorted by sort_big_segments_kernel
        {
            const unsigned int x = m[i];
            int j = i - 1;
            while (j >= 0 && m[j] > x) { m[j + 1] = m[j]; --j; }
            m[j + 1] = x;
        }
        float sx = 0.0f, sy = 0.0f, sz = 0.0f; // vnormal.setZero()
        for (int i = 0; i < n; ++i)
        {
            const float *f = face_n + 3 * (size_t)(m[i] / 3);
            sx = fadd(sx, f[0]); sy = fadd(sy, f[1]); sz = fadd(sz, f[2]);
        }
        normalize3(sx, sy, sz);
        normals[3 * (size_t)v] = sx; normals[3 * (size_t)v + 1] = sy; normals[3 * (size_t)v + 2] = sz;
    }
}

// bump allocator over one cudaMalloc: every scratch array of a call lives in it
struct Arena
{
    char *base = nullptr;
    size_t used = 0, cap = 0;
    template <typename T> T *take(size_t n)
    {
        used = (used + 255) & ~(size_t)255;
        T *p = (T *)(base + used);
        used += n * sizeof(T);
        return p;
    }
};

static int grid_for(size_t n, int sm_count, int per_sm = 8)
{
    const size_t need = (n + 255) / 256;
    const size_t cap = (size_t)sm_count * per_sm;
    return (int)(need < cap ? (need ? need : 1) : cap);
}

// device-resident core of ClusteringSimplify; inputs and outputs are device pointers
int clustering_simplify_device(int sm_count, cudaStream_t s, const float *d_points, const float *d_colors, size_t nv, const unsigned int *d_tri,
                               size_t nt, float grid_len, float **o_points, float **o_colors, unsigned int **o_tri, size_t *o_nv, size_t *o_nt)
{
    *o_points = *o_colors = nullptr; *o_tri = nullptr; *o_nv = *o_nt = 0;
    if (nt == 0 || nv == 0) return OPB_OK;
    if (nt > 0x2AAAAAAAu || nv > 0x7FFFFFF0u) { set_error("mesh too large for 32-bit reference indices"); return OPB_ERR_CAPACITY; }
    const size_t nref = 3 * nt;
    size_t cap = 1024;
    while (cap < 2 * nref) cap <<= 1;
    Arena A;
    A.cap = cap * 12 + nref * 4 * 2 + nref * 4 * 3 /* cell arrays, worst case one cell per reference */ + nref * 4 * 4 + nt * 4 * 5 + nv * 4 * 5 +
            nv * 12 + nref * 4 * 2 /* big-segment list and sort scratch */ + 64 * 256;
    OPB_CUDA(cudaMalloc(&A.base, A.cap));
    ClusterDev d;
    d.points = d_points; d.tri = d_tri; d.nv = (int)nv; d.nt = (int)nt; d.nref = (int)nref; d.grid_len = grid_len;
    d.keys = A.take<unsigned long long>(cap);
    d.vals = A.take<unsigned int>(cap);
    d.cap_mask = (unsigned int)(cap - 1);
    unsigned int *counters = A.take<unsigned int>(8); // n_cells, bad, n_keep, n_used
    d.n_cells = counters; d.bad = (int *)(counters + 1);
    d.ref_cell = A.take<unsigned int>(nref);
    d.members = A.take<unsigned int>(nref);
    d.cell_off = A.take<unsigned int>(nref + 1);
    d.cell_fill = A.take<unsigned int>(nref);
    d.cell_rep = A.take<unsigned int>(nref);
    d.cell_mean = A.take<float>(3 * nref);
    unsigned int *tri_new = A.take<unsigned int>(nref), *keep = A.take<unsigned int>(nt + 1), *vertex_used = A.take<unsigned int>(nv + 1);
    unsigned int *big_list = A.take<unsigned int>(nref / kSmallSegment + 1), *sort_scratch = A.take<unsigned int>(nref);
    float *points2 = A.take<float>(3 * nv);
    int rc = OPB_OK;
    float *out_p = nullptr, *out_c = nullptr;
    unsigned int *out_t = nullptr;
    do
    {
#define OPB_TRY(expr) if ((expr) != cudaSuccess) { set_error("%s failed: %s", #expr, cudaGetErrorString(cudaGetLastError())); rc = OPB_ERR_CUDA; break; }
        if (A.used > A.cap) { set_error("internal: scratch arena too small"); rc = OPB_ERR_CAPACITY; break; }
        OPB_TRY(cudaMemsetAsync(d.keys, 0xFF, cap * sizeof(unsigned long long), s));
        OPB_TRY(cudaMemsetAsync(counters, 0, 8 * sizeof(unsigned int), s));
        OPB_TRY(cudaMemsetAsync(d.cell_off, 0, (nref + 1) * sizeof(unsigned int), s));
        OPB_TRY(cudaMemsetAsync(d.cell_fill, 0, nref * sizeof(unsigned int), s));
        OPB_TRY(cudaMemsetAsync(vertex_used, 0, (nv + 1) * sizeof(unsigned int), s));
        cluster_insert_kernel<<<grid_for(nref, sm_count), 256, 0, s>>>(d);
        unsigned int h_counters[8];
        OPB_TRY(cudaMemcpyAsync(h_counters, counters, sizeof(h_counters), cudaMemcpyDeviceToHost, s));
        OPB_TRY(cudaStreamSynchronize(s));
        if (h_counters[1]) { set_error("mesh has a triangle naming a missing vertex, or a vertex outside +-2^20 grid cells"); rc = OPB_ERR_INVALID; break; }
        const int n_cells = (int)h_counters[0];
        cluster_lookup_kernel<<<grid_for(nref, sm_count), 256, 0, s>>>(d);
        post_scan_kernel<<<1, 1024, 0, s>>>(d.cell_off, n_cells, counters + 4);
        cluster_scatter_kernel<<<grid_for(nref, sm_count), 256, 0, s>>>(d);
        find_big_segments_kernel<<<grid_for(n_cells, sm_count), 256, 0, s>>>(d.cell_fill, n_cells, big_list, counters + 5);
        sort_big_segments_kernel<<<sm_count * 2, 1024, 0, s>>>(d.members, sort_scratch, d.cell_off, d.cell_fill, big_list, counters + 5);
        cluster_reduce_kernel<<<grid_for((size_t)n_cells * 2, sm_count, 16), 128, 0, s>>>(d, n_cells);
        cluster_triangles_kernel<<<grid_for(nt, sm_count), 256, 0, s>>>(d, tri_new, keep);
        post_scan_kernel<<<1, 1024, 0, s>>>(keep, (int)nt, counters + 2);
        OPB_TRY(cudaMemcpyAsync(h_counters, counters, sizeof(h_counters), cudaMemcpyDeviceToHost, s));
        OPB_TRY(cudaStreamSynchronize(s));
        const unsigned int n_keep = h_counters[2];
        OPB_TRY(cudaMalloc(&out_t, (size_t)(n_keep ? n_keep : 1) * 3 * sizeof(unsigned int)));
        cluster_compact_triangles_kernel<<<grid_for(nt, sm_count), 256, 0, s>>>(tri_new, keep, (int)nt, n_keep, out_t, vertex_used);
        OPB_TRY(cudaMemcpyAsync(points2, d_points, nv * 3 * sizeof(float), cudaMemcpyDeviceToDevice, s));
        cluster_move_kernel<<<grid_for(n_cells, sm_count), 256, 0, s>>>(d, n_cells, points2);
        post_scan_kernel<<<1, 1024, 0, s>>>(vertex_used, (int)nv, counters + 3);
        OPB_TRY(cudaMemcpyAsync(h_counters, counters, sizeof(h_counters), cudaMemcpyDeviceToHost, s));
        OPB_TRY(cudaStreamSynchronize(s));
        const unsigned int n_used = h_counters[3];
        OPB_TRY(cudaMalloc(&out_p, (size_t)(n_used ? n_used : 1) * 3 * sizeof(float)));
        if (d_colors) OPB_TRY(cudaMalloc(&out_c, (size_t)(n_used ? n_used : 1) * 3 * sizeof(float)));
        cluster_gather_kernel<<<grid_for(nv, sm_count), 256, 0, s>>>(points2, d_colors, vertex_used, (int)nv, n_used, out_p, out_c);
        cluster_remap_kernel<<<grid_for((size_t)n_keep * 3, sm_count), 256, 0, s>>>(out_t, (size_t)n_keep * 3, vertex_used);
        OPB_TRY(cudaGetLastError());
        OPB_TRY(cudaStreamSynchronize(s));
        *o_points = out_p; *o_colors = out_c; *o_tri = out_t; *o_nv = n_used; *o_nt = n_keep;
        out_p = out_c = nullptr; out_t = nullptr;
#undef OPB_TRY
    } while (0);
    cudaFree(A.base); cudaFree(out_p); cudaFree(out_c); cudaFree(out_t);
    return rc;
}

// copies a device result into malloc'ed host buffers (released by the caller with opb_free) and frees the device copies
int mesh_result_to_host(cudaStream_t s, float *d_points, float *d_colors, unsigned int *d_tri, size_t nv, size_t nt, float **points, float **colors,
                        uint32_t **triangles)
{
    *points = (float *)malloc((nv ? nv : 1) * 3 * sizeof(float));
    *triangles = (uint32_t *)malloc((nt ? nt : 1) * 3 * sizeof(uint32_t));
    if (colors) *colors = d_colors ? (float *)malloc((nv ? nv : 1) * 3 * sizeof(float)) : nullptr;
    cudaError_t e = cudaSuccess;
    if (!*points || !*triangles || (colors && d_colors && !*colors)) e = cudaErrorMemoryAllocation;
    if (e == cudaSuccess && nv) e = cudaMemcpyAsync(*points, d_points, nv * 3 * sizeof(float), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && nv && colors && d_colors) e = cudaMemcpyAsync(*colors, d_colors, nv * 3 * sizeof(float), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && nt) e = cudaMemcpyAsync(*triangles, d_tri, nt * 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d_points); cudaFree(d_colors); cudaFree(d_tri);
    if (e != cudaSuccess)
    {
        free(*points); free(*triangles);
        if (colors) free(*colors);
        *points = nullptr; *triangles = nullptr;
        if (colors) *colors = nullptr;
        set_error("mesh download failed: %s", cudaGetErrorString(e));
        return OPB_ERR_CUDA;
    }
    return OPB_OK;
}
} // namespace opb

using namespace opb;

static int post_device(int device, int *sm_count)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
        cudaGetLastError();
        set_error("no CUDA device: onepiece_b200 has no CPU path");
        return OPB_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (%d devices)", device, ndev); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(device));
    OPB_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, device));
    return OPB_OK;
}

extern "C"
{
int opb_mesh_clustering_simplify(int device, const float *points, const float *colors, size_t nv, const uint32_t *triangles, size_t nt,
                                 float grid_len, float **out_points, float **out_colors, uint32_t **out_triangles, size_t *out_nv, size_t *out_nt)
{
    if (!points || !triangles || !out_points || !out_triangles || !out_nv || !out_nt || (colors && !out_colors))
    {
        set_error("NULL argument");
        return OPB_ERR_INVALID;
    }
    *out_points = nullptr; *out_triangles = nullptr; *out_nv = *out_nt = 0;
    if (out_colors) *out_colors = nullptr;
    if (!(grid_len > 0))
    {
        // MeshSimplification.cpp:584-588: message, mesh returned unchanged
        set_error("[ClusteringMeshSimplification]::[ERROR]::Grid length cannot be less than 0.");
        return OPB_ERR_INVALID;
    }
    int sm = 0;
    int rc = post_device(device, &sm);
    if (rc) return rc;
    float *d_p = nullptr, *d_c = nullptr;
    unsigned int *d_t = nullptr;
    cudaError_t e = cudaMalloc(&d_p, (nv ? nv : 1) * 3 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_t, (nt ? nt : 1) * 3 * sizeof(unsigned int));
    if (e == cudaSuccess && colors) e = cudaMalloc(&d_c, (nv ? nv : 1) * 3 * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(d_p, points, nv * 3 * sizeof(float), cudaMemcpyDefault);
    if (e == cudaSuccess) e = cudaMemcpy(d_t, triangles, nt * 3 * sizeof(unsigned int), cudaMemcpyDefault);
    if (e == cudaSuccess && colors) e = cudaMemcpy(d_c, colors, nv * 3 * sizeof(float), cudaMemcpyDefault);
    float *r_p = nullptr, *r_c = nullptr;
    unsigned int *r_t = nullptr;
    size_t r_nv = 0, r_nt = 0;
    if (e != cudaSuccess) { set_error("mesh upload failed: %s", cudaGetErrorString(e)); rc = OPB_ERR_CUDA; }
    else rc = clustering_simplify_device(sm, nullptr, d_p, d_c, nv, d_t, nt, grid_len, &r_p, &r_c, &r_t, &r_nv, &r_nt);
    cudaFree(d_p); cudaFree(d_c); cudaFree(d_t);
    if (rc) return rc;
    rc = mesh_result_to_host(nullptr, r_p, r_c, r_t, r_nv, r_nt, out_points, out_colors, out_triangles);
    if (rc) return rc;
    *out_nv = r_nv; *out_nt = r_nt;
    return OPB_OK;
}

int opb_pointcloud_downsample(int device, const float *points, const float *colors, const float *normals, size_t n, float grid_len,
                              float **out_points, float **out_colors, float **out_normals, size_t *out_n)
{
    if (!points || !out_points || !out_n || (colors && !out_colors) || (normals && !out_normals)) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    *out_points = nullptr; *out_n = 0;
    if (out_colors) *out_colors = nullptr;
    if (out_normals) *out_normals = nullptr;
    if (!(grid_len > 0)) { set_error("grid_len must be > 0"); return OPB_ERR_INVALID; }
    if (n == 0) return OPB_OK;
    if (n > 0x7FFFFFF0u) { set_error("point cloud too large for 32-bit indices"); return OPB_ERR_CAPACITY; }
    int sm = 0;
    int rc = post_device(device, &sm);
    if (rc) return rc;
    size_t cap = 1024;
    while (cap < 2 * n) cap <<= 1;
    const int n_attr = 1 + (colors ? 1 : 0) + (normals ? 1 : 0);
    Arena A;
    // inputs (n_attr x 12 n), cell_mean / cell_attr / d_out (3 x 12 n), eight index arrays (4 n each, big_list n / 16), the hash table
    A.cap = cap * 12 + n * 4 * 10 + n * 12 * (size_t)(n_attr + 4) + 64 * 256;
    OPB_CUDA(cudaMalloc(&A.base, A.cap));
    ClusterDev d;
    float *d_in[3] = {A.take<float>(3 * n), colors ? A.take<float>(3 * n) : nullptr, normals ? A.take<float>(3 * n) : nullptr};
    const float *h_in[3] = {points, colors, normals};
    d.points = d_in[0]; d.tri = nullptr; d.nv = (int)n; d.nt = 0; d.nref = (int)n; d.grid_len = grid_len;
    d.keys = A.take<unsigned long long>(cap);
    d.vals = A.take<unsigned int>(cap);
    d.cap_mask = (unsigned int)(cap - 1);
    unsigned int *counters = A.take<unsigned int>(8);
    d.n_cells = counters; d.bad = (int *)(counters + 1);
    d.ref_cell = A.take<unsigned int>(n);
    d.members = A.take<unsigned int>(n);
    d.cell_off = A.take<unsigned int>(n + 1);
    d.cell_fill = A.take<unsigned int>(n);
    d.cell_rep = A.take<unsigned int>(n);
    d.cell_mean = A.take<float>(3 * n);
    unsigned int *first_flag = A.take<unsigned int>(n + 1);
    float *cell_attr = A.take<float>(3 * n), *d_out = A.take<float>(3 * n);
    unsigned int *big_list = A.take<unsigned int>(n / kSmallSegment + 1), *sort_scratch = A.take<unsigned int>(n);
    cudaStream_t s = nullptr;
    float *h_out[3] = {nullptr, nullptr, nullptr};
    do
    {
#define OPB_TRY(expr) if ((expr) != cudaSuccess) { set_error("%s failed: %s", #expr, cudaGetErrorString(cudaGetLastError())); rc = OPB_ERR_CUDA; break; }
        if (A.used > A.cap) { set_error("internal: scratch arena too small"); rc = OPB_ERR_CAPACITY; break; }
        for (int a = 0; a < 3 && rc == OPB_OK; ++a)
            if (d_in[a] && cudaMemcpyAsync(d_in[a], h_in[a], n * 3 * sizeof(float), cudaMemcpyDefault, s) != cudaSuccess)
            {
                set_error("point cloud upload failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = OPB_ERR_CUDA;
            }
        if (rc) break;
        OPB_TRY(cudaMemsetAsync(d.keys, 0xFF, cap * sizeof(unsigned long long), s));
        OPB_TRY(cudaMemsetAsync(counters, 0, 8 * sizeof(unsigned int), s));
        OPB_TRY(cudaMemsetAsync(d.cell_off, 0, (n + 1) * sizeof(unsigned int), s));
        OPB_TRY(cudaMemsetAsync(d.cell_fill, 0, n * sizeof(unsigned int), s));
        OPB_TRY(cudaMemsetAsync(first_flag, 0, (n + 1) * sizeof(unsigned int), s));
        cluster_insert_kernel<<<grid_for(n, sm), 256, 0, s>>>(d);
        unsigned int h_counters[8];
        OPB_TRY(cudaMemcpyAsync(h_counters, counters, sizeof(h_counters), cudaMemcpyDeviceToHost, s));
        OPB_TRY(cudaStreamSynchronize(s));
        if (h_counters[1]) { set_error("point cloud has a point outside +-2^20 grid cells (or NaN)"); rc = OPB_ERR_INVALID; break; }
        const int n_cells = (int)h_counters[0];
        cluster_lookup_kernel<<<grid_for(n, sm), 256, 0, s>>>(d);
        post_scan_kernel<<<1, 1024, 0, s>>>(d.cell_off, n_cells, counters + 4);
        cluster_scatter_kernel<<<grid_for(n, sm), 256, 0, s>>>(d);
        find_big_segments_kernel<<<grid_for(n_cells, sm), 256, 0, s>>>(d.cell_fill, n_cells, big_list, counters + 5);
        sort_big_segments_kernel<<<sm * 2, 1024, 0, s>>>(d.members, sort_scratch, d.cell_off, d.cell_fill, big_list, counters + 5);
        cluster_reduce_kernel<<<grid_for((size_t)n_cells * 2, sm, 16), 128, 0, s>>>(d, n_cells); // sorts the members; mean position
        downsample_mark_kernel<<<grid_for(n_cells, sm), 256, 0, s>>>(d, n_cells, first_flag);
        post_scan_kernel<<<1, 1024, 0, s>>>(first_flag, (int)n, counters + 2);
        OPB_TRY(cudaGetLastError());
        for (int a = 0; a < 3 && rc == OPB_OK; ++a)
        {
            if (!d_in[a]) continue;
            const float *means = d.cell_mean;
            if (a > 0)
            {
                downsample_attribute_kernel<<<grid_for((size_t)n_cells * 2, sm, 16), 128, 0, s>>>(d, n_cells, d_in[a], cell_attr);
                means = cell_attr;
            }
            downsample_place_kernel<<<grid_for(n_cells, sm), 256, 0, s>>>(d, n_cells, first_flag, means, d_out);
            h_out[a] = (float *)malloc((size_t)n_cells * 3 * sizeof(float) + 4);
            if (!h_out[a]) { set_error("host allocation failed"); rc = OPB_ERR_CAPACITY; break; }
            if (cudaMemcpyAsync(h_out[a], d_out, (size_t)n_cells * 3 * sizeof(float), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
                cudaStreamSynchronize(s) != cudaSuccess)
            {
                set_error("down-sampling failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = OPB_ERR_CUDA;
            }
        }
        if (rc) break;
        *out_points = h_out[0];
        if (colors) *out_colors = h_out[1];
        if (normals) *out_normals = h_out[2];
        *out_n = (size_t)n_cells;
        h_out[0] = h_out[1] = h_out[2] = nullptr;
#undef OPB_TRY
    } while (0);
    cudaFree(A.base);
    for (int a = 0; a < 3; ++a) free(h_out[a]);
    return rc;
}

int opb_mesh_compute_normals(int device, const float *points, size_t nv, const uint32_t *triangles, size_t nt, float *normals)
{
    if (!points || !triangles || !normals) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    if (nv == 0) return OPB_OK;
    if (nt > 0x2AAAAAAAu || nv > 0x7FFFFFF0u) { set_error("mesh too large for 32-bit reference indices"); return OPB_ERR_CAPACITY; }
    int sm = 0;
    int rc = post_device(device, &sm);
    if (rc) return rc;
    const size_t nref = 3 * nt;
    Arena A;
    A.cap = nv * 12 * 2 + nref * 4 * 4 + nt * 12 + nv * 4 * 2 + 64 * 256;
    OPB_CUDA(cudaMalloc(&A.base, A.cap));
    float *d_p = A.take<float>(3 * nv), *d_n = A.take<float>(3 * nv), *face_n = A.take<float>(3 * (nt ? nt : 1));
    unsigned int *d_t = A.take<unsigned int>(nref ? nref : 1), *members = A.take<unsigned int>(nref ? nref : 1);
    unsigned int *v_off = A.take<unsigned int>(nv + 1), *v_fill = A.take<unsigned int>(nv + 1), *counters = A.take<unsigned int>(8);
    unsigned int *big_list = A.take<unsigned int>(nref / kSmallSegment + 2), *sort_scratch = A.take<unsigned int>(nref ? nref : 1);
    cudaStream_t s = nullptr;
    do
    {
#define OPB_TRY(expr) if ((expr) != cudaSuccess) { set_error("%s failed: %s", #expr, cudaGetErrorString(cudaGetLastError())); rc = OPB_ERR_CUDA; break; }
        OPB_TRY(cudaMemcpyAsync(d_p, points, nv * 3 * sizeof(float), cudaMemcpyDefault, s));
        OPB_TRY(cudaMemcpyAsync(d_t, triangles, nref * sizeof(unsigned int), cudaMemcpyDefault, s));
        OPB_TRY(cudaMemsetAsync(v_off, 0, (nv + 1) * sizeof(unsigned int), s));
        OPB_TRY(cudaMemsetAsync(v_fill, 0, (nv + 1) * sizeof(unsigned int), s));
        OPB_TRY(cudaMemsetAsync(counters, 0, 8 * sizeof(unsigned int), s));
        if (nt) normals_face_kernel<<<grid_for(nt, sm), 256, 0, s>>>(d_p, d_t, (int)nt, (int)nv, face_n, v_off, (int *)(counters + 1));
        unsigned int h_counters[8];
        OPB_TRY(cudaMemcpyAsync(h_counters, counters, sizeof(h_counters), cudaMemcpyDeviceToHost, s));
        OPB_TRY(cudaStreamSynchronize(s));
        if (h_counters[1]) { set_error("mesh has a triangle naming a missing vertex"); rc = OPB_ERR_INVALID; break; }
        post_scan_kernel<<<1, 1024, 0, s>>>(v_off, (int)nv, counters);
        if (nt) normals_scatter_kernel<<<grid_for(nref, sm), 256, 0, s>>>(d_t, (int)nt, v_off, v_fill, members);
        find_big_segments_kernel<<<grid_for(nv, sm), 256, 0, s>>>(v_fill, (int)nv, big_list, counters + 5);
        sort_big_segments_kernel<<<sm * 2, 1024, 0, s>>>(members, sort_scratch, v_off, v_fill, big_list, counters + 5);
        normals_vertex_kernel<<<grid_for(nv * 2, sm, 16), 128, 0, s>>>(face_n, v_off, v_fill, members, (int)nv, d_n);
        OPB_TRY(cudaGetLastError());
        OPB_TRY(cudaMemcpyAsync(normals, d_n, nv * 3 * sizeof(float), cudaMemcpyDefault, s));
        OPB_TRY(cudaStreamSynchronize(s));
#undef OPB_TRY
    } while (0);
    cudaFree(A.base);
    return rc;
}
} // extern "C"
