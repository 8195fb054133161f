// K7-K9: ICP (point-to-plane and point-to-point) on the device.
//
// Reference path rebuilt here (file:line relative to the reference tree):
//   registration::PointToPlane        src/Registration/ICP.cpp:146-224
//   registration::PointToPoint        src/Registration/ICP.cpp:31-107
//   CountInliers                      src/Registration/ICP.cpp:9-30
//   EstimateRigidTransformationPointToPlane  ICP.cpp:108-144   (6x6 J^T J, JacobiSVD solve, Se3ToSE3)
//   geometry::TransformPoints         src/Geometry/Geometry.cpp:19-27
//   geometry::EstimateRigidTransformation    src/Geometry/Geometry.cpp:107-151 (Kabsch; also result.T)
//   KDTree<>::KnnSearch(k=1)          src/Geometry/KDTree.h:177-196 (nanoflann, exact, L2_Simple distance)
//
// Design.  The k-d tree is replaced by a uniform grid over the target cloud (counting sort by cell, points
// stored cell-contiguous as float4 {x,y,z,index}); a query walks Chebyshev rings of cells around its home cell
// and stops as soon as the best squared distance is provably minimal, so the answer is the exact nearest
// neighbour (distance formula and summation order of nanoflann's L2_Simple_Adaptor).  The walk is capped at the
// inlier threshold: a neighbour farther than that can never pass CountInliers, so it is reported as "none".
// An ICP iteration is three launches.  icp_certify_kernel: which queries provably keep the nearest neighbour of their last
// full search (the pose barely moves once the iteration has converged) -- they are answered from the stored index, the rest
// go on a work list.  icp_search_kernel: transform + exact nearest neighbour + certificate for the queries on the list.
// icp_accumulate_kernel: inlier test + Jacobian row +
// the 30-scalar reduction (per-thread fp64 accumulators, warp shuffles, fixed-order CTA partials -> deterministic) and, in
// the last CTA to finish, the 6x6 solve, the SE(3) exponential and the pose update -- the host is not involved until the end.
#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/onepiece_b200.h"
#include "opb_cloud_host.h"
#include "opb_common.cuh"
#include "opb_fitplane.cuh"
#include "opb_linalg.h"

namespace opb
{
constexpr int kIcpThreads = 256;
constexpr int kPacket = 32;             // doubles per reduction packet (29 used for point-to-plane, 17 for point-to-point)
constexpr unsigned int kMaxCells = 1u << 26;

struct IcpGrid
{
    float origin[3];
    float h, inv_h;
    int dim[3];
};


struct IcpState // device-resident solver state
{
    float T[16];        // current source->target transform, column-major float (start_T)
    float T_prev[16];   // the pose the last solving pass searched under (the closing CountInliers re-tests ITS neighbours)
    double packet[kPacket];
    int iteration;
    IcpGrid grid;
    float bbox_lo[3], bbox_hi[3];
    unsigned int bbox_enc[6];
    unsigned long long n_inliers;
    double sum_error;
    unsigned int blocks_done;
    unsigned long long n_inliers_local; // this rank's share when the source is split across ranks (== n_inliers otherwise)
    unsigned int wl_count;              // queries of the current pass whose nearest neighbour could not be certified
    unsigned long long searched_total;  // full searches done over the whole call (profiling)
    unsigned int searched_per_pass[64]; // ... and per pass (the first 64)
    unsigned int pass;
    // per-pass time stamps of the persistent loop (globaltimer ns), the first kStampPasses passes.  CTA 0: [0] pass start, [1] own
    // points certified / searched, [2] own points accumulated and partial published, [3] released into the next pass.  The CTA
    // that finished last: [4] knows it is last, [5] partials summed (+ exchanged), [6] solved and pose updated.
    unsigned long long stamps[48][8];
};
constexpr int kStampPasses = 48;
__device__ __forceinline__ void icp_stamp(IcpState *st, unsigned int pass, int slot)
{
    if (pass < (unsigned int)kStampPasses) st->stamps[pass][slot] = global_timer_ns();
}

// ---------------------------------------------------------------------------------------------------------
// cross-GPU exchange of the reduction packet: when the source points of one registration are split
// across ranks, every rank reduces its share to the same 30-scalar packet and the packets are summed across ranks INSIDE the
// reduction kernel's tail -- the last CTA stores its packet into every peer's mailbox over NVLink (plain peer stores into
// cudaIpc-mapped memory), raises a flag there, waits for the flags of all ranks in its own mailbox, adds the packets in rank
// order (so every rank gets the bit-identical sum and solves the identical 6x6 system redundantly) and goes on to the solve.
// No NCCL call, no host round trip: the 31 iterations of a call stay one uninterrupted stream of launches.
// ---------------------------------------------------------------------------------------------------------
constexpr int kMaxRanks = 16;
struct IcpMailbox
{
    unsigned long long flag[2][kMaxRanks];   // [parity of the exchange number][writer rank] = exchange number
    double data[2][kMaxRanks][kPacket];      // [parity][writer rank][component]
    unsigned long long epoch;                // exchanges completed by the owner of this mailbox (local use only)
    int error;                               // set when a peer did not answer within the time limit
};
struct IcpComm
{
    IcpMailbox *box[kMaxRanks]; // box[rank] is this rank's own mailbox, the others are peer mappings
    int rank, world;
};


// Sum of packet[0..n) over all ranks, in rank order; called by every thread of ONE CTA (>= 32 threads) on every rank the same
// number of times.  Two mailbox halves alternate: a rank can be at most one exchange ahead of a peer, because finishing an
// exchange needs the peer's packet of that exchange.
__device__ void comm_allreduce(const IcpComm &cm, double *packet, int n)
{
    if (cm.world <= 1) return;
    IcpMailbox *mine = cm.box[cm.rank];
    __shared__ unsigned long long s_epoch;
    if (threadIdx.x == 0) s_epoch = mine->epoch + 1;
    __syncthreads();
    const unsigned long long ep = s_epoch;
    const int par = (int)(ep & 1);
    for (int idx = threadIdx.x; idx < cm.world * n; idx += blockDim.x)
    {
        const int r = idx / n, k = idx - r * n;
        cm.box[r]->data[par][cm.rank][k] = packet[k];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < cm.world)
    {
        *(volatile unsigned long long *)&cm.box[threadIdx.x]->flag[par][cm.rank] = ep;
        volatile unsigned long long *f = &mine->flag[par][threadIdx.x];
        const unsigned long long t0 = global_timer_ns();
        while (*f < ep) // 4 s: a peer died or never made the call; once that happened the call is lost, do not wait again
            if (*(volatile int *)&mine->error || global_timer_ns() - t0 > 4000000000ull) { mine->error = 1; break; }
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < n)
    {
        double s = 0.0;
        for (int r = 0; r < cm.world; ++r) s += *(volatile double *)&mine->data[par][r][threadIdx.x];
        packet[threadIdx.x] = s;
    }
    if (threadIdx.x == 0) mine->epoch = ep;
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------
// grid construction
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void icp_bbox_kernel_body(const float *pts, int n, IcpState *st)
{
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
            const float v = pts[3 * i + a];
            if (v == v && fabsf(v) < FLT_MAX) { lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); }
        }
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if ((threadIdx.x & 31) == 0)
        {
            if (lo[a] != FLT_MAX) atomicMin(&st->bbox_enc[a], float_to_ordered(lo[a]));
            if (hi[a] != -FLT_MAX) atomicMax(&st->bbox_enc[3 + a], float_to_ordered(hi[a]));
        }
    }
}
__global__ void icp_bbox_kernel(const float *pts, int n, IcpState *st) { icp_bbox_kernel_body(pts, n, st); }

// one thread: cell size such that the grid has at most kMaxCells cells and about two points per occupied
// cell for surface-like clouds
__device__ __forceinline__ void icp_grid_setup_kernel_body(IcpState *st, int n, float min_cell, unsigned int max_cells)
{
    float ext[3];
    for (int a = 0; a < 3; ++a)
    {
        float lo = ordered_to_float(st->bbox_enc[a]), hi = ordered_to_float(st->bbox_enc[3 + a]);
        if (!(hi >= lo)) { lo = 0.0f; hi = 0.0f; }
        st->bbox_lo[a] = lo;
        st->bbox_hi[a] = hi;
        ext[a] = hi - lo;
    }
    const float emax = fmaxf(ext[0], fmaxf(ext[1], ext[2]));
    // surface heuristic: area ~ product of the two largest extents, spacing ~ sqrt(area / n)
    const float emin = fminf(ext[0], fminf(ext[1], ext[2]));
    const float emid = ext[0] + ext[1] + ext[2] - emax - emin;
    float h = 2.0f * sqrtf(fmaxf(emax * emid, 1e-12f) / (float)(n > 0 ? n : 1));
    h = fmaxf(h, min_cell);
    h = fmaxf(h, emax * 1e-6f + 1e-9f);
    for (;;)
    {
        double cells = 1.0;
        for (int a = 0; a < 3; ++a) cells *= floor((double)ext[a] / h) + 1.0;
        if (cells <= (double)max_cells) break;
        h *= 1.26f;
    }
    st->grid.h = h;
    st->grid.inv_h = 1.0f / h;
    for (int a = 0; a < 3; ++a)
    {
        st->grid.origin[a] = st->bbox_lo[a];
        st->grid.dim[a] = (int)floorf(ext[a] / h) + 1;
    }
}
__global__ void icp_grid_setup_kernel(IcpState *st, int n, float min_cell, unsigned int max_cells) { icp_grid_setup_kernel_body(st, n, min_cell, max_cells); }

__device__ __forceinline__ int cell_coord(float v, float origin, float inv_h, int dim)
{
    const int c = (int)floorf((v - origin) * inv_h);
    return c < 0 ? 0 : (c >= dim ? dim - 1 : c);
}
__device__ __forceinline__ unsigned int cell_of(const IcpGrid &g, float x, float y, float z)
{
    const int cx = cell_coord(x, g.origin[0], g.inv_h, g.dim[0]);
    const int cy = cell_coord(y, g.origin[1], g.inv_h, g.dim[1]);
    const int cz = cell_coord(z, g.origin[2], g.inv_h, g.dim[2]);
    return (unsigned int)cx + (unsigned int)g.dim[0] * ((unsigned int)cy + (unsigned int)g.dim[1] * (unsigned int)cz);
}

__device__ __forceinline__ void icp_count_kernel_body(const float *pts, int n, const IcpState *st, unsigned int *cell_count, unsigned int *point_cell)
{
    const IcpGrid g = st->grid;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
        unsigned int c = 0xFFFFFFFFu; // non-finite points are left out of the grid (they can never be nearest)
        if (x == x && y == y && z == z && fabsf(x) < FLT_MAX && fabsf(y) < FLT_MAX && fabsf(z) < FLT_MAX)
        {
            c = cell_of(g, x, y, z);
            atomicAdd(&cell_count[c], 1u);
        }
        point_cell[i] = c;
    }
}
__global__ void icp_count_kernel(const float *pts, int n, const IcpState *st, unsigned int *cell_count, unsigned int *point_cell) { icp_count_kernel_body(pts, n, st, cell_count, point_cell); }

// exclusive scan of the cell counts -> cell_start (n_cells + 1 entries), three phases over tiles of 8192 cells:
// tile sums, scan of the tile sums by one CTA, per-tile scan with the tile offset
constexpr unsigned int kScanTile = 1024u * 8u;
constexpr unsigned int kMaxTiles = kMaxCells / kScanTile + 1;

__device__ __forceinline__ unsigned int icp_n_cells(const IcpState *st)
{
    return (unsigned int)st->grid.dim[0] * (unsigned int)st->grid.dim[1] * (unsigned int)st->grid.dim[2];
}
__device__ __forceinline__ void icp_clear_kernel_body(const IcpState *st, unsigned int *cell_count)
{
    const unsigned int n = icp_n_cells(st) + 1;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) cell_count[i] = 0u;
}
__global__ void icp_clear_kernel(const IcpState *st, unsigned int *cell_count) { icp_clear_kernel_body(st, cell_count); }
// block-wide inclusive scan helper over 1024 threads; returns the inclusive value, total in *total
__device__ __forceinline__ unsigned int block_inclusive_scan_1024(unsigned int v, unsigned int *warp_sums, unsigned int *total)
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
    *total = warp_sums[31];
    return inc + (warp ? warp_sums[warp - 1] : 0u);
}
__device__ __forceinline__ void icp_tile_sums_kernel_body(const IcpState *st, const unsigned int *cell_count, unsigned int *tile_sums)
{
    __shared__ unsigned int warp_sums[32];
    const unsigned int n = icp_n_cells(st);
    for (unsigned int tile = blockIdx.x; tile * kScanTile < n; tile += gridDim.x)
    {
        const unsigned int first = tile * kScanTile + threadIdx.x * 8u;
        unsigned int local = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) local += first + k < n ? cell_count[first + k] : 0u;
        unsigned int total;
        block_inclusive_scan_1024(local, warp_sums, &total);
        if (threadIdx.x == 0) tile_sums[tile] = total;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(1024) icp_tile_sums_kernel(const IcpState *st, const unsigned int *cell_count, unsigned int *tile_sums) { icp_tile_sums_kernel_body(st, cell_count, tile_sums); }
__device__ __forceinline__ void icp_scan_tiles_kernel_body(const IcpState *st, unsigned int *tile_sums)
{
    __shared__ unsigned int warp_sums[32];
    const unsigned int n_tiles = (icp_n_cells(st) + kScanTile - 1) / kScanTile; // <= kMaxTiles <= 1025
    unsigned int carry = 0;
    for (unsigned int base = 0; base < n_tiles; base += 1024u)
    {
        const unsigned int i = base + threadIdx.x;
        const unsigned int v = i < n_tiles ? tile_sums[i] : 0u;
        unsigned int total;
        const unsigned int inc = block_inclusive_scan_1024(v, warp_sums, &total);
        if (i < n_tiles) tile_sums[i] = carry + inc - v;
        carry += total;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(1024) icp_scan_tiles_kernel(const IcpState *st, unsigned int *tile_sums) { icp_scan_tiles_kernel_body(st, tile_sums); }
__device__ __forceinline__ void icp_scan_apply_kernel_body(const IcpState *st, const unsigned int *cell_count,
                                                              const unsigned int *tile_sums, unsigned int *cell_start)
{
    __shared__ unsigned int warp_sums[32];
    const unsigned int n = icp_n_cells(st);
    for (unsigned int tile = blockIdx.x; tile * kScanTile < n; tile += gridDim.x)
    {
        const unsigned int first = tile * kScanTile + threadIdx.x * 8u;
        unsigned int v[8], local = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { v[k] = first + k < n ? cell_count[first + k] : 0u; local += v[k]; }
        unsigned int total;
        const unsigned int inc = block_inclusive_scan_1024(local, warp_sums, &total);
        unsigned int run = tile_sums[tile] + inc - local;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (first + k < n) { cell_start[first + k] = run; run += v[k]; }
        // the thread that owns the last cell also writes the end sentinel (= number of points in the grid)
        if (first < n && first + 8 >= n) cell_start[n] = run;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(1024) icp_scan_apply_kernel(const IcpState *st, const unsigned int *cell_count,
                                                              const unsigned int *tile_sums, unsigned int *cell_start) { icp_scan_apply_kernel_body(st, cell_count, tile_sums, cell_start); }

// scatter into cell order (the order inside a cell is irrelevant: ties are broken by original index in the search)
__device__ __forceinline__ void icp_scatter_kernel_body(const float *pts, int n, const unsigned int *point_cell, const unsigned int *cell_start,
                                   unsigned int *cell_count, float4 *sorted)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const unsigned int c = point_cell[i];
        if (c == 0xFFFFFFFFu) continue;
        const unsigned int pos = cell_start[c] + atomicSub(&cell_count[c], 1u) - 1u;
        sorted[pos] = make_float4(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], __int_as_float(i));
    }
}
__global__ void icp_scatter_kernel(const float *pts, int n, const unsigned int *point_cell, const unsigned int *cell_start,
                                   unsigned int *cell_count, float4 *sorted) { icp_scatter_kernel_body(pts, n, point_cell, cell_start, cell_count, sorted); }

// The whole grid construction as ONE cooperative launch: the eight phases above separated by grid-wide barriers instead of
// kernel boundaries (eight launches of a few microseconds of work each cost more in launch latency than in work).
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int &generation)
{
    ++generation;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        __threadfence();
        atomicAdd(counter, 1u);
        while (*(volatile unsigned int *)counter < generation * gridDim.x) { }
        __threadfence();
    }
    __syncthreads();
}
__global__ void __launch_bounds__(1024, 1) icp_grid_build_kernel(const float *pts, int n, IcpState *st, float min_cell, unsigned int max_cells,
                                                                 unsigned int *cell_count, unsigned int *point_cell, unsigned int *tile_sums,
                                                                 unsigned int *cell_start, float4 *sorted, unsigned int *sync)
{
    unsigned int generation = 0;
    icp_bbox_kernel_body(pts, n, st);
    grid_barrier(sync, generation);
    if (blockIdx.x == 0 && threadIdx.x == 0) icp_grid_setup_kernel_body(st, n, min_cell, max_cells);
    grid_barrier(sync, generation);
    icp_clear_kernel_body(st, cell_count);
    grid_barrier(sync, generation);
    icp_count_kernel_body(pts, n, st, cell_count, point_cell);
    grid_barrier(sync, generation);
    icp_tile_sums_kernel_body(st, cell_count, tile_sums);
    grid_barrier(sync, generation);
    if (blockIdx.x == 0) icp_scan_tiles_kernel_body(st, tile_sums);
    grid_barrier(sync, generation);
    icp_scan_apply_kernel_body(st, cell_count, tile_sums, cell_start);
    grid_barrier(sync, generation);
    icp_scatter_kernel_body(pts, n, point_cell, cell_start, cell_count, sorted);
}

// ---------------------------------------------------------------------------------------------------------
// per-iteration kernel
// ---------------------------------------------------------------------------------------------------------
// nanoflann L2_Simple_Adaptor::evalMetric: result += diff*diff over the 3 dimensions, in order
__device__ __forceinline__ float dist2_nanoflann(float ax, float ay, float az, float bx, float by, float bz)
{
    const float dx = fsub(ax, bx), dy = fsub(ay, by), dz = fsub(az, bz);
    return fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
}

struct IcpArgs
{
    const float *src;       // ns x 3 (already scaled)
    const float *tgt;       // nt x 3
    const float *nrm;       // nt x 3 or nullptr (point-to-point)
    const unsigned int *cell_start;
    const float4 *sorted;
    int *nn;                // ns
    double *partials;       // gridDim.x x kPacket
    IcpState *st;
    int ns;
    float search_radius;
    double sq_threshold;
    int final_pass;         // 1: only CountInliers (rmse + pairs), no solve
    int keep_far;           // 1 in the last solving pass: nn keeps the true nearest neighbour even beyond the inlier radius (or -2 =
                            // "none within the search radius, true nearest unknown"), because the closing CountInliers re-tests
                            // exactly these neighbours under the NEXT pose (ICP.cpp:90,206)
    int *pairs;             // final pass: inlier flags are turned into pairs by the compaction kernel
    unsigned char *inlier;  // ns flags
    IcpComm comm;           // world <= 1: single GPU
    // nearest-neighbour certificates (see icp_certify_kernel)
    float4 *qref;           // ns: query position at the point's last full search; w = displacement budget (NaN: none yet)
    int2 *nn_ref;           // ns: nearest neighbour found by that search (-1 = none within the inlier radius) and the runner-up
    float *budget2;         // ns: displacement budget of the two-candidate certificate
    float4 *rec;            // 2 x ns (second persistent form): the nearest neighbour's point and normal {tx,ty,tz,nx},{ny,nz,-,-}; tx = NaN: none
    unsigned int *worklist; // ns
    float guard;            // extra radius every full search covers beyond its nearest neighbour, in grid cells
    int certify;            // 0: every pass searches every point (reference behaviour of the search, for A/B tests)
    int seed_walks;         // 1: a full search starts from the previous pass's neighbour as its bound (same result, shorter walk)
    float unscale;          // ICPParameter::scaling: the closing pass divides the (scaled) coordinates by it again
};

// geometry::TransformPoints: T * (x,y,z,1) then divide by w (Geometry.cpp:19-27)
__device__ __forceinline__ void transform_point(const float *T, float sx, float sy, float sz, float &px, float &py, float &pz)
{
    const float w = row_xyz1(T[3], T[7], T[11], T[15], sx, sy, sz);
    px = fdiv(row_xyz1(T[0], T[4], T[8], T[12], sx, sy, sz), w);
    py = fdiv(row_xyz1(T[1], T[5], T[9], T[13], sx, sy, sz), w);
    pz = fdiv(row_xyz1(T[2], T[6], T[10], T[14], sx, sy, sz), w);
}

// scans cells [xa, xb] (clipped to the grid) of row (cy, cz) into the running best (distance, index); ties go to the
// smaller target index
// running result of a walk: the two nearest points seen (distance, target index; ties go to the smaller index) and the
// third smallest distance
struct Nearest2
{
    float d1, d2, d3;
    int i1, i2;
};
__device__ __forceinline__ void nearest2_push(Nearest2 &b, float d, int ti)
{
    if (d < b.d1 || (d == b.d1 && ti < b.i1)) { b.d3 = b.d2; b.d2 = b.d1; b.i2 = b.i1; b.d1 = d; b.i1 = ti; }
    else if (d < b.d2 || (d == b.d2 && ti < b.i2)) { b.d3 = b.d2; b.d2 = d; b.i2 = ti; }
    else if (d < b.d3) b.d3 = d;
}
__device__ __forceinline__ void scan_cells(const IcpGrid &g, const unsigned int *__restrict__ cell_start, const float4 *__restrict__ sorted,
                                           int xa, int xb, int cy, int cz, float qx, float qy, float qz, Nearest2 &b)
{
    xa = max(xa, 0);
    xb = min(xb, g.dim[0] - 1);
    if (xa > xb || cy < 0 || cy >= g.dim[1] || cz < 0 || cz >= g.dim[2]) return;
    const unsigned int row = (unsigned int)g.dim[0] * ((unsigned int)cy + (unsigned int)g.dim[1] * (unsigned int)cz);
    const unsigned int s = __ldg(&cell_start[row + xa]), e = __ldg(&cell_start[row + xb + 1]);
    unsigned int k = s;
    for (; k < e; ++k)
    {
        const float4 t = __ldg(&sorted[k]);
        nearest2_push(b, dist2_nanoflann(qx, qy, qz, t.x, t.y, t.z), __float_as_int(t.w));
    }
}

// One query against the grid.  Cells of a grid row (fixed y, z) are contiguous in the sorted array, so the search works
// on rows: the home cell first, then the rows of growing Chebyshev rings in (y, z) around it.  A row is visited once, and
// only the x-interval of it that can still hold a point at least as close as the current best (or the inlier radius) --
// the pruning of a k-d tree descent, on a grid.  Ring r complete means every point within (r + m) cells of the query has
// been seen, m = distance to the nearest y/z face of the home cell; the walk stops there once the best is that close.
// All gap arithmetic is in cell units and shrunk by `slack` because cell assignment itself is rounded in float.
struct RowPruner
{
    float fx, ay, az;       // x position in cell units; position inside the home cell along y, z
    float inv_h2, slack;
    __device__ __forceinline__ float gap(float a, int d) const
    {
        if (d == 0) return 0.0f;
        return fmaxf((d < 0 ? a : 1.0f - a) + (float)(abs(d) - 1) - slack, 0.0f);
    }
    // x-interval [xlo, xhi] of row (dy, dz) worth scanning given the squared search bound eb; false if none
    __device__ __forceinline__ bool interval(int dy, int dz, float eb, int &xlo, int &xhi) const
    {
        const float gy = gap(ay, dy), gz = gap(az, dz);
        const float rem = eb * inv_h2 - (gy * gy + gz * gz);
        if (!(rem >= 0.0f)) return false;
        const float half = sqrtf(rem) + slack;
        xlo = (int)floorf(fx - half);
        xhi = (int)floorf(fx + half);
        return true;
    }
};

struct NnResult
{
    int index;      // exact nearest neighbour, -1 if none within the inlier radius
    int index2;     // second nearest (valid when budget2 > 0)
    float budget;   // how far the query may move and keep `index` as its strictly nearest neighbour (or keep having none)
    float budget2;  // ... and keep {index, index2} as its two nearest (see icp_certify_kernel)
};

// Squared radius the walk has to cover: the best distance so far plus the guard, never beyond inlier radius + guard.
__device__ __forceinline__ float search_bound(float bd, float guard, float cap_g2)
{
    const float r = sqrtf(bd) + guard; // +inf stays +inf
    return fminf(r * r, cap_g2);
}

// Exact nearest neighbour with a certificate.  The walk covers every target point within C = sqrt(d1) + guard of the query
// (d1 = squared distance of the nearest one; C is at most inlier radius + guard), so besides the nearest neighbour it knows
// a lower bound L2 on the distance of every OTHER point: the second smallest distance seen, or C.  A query that has moved by
// less than (L2 - sqrt(d1)) / 2 still has the same, strictly nearest neighbour -- no search needed.  If the second nearest
// point lies inside C as well, the same argument one level up (L3 = third smallest distance seen, or C) says how far the
// query may move before a third point can interfere: until then the answer is the nearer of the two stored candidates, which
// settles the queries that sit on the bisector of two target points.  A query with no neighbour within the inlier radius has
// none as long as it moved by less than (nearest distance seen or C) - inlier radius.  All budgets are shrunk by 1e-4
// relative and 1 um absolute, orders of magnitude above the float rounding of the distance arithmetic they stand in for.
// `seed_d2`: squared distance from the query to ANY point of this target cloud (the neighbour of the previous pass), +inf if none is
// known.  It only tightens the radius the walk starts with -- the nearest point is at most that far, so everything within
// sqrt(d1) + guard is still covered and the result, certificate included, is the one of the unseeded walk.
__device__ NnResult grid_nearest(const IcpGrid &g, const unsigned int *__restrict__ cell_start, const float4 *__restrict__ sorted, float qx,
                                 float qy, float qz, float radius, float guard, float seed_d2 = HUGE_VALF)
{
    NnResult out;
    out.index = out.index2 = -1;
    out.budget = out.budget2 = -1.0f;
    const float fx = (qx - g.origin[0]) * g.inv_h, fy = (qy - g.origin[1]) * g.inv_h, fz = (qz - g.origin[2]) * g.inv_h;
    const float bound = (float)(1 << 20);
    if (!(qx == qx && qy == qy && qz == qz && fabsf(fx) < bound && fabsf(fy) < bound && fabsf(fz) < bound)) return out;
    // home cell, not clamped: a query outside the grid starts where it is
    const int hx = (int)floorf(fx), hy = (int)floorf(fy), hz = (int)floorf(fz);
    RowPruner pr;
    pr.fx = fx; pr.ay = fy - hy; pr.az = fz - hz;
    pr.inv_h2 = g.inv_h * g.inv_h;
    pr.slack = 1e-3f;
    const float r2cap = radius * radius; // a neighbour farther than this is reported as "none" anyway
    const float cap_g = radius + guard, cap_g2 = cap_g * cap_g;
    const float inf = __int_as_float(0x7f800000);
    Nearest2 nb;
    nb.d1 = nb.d2 = nb.d3 = inf;
    nb.i1 = nb.i2 = -1;
    int xlo, xhi;
    scan_cells(g, cell_start, sorted, hx, hx, hy, hz, qx, qy, qz, nb);
    // ring 0: the rest of the home row
    if (pr.interval(0, 0, search_bound(fminf(nb.d1, seed_d2), guard, cap_g2), xlo, xhi))
    {
        if (xlo < hx) scan_cells(g, cell_start, sorted, xlo, hx - 1, hy, hz, qx, qy, qz, nb);
        if (xhi > hx) scan_cells(g, cell_start, sorted, hx + 1, xhi, hy, hz, qx, qy, qz, nb);
    }
    const float m_yz = fminf(fminf(pr.ay, 1.0f - pr.ay), fminf(pr.az, 1.0f - pr.az));
    const float radius_cells = cap_g * g.inv_h;
    const int r_max = (int)ceilf(radius_cells) + 1;
    for (int r = 1; r <= r_max; ++r)
    {
        // everything within `covered` cells has been seen once ring r-1 is complete
        const float covered = (float)(r - 1) + m_yz - pr.slack;
        float eb = search_bound(fminf(nb.d1, seed_d2), guard, cap_g2);
        if (covered > 0.0f && nb.i1 >= 0 && eb * pr.inv_h2 <= covered * covered) break;
        if (covered > radius_cells) break;
        if (r == 1)
        {
            // the eight rows around the home row, nearest first; the bound shrinks as better points turn up
            unsigned int mask = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b)
            {
                const int dy = (b == 0 || b == 4 || b == 5) ? -1 : ((b == 1 || b == 6 || b == 7) ? 1 : 0);
                const int dz = (b == 2 || b == 4 || b == 6) ? -1 : ((b == 3 || b == 5 || b == 7) ? 1 : 0);
                if (pr.interval(dy, dz, eb, xlo, xhi)) mask |= 1u << b;
            }
            while (mask)
            {
                const int b = __ffs(mask) - 1;
                mask &= mask - 1;
                // bit b -> (dy, dz):  0:(-1,0) 1:(+1,0) 2:(0,-1) 3:(0,+1) 4:(-1,-1) 5:(-1,+1) 6:(+1,-1) 7:(+1,+1)
                const int dy = (b == 0 || b == 4 || b == 5) ? -1 : ((b == 1 || b == 6 || b == 7) ? 1 : 0);
                const int dz = (b == 2 || b == 4 || b == 6) ? -1 : ((b == 3 || b == 5 || b == 7) ? 1 : 0);
                if (pr.interval(dy, dz, eb, xlo, xhi))
                {
                    scan_cells(g, cell_start, sorted, xlo, xhi, hy + dy, hz + dz, qx, qy, qz, nb);
                    eb = search_bound(fminf(nb.d1, seed_d2), guard, cap_g2);
                }
            }
        }
        else
        {
            for (int dz = -r; dz <= r; ++dz)
                for (int dy = -r; dy <= r; dy += (abs(dz) == r ? 1 : 2 * r))
                    if (pr.interval(dy, dz, eb, xlo, xhi))
                    {
                        scan_cells(g, cell_start, sorted, xlo, xhi, hy + dy, hz + dz, qx, qy, qz, nb);
                        eb = search_bound(fminf(nb.d1, seed_d2), guard, cap_g2);
                    }
        }
    }
    const float shrink = 1.0f - 1e-4f, abs_slack = 1e-6f;
    if (nb.i1 >= 0 && nb.d1 <= r2cap)
    {
        const float r1 = sqrtf(nb.d1), r2 = sqrtf(nb.d2);
        const float C = fminf(r1 + guard, cap_g);       // everything closer than this has been seen
        out.index = nb.i1;
        out.budget = 0.5f * (fminf(r2, C) - r1) * shrink - abs_slack;
        if (nb.i2 >= 0 && r2 < C)
        {
            out.index2 = nb.i2;
            out.budget2 = 0.5f * (fminf(sqrtf(nb.d3), C) - r2) * shrink - abs_slack;
        }
    }
    else
    {
        const float nearest = nb.i1 >= 0 ? fminf(sqrtf(nb.d1), cap_g) : cap_g; // nothing is closer than this
        out.budget = (nearest - radius) * shrink - abs_slack;
    }
    return out;
}

// K7a: which queries keep their nearest neighbour.  One thread per source point: the query is transformed with the current
// pose exactly as the search would, compared with where it stood at its last full search, and if it moved by less than that
// search's budget the stored neighbour (or the nearer of the two stored candidates) is still the exact answer, re-tested
// against the inlier radius with the search's own distance arithmetic; otherwise the point goes on the work list of the
// search kernel.  A pass over already converged iterations thus costs one streaming read instead of a grid walk per point.
__global__ void __launch_bounds__(kIcpThreads) icp_certify_kernel(IcpArgs a)
{
    __shared__ float sT[16];
    if (threadIdx.x < 16) sT[threadIdx.x] = a.st->T[threadIdx.x];
    __syncthreads();
    const float r2cap = a.search_radius * a.search_radius;
    const int lane = threadIdx.x & 31;
    const int n_round = (a.ns + 31) & ~31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x)
    {
        bool todo = false;
        if (i < a.ns)
        {
            float px, py, pz;
            transform_point(sT, a.src[3 * i], a.src[3 * i + 1], a.src[3 * i + 2], px, py, pz);
            const float4 q = a.qref[i];
            const float dx = px - q.x, dy = py - q.y, dz = pz - q.z;
            const float moved = sqrtf(dx * dx + dy * dy + dz * dz) * (1.0f + 1e-6f);
            int nn = -1;
            if (moved < q.w) // false for the NaN budget of a point that was never searched
            {
                const int j = a.nn_ref[i].x;
                if (j >= 0)
                {
                    const float d = dist2_nanoflann(px, py, pz, __ldg(&a.tgt[3 * j]), __ldg(&a.tgt[3 * j + 1]), __ldg(&a.tgt[3 * j + 2]));
                    if (!(d > r2cap) || a.keep_far) nn = j;
                }
                else if (a.keep_far) nn = -2;
                a.nn[i] = nn;
            }
            else if (moved < a.budget2[i])
            {
                const int2 jj = a.nn_ref[i];
                const float da = dist2_nanoflann(px, py, pz, __ldg(&a.tgt[3 * jj.x]), __ldg(&a.tgt[3 * jj.x + 1]), __ldg(&a.tgt[3 * jj.x + 2]));
                const float db = dist2_nanoflann(px, py, pz, __ldg(&a.tgt[3 * jj.y]), __ldg(&a.tgt[3 * jj.y + 1]), __ldg(&a.tgt[3 * jj.y + 2]));
                const bool first = da < db || (da == db && jj.x < jj.y);
                if (!((first ? da : db) > r2cap) || a.keep_far) nn = first ? jj.x : jj.y;
                a.nn[i] = nn;
            }
            else
                todo = true;
        }
        // warp-aggregated append
        const unsigned int m = __ballot_sync(0xffffffffu, todo);
        if (m)
        {
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(&a.st->wl_count, (unsigned int)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (todo) a.worklist[base + __popc(m & ((1u << lane) - 1u))] = (unsigned int)i;
        }
    }
}

// K7b: exact nearest neighbour (and certificate) of the queries on the work list, one thread per query
__global__ void __launch_bounds__(kIcpThreads) icp_search_kernel(IcpArgs a)
{
    __shared__ float sT[16];
    if (threadIdx.x < 16) sT[threadIdx.x] = a.st->T[threadIdx.x];
    __syncthreads();
    const IcpGrid g = a.st->grid;
    const int n = (int)a.st->wl_count;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < n; w += gridDim.x * blockDim.x)
    {
        const int i = (int)a.worklist[w];
        float px, py, pz;
        transform_point(sT, a.src[3 * i], a.src[3 * i + 1], a.src[3 * i + 2], px, py, pz);
        const NnResult r = grid_nearest(g, a.cell_start, a.sorted, px, py, pz, a.search_radius, a.certify ? a.guard * g.h : 0.0f);
        a.nn[i] = r.index < 0 && a.keep_far ? -2 : r.index;
        a.nn_ref[i] = make_int2(r.index, r.index2);
        a.qref[i] = make_float4(px, py, pz, a.certify ? r.budget : -1.0f);
        a.budget2[i] = a.certify ? r.budget2 : -1.0f;
    }
}

// The closing CountInliers (ICP.cpp:90,206) runs no search: it re-tests corresponding_index of the last loop iteration -- the
// nearest neighbours under the PREVIOUS pose, at whatever distance -- with the final pose.  The last solving pass therefore
// leaves in nn the true nearest neighbour where it knows it (keep_far) and -2 where all it knows is "nothing within the search
// radius".  Such a point can only become an inlier if the pose update moved it by more than its clearance: the new position
// is within m of the old one, so a neighbour that passes the test now was within threshold + m then.  Points whose clearance
// (the "none" budget of their last full search) covers the move have no candidate; the others get an exact search of that
// radius around their PREVIOUS position.  Kept out of line: it is the rare path of one pass.
__device__ __noinline__ int final_resolve_far(const IcpArgs &a, const IcpGrid &g, const float *T_prev, const float *T, int i)
{
    const float sx = a.src[3 * i], sy = a.src[3 * i + 1], sz = a.src[3 * i + 2];
    float ox, oy, oz, px, py, pz;
    transform_point(T_prev, sx, sy, sz, ox, oy, oz);
    transform_point(T, sx, sy, sz, px, py, pz);
    const float4 q = a.qref[i];
    const float ax = ox - q.x, ay = oy - q.y, az = oz - q.z, bx = px - ox, by = py - oy, bz = pz - oz;
    const float moved = sqrtf(ax * ax + ay * ay + az * az) * (1.0f + 1e-6f);
    const float m = sqrtf(bx * bx + by * by + bz * bz) * (1.0f + 1e-6f);
    if (moved + m < q.w) return -1; // also false for NaN
    if (!(m == m) || !(fabsf(m) < 1e30f)) return -1;
    return grid_nearest(g, a.cell_start, a.sorted, ox, oy, oz, a.search_radius + m * (1.0f + 1e-3f) + 1e-6f, 0.0f).index;
}
__global__ void __launch_bounds__(kIcpThreads) icp_final_resolve_kernel(IcpArgs a)
{
    __shared__ float sT[32];
    if (threadIdx.x < 16) sT[threadIdx.x] = a.st->T[threadIdx.x];
    else if (threadIdx.x < 32) sT[threadIdx.x] = a.st->T_prev[threadIdx.x - 16];
    __syncthreads();
    const IcpGrid g = a.st->grid;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.ns; i += gridDim.x * blockDim.x)
        if (a.nn[i] == -2) a.nn[i] = final_resolve_far(a, g, sT + 16, sT, i);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// adds the warp-wide sum of v to the warp's running total of packet component k
__device__ __forceinline__ void warp_accumulate(double (*s_part)[kPacket], int warp, int lane, int k, double v)
{
    v = warp_sum(v);
    if (lane == 0) s_part[warp][k] += v;
}

// the pose update of one iteration (ICP.cpp:78-86, 137-143, 198) from the 30-scalar packet: T <- dT * T, T_prev <- T.  One thread.
// Returns false when nothing was updated (point-to-point without inliers: the reference would produce NaNs there).
__device__ bool icp_solve_core(const double *pk, bool plane, float *T, float *T_prev)
{
    for (int e = 0; e < 16; ++e) T_prev[e] = T[e];
    double dT[16];
    if (plane)
    {
        double JTJ[36], nJTr[6], x[6];
        // (fully unrolled: every index is a compile-time constant, the system stays in registers on the common path)
#pragma unroll
        for (int p = 0; p < 6; ++p)
#pragma unroll
            for (int q = p; q < 6; ++q)
            {
                const double v = pk[p * 6 - (p * (p - 1)) / 2 + (q - p)]; // position of (p, q) in the row-wise upper triangle
                JTJ[p * 6 + q] = v;
                JTJ[q * 6 + p] = v;
            }
#pragma unroll
        for (int p = 0; p < 6; ++p) nJTr[p] = -pk[21 + p];
        linalg::solve_normal_equations6(JTJ, nJTr, x);
        // the reference's x is float32
#pragma unroll
        for (int p = 0; p < 6; ++p) x[p] = (double)(float)x[p];
        linalg::se3_exp(x, dT);
    }
    else
    {
        if (pk[29] < 0.5) return false; // no inliers: leave T unchanged
        linalg::kabsch_from_sums(pk[29], &pk[0], &pk[3], &pk[6], dT);
    }
    // start_T = tmp_T * start_T in float (ICP.cpp:86,198)
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
    return true;
}
// ... on the device-resident state, run by one thread of the last CTA (separate-launch form and first persistent form)
__device__ void icp_solve_and_update(const IcpArgs &a, IcpState *st)
{
    st->sum_error = st->packet[28];
    st->n_inliers = (unsigned long long)(st->packet[29] + 0.5);
    if (a.final_pass) return;
    st->iteration += 1;
    float T[16], Tp[16];
    for (int e = 0; e < 16; ++e) T[e] = st->T[e];
    const bool moved = icp_solve_core(st->packet, a.nrm != nullptr, T, Tp);
    for (int e = 0; e < 16; ++e) st->T_prev[e] = Tp[e];
    if (moved)
        for (int e = 0; e < 16; ++e) st->T[e] = T[e];
}

// K8: CountInliers test, Jacobian row and the 30-scalar packet over the pairs (i, nn[i]).  Per-thread double accumulators
// -> warp shuffles -> one partial per CTA in a fixed slot; the last CTA to finish sums the partials in a fixed order
// (deterministic), solves and updates the pose.  PLANE: point-to-plane rows, else the Kabsch sums of point-to-point.
struct IcpShared
{
    double part[kIcpThreads / 32][kPacket];
    float T[16];
    float T_prev[16];
    bool last;
};
// returns true in every thread of the CTA that finished last (the one that summed the partials, solved and updated the pose)
template <bool PLANE>
__device__ __forceinline__ bool accumulate_pass(const IcpArgs &a, IcpShared &sh)
{
    double (*s_part)[kPacket] = sh.part;
    float *sT = sh.T;
    bool &s_last = sh.last;
    __syncthreads(); // the previous user of the shared block is done
    if (threadIdx.x < 16) sT[threadIdx.x] = a.st->T[threadIdx.x];
    __syncthreads();
    const float *T = sT; // column-major
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NA = PLANE ? 27 : 15;
    double acc[NA], err_sum = 0.0, cnt = 0.0;
#pragma unroll
    for (int k = 0; k < NA; ++k) acc[k] = 0.0;
    const bool sums = !a.final_pass;
    const int *nn = a.nn;
    const float *__restrict__ src = a.src;
    const float *__restrict__ tgt = a.tgt;
    const float *__restrict__ nrm = a.nrm;
    // two points per trip: their index / coordinate loads are issued together (the trips are latency-bound gathers)
#pragma unroll 2
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.ns; i += gridDim.x * blockDim.x)
    {
        const int j = __ldcg(&nn[i]); // written by this very thread earlier in the pass (or the pass before): not the read-only path
        bool inl = false;
        if (j >= 0)
        {
            const float sx = __ldg(&src[3 * i]), sy = __ldg(&src[3 * i + 1]), sz = __ldg(&src[3 * i + 2]);
            const float tx = __ldg(&tgt[3 * j]), ty = __ldg(&tgt[3 * j + 1]), tz = __ldg(&tgt[3 * j + 2]);
            // CountInliers (ICP.cpp:19-21): (R*s + t - target).squaredNorm() in float, compared as double
            const float ex = fsub(fadd(fadd(fmul(T[0], sx), fadd(fmul(T[4], sy), fmul(T[8], sz))), T[12]), tx);
            const float ey = fsub(fadd(fadd(fmul(T[1], sx), fadd(fmul(T[5], sy), fmul(T[9], sz))), T[13]), ty);
            const float ez = fsub(fadd(fadd(fmul(T[2], sx), fadd(fmul(T[6], sy), fmul(T[10], sz))), T[14]), tz);
            const double err = (double)fadd(fmul(ex, ex), fadd(fmul(ey, ey), fmul(ez, ez)));
            inl = err < a.sq_threshold;
            if (inl)
            {
                err_sum += err;
                cnt += 1.0;
                if (sums)
                {
                    float px, py, pz;
                    transform_point(T, sx, sy, sz, px, py, pz);
                    if (PLANE)
                    {
                        // EstimateRigidTransformationPointToPlane (ICP.cpp:121-136): row = [n ; s' x n], r = n.s' - n.t with
                        // s' the transformed source point; all in float like the reference, summed in double
                        const float nx = __ldg(&nrm[3 * j]), ny = __ldg(&nrm[3 * j + 1]), nz = __ldg(&nrm[3 * j + 2]);
                        const float r = fsub(dot3(nx, ny, nz, px, py, pz), dot3(nx, ny, nz, tx, ty, tz));
                        const float row[6] = {nx, ny, nz, fsub(fmul(py, nz), fmul(pz, ny)), fsub(fmul(pz, nx), fmul(px, nz)),
                                              fsub(fmul(px, ny), fmul(py, nx))};
                        int k = 0;
#pragma unroll
                        for (int p = 0; p < 6; ++p)
#pragma unroll
                            for (int q = p; q < 6; ++q) acc[k++] += (double)fmul(row[p], row[q]);
#pragma unroll
                        for (int p = 0; p < 6; ++p) acc[21 + p] += (double)fmul(r, row[p]);
                    }
                    else
                    {
                        // PointToPoint (ICP.cpp:78-84): Kabsch sums over (transformed source, target)
                        const double P[3] = {px, py, pz}, Q[3] = {tx, ty, tz};
#pragma unroll
                        for (int c = 0; c < 3; ++c) { acc[c] += P[c]; acc[3 + c] += Q[c]; }
#pragma unroll
                        for (int p = 0; p < 3; ++p)
#pragma unroll
                            for (int q = 0; q < 3; ++q) acc[6 + 3 * p + q] += P[p] * Q[q];
                    }
                }
            }
        }
        if (a.final_pass) a.inlier[i] = inl;
    }
    if (lane < 30) s_part[warp][lane] = 0.0;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NA; ++k)
    {
        const double v = warp_sum(acc[k]);
        if (lane == 0) s_part[warp][k] = v;
    }
    err_sum = warp_sum(err_sum);
    cnt = warp_sum(cnt);
    if (lane == 0) { s_part[warp][28] = err_sum; s_part[warp][29] = cnt; }
    __syncthreads();
    if (threadIdx.x < 30)
    {
        double v = 0.0;
        for (int w = 0; w < kIcpThreads / 32; ++w) v += s_part[w][threadIdx.x];
        a.partials[(size_t)blockIdx.x * kPacket + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&a.st->blocks_done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    if (threadIdx.x == 0) icp_stamp(a.st, a.st->pass, 4);
    {
        // 8 interleaved chains per component, then the chains in order: deterministic
        const int k = threadIdx.x & 31, chain = threadIdx.x >> 5;
        double v = 0.0;
        if (k < 30)
        {
            constexpr int kChains = kIcpThreads / 32, kInFlight = 40;
            if (gridDim.x <= kChains * kInFlight)
            {
                // all loads of the chain in flight at once, then the additions in the chain's fixed order (x + 0.0 == x)
                double t[kInFlight];
#pragma unroll
                for (int u = 0; u < kInFlight; ++u)
                {
                    const unsigned int b = chain + u * kChains;
                    t[u] = b < gridDim.x ? __ldcg(&a.partials[(size_t)b * kPacket + k]) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < kInFlight; ++u) v += t[u];
            }
            else
            {
#pragma unroll 8
                for (unsigned int b = chain; b < gridDim.x; b += kChains) v += __ldcg(&a.partials[(size_t)b * kPacket + k]);
            }
        }
        __syncthreads();
        s_part[chain][k] = v;
        __syncthreads();
        if (threadIdx.x < 30)
        {
            double tot = 0.0;
            for (int c = 0; c < kIcpThreads / 32; ++c) tot += s_part[c][threadIdx.x];
            a.st->packet[threadIdx.x] = tot;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) a.st->n_inliers_local = (unsigned long long)(a.st->packet[29] + 0.5);
    comm_allreduce(a.comm, a.st->packet, 30);
    if (threadIdx.x != 0) return true;
    icp_stamp(a.st, a.st->pass, 5);
    a.st->blocks_done = 0;
    a.st->searched_total += a.st->wl_count;
    if (a.st->pass < 64) a.st->searched_per_pass[a.st->pass] = a.st->wl_count;
    a.st->wl_count = 0; // the next pass builds its own work list
    icp_solve_and_update(a, a.st);
    icp_stamp(a.st, a.st->pass, 6);
    return true;
}
template <bool PLANE>
__global__ void __launch_bounds__(kIcpThreads) icp_accumulate_kernel(IcpArgs a)
{
    __shared__ IcpShared sh;
    if (accumulate_pass<PLANE>(a, sh) && threadIdx.x == 0) a.st->pass += 1;
}

// The whole iteration loop as ONE persistent launch (cooperative: every CTA is resident).  A pass is what the three kernels
// above do, without their launch gaps and without the round trip of the work list: every thread certifies its own points
// and searches the ones that fail on the spot, then accumulates the same points; the last CTA to arrive sums the partials,
// exchanges the packet with the peer ranks, solves, updates the pose and releases the grid into the next pass (one
// grid-wide barrier per pass, built on the pass counter).  Point-to-thread assignment, per-CTA partials and the order of
// every sum are those of the separate kernels, so the results are bit-identical to theirs.
template <bool PLANE>
__global__ void __launch_bounds__(kIcpThreads, 2) icp_loop_kernel(IcpArgs a, int n_pass)
{
    __shared__ IcpShared sh;
    const IcpGrid g = a.st->grid;
    const float r2cap = a.search_radius * a.search_radius;
    const float guard = a.certify ? a.guard * g.h : 0.0f;
    for (int pass = 0; pass < n_pass; ++pass)
    {
        a.final_pass = pass == n_pass - 1;
        a.keep_far = pass == n_pass - 2;
        if (blockIdx.x == 0 && threadIdx.x == 0) icp_stamp(a.st, pass, 0);
        __syncthreads();
        if (threadIdx.x < 16) sh.T[threadIdx.x] = a.st->T[threadIdx.x];
        else if (threadIdx.x < 32 && a.final_pass) sh.T_prev[threadIdx.x - 16] = a.st->T_prev[threadIdx.x - 16];
        __syncthreads();
        unsigned int searched = 0;
        if (a.final_pass)
        {
            // no search: the neighbours of the last solving pass, re-tested under the final pose by accumulate_pass
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.ns; i += gridDim.x * blockDim.x)
                if (a.nn[i] == -2) { a.nn[i] = final_resolve_far(a, g, sh.T_prev, sh.T, i); ++searched; }
        }
        else
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.ns; i += gridDim.x * blockDim.x)
        {
            float px, py, pz;
            transform_point(sh.T, a.src[3 * i], a.src[3 * i + 1], a.src[3 * i + 2], px, py, pz);
            const float4 q = a.qref[i];
            const float dx = px - q.x, dy = py - q.y, dz = pz - q.z;
            const float moved = sqrtf(dx * dx + dy * dy + dz * dz) * (1.0f + 1e-6f);
            int nn = -1;
            if (moved < q.w)
            {
                const int j = a.nn_ref[i].x;
                if (j >= 0)
                {
                    const float d = dist2_nanoflann(px, py, pz, __ldg(&a.tgt[3 * j]), __ldg(&a.tgt[3 * j + 1]), __ldg(&a.tgt[3 * j + 2]));
                    if (!(d > r2cap) || a.keep_far) nn = j;
                }
                else if (a.keep_far) nn = -2;
            }
            else if (moved < a.budget2[i])
            {
                const int2 jj = a.nn_ref[i];
                const float da = dist2_nanoflann(px, py, pz, __ldg(&a.tgt[3 * jj.x]), __ldg(&a.tgt[3 * jj.x + 1]), __ldg(&a.tgt[3 * jj.x + 2]));
                const float db = dist2_nanoflann(px, py, pz, __ldg(&a.tgt[3 * jj.y]), __ldg(&a.tgt[3 * jj.y + 1]), __ldg(&a.tgt[3 * jj.y + 2]));
                const bool first = da < db || (da == db && jj.x < jj.y);
                if (!((first ? da : db) > r2cap) || a.keep_far) nn = first ? jj.x : jj.y;
            }
            else
            {
                const NnResult r = grid_nearest(g, a.cell_start, a.sorted, px, py, pz, a.search_radius, guard);
                nn = r.index < 0 && a.keep_far ? -2 : r.index;
                a.nn_ref[i] = make_int2(r.index, r.index2);
                a.qref[i] = make_float4(px, py, pz, a.certify ? r.budget : -1.0f);
                a.budget2[i] = a.certify ? r.budget2 : -1.0f;
                ++searched;
            }
            a.nn[i] = nn;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) searched += __shfl_xor_sync(0xffffffffu, searched, o);
        if ((threadIdx.x & 31) == 0 && searched) atomicAdd(&a.st->wl_count, searched);
        if (blockIdx.x == 0 && threadIdx.x == 0) icp_stamp(a.st, pass, 1);
        // every thread reads back only the nn entries it wrote itself: no grid-wide ordering is needed before this
        const bool last = accumulate_pass<PLANE>(a, sh);
        if (blockIdx.x == 0 && threadIdx.x == 0) icp_stamp(a.st, pass, 2);
        if (pass == n_pass - 1) break;
        if (threadIdx.x == 0)
        {
            if (last)
            {
                __threadfence();
                atomicExch(&a.st->pass, (unsigned int)(pass + 1)); // releases the grid
            }
            else
                while (*(volatile unsigned int *)&a.st->pass <= (unsigned int)pass) { }
            __threadfence();
            if (blockIdx.x == 0) icp_stamp(a.st, pass, 3);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// The pass loop, second persistent form (the default).  What the first form spends its time on, measured per converged pass
// (27 us, profiles/r02_icp_phases_before.json): 5 us certifying, 8 us accumulating 29 double sums per thread and folding them with
// 435 shuffle instructions per warp, then 14 us in which 295 CTAs wait for one to sum 296 partials (7 us), solve (3 us) and
// release them.  This form removes the serial owner and the per-thread sums:
//   * one CTA of 1024 threads per SM, one point per lane and trip; certified points go straight from the certificate to their
//     Jacobian row (one loop, the neighbour index never leaves the registers), the others are searched on the spot by an
//     out-of-line call that keeps the exact grid walk's registers out of the loop;
//   * the sums are an 8x8 outer-product accumulation: every point contributes c c^T with c = (n, s' x n, r, 1) for point-to-plane
//     (J^T J, J^T r and the count are entries of that matrix), c = (s', 1, t, 0) for point-to-point (the Kabsch sums), c = (err, 1)
//     for the closing CountInliers.  A warp stages its 32 vectors in shared memory and folds them with eight DMMA.8x8x4
//     instructions -- the FP64 tensor-core op used as a reduction primitive, exact products of floats, double accumulation --
//     into a fragment of two doubles per lane that lives in registers for the whole pass: no per-thread accumulators, no shuffles;
//   * every CTA publishes its 8x8 partial, announces it on one counter, waits until all have, and then EVERY CTA sums all
//     partials in the same fixed order and solves the same 6x6 system redundantly: nobody waits for an owner, the pose never
//     travels through global memory, and the result is deterministic and identical on all CTAs.
// When the source is split across ranks the exchange needs an owner again: CTA 0 sums, exchanges the packet with the peers
// (comm_allreduce) and publishes the combined packet; the others wait for it.
// ---------------------------------------------------------------------------------------------------------
constexpr int kLoop2Threads = 1024;
constexpr int kLoop2Warps = kLoop2Threads / 32;
constexpr int kLoop2Chunks = kLoop2Threads / 64; // groups that each sum every 16th partial

// entry (row * 8 + col) of the 8x8 sum matrix that feeds component k of the 30-scalar packet, -1: none.
// mode 0 point-to-plane, 1 point-to-point, 2 closing CountInliers
__host__ __device__ constexpr int packet_source(int mode, int k)
{
    if (mode == 0)
    {
        if (k < 21)
        {   // upper triangle of J^T J, row by row
            int p = 0, first = 0;
            while (k >= first + (6 - p)) { first += 6 - p; ++p; }
            return p * 8 + p + (k - first);
        }
        if (k < 27) return (k - 21) * 8 + 6; // J^T r
        return k == 29 ? 63 : -1;            // count
    }
    if (mode == 1)
    {
        if (k < 3) return k * 8 + 3;               // sum s'
        if (k < 6) return 3 * 8 + 4 + (k - 3);     // sum t
        if (k < 15) return ((k - 6) / 3) * 8 + 4 + (k - 6) % 3; // sum s' t^T
        return k == 29 ? 3 * 8 + 3 : -1;
    }
    // closing pass, c = (s, 1, t, err) with the ORIGINAL clouds: the Kabsch sums of result.T (ICP.cpp:93-105,208-221), sum err, count
    if (k < 3) return k * 8 + 3;               // sum s
    if (k < 6) return 3 * 8 + 4 + (k - 3);     // sum t
    if (k < 15) return ((k - 6) / 3) * 8 + 4 + (k - 6) % 3; // sum s t^T
    if (k == 15) return 3 * 8 + 3;             // count (where kabsch_from_sums' caller expects it)
    return k == 28 ? 3 * 8 + 7 : (k == 29 ? 3 * 8 + 3 : -1); // sum err, count
}
__host__ __device__ constexpr unsigned long long packet_need_mask(int mode)
{
    unsigned long long m = 0;
    for (int k = 0; k < 30; ++k)
        if (packet_source(mode, k) >= 0) m |= 1ull << packet_source(mode, k);
    return m;
}

struct Loop2Shared
{
    union
    {
        float stage[kLoop2Warps][32 * 8]; // during the trips: per warp, 32 points x 8 components
        double wsum[kLoop2Warps][64];     // after the trips: per-warp 8x8 sums
        double chunk[kLoop2Chunks][64];   // partial sums of the cross-CTA reduction
    } u;
    double sum64[64];   // the 8x8 sums over all CTAs
    double packet[32];  // ... as the 30-scalar packet
    float T[16], T_prev[16];
    IcpGrid grid;
};

// the exact search of one query, out of line (its registers stay out of the streaming loop); updates the certificate
__device__ __noinline__ int loop2_search(const IcpArgs &a, const IcpGrid *g, float guard, int i, float px, float py, float pz, int keep_far, bool seeded)
{
    // from the second pass on the point's record holds the neighbour the previous search of THIS call found (pass 0 searches every
    // point: the certificates start out invalid); its distance from where the query stands now bounds the walk from the start
    float seed = __int_as_float(0x7f800000);
    if (seeded)
    {
        const float4 p0 = a.rec[2 * i];
        if (p0.x == p0.x) seed = dist2_nanoflann(px, py, pz, p0.x, p0.y, p0.z);
    }
    const NnResult r = grid_nearest(*g, a.cell_start, a.sorted, px, py, pz, a.search_radius, guard, seed);
    a.nn_ref[i] = make_int2(r.index, r.index2);
    a.qref[i] = make_float4(px, py, pz, a.certify ? r.budget : -1.0f);
    a.budget2[i] = a.certify ? r.budget2 : -1.0f;
    // the neighbour's point and normal travel with the certificate: a certified pass then streams 60 contiguous bytes per point
    // (source, certificate, record) instead of gathering two 32-byte sectors per point from the target arrays
    float4 r0 = make_float4(__int_as_float(0x7fc00000), 0.0f, 0.0f, 0.0f), r1 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (r.index >= 0)
    {
        r0.x = __ldg(&a.tgt[3 * r.index]); r0.y = __ldg(&a.tgt[3 * r.index + 1]); r0.z = __ldg(&a.tgt[3 * r.index + 2]);
        if (a.nrm) { r0.w = __ldg(&a.nrm[3 * r.index]); r1.x = __ldg(&a.nrm[3 * r.index + 1]); r1.y = __ldg(&a.nrm[3 * r.index + 2]); }
    }
    a.rec[2 * i] = r0;
    a.rec[2 * i + 1] = r1;
    return r.index < 0 && keep_far ? -2 : r.index;
}
// geometry::TransformPoints for a rigid pose: w = ((0*x + 0*y) + 0*z) + 1 == 1 exactly, and x / 1 == x, so the three divisions
// are skipped whenever w is exactly one (any other w takes them)
__device__ __forceinline__ void transform_point_w1(const float *T, float sx, float sy, float sz, float &px, float &py, float &pz)
{
    const float w = row_xyz1(T[3], T[7], T[11], T[15], sx, sy, sz);
    px = row_xyz1(T[0], T[4], T[8], T[12], sx, sy, sz);
    py = row_xyz1(T[1], T[5], T[9], T[13], sx, sy, sz);
    pz = row_xyz1(T[2], T[6], T[10], T[14], sx, sy, sz);
    if (w != 1.0f) { px = fdiv(px, w); py = fdiv(py, w); pz = fdiv(pz, w); }
}

// the inlier test and the point's 8 components (zeros unless it is an inlier); T = rows 0..2 of the pose, T[3 * col + row]
template <bool PLANE>
__device__ __forceinline__ bool loop2_components_of(const IcpArgs &a, const float *T, bool final_pass, float tx, float ty, float tz, float nx, float ny,
                                                    float nz, float sx, float sy, float sz, float px, float py, float pz, float *comp);
template <bool PLANE>
__device__ __forceinline__ bool loop2_components(const IcpArgs &a, const float *T, bool final_pass, int nn, float sx, float sy, float sz, float px,
                                                 float py, float pz, float *comp)
{
#pragma unroll
    for (int k = 0; k < 8; ++k) comp[k] = 0.0f;
    if (nn < 0) return false;
    const float tx = __ldg(&a.tgt[3 * nn]), ty = __ldg(&a.tgt[3 * nn + 1]), tz = __ldg(&a.tgt[3 * nn + 2]);
    // the normal travels with the target point, not after the inlier test: one round trip to L2 instead of two
    float nx = 0.0f, ny = 0.0f, nz = 0.0f;
    if (PLANE && !final_pass) { nx = __ldg(&a.nrm[3 * nn]); ny = __ldg(&a.nrm[3 * nn + 1]); nz = __ldg(&a.nrm[3 * nn + 2]); }
    return loop2_components_of<PLANE>(a, T, final_pass, tx, ty, tz, nx, ny, nz, sx, sy, sz, px, py, pz, comp);
}
// ... given the neighbour's point (and normal): comp must be zero on entry
template <bool PLANE>
__device__ __forceinline__ bool loop2_components_of(const IcpArgs &a, const float *T, bool final_pass, float tx, float ty, float tz, float nx, float ny,
                                                    float nz, float sx, float sy, float sz, float px, float py, float pz, float *comp)
{
    // CountInliers (ICP.cpp:19-21): (R*s + t - target).squaredNorm() in float, compared as double
    const float ex = fsub(fadd(fadd(fmul(T[0], sx), fadd(fmul(T[3], sy), fmul(T[6], sz))), T[9]), tx);
    const float ey = fsub(fadd(fadd(fmul(T[1], sx), fadd(fmul(T[4], sy), fmul(T[7], sz))), T[10]), ty);
    const float ez = fsub(fadd(fadd(fmul(T[2], sx), fadd(fmul(T[5], sy), fmul(T[8], sz))), T[11]), tz);
    const float err = fadd(fmul(ex, ex), fadd(fmul(ey, ey), fmul(ez, ez)));
    if (!((double)err < a.sq_threshold)) return false;
    if (final_pass)
    {   // un-scaled originals (ICP.cpp:93-99,208-214 divide the scaled copies again)
        comp[0] = sx; comp[1] = sy; comp[2] = sz; comp[3] = 1.0f;
        comp[4] = tx; comp[5] = ty; comp[6] = tz; comp[7] = err;
        if (a.unscale != 1.0f)
        {
            comp[0] = fdiv(sx, a.unscale); comp[1] = fdiv(sy, a.unscale); comp[2] = fdiv(sz, a.unscale);
            comp[4] = fdiv(tx, a.unscale); comp[5] = fdiv(ty, a.unscale); comp[6] = fdiv(tz, a.unscale);
        }
    }
    else if (PLANE)
    {
        // EstimateRigidTransformationPointToPlane (ICP.cpp:121-136): row = [n ; s' x n], r = n.s' - n.t, in float
        comp[0] = nx; comp[1] = ny; comp[2] = nz;
        comp[3] = fsub(fmul(py, nz), fmul(pz, ny));
        comp[4] = fsub(fmul(pz, nx), fmul(px, nz));
        comp[5] = fsub(fmul(px, ny), fmul(py, nx));
        comp[6] = fsub(dot3(nx, ny, nz, px, py, pz), dot3(nx, ny, nz, tx, ty, tz));
        comp[7] = 1.0f;
    }
    else
    {   // PointToPoint (ICP.cpp:78-84): Kabsch sums over (transformed source, target)
        comp[0] = px; comp[1] = py; comp[2] = pz; comp[3] = 1.0f;
        comp[4] = tx; comp[5] = ty; comp[6] = tz;
    }
    return true;
}
__device__ __forceinline__ void loop2_transform(const float *T, bool rigid, const float *T_full, float sx, float sy, float sz, float &px, float &py,
                                                float &pz)
{
    if (rigid)
    {   // the products with the zero row still decide whether w is 1 or NaN (non-finite source points)
        px = row_xyz1(T[0], T[3], T[6], T[9], sx, sy, sz);
        py = row_xyz1(T[1], T[4], T[7], T[10], sx, sy, sz);
        pz = row_xyz1(T[2], T[5], T[8], T[11], sx, sy, sz);
        const float w = row_xyz1(0.0f, 0.0f, 0.0f, 1.0f, sx, sy, sz);
        if (w != 1.0f) { px = fdiv(px, w); py = fdiv(py, w); pz = fdiv(pz, w); }
    }
    else
        transform_point(T_full, sx, sy, sz, px, py, pz);
}

template <bool PLANE>
__global__ void __launch_bounds__(kLoop2Threads, 1) icp_loop2_kernel(const __grid_constant__ IcpArgs a, int n_pass, double *partials, unsigned int *sync)
{
    __shared__ Loop2Shared sh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_cta = gridDim.x;
    if (threadIdx.x < 16) sh.T[threadIdx.x] = a.st->T[threadIdx.x];
    if (threadIdx.x == 32) sh.grid = a.st->grid;
    __syncthreads();
    const float guard = a.certify ? a.guard * sh.grid.h : 0.0f;
    const int n_trips = (a.ns + 31) >> 5;
    // trips of this warp: first, first + step, ...  Consecutive trips go to different CTAs, so that a cluster of points that need
    // a full search (they come in spatial clusters) is spread over the grid instead of holding up one CTA.
    const int first_trip = warp * n_cta + blockIdx.x, trip_step = kLoop2Warps * n_cta;
    for (int pass = 0; pass < n_pass; ++pass)
    {
        const bool final_pass = pass == n_pass - 1;
        const int keep_far = pass == n_pass - 2;
        if (blockIdx.x == 0 && threadIdx.x == 0) icp_stamp(a.st, pass, 0);
        float T[12]; // rows 0..2 of the pose: T[3 * c + r]
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int r = 0; r < 3; ++r) T[3 * c + r] = sh.T[4 * c + r];
        const bool rigid = sh.T[3] == 0.0f && sh.T[7] == 0.0f && sh.T[11] == 0.0f && sh.T[15] == 1.0f;
        double c0 = 0.0, c1 = 0.0;
        float *stage = sh.u.stage[warp];
        // ---- streaming loop: certified points go from the certificate straight to their row; the others are noted ----
        unsigned int todo = 0; // bit k: the point of this lane in the warp's k-th trip needs the out-of-line path
        int k = 0;
        for (int trip = first_trip; trip < n_trips; trip += trip_step, ++k)
        {
            const int i = trip * 32 + lane;
            float comp[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
            if (i < a.ns)
            {
                const float sx = __ldg(&a.src[3 * i]), sy = __ldg(&a.src[3 * i + 1]), sz = __ldg(&a.src[3 * i + 2]);
                int nn = -1;
                bool later = false;
                float px = 0.0f, py = 0.0f, pz = 0.0f;
                bool by_record = false;
                float4 r0 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), r1 = r0;
                if (!final_pass)
                {
                    const float4 q = a.qref[i];
                    r0 = a.rec[2 * i];     // with the certificate, not after it: nothing in the certified path depends on a loaded index
                    r1 = a.rec[2 * i + 1];
                    loop2_transform(T, rigid, sh.T, sx, sy, sz, px, py, pz);
                    const float dx = px - q.x, dy = py - q.y, dz = pz - q.z;
                    const float moved2 = (dx * dx + dy * dy + dz * dz) * (1.0f + 4e-6f);
                    if (q.w > 0.0f && moved2 < q.w * q.w)
                    {   // the certified strictly nearest neighbour (beyond the inlier radius the inlier test rejects it)
                        by_record = true;
                        if (keep_far)
                        {
                            nn = a.nn_ref[i].x;
                            if (nn < 0) nn = -2;
                        }
                    }
                    else
                    {
                        const int2 jj = a.nn_ref[i];
                        const float b2 = a.budget2[i];
                        if (b2 > 0.0f && moved2 < b2 * b2)
                        {
                            const float da = dist2_nanoflann(px, py, pz, __ldg(&a.tgt[3 * jj.x]), __ldg(&a.tgt[3 * jj.x + 1]), __ldg(&a.tgt[3 * jj.x + 2]));
                            const float db = dist2_nanoflann(px, py, pz, __ldg(&a.tgt[3 * jj.y]), __ldg(&a.tgt[3 * jj.y + 1]), __ldg(&a.tgt[3 * jj.y + 2]));
                            nn = (da < db || (da == db && jj.x < jj.y)) ? jj.x : jj.y;
                        }
                        else
                            later = true;
                    }
                    if (keep_far && !later) a.nn[i] = nn; // the closing CountInliers re-tests exactly these (ICP.cpp:90,206)
                }
                else
                {
                    nn = a.nn[i];
                    later = nn == -2;
                }
                if (later) todo |= 1u << (k & 31);
                else if (by_record)
                {
                    if (r0.x == r0.x) loop2_components_of<PLANE>(a, T, false, r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, sx, sy, sz, px, py, pz, comp);
                }
                else
                {
                    const bool inl = loop2_components<PLANE>(a, T, final_pass, nn, sx, sy, sz, px, py, pz, comp);
                    if (final_pass) a.inlier[i] = inl;
                }
            }
            warp_fold_outer8(stage, lane, comp, c0, c1);
        }
        // ---- the noted points: exact search (or, in the closing pass, the wider re-search), then the same fold ----
        unsigned int searched = 0;
        double d0 = 0.0, d1 = 0.0; // own accumulators: these live across the out-of-line calls, the streaming loop's must not
        while (__any_sync(0xffffffffu, todo != 0u))
        {
            float comp[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
            if (todo)
            {
                const int kk = __ffs(todo) - 1;
                todo &= todo - 1u;
                const int i = (first_trip + kk * trip_step) * 32 + lane;
                const float sx = __ldg(&a.src[3 * i]), sy = __ldg(&a.src[3 * i + 1]), sz = __ldg(&a.src[3 * i + 2]);
                float px = 0.0f, py = 0.0f, pz = 0.0f;
                int nn;
                if (!final_pass)
                {
                    loop2_transform(T, rigid, sh.T, sx, sy, sz, px, py, pz);
                    nn = loop2_search(a, &sh.grid, guard, i, px, py, pz, keep_far, pass > 0 && a.seed_walks);
                    if (keep_far) a.nn[i] = nn;
                }
                else
                {
                    nn = final_resolve_far(a, sh.grid, sh.T_prev, sh.T, i);
                    a.nn[i] = nn;
                }
                ++searched;
                const bool inl = loop2_components<PLANE>(a, T, final_pass, nn, sx, sy, sz, px, py, pz, comp);
                if (final_pass) a.inlier[i] = inl;
            }
            warp_fold_outer8(stage, lane, comp, d0, d1);
        }
        c0 += d0;
        c1 += d1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) searched += __shfl_xor_sync(0xffffffffu, searched, o);
        if (lane == 0 && searched)
        {
            atomicAdd(&a.st->searched_total, (unsigned long long)searched);
            if (pass < 64) atomicAdd(&a.st->searched_per_pass[pass], searched);
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) icp_stamp(a.st, pass, 1);
        // ---- CTA partial: lane L of a warp holds entries 2L, 2L+1 of the row-major 8x8 matrix ----
        __syncthreads(); // every warp is done with its staging area (aliased below)
        *reinterpret_cast<double2 *>(&sh.u.wsum[warp][2 * lane]) = make_double2(c0, c1);
        __syncthreads();
        const unsigned long long need = final_pass ? packet_need_mask(2) : packet_need_mask(PLANE ? 0 : 1);
        const int e = threadIdx.x & 63, grp = threadIdx.x >> 6;
        const bool needed = (need >> e) & 1ull;
        double v = 0.0;
        if (needed && grp < 8)
            v = (sh.u.wsum[4 * grp][e] + sh.u.wsum[4 * grp + 1][e]) + (sh.u.wsum[4 * grp + 2][e] + sh.u.wsum[4 * grp + 3][e]);
        __syncthreads();
        if (grp < 8) sh.u.chunk[grp][e] = v;
        __syncthreads();
        double *mine = partials + ((size_t)(pass & 1) * n_cta + blockIdx.x) * 64;
        if (threadIdx.x < 64 && needed)
        {
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += sh.u.chunk[k][e];
            __stcg(&mine[e], t);
        }
        __threadfence();
        __syncthreads();
        if (blockIdx.x == 0 && threadIdx.x == 0) icp_stamp(a.st, pass, 2);
        // ---- arrive, wait for everybody ----
        if (threadIdx.x == 0)
        {
            atomicAdd(sync, 1u);
            const unsigned int want = (unsigned int)(pass + 1) * (unsigned int)n_cta;
            while (*(volatile unsigned int *)sync < want) { }
            __threadfence();
        }
        __syncthreads();
        if (blockIdx.x == 0 && threadIdx.x == 0) icp_stamp(a.st, pass, 4);
        // ---- all partials -> the 8x8 sums, in one fixed order (every CTA, or CTA 0 alone when ranks exchange packets) ----
        const bool owner_mode = a.comm.world > 1;
        if (!owner_mode || blockIdx.x == 0)
        {
            const double *all = partials + (size_t)(pass & 1) * n_cta * 64;
            double t = 0.0;
            if (needed)
            {
                if (n_cta <= kLoop2Chunks * 10)
                {
                    double ld[10];
#pragma unroll
                    for (int u = 0; u < 10; ++u)
                    {
                        const int b = grp + u * kLoop2Chunks;
                        ld[u] = b < n_cta ? __ldcg(&all[(size_t)b * 64 + e]) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 10; ++u) t += ld[u];
                }
                else
                    for (int b = grp; b < n_cta; b += kLoop2Chunks) t += __ldcg(&all[(size_t)b * 64 + e]);
            }
            sh.u.chunk[grp][e] = t;
            __syncthreads();
            if (threadIdx.x < 64)
            {
                double tot = 0.0;
                if (needed)
#pragma unroll
                    for (int k = 0; k < kLoop2Chunks; ++k) tot += sh.u.chunk[k][e];
                sh.sum64[e] = tot;
            }
            __syncthreads();
            if (threadIdx.x < 32)
            {
                const int mode = final_pass ? 2 : (PLANE ? 0 : 1);
                int src = -1;
                if (threadIdx.x < 30) src = mode == 0 ? packet_source(0, threadIdx.x) : (mode == 1 ? packet_source(1, threadIdx.x) : packet_source(2, threadIdx.x));
                sh.packet[threadIdx.x] = src >= 0 ? sh.sum64[src] : 0.0;
            }
            __syncthreads();
        }
        if (owner_mode)
        {
            if (blockIdx.x == 0)
            {
                if (threadIdx.x == 0) a.st->n_inliers_local = (unsigned long long)(sh.packet[29] + 0.5);
                comm_allreduce(a.comm, sh.packet, 30);
                if (threadIdx.x < 30) __stcg(&a.st->packet[threadIdx.x], sh.packet[threadIdx.x]);
                __threadfence();
                __syncthreads();
                if (threadIdx.x == 0) atomicExch(&a.st->pass, (unsigned int)(pass + 1));
            }
            else
            {
                if (threadIdx.x == 0)
                {
                    while (*(volatile unsigned int *)&a.st->pass <= (unsigned int)pass) { }
                    __threadfence();
                }
                __syncthreads();
                if (threadIdx.x < 30) sh.packet[threadIdx.x] = __ldcg(&a.st->packet[threadIdx.x]);
                __syncthreads();
            }
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) icp_stamp(a.st, pass, 5);
        // ---- the same solve on every CTA ----
        if (threadIdx.x == 0 && !final_pass)
            icp_solve_core(sh.packet, PLANE, sh.T, sh.T_prev);
        if (blockIdx.x == 0 && threadIdx.x == 0) { icp_stamp(a.st, pass, 6); icp_stamp(a.st, pass, 3); }
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        IcpState *st = a.st;
        st->sum_error = sh.packet[28];
        st->n_inliers = (unsigned long long)(sh.packet[29] + 0.5);
        if (a.comm.world <= 1) st->n_inliers_local = st->n_inliers;
        st->iteration = (int)(n_pass - 1);
        for (int k = 0; k < 16; ++k) { st->T[k] = sh.T[k]; st->T_prev[k] = sh.T_prev[k]; }
        for (int k = 0; k < 30; ++k) st->packet[k] = sh.packet[k];
        st->pass = (unsigned int)n_pass;
    }
}

// final Kabsch sums over the inlier pairs of the ORIGINAL (unscaled) clouds (per-CTA partials, 16 components)
__global__ void __launch_bounds__(kIcpThreads) icp_final_sums_kernel(const float *src, const float *tgt, const int *nn,
                                                                     const unsigned char *inlier, int ns, float scaling,
                                                                     double *partials)
{
    __shared__ double s_part[kIcpThreads / 32][kPacket];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane < 16) s_part[warp][lane] = 0.0;
    __syncwarp();
    const int stride = gridDim.x * blockDim.x;
    for (int base = blockIdx.x * blockDim.x + warp * 32; base < ns; base += stride)
    {
        const int i = base + lane;
        const bool inl = i < ns && inlier[i];
        if (!__any_sync(0xffffffffu, inl)) continue;
        double S[3] = {0, 0, 0}, Q[3] = {0, 0, 0};
        if (inl)
        {
            const int j = nn[i];
            float s[3] = {src[3 * i], src[3 * i + 1], src[3 * i + 2]}, t[3] = {tgt[3 * j], tgt[3 * j + 1], tgt[3 * j + 2]};
            if (scaling != 1.0f)
                for (int c = 0; c < 3; ++c) { s[c] = fdiv(s[c], scaling); t[c] = fdiv(t[c], scaling); } // ICP.cpp:93-99,208-214
            for (int c = 0; c < 3; ++c) { S[c] = s[c]; Q[c] = t[c]; }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) { warp_accumulate(s_part, warp, lane, c, S[c]); warp_accumulate(s_part, warp, lane, 3 + c, Q[c]); }
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = 0; q < 3; ++q) warp_accumulate(s_part, warp, lane, 6 + 3 * p + q, S[p] * Q[q]);
        warp_accumulate(s_part, warp, lane, 15, inl ? 1.0 : 0.0);
    }
    __syncthreads();
    if (threadIdx.x < 16)
    {
        double v = 0.0;
        for (int w = 0; w < kIcpThreads / 32; ++w) v += s_part[w][threadIdx.x];
        partials[(size_t)blockIdx.x * kPacket + threadIdx.x] = v;
    }
}
__global__ void __launch_bounds__(kIcpThreads) icp_final_reduce_kernel(const double *partials, int n_partials, IcpState *st, IcpComm comm)
{
    __shared__ double s_part[kIcpThreads / 32][kPacket];
    const int k = threadIdx.x & 31, chain = threadIdx.x >> 5;
    double s = 0.0;
    if (k < 16)
        for (int b = chain; b < n_partials; b += kIcpThreads / 32) s += partials[(size_t)b * kPacket + k];
    s_part[chain][k] = s;
    __syncthreads();
    if (threadIdx.x < 16)
    {
        double tot = 0.0;
        for (int c = 0; c < kIcpThreads / 32; ++c) tot += s_part[c][threadIdx.x];
        st->packet[threadIdx.x] = tot;
    }
    __syncthreads();
    comm_allreduce(comm, st->packet, 16);
}

// ordered compaction of inlier pairs (source index ascending, like the reference's push_back loop): inliers per tile of 1024
// points, exclusive scan of the tile counts by one CTA, then every tile places its own pairs
constexpr int kPairTile = 1024;
__global__ void __launch_bounds__(kPairTile) icp_pair_count_kernel(const unsigned char *inlier, int ns, unsigned int *tile_counts)
{
    const int i = blockIdx.x * kPairTile + threadIdx.x;
    const int n = __syncthreads_count(i < ns && inlier[i]);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = (unsigned int)n;
}
__global__ void __launch_bounds__(1024) icp_pair_scan_kernel(unsigned int *tile_counts, int n_tiles)
{
    __shared__ unsigned int warp_sums[32];
    __shared__ unsigned int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += 1024)
    {
        const int i = base + threadIdx.x;
        const unsigned int v = i < n_tiles ? tile_counts[i] : 0u;
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
        const unsigned int excl = carry + (warp ? warp_sums[warp - 1] : 0u) + inc - v;
        if (i < n_tiles) tile_counts[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(kPairTile) icp_pair_write_kernel(const int *nn, const unsigned char *inlier, int ns, const unsigned int *tile_off,
                                                                   int *pairs, unsigned long long cap)
{
    __shared__ unsigned int warp_sums[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * kPairTile + threadIdx.x;
    const unsigned int f = i < ns ? inlier[i] : 0u;
    const unsigned int ballot = __ballot_sync(0xffffffffu, f != 0);
    if (lane == 0) warp_sums[warp] = (unsigned int)__popc(ballot);
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
        warp_sums[lane] = w; // inclusive
    }
    __syncthreads();
    if (!f) return;
    const unsigned long long pos = (unsigned long long)tile_off[blockIdx.x] + (warp ? warp_sums[warp - 1] : 0u) + __popc(ballot & ((1u << lane) - 1u));
    if (pos < cap) { pairs[2 * pos] = i; pairs[2 * pos + 1] = nn[i]; }
}

__global__ void icp_scale_kernel(float *p, size_t n, float s)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = fmul(p[i], s);
}

// ---------------------------------------------------------------------------------------------------------
// PointCloud::EstimateNormals (src/Geometry/PointCloud.cpp:102-144) on the grid of the ICP search: the step in front of
// registration::PointToPlane when the clouds come without normals (example/ICPTest.cpp:27-33).
//   KDTree::KnnRadiusSearch  src/Geometry/KDTree.h:230-255   knn nearest (ascending squared distance, the point itself first),
//                                                           cut where the SQUARED distance exceeds `radius`
//   geometry::FitPlane       src/Geometry/Geometry.cpp:172-199   float mean, float covariance, JacobiSVD, third column of U
//   Eigen 3.3.7 JacobiSVD<MatrixXf> for a 3x3 matrix: SVD/JacobiSVD.h:660-780, misc/RealSvd2x2.h:19-50, Jacobi/Jacobi.h:83-113
// One thread per point: exact k-nearest walk over grid rows (bound = the current k-th distance), the list kept sorted by
// (distance, index) in local memory; then the plane fit in the reference's float operation order.
// ---------------------------------------------------------------------------------------------------------
constexpr int kKnnMax = 64;
constexpr int kKnnThreads = 128;
// the sorted candidate list of a thread lives in shared memory, entry k of thread t at [k * kKnnThreads + t] (a list in local
// memory costs a DRAM round trip per shifted entry: 289 MB of writes for a 640x480 cloud)
struct KnnList
{
    float *d;
    int *i;
    int n, k;
    __device__ __forceinline__ float &dist(int e) { return d[e * kKnnThreads]; }
    __device__ __forceinline__ int &index(int e) { return i[e * kKnnThreads]; }
};
__device__ __forceinline__ void knn_push(KnnList &L, float d, int idx)
{
    if (L.n == L.k)
    {
        const float wd = L.dist(L.n - 1);
        if (!(d < wd || (d == wd && idx < L.index(L.n - 1)))) return;
    }
    int pos = L.n < L.k ? L.n : L.k - 1;
    while (pos > 0 && (L.dist(pos - 1) > d || (L.dist(pos - 1) == d && L.index(pos - 1) > idx)))
    {
        L.dist(pos) = L.dist(pos - 1);
        L.index(pos) = L.index(pos - 1);
        --pos;
    }
    L.dist(pos) = d;
    L.index(pos) = idx;
    if (L.n < L.k) ++L.n;
}
__device__ __forceinline__ void knn_scan_cells(const IcpGrid &g, const unsigned int *__restrict__ cell_start, const float4 *__restrict__ sorted,
                                               int xa, int xb, int cy, int cz, float qx, float qy, float qz, float cap2, KnnList &L)
{
    xa = max(xa, 0);
    xb = min(xb, g.dim[0] - 1);
    if (xa > xb || cy < 0 || cy >= g.dim[1] || cz < 0 || cz >= g.dim[2]) return;
    const unsigned int row = (unsigned int)g.dim[0] * ((unsigned int)cy + (unsigned int)g.dim[1] * (unsigned int)cz);
    const unsigned int s = __ldg(&cell_start[row + xa]), e = __ldg(&cell_start[row + xb + 1]);
    for (unsigned int k = s; k < e; ++k)
    {
        const float4 t = __ldg(&sorted[k]);
        const float d = dist2_nanoflann(qx, qy, qz, t.x, t.y, t.z);
        if (!(d > cap2)) knn_push(L, d, __float_as_int(t.w));
    }
}
__global__ void __launch_bounds__(kKnnThreads) estimate_normals_kernel(const float *__restrict__ pts, int n, const IcpState *st,
                                                               const unsigned int *__restrict__ cell_start, const float4 *__restrict__ sorted,
                                                               float radius, int knn, float *__restrict__ normals)
{
    extern __shared__ float knn_smem[]; // knn x kKnnThreads distances, then as many indices
    const IcpGrid g = st->grid;
    const float cap2 = radius;            // the reference compares the squared distance with `radius` itself
    const float cap = sqrtf(radius) * (1.0f + 1e-6f);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const float qx = pts[3 * i], qy = pts[3 * i + 1], qz = pts[3 * i + 2];
        KnnList L;
        L.d = knn_smem + threadIdx.x;
        L.i = reinterpret_cast<int *>(knn_smem + knn * kKnnThreads) + threadIdx.x;
        L.n = 0;
        L.k = knn;
        const float fx = (qx - g.origin[0]) * g.inv_h, fy = (qy - g.origin[1]) * g.inv_h, fz = (qz - g.origin[2]) * g.inv_h;
        float nrm[3] = {0.0f, 0.0f, 0.0f};
        if (qx == qx && qy == qy && qz == qz && fabsf(fx) < 1048576.0f && fabsf(fy) < 1048576.0f && fabsf(fz) < 1048576.0f)
        {
            const int hx = (int)floorf(fx), hy = (int)floorf(fy), hz = (int)floorf(fz);
            RowPruner pr;
            pr.fx = fx; pr.ay = fy - hy; pr.az = fz - hz;
            pr.inv_h2 = g.inv_h * g.inv_h;
            pr.slack = 1e-3f;
            const float m_yz = fminf(fminf(pr.ay, 1.0f - pr.ay), fminf(pr.az, 1.0f - pr.az));
            const float cap_cells = cap * g.inv_h;
            const int r_max = (int)ceilf(cap_cells) + 1;
            for (int r = 0; r <= r_max; ++r)
            {
                if (r > 0)
                {
                    // everything within `covered` cells has been seen once ring r-1 is complete
                    const float covered = (float)(r - 1) + m_yz - pr.slack;
                    if (covered > 0.0f && L.n == L.k && L.dist(L.n - 1) * pr.inv_h2 <= covered * covered) break;
                    if (covered > cap_cells) break;
                }
                for (int dz = -r; dz <= r; ++dz)
                    for (int dy = -r; dy <= r; dy += (abs(dz) == r || r == 0 ? 1 : 2 * r))
                    {
                        // bound: the k-th distance once the list is full, the radius before; ties at the bound must be seen
                        const float eb = L.n == L.k ? fminf(L.dist(L.n - 1), cap * cap) : cap * cap;
                        int xlo, xhi;
                        if (pr.interval(dy, dz, eb * (1.0f + 1e-6f), xlo, xhi)) knn_scan_cells(g, cell_start, sorted, xlo, xhi, hy + dy, hz + dz, qx, qy, qz, cap2, L);
                    }
            }
            fit_plane_normal(pts, L.n, [&](int k) { return L.index(k); }, nrm); // FitPlane (Geometry.cpp:172-199)
        }
        normals[3 * i] = nrm[0]; normals[3 * i + 1] = nrm[1]; normals[3 * i + 2] = nrm[2];
    }
}

} // namespace opb

using namespace opb;

struct opb_icp
{
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    // device buffers (grown on demand)
    float *d_src = nullptr, *d_tgt = nullptr, *d_nrm = nullptr;
    size_t cap_src = 0, cap_tgt = 0;
    float4 *d_sorted = nullptr;
    unsigned int *d_point_cell = nullptr, *d_cell_count = nullptr, *d_tile_sums = nullptr, *d_cell_start = nullptr;
    int *d_nn = nullptr, *d_pairs = nullptr;
    unsigned char *d_inlier = nullptr;
    float4 *d_qref = nullptr;          // nearest-neighbour certificates (icp_certify_kernel)
    int2 *d_nn_ref = nullptr;
    float *d_budget2 = nullptr;
    float4 *d_rec = nullptr;
    unsigned int *d_pair_tiles = nullptr; // inlier count / offset per tile of 1024 source points
    unsigned int *d_worklist = nullptr;
    unsigned long long last_searched = 0; // full searches of the last call (of ns * (max_iteration + 1) queries)
    unsigned int last_searched_per_pass[64] = {0};
    int last_launches = 0;                // kernels launched by the last call
    double *d_partials = nullptr;
    size_t cap_partials = 0;
    IcpState *d_state = nullptr;
    IcpState *h_state = nullptr; // pinned
    double *h_sums = nullptr;    // pinned: the 16 Kabsch sums of result.T (separate-launch finaliser)
    // cross-GPU exchange (opb_icp_comm_*)
    IcpMailbox *d_mailbox = nullptr;
    IcpComm comm = {};
    // timing
    bool profiling = false;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    int coop_ctas_per_sm = 0; // resident CTAs per SM of the persistent loop kernel; 0: cooperative launch unavailable
    bool loop2_ok = false;    // the second persistent form (one CTA of 1024 threads per SM) can be launched cooperatively
    double *d_partials2 = nullptr; // 2 (pass parity) x SMs x 64: per-CTA 8x8 sums of icp_loop2_kernel
    unsigned int *d_loop_sync = nullptr;
    int grid_ctas_per_sm = 0; // the same for the fused grid construction
    unsigned int *d_grid_sync = nullptr;
    bool peers_share_device = false; // a peer rank runs on this very GPU: two persistent grids could not be resident together
    // uploads that overlap the grid construction
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy[3] = {nullptr, nullptr, nullptr};
    // opb_icp_set_async_pairs: the call returns when pose and counters are on the host; the pair list follows on the copy stream
    bool async_pairs = false, pairs_pending = false;
    cudaEvent_t ev_pose = nullptr, ev_pairs_ready = nullptr, ev_pairs_done = nullptr;
    float last_build_ms = 0, last_iter_ms = 0;
};

static int icp_reserve(opb_icp *c, size_t ns, size_t nt)
{
    if (ns > c->cap_src)
    {
        cudaFree(c->d_src); cudaFree(c->d_nn); cudaFree(c->d_pairs); cudaFree(c->d_inlier);
        cudaFree(c->d_qref); cudaFree(c->d_nn_ref); cudaFree(c->d_worklist); cudaFree(c->d_budget2); cudaFree(c->d_rec);
        c->d_budget2 = nullptr; c->d_rec = nullptr;
        c->d_src = nullptr; c->d_nn = nullptr; c->d_pairs = nullptr; c->d_inlier = nullptr; c->cap_src = 0;
        c->d_qref = nullptr; c->d_nn_ref = nullptr; c->d_worklist = nullptr;
        OPB_CUDA(cudaMalloc(&c->d_qref, ns * sizeof(float4)));
        OPB_CUDA(cudaMalloc(&c->d_nn_ref, ns * sizeof(int2)));
        OPB_CUDA(cudaMalloc(&c->d_budget2, ns * sizeof(float)));
        OPB_CUDA(cudaMalloc(&c->d_rec, ns * 2 * sizeof(float4)));
        cudaFree(c->d_pair_tiles);
        c->d_pair_tiles = nullptr;
        OPB_CUDA(cudaMalloc(&c->d_pair_tiles, (ns / kPairTile + 2) * sizeof(unsigned int)));
        OPB_CUDA(cudaMalloc(&c->d_worklist, ns * sizeof(unsigned int)));
        OPB_CUDA(cudaMalloc(&c->d_src, ns * 3 * sizeof(float)));
        OPB_CUDA(cudaMalloc(&c->d_nn, ns * sizeof(int)));
        OPB_CUDA(cudaMalloc(&c->d_pairs, ns * 2 * sizeof(int)));
        OPB_CUDA(cudaMalloc(&c->d_inlier, ns));
        c->cap_src = ns;
    }
    if (nt > c->cap_tgt)
    {
        cudaFree(c->d_tgt); cudaFree(c->d_nrm); cudaFree(c->d_sorted); cudaFree(c->d_point_cell);
        c->d_tgt = nullptr; c->d_nrm = nullptr; c->d_sorted = nullptr; c->d_point_cell = nullptr; c->cap_tgt = 0;
        OPB_CUDA(cudaMalloc(&c->d_tgt, nt * 3 * sizeof(float)));
        OPB_CUDA(cudaMalloc(&c->d_nrm, nt * 3 * sizeof(float)));
        OPB_CUDA(cudaMalloc(&c->d_sorted, nt * sizeof(float4)));
        OPB_CUDA(cudaMalloc(&c->d_point_cell, nt * sizeof(unsigned int)));
        c->cap_tgt = nt;
    }
    const size_t blocks = (ns + kIcpThreads - 1) / kIcpThreads + 1;
    if (blocks > c->cap_partials)
    {
        cudaFree(c->d_partials);
        c->d_partials = nullptr; c->cap_partials = 0;
        OPB_CUDA(cudaMalloc(&c->d_partials, blocks * kPacket * sizeof(double)));
        c->cap_partials = blocks;
    }
    return OPB_OK;
}

// uniform grid over the nt points in c->d_tgt (bounding box in c->d_state already reset): cell_start + cell-sorted points
static int icp_build_grid(opb_icp *c, const float *d_tgt, size_t nt)
{
    cudaStream_t s = c->stream;
    // Cap on the number of grid cells: the construction streams over all of them (clear, count, scan) while only a few
    // passes of a registration still walk the grid, so about 16 cells per target point is the measured sweet spot
    // (640x480 frame: construction 0.134 -> 0.105 ms, pass loop +0.01 ms).  OPB_ICP_MAX_CELLS overrides it.
    static const long long k_cells_env = getenv("OPB_ICP_MAX_CELLS") ? atoll(getenv("OPB_ICP_MAX_CELLS")) : 0;
    unsigned long long max_cells = k_cells_env > 0 ? (unsigned long long)k_cells_env : 16ull * nt;
    if (max_cells < (1u << 20)) max_cells = 1u << 20;
    if (max_cells > kMaxCells) max_cells = kMaxCells;
    static const int k_fused = getenv("OPB_ICP_FUSED_GRID") ? atoi(getenv("OPB_ICP_FUSED_GRID")) : 1;
    if (k_fused && c->grid_ctas_per_sm > 0 && !c->peers_share_device)
    {
        // one cooperative launch, grid-wide barriers between the phases (not when a peer rank shares this GPU: its kernels may
        // be spinning on our packet while ours waits for the whole device to be free)
        OPB_CUDA(cudaMemsetAsync(c->d_grid_sync, 0, sizeof(unsigned int), s));
        const float *pts = d_tgt;
        int n = (int)nt;
        float min_cell = 0.0f;
        unsigned int mc = (unsigned int)max_cells;
        void *kargs[] = {(void *)&pts, (void *)&n, (void *)&c->d_state, (void *)&min_cell, (void *)&mc, (void *)&c->d_cell_count,
                         (void *)&c->d_point_cell, (void *)&c->d_tile_sums, (void *)&c->d_cell_start, (void *)&c->d_sorted, (void *)&c->d_grid_sync};
        OPB_CUDA(cudaLaunchCooperativeKernel((const void *)icp_grid_build_kernel, dim3(c->sm_count * c->grid_ctas_per_sm), dim3(1024), kargs, 0, s));
        return OPB_OK;
    }
    const int nb_t = (int)((nt + 255) / 256) < c->sm_count * 8 ? (int)((nt + 255) / 256) : c->sm_count * 8;
    icp_bbox_kernel<<<nb_t, 256, 0, s>>>(d_tgt, (int)nt, c->d_state);
    icp_grid_setup_kernel<<<1, 1, 0, s>>>(c->d_state, (int)nt, 0.0f, (unsigned int)max_cells);
    icp_clear_kernel<<<c->sm_count * 8, 256, 0, s>>>(c->d_state, c->d_cell_count);
    icp_count_kernel<<<nb_t, 256, 0, s>>>(d_tgt, (int)nt, c->d_state, c->d_cell_count, c->d_point_cell);
    icp_tile_sums_kernel<<<c->sm_count * 2, 1024, 0, s>>>(c->d_state, c->d_cell_count, c->d_tile_sums);
    icp_scan_tiles_kernel<<<1, 1024, 0, s>>>(c->d_state, c->d_tile_sums);
    icp_scan_apply_kernel<<<c->sm_count * 2, 1024, 0, s>>>(c->d_state, c->d_cell_count, c->d_tile_sums, c->d_cell_start);
    icp_scatter_kernel<<<nb_t, 256, 0, s>>>(d_tgt, (int)nt, c->d_point_cell, c->d_cell_start, c->d_cell_count, c->d_sorted);
    OPB_CUDA(cudaGetLastError());
    return OPB_OK;
}

extern "C"
{
int opb_icp_create(int device, void *stream, opb_icp **out)
{
    if (!out) { set_error("out is NULL"); return OPB_ERR_INVALID; }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
        cudaGetLastError();
        set_error("no CUDA device: onepiece_b200 has no CPU path");
        return OPB_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (%d devices)", device, ndev); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(device));
    opb_icp *c = new opb_icp();
    c->device = device;
    cudaDeviceProp prop;
    OPB_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    if (stream) c->stream = (cudaStream_t)stream;
    else { OPB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    cudaError_t e = cudaMalloc(&c->d_state, sizeof(IcpState));
    if (e == cudaSuccess) e = cudaHostAlloc(&c->h_state, sizeof(IcpState), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc(&c->h_sums, 17 * sizeof(double), cudaHostAllocDefault); // 16 sums + the mailbox error word
    if (e == cudaSuccess) e = cudaMalloc(&c->d_cell_count, ((size_t)kMaxCells + 1) * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_tile_sums, ((size_t)kMaxTiles + 1) * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_cell_start, ((size_t)kMaxCells + 2) * sizeof(unsigned int));
    for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = cudaEventCreate(&c->ev[i]);
    for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_pose, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_pairs_ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_pairs_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess)
    {
        // Load every kernel of the pass loop now.  With CUDA's lazy module loading the FIRST launch of a kernel takes a context-wide
        // lock and may wait for the device to drain; when two workspaces of one process exchange packets (a kernel of one spins
        // until the other's has run), a first launch in the middle of a call would stall the other thread's launches behind that
        // lock and the exchange would never complete.
        cudaFuncAttributes fa;
        const void *fns[] = {(const void *)icp_certify_kernel, (const void *)icp_search_kernel, (const void *)icp_final_resolve_kernel,
                             (const void *)icp_accumulate_kernel<true>, (const void *)icp_accumulate_kernel<false>,
                             (const void *)icp_final_sums_kernel, (const void *)icp_final_reduce_kernel, (const void *)icp_pair_count_kernel,
                             (const void *)icp_pair_scan_kernel, (const void *)icp_pair_write_kernel, (const void *)icp_scale_kernel,
                             (const void *)icp_bbox_kernel, (const void *)icp_grid_setup_kernel, (const void *)icp_clear_kernel,
                             (const void *)icp_count_kernel, (const void *)icp_tile_sums_kernel, (const void *)icp_scan_tiles_kernel,
                             (const void *)icp_scan_apply_kernel, (const void *)icp_scatter_kernel, (const void *)icp_grid_build_kernel,
                             (const void *)icp_loop_kernel<true>, (const void *)icp_loop_kernel<false>,
                             (const void *)icp_loop2_kernel<true>, (const void *)icp_loop2_kernel<false>};
        for (const void *fn : fns)
            if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, fn);
    }
    if (e == cudaSuccess)
    {
        int coop = 0, occ_plane = 0, occ_point = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
        if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_plane, icp_loop_kernel<true>, kIcpThreads, 0) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_point, icp_loop_kernel<false>, kIcpThreads, 0) == cudaSuccess)
            c->coop_ctas_per_sm = occ_plane < occ_point ? occ_plane : occ_point;
        int occ2a = 0, occ2b = 0;
        if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2a, icp_loop2_kernel<true>, kLoop2Threads, 0) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2b, icp_loop2_kernel<false>, kLoop2Threads, 0) == cudaSuccess && occ2a >= 1 && occ2b >= 1 &&
            cudaMalloc(&c->d_partials2, (size_t)2 * c->sm_count * 64 * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&c->d_loop_sync, sizeof(unsigned int)) == cudaSuccess)
            c->loop2_ok = true;
        int occ_grid = 0;
        if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_grid, icp_grid_build_kernel, 1024, 0) == cudaSuccess &&
            cudaMalloc(&c->d_grid_sync, sizeof(unsigned int)) == cudaSuccess)
            c->grid_ctas_per_sm = occ_grid > 2 ? 2 : occ_grid;
        cudaGetLastError();
    }
    if (e != cudaSuccess)
    {
        set_error("ICP workspace allocation failed: %s", cudaGetErrorString(e));
        opb_icp_destroy(c);
        return OPB_ERR_CUDA;
    }
    *out = c;
    return OPB_OK;
}

void opb_icp_destroy(opb_icp *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->d_src); cudaFree(c->d_tgt); cudaFree(c->d_nrm); cudaFree(c->d_sorted); cudaFree(c->d_point_cell);
    cudaFree(c->d_cell_count); cudaFree(c->d_tile_sums); cudaFree(c->d_cell_start); cudaFree(c->d_nn); cudaFree(c->d_pairs);
    cudaFree(c->d_inlier); cudaFree(c->d_partials); cudaFree(c->d_state); cudaFree(c->d_mailbox);
    cudaFree(c->d_qref); cudaFree(c->d_nn_ref); cudaFree(c->d_worklist); cudaFree(c->d_budget2); cudaFree(c->d_rec); cudaFree(c->d_pair_tiles);
    cudaFree(c->d_grid_sync); cudaFree(c->d_partials2); cudaFree(c->d_loop_sync);
    if (c->h_state) cudaFreeHost(c->h_state);
    if (c->h_sums) cudaFreeHost(c->h_sums);
    for (int i = 0; i < 3; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 3; ++i) if (c->ev_copy[i]) cudaEventDestroy(c->ev_copy[i]);
    if (c->ev_pose) cudaEventDestroy(c->ev_pose);
    if (c->ev_pairs_ready) cudaEventDestroy(c->ev_pairs_ready);
    if (c->ev_pairs_done) cudaEventDestroy(c->ev_pairs_done);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    cudaGetLastError();
    delete c;
}

int opb_icp_set_profiling(opb_icp *c, int on)
{
    if (!c) { set_error("icp is NULL"); return OPB_ERR_INVALID; }
    c->profiling = on != 0;
    return OPB_OK;
}
int opb_icp_last_timing(opb_icp *c, float *grid_build_ms, float *iterations_ms)
{
    if (!c) { set_error("icp is NULL"); return OPB_ERR_INVALID; }
    if (grid_build_ms) *grid_build_ms = c->last_build_ms;
    if (iterations_ms) *iterations_ms = c->last_iter_ms;
    return OPB_OK;
}

// src/tgt/nrm may be host or device pointers (cudaMemcpyDefault).  borrow: they are device arrays that stay valid and unchanged
// for the whole (synchronous) call -- opb_cloud buffers -- and are used where they lie (no copy); ready_a / ready_b: events the
// work stream has to wait for before it reads them.
static int icp_run(opb_icp *c, const float *src, size_t ns, const float *tgt, const float *nrm, size_t nt, const float *init_T,
                   const opb_icp_params *par, opb_icp_result *res, int32_t *pairs, size_t pairs_cap, bool point_to_plane,
                   bool borrow = false, cudaEvent_t ready_a = nullptr, cudaEvent_t ready_b = nullptr)
{
    if (!c || !src || !tgt || !init_T || !par || !res) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    memset(res, 0, sizeof(*res));
    for (int i = 0; i < 4; ++i) res->T[i * 5] = 1.0f;
    if (ns > 0x7FFFFFF0u || nt > 0x7FFFFFF0u) { set_error("clouds above 2^31 points are not supported"); return OPB_ERR_INVALID; }
    if (point_to_plane && (!nrm || par->scaling != 1.0))
    {
        // ICP.cpp:159-163: message + default RegistrationResult
        set_error("[ERROR]::[ICPPointToPlane]::target point cloud need to have normals.");
        res->status = OPB_ERR_INVALID;
        return OPB_ERR_INVALID;
    }
    if (nt == 0 || (ns == 0 && c->comm.world <= 1)) { set_error("empty point cloud"); res->status = OPB_ERR_INVALID; return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(c->device));
    if (borrow && par->scaling != 1.0) borrow = false; // PointToPoint scales its own copies of both clouds (ICP.cpp:33-42)
    int rc = icp_reserve(c, ns, nt); // (borrowed clouds leave the workspace's own copies unused: 12 bytes per point)
    if (rc) return rc;
    cudaStream_t s = c->stream;
    const float *d_src = borrow ? src : c->d_src, *d_tgt = borrow ? tgt : c->d_tgt, *d_nrm = borrow ? nrm : c->d_nrm;
    if (c->pairs_pending) { OPB_CUDA(cudaStreamWaitEvent(s, c->ev_pairs_done, 0)); c->pairs_pending = false; } // d_pairs is still being copied out
    if (ready_a) OPB_CUDA(cudaStreamWaitEvent(s, ready_a, 0));
    if (ready_b) OPB_CUDA(cudaStreamWaitEvent(s, ready_b, 0));
    // The target goes first on the work stream (the grid is built from it); the source and the normals follow on the copy
    // stream and arrive under the grid construction: the first certify pass waits for the source, the first accumulation
    // for the normals.  (Calls are synchronous, so nothing of an earlier call can still be reading these buffers.)
    if (!borrow) OPB_CUDA(cudaMemcpyAsync(c->d_tgt, tgt, nt * 3 * sizeof(float), cudaMemcpyDefault, s));
    OPB_CUDA(cudaEventRecord(c->ev_copy[0], s));
    OPB_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_copy[0], 0)); // keep the link for the target alone first
    if (ns && !borrow) OPB_CUDA(cudaMemcpyAsync(c->d_src, src, ns * 3 * sizeof(float), cudaMemcpyDefault, c->copy_stream));
    OPB_CUDA(cudaEventRecord(c->ev_copy[1], c->copy_stream));
    if (point_to_plane && !borrow) OPB_CUDA(cudaMemcpyAsync(c->d_nrm, nrm, nt * 3 * sizeof(float), cudaMemcpyDefault, c->copy_stream));
    OPB_CUDA(cudaEventRecord(c->ev_copy[2], c->copy_stream));
    const float scaling = (float)par->scaling;
    if (par->scaling != 1.0)
    {   // PointToPoint scales both clouds (ICP.cpp:36-42)
        OPB_CUDA(cudaStreamWaitEvent(s, c->ev_copy[1], 0));
        icp_scale_kernel<<<c->sm_count * 4, 256, 0, s>>>(c->d_src, ns * 3, scaling);
        icp_scale_kernel<<<c->sm_count * 4, 256, 0, s>>>(c->d_tgt, nt * 3, scaling);
    }
    // state
    IcpState *h = c->h_state;
    memset(h, 0, sizeof(IcpState));
    memcpy(h->T, init_T, 16 * sizeof(float));
    for (int a = 0; a < 3; ++a) { h->bbox_enc[a] = 0xFFFFFFFFu; h->bbox_enc[3 + a] = 0u; }
    OPB_CUDA(cudaMemcpyAsync(c->d_state, h, sizeof(IcpState), cudaMemcpyHostToDevice, s));
    if (c->profiling) OPB_CUDA(cudaEventRecord(c->ev[0], s));
    // grid over the target (opb_icp_estimate_normals builds the same grid over its cloud)
    rc = icp_build_grid(c, d_tgt, nt);
    if (rc) return rc;
    if (c->profiling) OPB_CUDA(cudaEventRecord(c->ev[1], s));
    // iterations
    IcpArgs a;
    a.src = d_src; a.tgt = d_tgt; a.nrm = point_to_plane ? d_nrm : nullptr;
    a.cell_start = c->d_cell_start; a.sorted = c->d_sorted; a.nn = c->d_nn; a.partials = c->d_partials; a.st = c->d_state;
    a.ns = (int)ns;
    a.search_radius = (float)(par->threshold * (1.0 + 1e-3)) + 1e-6f;
    a.sq_threshold = par->threshold * par->threshold;
    a.final_pass = 0; a.keep_far = 0; a.pairs = c->d_pairs; a.inlier = c->d_inlier;
    a.comm = c->comm;
    // nearest-neighbour certificates: a NaN budget marks "never searched"
    a.qref = c->d_qref; a.nn_ref = c->d_nn_ref; a.worklist = c->d_worklist; a.budget2 = c->d_budget2; a.rec = c->d_rec;
    static const float k_guard = getenv("OPB_ICP_GUARD") ? (float)atof(getenv("OPB_ICP_GUARD")) : 0.125f;
    static const int k_certify = getenv("OPB_ICP_CERTIFY") ? atoi(getenv("OPB_ICP_CERTIFY")) : 1;
    static const int k_seed = getenv("OPB_ICP_SEED") ? atoi(getenv("OPB_ICP_SEED")) : 1;
    a.guard = k_guard; a.certify = k_certify; a.seed_walks = k_seed; a.unscale = scaling;
    if (ns) OPB_CUDA(cudaMemsetAsync(c->d_qref, 0xFF, ns * sizeof(float4), s));
    if (ns) OPB_CUDA(cudaMemsetAsync(c->d_nn, 0xFF, ns * sizeof(int), s)); // corresponding_index(n, -1) (ICP.cpp:58,174)
    const int nb_need = ns ? (int)((ns + kIcpThreads - 1) / kIcpThreads) : 1; // an empty share still takes part in the exchange
    // developer knobs for grid-size sweeps (CTAs per SM); the defaults are the measured optimum on B200
    static const int k_search = getenv("OPB_ICP_SEARCH_CTAS") ? atoi(getenv("OPB_ICP_SEARCH_CTAS")) : 8;
    static const int k_accum = getenv("OPB_ICP_ACCUM_CTAS") ? atoi(getenv("OPB_ICP_ACCUM_CTAS")) : 2;
    const int nb_s = nb_need < c->sm_count * k_search ? nb_need : c->sm_count * k_search; // search: grid-stride over the points
    const int nb_a = nb_need < c->sm_count * k_accum ? nb_need : c->sm_count * k_accum;   // accumulate: few partials for the last CTA to sum
    static const int k_cert = getenv("OPB_ICP_CERTIFY_CTAS") ? atoi(getenv("OPB_ICP_CERTIFY_CTAS")) : 8;
    const int nb_c = nb_need < c->sm_count * k_cert ? nb_need : c->sm_count * k_cert;
    static const int k_persistent = getenv("OPB_ICP_PERSISTENT") ? atoi(getenv("OPB_ICP_PERSISTENT")) : 1;
    bool looped = false, looped2 = false;
    c->last_launches = 8 + 3 * (par->max_iteration + 1) + 2 + (pairs && pairs_cap && ns ? 3 : 0) + (par->scaling != 1.0 ? 2 : 0);
    if (k_persistent == 1 && c->loop2_ok && !c->peers_share_device && (ns + 31) / 32 <= (size_t)32 * kLoop2Warps * c->sm_count)
    {
        // second persistent form: one CTA of 1024 threads per SM, all passes in one cooperative launch
        const int nb_trips = ns ? (int)((ns + 31) / 32 + kLoop2Warps - 1) / kLoop2Warps : 1;
        const int nb_l = nb_trips < c->sm_count ? nb_trips : c->sm_count;
        int n_pass = par->max_iteration + 1;
        OPB_CUDA(cudaStreamWaitEvent(s, c->ev_copy[1], 0)); // source points uploaded
        OPB_CUDA(cudaStreamWaitEvent(s, c->ev_copy[2], 0)); // target normals uploaded
        OPB_CUDA(cudaMemsetAsync(c->d_loop_sync, 0, sizeof(unsigned int), s));
        void *kargs[] = {(void *)&a, (void *)&n_pass, (void *)&c->d_partials2, (void *)&c->d_loop_sync};
        const void *fn = point_to_plane ? (const void *)icp_loop2_kernel<true> : (const void *)icp_loop2_kernel<false>;
        OPB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(nb_l), dim3(kLoop2Threads), kargs, 0, s));
        looped = looped2 = true; // (its closing pass also forms the Kabsch sums of result.T: no finaliser kernels)
        c->last_launches = (c->grid_ctas_per_sm > 0 && !c->peers_share_device ? 1 : 8) + 1 + (pairs && pairs_cap && ns ? 3 : 0) + (par->scaling != 1.0 ? 2 : 0);
    }
    else if (k_persistent && c->coop_ctas_per_sm > 0 && !c->peers_share_device)
    {
        // one cooperative launch for all passes; grid = the accumulate grid (2 CTAs per SM), which fixes the order of the sums
        const int per_sm = c->coop_ctas_per_sm < k_accum ? c->coop_ctas_per_sm : k_accum;
        const int nb_l = nb_need < c->sm_count * per_sm ? nb_need : c->sm_count * per_sm;
        int n_pass = par->max_iteration + 1;
        OPB_CUDA(cudaStreamWaitEvent(s, c->ev_copy[1], 0)); // source points uploaded
        OPB_CUDA(cudaStreamWaitEvent(s, c->ev_copy[2], 0)); // target normals uploaded
        void *kargs[] = {(void *)&a, (void *)&n_pass};
        const void *fn = point_to_plane ? (const void *)icp_loop_kernel<true> : (const void *)icp_loop_kernel<false>;
        OPB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(nb_l), dim3(kIcpThreads), kargs, 0, s));
        looped = true;
        c->last_launches = (c->grid_ctas_per_sm > 0 && !c->peers_share_device ? 1 : 8) + 1 + 2 + (pairs && pairs_cap && ns ? 3 : 0) + (par->scaling != 1.0 ? 2 : 0);
    }
    for (int it = 0; !looped && it <= par->max_iteration; ++it)
    {
        // the last pass is the final CountInliers with the final T (ICP.cpp:90-91,206-207)
        a.final_pass = it == par->max_iteration;
        a.keep_far = it == par->max_iteration - 1;
        if (it == 0) OPB_CUDA(cudaStreamWaitEvent(s, c->ev_copy[1], 0)); // source points uploaded
        if (a.final_pass)
            icp_final_resolve_kernel<<<nb_c, kIcpThreads, 0, s>>>(a); // no search: last pass's neighbours under the final pose
        else
        {
            icp_certify_kernel<<<nb_c, kIcpThreads, 0, s>>>(a);
            icp_search_kernel<<<nb_s, kIcpThreads, 0, s>>>(a);
        }
        if (it == 0) OPB_CUDA(cudaStreamWaitEvent(s, c->ev_copy[2], 0)); // target normals uploaded
        if (point_to_plane) icp_accumulate_kernel<true><<<nb_a, kIcpThreads, 0, s>>>(a);
        else icp_accumulate_kernel<false><<<nb_a, kIcpThreads, 0, s>>>(a);
    }
    OPB_CUDA(cudaMemcpyAsync(h, c->d_state, sizeof(IcpState), cudaMemcpyDeviceToHost, s));
    if (!looped2)
    {
        icp_final_sums_kernel<<<nb_a, kIcpThreads, 0, s>>>(d_src, d_tgt, c->d_nn, c->d_inlier, (int)ns, scaling, c->d_partials);
        icp_final_reduce_kernel<<<1, kIcpThreads, 0, s>>>(c->d_partials, nb_a, c->d_state, c->comm);
        OPB_CUDA(cudaMemcpyAsync(c->h_sums, (const char *)c->d_state + offsetof(IcpState, packet), 16 * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    OPB_CUDA(cudaEventRecord(c->ev_pose, s));
    const size_t n_copy = pairs && pairs_cap && ns ? (pairs_cap < ns ? pairs_cap : ns) : 0;
    bool pairs_direct = false;
    if (n_copy)
    {
        const int n_tiles = (int)((ns + kPairTile - 1) / kPairTile);
        icp_pair_count_kernel<<<n_tiles, kPairTile, 0, s>>>(c->d_inlier, (int)ns, c->d_pair_tiles);
        icp_pair_scan_kernel<<<1, 1024, 0, s>>>(c->d_pair_tiles, n_tiles);
        icp_pair_write_kernel<<<n_tiles, kPairTile, 0, s>>>(c->d_nn, c->d_inlier, (int)ns, c->d_pair_tiles, c->d_pairs, (unsigned long long)pairs_cap);
        // Into page-locked (or device) memory the pairs follow on the same stream (one synchronisation for the whole call): the
        // first n_local_pairs entries are the result, the rest of the caller's buffer up to min(pairs_cap, ns) is scratch.
        // Into pageable memory the runtime would stage the copy and keep this thread inside the driver until the stream
        // has drained -- with a peer workspace of the same process still to launch the kernels ours waits for, that is a
        // deadlock (until the exchange times out); such buffers get exactly n_local_pairs entries after the synchronisation.
        cudaPointerAttributes attr;
        pairs_direct = cudaPointerGetAttributes(&attr, pairs) == cudaSuccess && attr.type != cudaMemoryTypeUnregistered;
        cudaGetLastError();
        const bool detached = pairs_direct && c->async_pairs && c->comm.world <= 1;
        if (detached)
        {   // the list leaves on the copy stream while the caller already works with the pose (opb_icp_wait_pairs collects it)
            OPB_CUDA(cudaEventRecord(c->ev_pairs_ready, s));
            OPB_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_pairs_ready, 0));
            OPB_CUDA(cudaMemcpyAsync(pairs, c->d_pairs, n_copy * 2 * sizeof(int), cudaMemcpyDefault, c->copy_stream));
            OPB_CUDA(cudaEventRecord(c->ev_pairs_done, c->copy_stream));
            c->pairs_pending = true;
        }
        else if (pairs_direct) OPB_CUDA(cudaMemcpyAsync(pairs, c->d_pairs, n_copy * 2 * sizeof(int), cudaMemcpyDefault, s));
    }
    if (c->profiling) OPB_CUDA(cudaEventRecord(c->ev[2], s));
    OPB_CUDA(cudaGetLastError());
    if (c->pairs_pending) OPB_CUDA(cudaEventSynchronize(c->ev_pose)); // pose, counters and Kabsch sums are on the host
    else OPB_CUDA(cudaStreamSynchronize(s));
    if (c->profiling)
    {
        if (c->pairs_pending) cudaEventSynchronize(c->ev[2]);
        cudaEventElapsedTime(&c->last_build_ms, c->ev[0], c->ev[1]);
        cudaEventElapsedTime(&c->last_iter_ms, c->ev[1], c->ev[2]);
    }
    if (c->comm.world > 1)
    {
        // (on the workspace's own stream, not the legacy stream: a command to the NULL stream is an implicit synchronisation point,
        // and a peer workspace of the same process may have a kernel spinning on this workspace's next packet)
        int *err_slot = (int *)(c->h_sums + 16);
        OPB_CUDA(cudaMemcpyAsync(err_slot, &c->d_mailbox->error, sizeof(int), cudaMemcpyDeviceToHost, s));
        OPB_CUDA(cudaStreamSynchronize(s));
        const int err = *err_slot;
        if (err)
        {
            set_error("ICP packet exchange timed out: a peer rank did not make the matching call");
            res->status = OPB_ERR_CUDA;
            return OPB_ERR_CUDA;
        }
    }
    c->last_searched = h->searched_total;
    memcpy(c->last_searched_per_pass, h->searched_per_pass, sizeof(c->last_searched_per_pass));
    res->n_inliers = (size_t)h->n_inliers;
    res->n_local_pairs = (size_t)h->n_inliers_local;
    res->rmse = sqrt(h->sum_error / (double)h->n_inliers); // CountInliers: sqrt(sum_error / inliers.size())
    res->iterations = h->iteration;
    memcpy(res->T_iterated, h->T, 16 * sizeof(float));
    // result.T = Kabsch over the final inlier pairs of the original clouds (ICP.cpp:103-105,221)
    const double *sums = looped2 ? h->packet : c->h_sums;
    if (sums[15] >= 0.5)
    {
        double Tk[16];
        linalg::kabsch_from_sums(sums[15], &sums[0], &sums[3], &sums[6], Tk);
        for (int r = 0; r < 4; ++r)
            for (int col = 0; col < 4; ++col) res->T[col * 4 + r] = (float)Tk[r * 4 + col];
    }
    else
        for (int e = 0; e < 16; ++e) res->T[e] = nanf(""); // the reference divides by zero pairs here
    if (n_copy && !pairs_direct)
    {
        const size_t n = res->n_local_pairs < pairs_cap ? res->n_local_pairs : pairs_cap;
        OPB_CUDA(cudaMemcpyAsync(pairs, c->d_pairs, n * 2 * sizeof(int), cudaMemcpyDeviceToHost, s)); // the stream is idle: returns when copied
        OPB_CUDA(cudaStreamSynchronize(s));
    }
    res->status = OPB_OK;
    return OPB_OK;
}

int opb_icp_set_async_pairs(opb_icp *c, int on)
{
    if (!c) { set_error("icp is NULL"); return OPB_ERR_INVALID; }
    c->async_pairs = on != 0;
    return OPB_OK;
}
int opb_icp_wait_pairs(opb_icp *c)
{
    if (!c) { set_error("icp is NULL"); return OPB_ERR_INVALID; }
    if (!c->pairs_pending) return OPB_OK;
    OPB_CUDA(cudaSetDevice(c->device));
    OPB_CUDA(cudaEventSynchronize(c->ev_pairs_done));
    c->pairs_pending = false;
    return OPB_OK;
}
int opb_icp_reserve(opb_icp *c, size_t n_source, size_t n_target)
{
    if (!c) { set_error("icp is NULL"); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(c->device));
    OPB_CUDA(cudaStreamSynchronize(c->stream));
    return icp_reserve(c, n_source, n_target);
}

void opb_icp_params_default(opb_icp_params *p)
{
    if (!p) return;
    p->max_iteration = 30; // ICP.h:16-18
    p->threshold = 0.2;
    p->scaling = 1.0;
}

int opb_icp_point_to_plane(opb_icp *c, const float *src_xyz, size_t ns, const float *tgt_xyz, const float *tgt_normals, size_t nt,
                           const float init_T[16], const opb_icp_params *params, opb_icp_result *result, int32_t *pairs,
                           size_t pairs_cap)
{
    return icp_run(c, src_xyz, ns, tgt_xyz, tgt_normals, nt, init_T, params, result, pairs, pairs_cap, true);
}
int opb_icp_point_to_point(opb_icp *c, const float *src_xyz, size_t ns, const float *tgt_xyz, size_t nt, const float init_T[16],
                           const opb_icp_params *params, opb_icp_result *result, int32_t *pairs, size_t pairs_cap)
{
    return icp_run(c, src_xyz, ns, tgt_xyz, nullptr, nt, init_T, params, result, pairs, pairs_cap, false);
}
static int icp_run_clouds(opb_icp *c, opb_cloud *source, opb_cloud *target, const float *init_T, const opb_icp_params *par,
                          opb_icp_result *res, int32_t *pairs, size_t pairs_cap, bool plane)
{
    if (!c || !source || !target) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    if (source->device != c->device || target->device != c->device) { set_error("clouds and workspace live on different devices"); return OPB_ERR_INVALID; }
    size_t ns = 0, nt = 0;
    int rc = cloud_wait(source, &ns);
    if (rc == OPB_OK) rc = cloud_wait(target, &nt);
    if (rc) return rc;
    return icp_run(c, source->d_xyz, ns, target->d_xyz, plane && target->has_normals ? target->d_nrm : nullptr, nt, init_T, par, res, pairs,
                   pairs_cap, plane, true);
}
int opb_icp_point_to_plane_clouds(opb_icp *c, opb_cloud *source, opb_cloud *target, const float init_T[16], const opb_icp_params *params,
                                  opb_icp_result *result, int32_t *pairs, size_t pairs_cap)
{
    return icp_run_clouds(c, source, target, init_T, params, result, pairs, pairs_cap, true);
}
int opb_icp_point_to_point_clouds(opb_icp *c, opb_cloud *source, opb_cloud *target, const float init_T[16], const opb_icp_params *params,
                                  opb_icp_result *result, int32_t *pairs, size_t pairs_cap)
{
    return icp_run_clouds(c, source, target, init_T, params, result, pairs, pairs_cap, false);
}
int opb_icp_last_search_count(opb_icp *c, uint64_t *full_searches)
{
    if (!c || !full_searches) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    *full_searches = c->last_searched;
    return OPB_OK;
}
int opb_icp_estimate_normals(opb_icp *c, const float *xyz, size_t n, float radius, int knn, float *normals)
{
    if (!c || !xyz || !normals) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    if (knn < 1 || knn > kKnnMax) { set_error("knn must be 1..%d", kKnnMax); return OPB_ERR_INVALID; }
    if (!(radius > 0)) { set_error("radius must be > 0"); return OPB_ERR_INVALID; }
    if (n == 0) return OPB_OK;
    if (n > 0x7FFFFFF0u) { set_error("clouds above 2^31 points are not supported"); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(c->device));
    int rc = icp_reserve(c, n, n);
    if (rc) return rc;
    cudaStream_t s = c->stream;
    OPB_CUDA(cudaMemcpyAsync(c->d_tgt, xyz, n * 3 * sizeof(float), cudaMemcpyDefault, s));
    IcpState *h = c->h_state;
    memset(h, 0, sizeof(IcpState));
    for (int a = 0; a < 3; ++a) { h->bbox_enc[a] = 0xFFFFFFFFu; h->bbox_enc[3 + a] = 0u; }
    OPB_CUDA(cudaMemcpyAsync(c->d_state, h, sizeof(IcpState), cudaMemcpyHostToDevice, s));
    rc = icp_build_grid(c, c->d_tgt, n);
    if (rc) return rc;
    // the normals land in the workspace's normal buffer, then go to wherever the caller's pointer lives
    const size_t knn_smem = (size_t)knn * kKnnThreads * (sizeof(float) + sizeof(int));
    OPB_CUDA(cudaFuncSetAttribute(estimate_normals_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kKnnMax * kKnnThreads * 8)));
    int per_sm = 0;
    OPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, estimate_normals_kernel, kKnnThreads, knn_smem));
    if (per_sm < 1) per_sm = 1;
    const int nb = (int)((n + kKnnThreads - 1) / kKnnThreads) < c->sm_count * per_sm * 2 ? (int)((n + kKnnThreads - 1) / kKnnThreads) : c->sm_count * per_sm * 2;
    estimate_normals_kernel<<<nb, kKnnThreads, knn_smem, s>>>(c->d_tgt, (int)n, c->d_state, c->d_cell_start, c->d_sorted, radius, knn, c->d_nrm);
    OPB_CUDA(cudaGetLastError());
    OPB_CUDA(cudaMemcpyAsync(normals, c->d_nrm, n * 3 * sizeof(float), cudaMemcpyDefault, s));
    OPB_CUDA(cudaStreamSynchronize(s));
    return OPB_OK;
}
int opb_icp_last_launch_count(opb_icp *c, int *launches)
{
    if (!c || !launches) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    *launches = c->last_launches;
    return OPB_OK;
}
int opb_icp_last_search_trace(opb_icp *c, uint32_t *per_pass, int cap)
{
    if (!c || !per_pass) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    for (int i = 0; i < cap && i < 64; ++i) per_pass[i] = c->last_searched_per_pass[i];
    return OPB_OK;
}
int opb_icp_comm_buffer(opb_icp *c, void **d_buffer, unsigned char ipc_handle[OPB_IPC_HANDLE_BYTES])
{
    if (!c || !d_buffer) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) == OPB_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
    OPB_CUDA(cudaSetDevice(c->device));
    if (!c->d_mailbox)
    {
        OPB_CUDA(cudaMalloc(&c->d_mailbox, sizeof(IcpMailbox)));
        OPB_CUDA(cudaMemset(c->d_mailbox, 0, sizeof(IcpMailbox)));
    }
    *d_buffer = c->d_mailbox;
    if (ipc_handle)
    {
        cudaIpcMemHandle_t h;
        OPB_CUDA(cudaIpcGetMemHandle(&h, c->d_mailbox));
        memcpy(ipc_handle, &h, sizeof(h));
    }
    return OPB_OK;
}
int opb_ipc_open(int device, const unsigned char ipc_handle[OPB_IPC_HANDLE_BYTES], void **d_ptr)
{
    if (!ipc_handle || !d_ptr) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    OPB_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return OPB_OK;
}
int opb_ipc_close(int device, void *d_ptr)
{
    if (!d_ptr) return OPB_OK;
    OPB_CUDA(cudaSetDevice(device));
    OPB_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return OPB_OK;
}
int opb_icp_comm_attach(opb_icp *c, int rank, int world, void *const *buffers)
{
    if (!c || !buffers) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) { set_error("rank %d of world %d is out of range (max %d ranks)", rank, world, kMaxRanks); return OPB_ERR_INVALID; }
    if (!c->d_mailbox || buffers[rank] != (void *)c->d_mailbox) { set_error("buffers[rank] must be this workspace's own opb_icp_comm_buffer"); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(c->device));
    OPB_CUDA(cudaStreamSynchronize(c->stream));
    OPB_CUDA(cudaMemset(c->d_mailbox, 0, sizeof(IcpMailbox)));
    c->peers_share_device = false;
    for (int r = 0; r < world; ++r)
    {
        if (!buffers[r]) { set_error("buffers[%d] is NULL", r); return OPB_ERR_INVALID; }
        c->comm.box[r] = (IcpMailbox *)buffers[r];
        // A peer on this very GPU (tests, oversubscription): its persistent loop kernel and ours could not be resident at the
        // same time, and each waits for the other's packet -- such workspaces run the loop as separate launches instead.
        cudaPointerAttributes attr;
        if (r != rank && cudaPointerGetAttributes(&attr, buffers[r]) == cudaSuccess && attr.device == c->device) c->peers_share_device = true;
    }
    cudaGetLastError();
    c->comm.rank = rank;
    c->comm.world = world;
    return OPB_OK;
}
int opb_icp_comm_detach(opb_icp *c)
{
    if (!c) { set_error("icp is NULL"); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(c->device));
    OPB_CUDA(cudaStreamSynchronize(c->stream));
    c->comm = IcpComm{};
    c->peers_share_device = false;
    return OPB_OK;
}
int opb_icp_last_stamps(opb_icp *c, uint64_t *stamps, int passes)
{
    if (!c || !stamps) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    for (int p = 0; p < passes && p < kStampPasses; ++p)
        for (int k = 0; k < 8; ++k) stamps[p * 8 + k] = c->h_state->stamps[p][k];
    return OPB_OK;
}
int opb_icp_last_prev_pose(opb_icp *c, float T[16])
{
    if (!c || !T) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    memcpy(T, c->h_state->T_prev, 16 * sizeof(float));
    return OPB_OK;
}
// nearest-neighbour indices of the LAST search (the last iteration's, which the closing CountInliers re-tests), for tests
int opb_icp_last_nn(opb_icp *c, int32_t *nn, size_t n)
{
    if (!c || !nn) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    if (n > c->cap_src) { set_error("n exceeds the last source size"); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(c->device));
    OPB_CUDA(cudaMemcpy(nn, c->d_nn, n * sizeof(int), cudaMemcpyDeviceToHost));
    return OPB_OK;
}
} // extern "C"
