// geometry::KDTree<3> on the device, with the reference's exact visiting order, and what the reference builds on it:
// PointCloud::EstimateNormals and registration::ComputeFPFHFeature.
//
// The reference searches through a (locally modified) nanoflann.  Which neighbours come back, and in which order, is not
// a function of the distances alone: equal distances are ordered by the tree traversal, the radius search stops after
// (size_t)(2.5 k) hits IN TRAVERSAL ORDER (with DenseSlam's own parameters that is the normal case, not a corner), and the
// hits are then ordered by libstdc++'s unstable std::sort.  The float sums downstream (FitPlane's covariance, the FPFH
// weighted histogram sum) depend on that order, so a bit-identical result needs the same tree and the same walk:
//
//   build     nanoflann.hpp:843-1010 (computeMinMax, divideTree, middleSplit_, planeSplit), leaf size 10 (KDTree.h:74)
//             level-synchronous: one CTA per node, min/max by block reduction, the two Hoare partition passes of planeSplit
//             reproduced exactly in parallel -- the k-th out-of-place element from the left always meets the k-th from the
//             right, so ranks from a block scan give the same permutation as the sequential pointer walk
//   search    nanoflann.hpp:1012-1028,1228-1295,1354-1417 (computeInitialDistances, searchLevel) as an explicit-stack walk,
//             one thread per query; KNNResultSet (:150-204) in shared memory, RadiusResultSet with the reference's
//             max_neighbors stop (:216-262) in a global scratch; std::sort restated (introsort, median of three, threshold
//             16, heap-sort fallback at depth 2 log2 n, final insertion sort: bits/stl_algo.h, bits/stl_heap.h)
//   users     KDTree.h:93-256 (KnnSearch, RadiusSearch, KnnRadiusSearch); PointCloud.cpp:102-144; 3DFeature.cpp:7-131
//
// All float arithmetic that decides a comparison uses the non-contracting intrinsics in the reference's operation order.
#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/onepiece_b200.h"
#include "opb_common.cuh"
#include "opb_fitplane.cuh"

namespace opb
{
constexpr int kLeafMax = 10;
constexpr int kBuildThreadsBig = 1024, kBuildThreadsMid = 256, kBuildThreadsSmall = 64; // CTA size by the level's largest node
constexpr int kBuildBigNode = 8192, kBuildSmallNode = 256;
constexpr int kMaxDepth = 64;      // search stack, 3-D trees (a 307,200-point frame is 23 levels deep)
constexpr int kMaxDepthRows = 256; // descriptor trees: middle splits of skewed histograms run 70-80 levels deep at 5,000 rows
constexpr int kMaxDim = 33;       // 3 (points, KDTree<3>) or 33 (FPFH descriptors, KDTree<33>)
constexpr int kQueryThreads = 128;
constexpr int kKnnCap = 64;       // k of the shared-memory k-nearest list
constexpr int kRadiusCapMax = 1024; // (int)(2.5 k) of the radius search
constexpr int kQueryBatch = 32768;

struct KdNode
{
    int left, right, child1, child2, divfeat;
    float divlow, divhigh;
    int level;
};
struct KdBuildCtl
{
    int n_nodes, queue_count[2], queue_max[2], max_level; // queue_max: the largest node waiting in that queue
    float root_lo[kMaxDim], root_hi[kMaxDim];
};
struct KdView
{
    const float *pts;
    const float4 *sorted;     // DIM 3: the points in vind order, w = the point's index: a leaf is one contiguous run
    const float *sorted_rows; // DIM 33: the rows in vind order
    const int *vind;
    const KdNode *nodes;
    float root_lo[kMaxDim], root_hi[kMaxDim];
    int n;
};
template <int DIM>
__device__ __forceinline__ float pick(const float *v, int i)
{
    if (DIM == 3) return i == 0 ? v[0] : i == 1 ? v[1] : v[2];
    return v[i];
}
template <int DIM>
__device__ __forceinline__ void put(float *v, int i, float x)
{
    if (DIM == 3)
    {
        if (i == 0) v[0] = x;
        else if (i == 1) v[1] = x;
        else v[2] = x;
    }
    else v[i] = x;
}

// ---------------------------------------------------------------------------------------------------------
// build
// ---------------------------------------------------------------------------------------------------------
template <int THREADS, class T, class Op>
__device__ __forceinline__ T block_reduce(T v, Op op, T *sh)
{
    constexpr int kBuildWarps = THREADS / 32;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    T r = sh[0];
#pragma unroll
    for (int w = 1; w < kBuildWarps; ++w) r = op(r, sh[w]);
    return r;
}
template <int THREADS>
__device__ __forceinline__ int block_exscan(int v, int *sh, int &total)
{
    constexpr int kBuildWarps = THREADS / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();
    if (lane == 31) sh[warp] = incl;
    __syncthreads();
    int base = 0;
    total = 0;
#pragma unroll
    for (int w = 0; w < kBuildWarps; ++w)
    {
        if (w < warp) base += sh[w];
        total += sh[w];
    }
    return base + incl - v;
}
// One Hoare pass of planeSplit (nanoflann.hpp:978-991 / :996-1008) over [begin, end): afterwards every element with `pred`
// sits below `boundary` (= begin + their count).  LE selects the second pass's predicate (<= cutval) over the first (<).
template <int THREADS, bool LE>
__device__ __forceinline__ void partition_pass(int *ind, float *key, int *pos, int begin, int end, int boundary, float cutval, int *sh)
{
    const int len = end - begin;
    const int seg = (len + THREADS - 1) / THREADS;
    const int a = min(begin + (int)threadIdx.x * seg, end), b = min(a + seg, end);
    int cl = 0, cr = 0;
    for (int i = a; i < b; ++i)
    {
        const bool p = LE ? key[i] <= cutval : key[i] < cutval;
        cl += (i < boundary && !p);
        cr += (i >= boundary && p);
    }
    int m, m2;
    int bl = block_exscan<THREADS>(cl, sh, m);
    int br = block_exscan<THREADS>(cr, sh, m2);
    int *pos_l = pos + begin, *pos_r = pos + begin + (len + 1) / 2; // m <= len / 2: the two lists cannot overlap
    for (int i = a; i < b; ++i)
    {
        const bool p = LE ? key[i] <= cutval : key[i] < cutval;
        if (i < boundary && !p) pos_l[bl++] = i;
        if (i >= boundary && p) pos_r[br++] = i;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < m; k += THREADS)
    {
        const int i = pos_l[k], j = pos_r[m - 1 - k];
        const int ti = ind[i]; ind[i] = ind[j]; ind[j] = ti;
        const float tk = key[i]; key[i] = key[j]; key[j] = tk;
    }
    __syncthreads();
}
__global__ void kd_iota_kernel(int *vind, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) vind[i] = i;
}
// divideTree (nanoflann.hpp:864-914) for one node per CTA; boxes = the LOOSE box handed down by the parent (it, not the tight
// one, picks the cut dimension and the middle), 2 DIM floats per node
template <int THREADS, int DIM>
__global__ void __launch_bounds__(THREADS) kd_split_kernel(const float *__restrict__ pts, int *vind, float *key, int *pos, KdNode *nodes,
                                                                 float *boxes, const int *__restrict__ queue_in, int *queue_out, KdBuildCtl *ctl,
                                                                 int out_slot)
{
    __shared__ float shf[THREADS / 32];
    __shared__ int shi[THREADS / 32];
    const int id = queue_in[blockIdx.x];
    const int left = nodes[id].left, right = nodes[id].right, count = right - left, level = nodes[id].level;
    int *ind = vind + left;
    float *ky = key + left;
    // computeMinMax of every coordinate
    float mn[DIM], mx[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) { mn[d] = FLT_MAX; mx[d] = -FLT_MAX; }
    for (int i = threadIdx.x; i < count; i += THREADS)
    {
        const int j = ind[i];
#pragma unroll
        for (int d = 0; d < DIM; ++d)
        {
            const float v = pts[(size_t)DIM * j + d];
            mn[d] = fminf(mn[d], v);
            mx[d] = fmaxf(mx[d], v);
        }
    }
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        mn[d] = block_reduce<THREADS>(mn[d], [](float x, float y) { return fminf(x, y); }, shf);
        mx[d] = block_reduce<THREADS>(mx[d], [](float x, float y) { return fmaxf(x, y); }, shf);
    }
    float lo[DIM], hi[DIM];
    if (id == 0)
    {
        // computeBoundingBox (:1324-1350): the root's box is the tight one
#pragma unroll
        for (int d = 0; d < DIM; ++d) { lo[d] = mn[d]; hi[d] = mx[d]; }
        if (threadIdx.x == 0)
            for (int d = 0; d < DIM; ++d) { ctl->root_lo[d] = mn[d]; ctl->root_hi[d] = mx[d]; }
    }
    else
    {
#pragma unroll
        for (int d = 0; d < DIM; ++d) { lo[d] = boxes[(size_t)2 * DIM * id + d]; hi[d] = boxes[(size_t)2 * DIM * id + DIM + d]; }
    }
    if (count <= kLeafMax) return; // only the root can arrive here as a leaf
    // middleSplit_ (:916-963)
    float max_span = fsub(hi[0], lo[0]);
#pragma unroll
    for (int d = 1; d < DIM; ++d)
    {
        const float span = fsub(hi[d], lo[d]);
        if (span > max_span) max_span = span;
    }
    const float thresh = fmul(fsub(1.0f, 0.00001f), max_span);
    float max_spread = -1.0f;
    int cutfeat = 0;
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        const float span = fsub(hi[d], lo[d]);
        if (span > thresh)
        {
            const float spread = fsub(mx[d], mn[d]);
            if (spread > max_spread) { cutfeat = d; max_spread = spread; }
        }
    }
    const float lo_c = pick<DIM>(lo, cutfeat), hi_c = pick<DIM>(hi, cutfeat), mn_c = pick<DIM>(mn, cutfeat), mx_c = pick<DIM>(mx, cutfeat);
    const float split_val = fdiv(fadd(lo_c, hi_c), 2.0f);
    const float cutval = split_val < mn_c ? mn_c : split_val > mx_c ? mx_c : split_val;
    // planeSplit (:974-1010): lim1 = #(< cutval), lim2 = lim1 + #(== cutval)
    int cl = 0, ce = 0;
    for (int i = threadIdx.x; i < count; i += THREADS)
    {
        const float v = pts[(size_t)DIM * ind[i] + cutfeat];
        ky[i] = v;
        cl += v < cutval;
        ce += v == cutval;
    }
    const int lim1 = block_reduce<THREADS>(cl, [](int x, int y) { return x + y; }, shi);
    const int lim2 = lim1 + block_reduce<THREADS>(ce, [](int x, int y) { return x + y; }, shi);
    partition_pass<THREADS, false>(ind, ky, pos + left, 0, count, lim1, cutval, shi);
    partition_pass<THREADS, true>(ind, ky, pos + left, lim1, count, lim2, cutval, shi);
    const int half = count / 2;
    const int idx = lim1 > half ? lim1 : lim2 < half ? lim2 : half;
    // the children's tight extent along the cut: divlow = left_bbox[cutfeat].high, divhigh = right_bbox[cutfeat].low (:904-905)
    float dl = -FLT_MAX, dh = FLT_MAX;
    for (int i = threadIdx.x; i < count; i += THREADS)
    {
        const float v = ky[i];
        if (i < idx) dl = fmaxf(dl, v);
        else dh = fminf(dh, v);
    }
    dl = block_reduce<THREADS>(dl, [](float x, float y) { return fmaxf(x, y); }, shf);
    dh = block_reduce<THREADS>(dh, [](float x, float y) { return fminf(x, y); }, shf);
    if (threadIdx.x == 0)
    {
        const int c = atomicAdd(&ctl->n_nodes, 2);
        KdNode a, b;
        a.left = left; a.right = left + idx; a.child1 = a.child2 = -1; a.divfeat = -1; a.divlow = a.divhigh = 0.0f; a.level = level + 1;
        b = a;
        b.left = left + idx; b.right = right;
        nodes[c] = a;
        nodes[c + 1] = b;
        float *box_a = boxes + (size_t)2 * DIM * c, *box_b = box_a + 2 * DIM;
        for (int d = 0; d < DIM; ++d)
        {
            box_a[d] = lo[d]; box_a[DIM + d] = d == cutfeat ? cutval : hi[d];
            box_b[d] = d == cutfeat ? cutval : lo[d]; box_b[DIM + d] = hi[d];
        }
        nodes[id].child1 = c; nodes[id].child2 = c + 1; nodes[id].divfeat = cutfeat;
        nodes[id].divlow = dl; nodes[id].divhigh = dh;
        if (idx > kLeafMax) queue_out[atomicAdd(&ctl->queue_count[out_slot], 1)] = c;
        if (count - idx > kLeafMax) queue_out[atomicAdd(&ctl->queue_count[out_slot], 1)] = c + 1;
        atomicMax(&ctl->queue_max[out_slot], max(idx, count - idx));
        atomicMax(&ctl->max_level, level + 1);
    }
}

__global__ void kd_reorder_rows_kernel(const float *__restrict__ pts, const int *__restrict__ vind, int n, int dim, float *__restrict__ rows)
{
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < (size_t)n * dim; e += (size_t)gridDim.x * blockDim.x)
        rows[e] = pts[(size_t)vind[e / dim] * dim + e % dim];
}
__global__ void kd_reorder_kernel(const float *__restrict__ pts, const int *__restrict__ vind, int n, float4 *__restrict__ sorted)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const int v = vind[i];
        sorted[i] = make_float4(pts[3 * v], pts[3 * v + 1], pts[3 * v + 2], __int_as_float(v));
    }
}

// ---------------------------------------------------------------------------------------------------------
// search
// ---------------------------------------------------------------------------------------------------------
// L2_Simple_Adaptor::evalMetric (:438-446): result += diff * diff, dimension by dimension
__device__ __forceinline__ float kd_dist2(const float *q, const float4 p)
{
    const float dx = fsub(q[0], p.x), dy = fsub(q[1], p.y), dz = fsub(q[2], p.z);
    return fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
}
// KNNResultSet (:150-204): a new point goes BEHIND the stored points of equal distance
struct KnnSet
{
    float *d;
    int *i;
    int count, capacity;
    __device__ __forceinline__ float &dist(int e) { return d[e * kQueryThreads]; }
    __device__ __forceinline__ int &index(int e) { return i[e * kQueryThreads]; }
    __device__ __forceinline__ void init(float *smem, int k)
    {
        d = smem + threadIdx.x;
        i = reinterpret_cast<int *>(smem + k * kQueryThreads) + threadIdx.x;
        count = 0;
        capacity = k;
        dist(k - 1) = FLT_MAX;
    }
    __device__ __forceinline__ float worst() { return dist(capacity - 1); }
    __device__ __forceinline__ bool add(float dd, int idx)
    {
        int e;
        for (e = count; e > 0; --e)
        {
            if (dist(e - 1) > dd)
            {
                if (e < capacity) { dist(e) = dist(e - 1); index(e) = index(e - 1); }
            }
            else break;
        }
        if (e < capacity) { dist(e) = dd; index(e) = idx; }
        if (count < capacity) count++;
        return true;
    }
};
// RadiusResultSet with the reference's early stop (:216-262); entries live in a global scratch, entry e of query slot s at
// [e * stride + s]
struct RadiusSet
{
    float *d;
    int *i;
    int stride, count, capacity;
    float radius;
    __device__ __forceinline__ float &dist(int e) { return d[(size_t)e * stride]; }
    __device__ __forceinline__ int &index(int e) { return i[(size_t)e * stride]; }
    __device__ __forceinline__ float worst() { return radius; }
    __device__ __forceinline__ bool add(float dd, int idx)
    {
        if (capacity > 0 && count >= capacity) return false;
        if (dd < radius) { dist(count) = dd; index(count) = idx; ++count; }
        return true;
    }
};
// one level of searchLevel's recursion, 16 bytes so that a frame moves as one vector access: `node` is the node to visit until
// it has been expanded and the far child afterwards; `value` is the node's mindistsq until the near child returns and the saved
// side distance afterwards
struct __align__(16) KdFrame
{
    int node, idx_stage;
    float value, cut;
};
// findNeighbors + searchLevel (:1228-1248, :1354-1417); false when the result set asked to stop
template <int DIM, class ResultSet>
__device__ bool kd_find(const KdView &t, ResultSet &rs, const float *q, float eps_error)
{
    if (t.n == 0) return true;
    float dists[DIM];
    float distsq = 0.0f;
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        dists[d] = 0.0f;
        if (q[d] < t.root_lo[d]) { const float x = fsub(q[d], t.root_lo[d]); dists[d] = fmul(x, x); distsq = fadd(distsq, dists[d]); }
        if (q[d] > t.root_hi[d]) { const float x = fsub(q[d], t.root_hi[d]); dists[d] = fmul(x, x); distsq = fadd(distsq, dists[d]); }
    }
    constexpr int kStack = DIM == 3 ? kMaxDepth : kMaxDepthRows;
    KdFrame st[kStack];
    int sp = 0;
    st[0].node = 0; st[0].idx_stage = 0; st[0].value = distsq; st[0].cut = 0.0f;
    while (sp >= 0)
    {
        KdFrame f = st[sp];
        const int stage = f.idx_stage & 3, idx = f.idx_stage >> 2;
        if (stage == 0)
        {
            const KdNode nd = t.nodes[f.node];
            if (nd.child1 < 0)
            {
                const float worst = rs.worst();
                for (int i = nd.left; i < nd.right; ++i)
                {
                    if (DIM == 3)
                    {
                        const float4 p = t.sorted[i];
                        const float d = kd_dist2(q, p);
                        if (d < worst)
                            if (!rs.add(d, __float_as_int(p.w))) return false;
                    }
                    else
                    {
                        // L2_Simple_Adaptor::evalMetric (:438-446): result += diff * diff, dimension by dimension
                        const float *row = t.sorted_rows + (size_t)i * DIM;
                        float d = 0.0f;
                        for (int e = 0; e < DIM; ++e)
                        {
                            const float diff = fsub(q[e], row[e]);
                            d = fadd(d, fmul(diff, diff));
                        }
                        if (d < worst)
                            if (!rs.add(d, t.vind[i])) return false;
                    }
                }
                --sp;
                continue;
            }
            const int cf = nd.divfeat;
            const float val = pick<DIM>(q, cf);
            const float diff1 = fsub(val, nd.divlow), diff2 = fsub(val, nd.divhigh);
            int best;
            if (fadd(diff1, diff2) < 0.0f) { best = nd.child1; f.node = nd.child2; f.cut = fmul(diff2, diff2); }
            else { best = nd.child2; f.node = nd.child1; f.cut = fmul(diff1, diff1); }
            f.idx_stage = (cf << 2) | 1;
            st[sp] = f;
            if (sp + 1 >= kStack) return true; // the host refuses trees deeper than the stack before any search
            ++sp;
            st[sp].node = best; st[sp].idx_stage = 0; st[sp].value = f.value; st[sp].cut = 0.0f;
        }
        else if (stage == 1)
        {
            const float dst = pick<DIM>(dists, idx);
            const float mind = fsub(fadd(f.value, f.cut), dst);
            put<DIM>(dists, idx, f.cut);
            if (fmul(mind, eps_error) <= rs.worst())
            {
                f.idx_stage = (idx << 2) | 2;
                f.value = dst;
                st[sp] = f;
                ++sp;
                st[sp].node = f.node; st[sp].idx_stage = 0; st[sp].value = mind; st[sp].cut = 0.0f;
            }
            else
            {
                put<DIM>(dists, idx, dst);
                --sp;
            }
        }
        else
        {
            put<DIM>(dists, idx, f.value);
            --sp;
        }
    }
    return true;
}

// libstdc++ std::sort over the radius hits, compared by distance only (IndexDist_Sorter, nanoflann.hpp:206-214)
struct HitPair { float d; int i; };
struct HitArray
{
    float *d;
    int *i;
    int stride;
    __device__ __forceinline__ HitPair get(int e) const { HitPair p; p.d = d[(size_t)e * stride]; p.i = i[(size_t)e * stride]; return p; }
    __device__ __forceinline__ float key(int e) const { return d[(size_t)e * stride]; }
    __device__ __forceinline__ void set(int e, HitPair p) const { d[(size_t)e * stride] = p.d; i[(size_t)e * stride] = p.i; }
    __device__ __forceinline__ void swap(int a, int b) const { const HitPair x = get(a), y = get(b); set(a, y); set(b, x); }
};
// __adjust_heap + __push_heap (bits/stl_heap.h) relative to `first`
__device__ void hit_adjust_heap(const HitArray &A, int first, int hole, int len, HitPair value)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2)
    {
        child = 2 * (child + 1);
        if (A.key(first + child) < A.key(first + child - 1)) child--;
        A.set(first + hole, A.get(first + child));
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2)
    {
        child = 2 * (child + 1);
        A.set(first + hole, A.get(first + child - 1));
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && A.key(first + parent) < value.d)
    {
        A.set(first + hole, A.get(first + parent));
        hole = parent;
        parent = (hole - 1) / 2;
    }
    A.set(first + hole, value);
}
__device__ void hit_heap_sort(const HitArray &A, int first, int last)
{
    const int len = last - first;
    if (len >= 2)
        for (int parent = (len - 2) / 2;; --parent)
        {
            hit_adjust_heap(A, first, parent, len, A.get(first + parent));
            if (parent == 0) break;
        }
    while (last - first > 1)
    {
        --last;
        const HitPair value = A.get(last);
        A.set(last, A.get(first));
        hit_adjust_heap(A, first, 0, last - first, value);
    }
}
__device__ __forceinline__ void hit_unguarded_linear_insert(const HitArray &A, int last)
{
    const HitPair val = A.get(last);
    int next = last - 1;
    while (val.d < A.key(next))
    {
        A.set(last, A.get(next));
        last = next;
        --next;
    }
    A.set(last, val);
}
__device__ void hit_insertion_sort(const HitArray &A, int first, int last)
{
    if (first == last) return;
    for (int i = first + 1; i != last; ++i)
    {
        if (A.key(i) < A.key(first))
        {
            const HitPair val = A.get(i);
            for (int k = i; k > first; --k) A.set(k, A.get(k - 1)); // move_backward
            A.set(first, val);
        }
        else hit_unguarded_linear_insert(A, i);
    }
}
__device__ void hit_std_sort(const HitArray &A, int n)
{
    if (n == 0) return;
    int lg = 0;
    for (int m = n; m > 1; m >>= 1) ++lg;
    // __introsort_loop: the two halves of a partition are independent, so a work stack visits them in any order; both
    // inherit the decremented depth budget
    int stk_first[48], stk_last[48], stk_depth[48];
    int sp = 0;
    stk_first[0] = 0; stk_last[0] = n; stk_depth[0] = 2 * lg;
    while (sp >= 0)
    {
        int first = stk_first[sp], last = stk_last[sp], depth = stk_depth[sp];
        --sp;
        while (last - first > 16)
        {
            if (depth == 0) { hit_heap_sort(A, first, last); break; }
            --depth;
            // __move_median_to_first(first, first + 1, mid, last - 1)
            const int a = first + 1, b = first + (last - first) / 2, c = last - 1;
            const float ka = A.key(a), kb = A.key(b), kc = A.key(c);
            if (ka < kb)
            {
                if (kb < kc) A.swap(first, b);
                else if (ka < kc) A.swap(first, c);
                else A.swap(first, a);
            }
            else if (ka < kc) A.swap(first, a);
            else if (kb < kc) A.swap(first, c);
            else A.swap(first, b);
            // __unguarded_partition(first + 1, last, pivot = *first)
            const float pivot = A.key(first);
            int lo = first + 1, hi = last;
            for (;;)
            {
                while (A.key(lo) < pivot) ++lo;
                --hi;
                while (pivot < A.key(hi)) --hi;
                if (!(lo < hi)) break;
                A.swap(lo, hi);
                ++lo;
            }
            if (sp + 1 < 48) { ++sp; stk_first[sp] = lo; stk_last[sp] = last; stk_depth[sp] = depth; }
            last = lo;
        }
    }
    // __final_insertion_sort
    if (n > 16)
    {
        hit_insertion_sort(A, 0, 16);
        for (int i = 16; i != n; ++i) hit_unguarded_linear_insert(A, i);
    }
    else hit_insertion_sort(A, 0, n);
}

// KnnSearch (mode 0) / KnnRadiusSearch (mode 2), KDTree.h:176-195,230-256: rows of k entries, -1 padded
__global__ void __launch_bounds__(kQueryThreads) kd_knn_kernel(KdView t, const float *__restrict__ queries, int nq, int mode, int k, float radius,
                                                               int *__restrict__ out_index, float *__restrict__ out_dist, int *__restrict__ out_count)
{
    extern __shared__ float knn_smem[];
    for (int qi = blockIdx.x * blockDim.x + threadIdx.x; qi < nq; qi += gridDim.x * blockDim.x)
    {
        const float q[3] = {queries[3 * qi], queries[3 * qi + 1], queries[3 * qi + 2]};
        KnnSet rs;
        rs.init(knn_smem, k);
        kd_find<3>(t, rs, q, 1.0f);
        int cnt = rs.count;
        if (mode == 2)
        {
            int in_radius = 0;
            for (; in_radius != cnt; ++in_radius)
                if (rs.dist(in_radius) > radius) break;
            cnt = in_radius;
        }
        out_count[qi] = cnt;
        for (int e = 0; e < k; ++e)
        {
            out_index[(size_t)qi * k + e] = e < cnt ? rs.index(e) : -1;
            out_dist[(size_t)qi * k + e] = e < cnt ? rs.dist(e) : -1.0f;
        }
    }
}
// RadiusSearch (KDTree.h:125-143): up to `cap` hits with dist^2 < radius in traversal order, std::sort, the first k kept.
// Queries q0 .. q0 + nq of one batch (taken in the tree's leaf order when they are the tree's own points: `order` = vind, so the
// threads of a warp walk the same part of the tree).  The hits of a thread live in shared memory (cap x blockDim entries) when
// scratch_d is NULL, else in a global scratch of cap x stride entries.
__global__ void __launch_bounds__(kQueryThreads) kd_radius_kernel(KdView t, const float *__restrict__ queries, const int *__restrict__ order, int q0,
                                                                  int nq, int k, int cap, float radius, float *scratch_d, int *scratch_i, int stride,
                                                                  int *__restrict__ out_index, float *__restrict__ out_dist,
                                                                  int *__restrict__ out_count)
{
    extern __shared__ float knn_smem[];
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nq; s += gridDim.x * blockDim.x)
    {
        const int qi = order ? order[q0 + s] : q0 + s;
        const float q[3] = {queries[3 * qi], queries[3 * qi + 1], queries[3 * qi + 2]};
        RadiusSet rs;
        if (scratch_d) { rs.d = scratch_d + s; rs.i = scratch_i + s; rs.stride = stride; }
        else { rs.d = knn_smem + threadIdx.x; rs.i = reinterpret_cast<int *>(knn_smem + cap * blockDim.x) + threadIdx.x; rs.stride = blockDim.x; }
        rs.count = 0; rs.capacity = cap; rs.radius = radius;
        kd_find<3>(t, rs, q, 1.0f + 1e-8f); // SearchParameter's eps (KDTree.h:19): 1 + 1e-8 rounds to 1 in float, like the reference's
        HitArray A;
        A.d = rs.d; A.i = rs.i; A.stride = rs.stride;
        hit_std_sort(A, rs.count);
        const int cnt = rs.count > k ? k : rs.count;
        out_count[qi] = cnt;
        for (int e = 0; e < k; ++e)
        {
            out_index[(size_t)qi * k + e] = e < cnt ? rs.index(e) : -1;
            if (out_dist) out_dist[(size_t)qi * k + e] = e < cnt ? rs.dist(e) : -1.0f;
        }
    }
}
// PointCloud::EstimateNormals (PointCloud.cpp:102-144): KnnRadiusSearch(knn, radius) around every point of the tree, FitPlane
__global__ void __launch_bounds__(kQueryThreads) kd_normals_kernel(KdView t, int k, float radius, float *__restrict__ normals)
{
    extern __shared__ float knn_smem[];
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < t.n; slot += gridDim.x * blockDim.x)
    {
        // queries in the tree's leaf order: the threads of a warp walk the same part of the tree
        const float4 self = t.sorted[slot];
        const int qi = __float_as_int(self.w);
        const float q[3] = {self.x, self.y, self.z};
        KnnSet rs;
        rs.init(knn_smem, k);
        kd_find<3>(t, rs, q, 1.0f);
        int in_radius = 0;
        for (; in_radius != rs.count; ++in_radius)
            if (rs.dist(in_radius) > radius) break;
        float nrm[3];
        fit_plane_normal(t.pts, in_radius, [&](int e) { return rs.index(e); }, nrm);
        normals[3 * qi] = nrm[0]; normals[3 * qi + 1] = nrm[1]; normals[3 * qi + 2] = nrm[2];
    }
}

// ---------------------------------------------------------------------------------------------------------
// FPFH (3DFeature.cpp:7-131)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float dot3e(const float *a, const float *b) { return fadd(fmul(a[0], b[0]), fadd(fmul(a[1], b[1]), fmul(a[2], b[2]))); } // Eigen: x + (y + z)
__device__ __forceinline__ void cross3e(const float *a, const float *b, float *o)
{
    o[0] = fsub(fmul(a[1], b[2]), fmul(a[2], b[1]));
    o[1] = fsub(fmul(a[2], b[0]), fmul(a[0], b[2]));
    o[2] = fsub(fmul(a[0], b[1]), fmul(a[1], b[0]));
}
// ComputePairDescriptor (:7-25) reduced to the three histogram bins ComputeSPFH takes from it (:63-75)
__device__ __forceinline__ void fpfh_pair_bins(const float *ps, const float *ns, const float *pt, const float *nt, int *bins)
{
    const float d[3] = {fsub(ps[0], pt[0]), fsub(ps[1], pt[1]), fsub(ps[2], pt[2])};
    const float distance = __fsqrt_rn(dot3e(d, d));
    const float diff[3] = {fsub(pt[0], ps[0]), fsub(pt[1], ps[1]), fsub(pt[2], ps[2])};
    const float dir[3] = {fdiv(diff[0], distance), fdiv(diff[1], distance), fdiv(diff[2], distance)};
    float v[3], w[3], desc[3] = {0.0f, 0.0f, 0.0f};
    cross3e(ns, dir, v);
    if (__fsqrt_rn(dot3e(v, v)) != 0.0f)
    {
        cross3e(ns, v, w);
        desc[1] = dot3e(v, nt);
        desc[2] = fdiv(dot3e(ns, diff), distance);
        // the reference's unqualified atan2 is the double one; its float rounding is what reaches the histogram
        desc[0] = (float)atan2((double)dot3e(w, nt), (double)dot3e(ns, nt));
    }
    const double kPi = 3.14159265358979323846;
    bins[0] = (int)floor(11 * ((double)desc[0] + kPi) / (2.0 * kPi));
    bins[1] = (int)floor((double)fmul(11.0f, fadd(desc[1], 1.0f)) / 2.0);
    bins[2] = (int)floor((double)fmul(11.0f, fadd(desc[2], 1.0f)) / 2.0);
#pragma unroll
    for (int b = 0; b < 3; ++b) bins[b] = bins[b] > 10 ? 10 : bins[b] < 0 ? 0 : bins[b];
}
// ComputeSPFH (:28-81) after the radius search: nbr rows hold the search result (first hit = the point itself, skipped).
// Every pair adds the same integer 100 / (points_num - 1) to one bin per third, so a bin is count x increment exactly.
__global__ void __launch_bounds__(kQueryThreads) fpfh_spfh_kernel(const float *__restrict__ pts, const float *__restrict__ nrm, int n, int knn,
                                                                  const int *__restrict__ nbr, const int *__restrict__ nbr_count,
                                                                  float *__restrict__ spfh)
{
    __shared__ unsigned char hits[33 * kQueryThreads];
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x)
    {
        const int i = base + threadIdx.x;
        for (int b = 0; b < 33; ++b) hits[b * kQueryThreads + threadIdx.x] = 0;
        if (i < n)
        {
            int points_num = nbr_count[i];
            float each = 0.0f;
            if (points_num - 1 > 0)
            {
                if (points_num > knn) points_num = knn;
                each = (float)(100 / (points_num - 1));
                const float ps[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]}, ns[3] = {nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]};
                for (int j = 1; j != points_num; ++j)
                {
                    const int t = nbr[(size_t)i * knn + j];
                    const float pt[3] = {pts[3 * t], pts[3 * t + 1], pts[3 * t + 2]}, nt[3] = {nrm[3 * t], nrm[3 * t + 1], nrm[3 * t + 2]};
                    int bins[3];
                    fpfh_pair_bins(ps, ns, pt, nt, bins);
                    hits[bins[0] * kQueryThreads + threadIdx.x]++;          // knn <= 256 neighbours: a byte holds the count
                    hits[(bins[1] + 11) * kQueryThreads + threadIdx.x]++;
                    hits[(bins[2] + 22) * kQueryThreads + threadIdx.x]++;
                }
            }
            for (int b = 0; b < 33; ++b) spfh[(size_t)i * 33 + b] = fmul((float)hits[b * kQueryThreads + threadIdx.x], each);
        }
    }
}
// ComputeFPFHFeature's second loop (:104-130): one warp per point, lane e owns histogram element e (lane 0 also element 32)
// and walks the neighbours in list order, so every element sees the reference's sequence of float additions
__global__ void __launch_bounds__(kQueryThreads) fpfh_combine_kernel(const float *__restrict__ pts, int n, int knn, const int *__restrict__ nbr,
                                                                     const int *__restrict__ nbr_count, const float *__restrict__ spfh,
                                                                     float *__restrict__ fpfh)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps)
    {
        int points_num = nbr_count[i];
        if (points_num > knn) points_num = knn;
        const float p[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
        float f = 0.0f, f32 = 0.0f;
        double sum[3] = {0.0, 0.0, 0.0};
        for (int j = 1; j < points_num; ++j)
        {
            const int t = nbr[(size_t)i * knn + j];
            const float d[3] = {fsub(p[0], pts[3 * t]), fsub(p[1], pts[3 * t + 1]), fsub(p[2], pts[3 * t + 2])};
            const float dist = __fsqrt_rn(dot3e(d, d));
            if (dist != 0.0f)
            {
                const float w_d = fdiv(1.0f, dist);
                const float s = spfh[(size_t)t * 33 + lane];
                f = fadd(f, fmul(w_d, s));
                const float s32 = spfh[(size_t)t * 33 + 32];
                f32 = fadd(f32, fmul(w_d, s32));
                // block<11,1>.sum() of integer-valued bins: exact in any order
                const float in0 = lane < 11 ? s : 0.0f, in1 = lane >= 11 && lane < 22 ? s : 0.0f, in2 = lane >= 22 ? s : 0.0f;
                float b0 = in0, b1 = in1, b2 = in2;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                {
                    b0 += __shfl_xor_sync(0xFFFFFFFFu, b0, o);
                    b1 += __shfl_xor_sync(0xFFFFFFFFu, b1, o);
                    b2 += __shfl_xor_sync(0xFFFFFFFFu, b2, o);
                }
                sum[0] += (double)b0; sum[1] += (double)b1; sum[2] += (double)(b2 + s32);
            }
        }
        const float scale0 = (float)(100.0 / sum[0]), scale1 = (float)(100.0 / sum[1]), scale2 = (float)(100.0 / sum[2]);
        const float scale = lane < 11 ? scale0 : lane < 22 ? scale1 : scale2;
        f = fadd(fmul(f, scale), spfh[(size_t)i * 33 + lane]);
        fpfh[(size_t)i * 33 + lane] = f;
        if (lane == 0) fpfh[(size_t)i * 33 + 32] = fadd(fmul(f32, scale2), spfh[(size_t)i * 33 + 32]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// registration::FeatureMatching3D (src/Registration/GlobalRegistration.cpp:29-73): nearest target feature of every source
// feature in 33 dimensions.  The reference asks a KDTree<33> for k = 1; the nearest neighbour is what an exhaustive scan finds,
// with nanoflann's metric (L2_Simple_Adaptor::evalMetric: result += diff * diff, dimension by dimension) and its acceptance rule
// (dist < worst, worst starting at FLT_MAX: a NaN feature matches nothing).  One thread per source feature, the targets pass
// through shared memory in tiles read as broadcasts.
// ---------------------------------------------------------------------------------------------------------
constexpr int kMatchThreads = 64, kMatchTile = 64, kFeatureDim = 33;
__global__ void __launch_bounds__(kMatchThreads) fpfh_match_kernel(const float *__restrict__ src, int ns, const float *__restrict__ tgt, int nt,
                                                                   int *__restrict__ nearest)
{
    __shared__ float tile[kMatchTile * kFeatureDim];
    const int i = blockIdx.x * kMatchThreads + threadIdx.x;
    float f[kFeatureDim];
#pragma unroll
    for (int e = 0; e < kFeatureDim; ++e) f[e] = i < ns ? src[(size_t)i * kFeatureDim + e] : 0.0f;
    float best = FLT_MAX;
    int best_j = -1;
    for (int j0 = 0; j0 < nt; j0 += kMatchTile)
    {
        const int m = min(kMatchTile, nt - j0);
        __syncthreads();
        for (int e = threadIdx.x; e < m * kFeatureDim; e += kMatchThreads) tile[e] = tgt[(size_t)j0 * kFeatureDim + e];
        __syncthreads();
        for (int j = 0; j < m; ++j)
        {
            float d = 0.0f;
#pragma unroll
            for (int e = 0; e < kFeatureDim; ++e)
            {
                const float diff = fsub(f[e], tile[j * kFeatureDim + e]);
                d = fadd(d, fmul(diff, diff));
            }
            if (d < best) { best = d; best_j = j0 + j; }
        }
    }
    if (i < ns) nearest[i] = best_j;
}
// the same question asked the reference's way: a KDTree<33> over the (finite) target descriptors, walked per source descriptor
// with a one-entry KNNResultSet -- so that exactly equidistant targets (duplicate descriptors on flat surfaces) come back in the
// reference's order too
struct Nearest1
{
    float best;
    int index;
    __device__ __forceinline__ float worst() { return best; }
    // KNNResultSet::addPoint with capacity 1 (:175-199): only a strictly closer point replaces the entry -- the leaf loop's
    // worst_dist is read once per leaf (:1361), so farther points of the same leaf do arrive here
    __device__ __forceinline__ bool add(float d, int i)
    {
        if (best > d) { best = d; index = i; }
        return true;
    }
};
__global__ void __launch_bounds__(kMatchThreads) fpfh_match_tree_kernel(KdView t, const float *__restrict__ src, int ns, int *__restrict__ nearest)
{
    const int i = blockIdx.x * kMatchThreads + threadIdx.x;
    if (i >= ns) return;
    float q[kFeatureDim];
    for (int e = 0; e < kFeatureDim; ++e) q[e] = src[(size_t)i * kFeatureDim + e];
    Nearest1 rs;
    rs.best = FLT_MAX;
    rs.index = -1;
    kd_find<kFeatureDim>(t, rs, q, 1.0f);
    nearest[i] = rs.index;
}
// ---------------------------------------------------------------------------------------------------------
// geometry::EstimateRigidTransformationRANSAC (src/Geometry/Ransac.cpp:7-41) on 3rdparty/GRANSAC/GRANSAC.hpp:71-131 with
// TransformationModel (src/Geometry/TransformationModel.hpp:28-96): every iteration takes eight distinct pairs, fits a rigid
// motion to them (EstimateRigidTransformation, float) and scores it by the number of pairs with |R a + t - b| < threshold;
// the FIRST iteration with the strictly largest score wins (GRANSAC.hpp:113-122) and its eight-point motion is the result.
// The iterations are independent: one thread per hypothesis, all threads of a warp stream the same pair at the same time
// (broadcast loads).  GRANSAC seeds its engines from std::random_device, so the reference's choice of samples is not
// reproducible even by itself; here the samples come from a counter-based generator (or from the caller, which is how the
// parity tests force the same hypotheses through the oracle).
// ---------------------------------------------------------------------------------------------------------
constexpr int kRansacSample = 8; // MIN_INLIER_SIZE_RANSAC_TRANSFORMATION, TransformationModel.hpp:5
__device__ __forceinline__ unsigned int ransac_mix(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull; // splitmix64
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return (unsigned int)((x ^ (x >> 31)) >> 32);
}
// eight distinct indices, each uniform over what is left: the distribution of the first eight entries of a uniform shuffle
__device__ __forceinline__ void ransac_draw(unsigned long long seed, int iteration, int n, int *s8)
{
    for (int k = 0; k < kRansacSample; ++k)
        for (unsigned int attempt = 0;; ++attempt)
        {
            const unsigned int r = ransac_mix(seed ^ ((unsigned long long)iteration << 24) ^ ((unsigned long long)k << 20) ^ attempt);
            const int idx = (int)(((unsigned long long)r * (unsigned long long)n) >> 32);
            bool seen = false;
            for (int j = 0; j < k; ++j) seen |= s8[j] == idx;
            if (!seen) { s8[k] = idx; break; }
        }
}
__device__ __forceinline__ bool ransac_is_inlier(const float *R, const float *t, const float *__restrict__ a, const float *__restrict__ b, int i,
                                                 double threshold)
{
    const float ax = a[3 * i], ay = a[3 * i + 1], az = a[3 * i + 2];
    float e[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
        e[r] = fsub(fadd(fadd(fmul(R[r * 3], ax), fadd(fmul(R[r * 3 + 1], ay), fmul(R[r * 3 + 2], az))), t[r]), b[3 * i + r]);
    const float err = __fsqrt_rn(fadd(fmul(e[0], e[0]), fadd(fmul(e[1], e[1]), fmul(e[2], e[2]))));
    return (double)err < threshold; // ComputeDistanceMeasure returns the float norm as a double (TransformationModel.hpp:37-52,83)
}
__device__ __forceinline__ void ransac_model(const float *__restrict__ a, const float *__restrict__ b, const int *s8, float *R, float *t)
{
    kabsch_f32(kRansacSample, [&](int k, float *pa, float *pb) {
        const int j = s8[k];
        pa[0] = a[3 * j]; pa[1] = a[3 * j + 1]; pa[2] = a[3 * j + 2];
        pb[0] = b[3 * j]; pb[1] = b[3 * j + 1]; pb[2] = b[3 * j + 2];
    }, R, t);
}
// best: (score << 32) | ~iteration, so that atomicMax keeps the highest score and, among equals, the earliest iteration
__global__ void __launch_bounds__(128) ransac_score_kernel(const float *__restrict__ a, const float *__restrict__ b, int n, int iterations,
                                                           double threshold, unsigned long long seed, const int *__restrict__ forced,
                                                           int *__restrict__ scores, unsigned long long *best)
{
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= iterations) return;
    int s8[kRansacSample];
    if (forced)
        for (int k = 0; k < kRansacSample; ++k) s8[k] = forced[(size_t)h * kRansacSample + k];
    else ransac_draw(seed, h, n, s8);
    float R[9], t[3];
    ransac_model(a, b, s8, R, t);
    int score = 0;
    for (int i = 0; i < n; ++i) score += ransac_is_inlier(R, t, a, b, i, threshold);
    scores[h] = score;
    atomicMax(best, ((unsigned long long)(unsigned int)score << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)h));
}
// the winner again: its sample, its motion (12 floats: R row-major, t) and one inlier flag per pair
__global__ void ransac_winner_kernel(const float *__restrict__ a, const float *__restrict__ b, int n, int winner, double threshold,
                                     unsigned long long seed, const int *__restrict__ forced, unsigned char *__restrict__ inlier,
                                     float *__restrict__ motion, int *__restrict__ sample)
{
    int s8[kRansacSample];
    if (forced)
        for (int k = 0; k < kRansacSample; ++k) s8[k] = forced[(size_t)winner * kRansacSample + k];
    else ransac_draw(seed, winner, n, s8);
    float R[9], t[3];
    ransac_model(a, b, s8, R, t);
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        for (int k = 0; k < 9; ++k) motion[k] = R[k];
        for (int k = 0; k < 3; ++k) motion[9 + k] = t[k];
        for (int k = 0; k < kRansacSample; ++k) sample[k] = s8[k];
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) inlier[i] = ransac_is_inlier(R, t, a, b, i, threshold);
}
// rows with a NaN or an infinity (isolated points' descriptors are 0 * inf): flags[0] = their number
__global__ void count_nonfinite_rows_kernel(const float *__restrict__ rows, int n, int dim, int *flags)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        bool bad = false;
        for (int e = 0; e < dim; ++e) bad |= !isfinite(rows[(size_t)i * dim + e]);
        if (bad) atomicAdd(flags, 1);
    }
}
} // namespace opb

using namespace opb;

struct opb_kdtree
{
    int device = 0, sm_count = 148, dim = 3;
    opb_kdtree *feature_tree = nullptr; // the KDTree<33> workspace of opb_kdtree_feature_matching, created on first use
    cudaStream_t stream = nullptr;
    size_t cap_points = 0, n = 0;
    float *d_pts = nullptr, *d_key = nullptr, *d_boxes = nullptr;
    float4 *d_sorted = nullptr;
    float *d_sorted_rows = nullptr;
    int *d_vind = nullptr, *d_pos = nullptr, *d_queue[2] = {nullptr, nullptr};
    KdNode *d_nodes = nullptr;
    KdBuildCtl *d_ctl = nullptr, *h_ctl = nullptr;
    int n_nodes = 0, max_level = 0;
    bool built = false;
    // query scratch
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0;
    void *d_aux[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t aux_bytes[4] = {0, 0, 0, 0};
};

static int kd_reserve(void **p, size_t *have, size_t want)
{
    if (*have >= want && *p) return OPB_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *have = 0;
    OPB_CUDA(cudaMalloc(p, want));
    *have = want;
    return OPB_OK;
}
static KdView kd_view(const opb_kdtree *t)
{
    KdView v;
    v.pts = t->d_pts; v.sorted = t->d_sorted; v.sorted_rows = t->d_sorted_rows; v.vind = t->d_vind; v.nodes = t->d_nodes; v.n = (int)t->n;
    for (int d = 0; d < kMaxDim; ++d) { v.root_lo[d] = t->h_ctl->root_lo[d]; v.root_hi[d] = t->h_ctl->root_hi[d]; }
    return v;
}
static int kd_grid(const opb_kdtree *t, size_t work, int per_sm)
{
    const size_t blocks = (work + kQueryThreads - 1) / kQueryThreads;
    const size_t cap = (size_t)t->sm_count * per_sm;
    return (int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

int opb_kdtree_create(int device, opb_kdtree **out)
{
    if (!out) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
    {
        set_error("no CUDA device %d (the library has no CPU path)", device);
        return OPB_ERR_CUDA;
    }
    OPB_CUDA(cudaSetDevice(device));
    opb_kdtree *t = new opb_kdtree();
    t->device = device;
    cudaDeviceGetAttribute(&t->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking) != cudaSuccess || cudaMalloc((void **)&t->d_ctl, sizeof(KdBuildCtl)) != cudaSuccess ||
        cudaMallocHost((void **)&t->h_ctl, sizeof(KdBuildCtl)) != cudaSuccess)
    {
        set_error("kd-tree workspace allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete t;
        return OPB_ERR_CUDA;
    }
    *out = t;
    return OPB_OK;
}
void opb_kdtree_destroy(opb_kdtree *t)
{
    if (!t) return;
    opb_kdtree_destroy(t->feature_tree);
    cudaSetDevice(t->device);
    cudaFree(t->d_sorted_rows);
    cudaFree(t->d_pts); cudaFree(t->d_key); cudaFree(t->d_boxes); cudaFree(t->d_vind); cudaFree(t->d_pos); cudaFree(t->d_sorted);
    cudaFree(t->d_queue[0]); cudaFree(t->d_queue[1]); cudaFree(t->d_nodes); cudaFree(t->d_ctl); cudaFree(t->d_scratch);
    for (int i = 0; i < 4; ++i) cudaFree(t->d_aux[i]);
    if (t->h_ctl) cudaFreeHost(t->h_ctl);
    if (t->stream) cudaStreamDestroy(t->stream);
    delete t;
}
// builds the tree of t->dim-dimensional rows
static int kd_build_rows(opb_kdtree *t, const float *xyz, size_t n)
{
    const size_t dim = (size_t)t->dim;
    if (n > 0x3FFFFFF0u / dim) { set_error("clouds above 2^30 values are not supported"); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(t->device));
    t->built = false;
    t->n = n;
    if (n > t->cap_points)
    {
        cudaFree(t->d_pts); cudaFree(t->d_key); cudaFree(t->d_boxes); cudaFree(t->d_vind); cudaFree(t->d_pos); cudaFree(t->d_sorted);
        cudaFree(t->d_queue[0]); cudaFree(t->d_queue[1]); cudaFree(t->d_nodes); cudaFree(t->d_sorted_rows);
        t->d_sorted = nullptr;
        t->d_sorted_rows = nullptr;
        t->d_pts = t->d_key = t->d_boxes = nullptr; t->d_vind = t->d_pos = t->d_queue[0] = t->d_queue[1] = nullptr; t->d_nodes = nullptr;
        t->cap_points = 0;
        const size_t cap = n + n / 8 + 1024, nodes = 2 * cap + 2;
        OPB_CUDA(cudaMalloc((void **)&t->d_pts, cap * dim * sizeof(float)));
        OPB_CUDA(cudaMalloc((void **)&t->d_key, cap * sizeof(float)));
        if (dim == 3) OPB_CUDA(cudaMalloc((void **)&t->d_sorted, cap * sizeof(float4)));
        else OPB_CUDA(cudaMalloc((void **)&t->d_sorted_rows, cap * dim * sizeof(float)));
        OPB_CUDA(cudaMalloc((void **)&t->d_vind, cap * sizeof(int)));
        OPB_CUDA(cudaMalloc((void **)&t->d_pos, cap * sizeof(int)));
        OPB_CUDA(cudaMalloc((void **)&t->d_queue[0], nodes * sizeof(int)));
        OPB_CUDA(cudaMalloc((void **)&t->d_queue[1], nodes * sizeof(int)));
        OPB_CUDA(cudaMalloc((void **)&t->d_nodes, nodes * sizeof(KdNode)));
        OPB_CUDA(cudaMalloc((void **)&t->d_boxes, nodes * 2 * dim * sizeof(float)));
        t->cap_points = cap;
    }
    cudaStream_t s = t->stream;
    memset(t->h_ctl, 0, sizeof(KdBuildCtl));
    t->n_nodes = 0;
    t->max_level = 0;
    if (n == 0) { t->built = true; return OPB_OK; }
    OPB_CUDA(cudaMemcpyAsync(t->d_pts, xyz, n * dim * sizeof(float), cudaMemcpyDefault, s));
    kd_iota_kernel<<<kd_grid(t, n, 8), kQueryThreads, 0, s>>>(t->d_vind, (int)n);
    KdNode root;
    root.left = 0; root.right = (int)n; root.child1 = root.child2 = -1; root.divfeat = -1; root.divlow = root.divhigh = 0.0f; root.level = 0;
    const int zero = 0;
    t->h_ctl->n_nodes = 1;
    OPB_CUDA(cudaMemcpyAsync(t->d_nodes, &root, sizeof(KdNode), cudaMemcpyHostToDevice, s));
    OPB_CUDA(cudaMemcpyAsync(t->d_queue[0], &zero, sizeof(int), cudaMemcpyHostToDevice, s));
    OPB_CUDA(cudaMemcpyAsync(t->d_ctl, t->h_ctl, sizeof(KdBuildCtl), cudaMemcpyHostToDevice, s));
    OPB_CUDA(cudaStreamSynchronize(s)); // root / zero are stack variables
    int in_count = 1, slot = 0, largest = (int)n;
    for (int level = 0; in_count > 0; ++level)
    {
        if (level > 4 * kMaxDepthRows) { set_error("kd-tree build did not terminate"); return OPB_ERR_UNSUPPORTED; }
        const int out_slot = slot ^ 1;
        OPB_CUDA(cudaMemsetAsync(&t->d_ctl->queue_count[out_slot], 0, sizeof(int), s));
        OPB_CUDA(cudaMemsetAsync(&t->d_ctl->queue_max[out_slot], 0, sizeof(int), s));
        // one CTA per node, sized for the largest node of the level: the top of the tree is a few very long nodes
#define OPB_KD_SPLIT(THREADS, DIM)                                                                                                                  \
    kd_split_kernel<THREADS, DIM><<<in_count, THREADS, 0, s>>>(t->d_pts, t->d_vind, t->d_key, t->d_pos, t->d_nodes, t->d_boxes, t->d_queue[slot], \
                                                               t->d_queue[out_slot], t->d_ctl, out_slot)
        if (dim == 3)
        {
            if (largest > kBuildBigNode) OPB_KD_SPLIT(kBuildThreadsBig, 3);
            else if (largest > kBuildSmallNode) OPB_KD_SPLIT(kBuildThreadsMid, 3);
            else OPB_KD_SPLIT(kBuildThreadsSmall, 3);
        }
        else if (largest > kBuildSmallNode) OPB_KD_SPLIT(kBuildThreadsMid, kFeatureDim); // descriptor sets are a few thousand rows
        else OPB_KD_SPLIT(kBuildThreadsSmall, kFeatureDim);
#undef OPB_KD_SPLIT
        OPB_CUDA(cudaGetLastError());
        OPB_CUDA(cudaMemcpyAsync(t->h_ctl, t->d_ctl, sizeof(KdBuildCtl), cudaMemcpyDeviceToHost, s));
        OPB_CUDA(cudaStreamSynchronize(s));
        in_count = t->h_ctl->queue_count[out_slot];
        largest = t->h_ctl->queue_max[out_slot];
        slot = out_slot;
    }
    if (dim == 3) kd_reorder_kernel<<<kd_grid(t, n, 8), kQueryThreads, 0, s>>>(t->d_pts, t->d_vind, (int)n, t->d_sorted);
    else kd_reorder_rows_kernel<<<kd_grid(t, n * dim, 8), kQueryThreads, 0, s>>>(t->d_pts, t->d_vind, (int)n, (int)dim, t->d_sorted_rows);
    OPB_CUDA(cudaGetLastError());
    t->n_nodes = t->h_ctl->n_nodes;
    t->max_level = t->h_ctl->max_level;
    const int stack_levels = dim == 3 ? kMaxDepth : kMaxDepthRows;
    if (t->max_level + 2 > stack_levels)
    {
        set_error("kd-tree of depth %d exceeds the search stack (%d levels)", t->max_level, stack_levels - 2);
        return OPB_ERR_UNSUPPORTED;
    }
    t->built = true;
    return OPB_OK;
}
int opb_kdtree_build(opb_kdtree *t, const float *xyz, size_t n)
{
    if (!t || (!xyz && n)) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    return kd_build_rows(t, xyz, n);
}
int opb_kdtree_dump(opb_kdtree *t, int32_t *vind, int32_t *node_ints, float *node_floats, float root_box[6], size_t *n_nodes)
{
    if (!t || !n_nodes) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    if (!t->built) { set_error("opb_kdtree_build has not succeeded"); return OPB_ERR_INVALID; }
    *n_nodes = (size_t)t->n_nodes;
    if (!vind && !node_ints && !node_floats) return OPB_OK;
    OPB_CUDA(cudaSetDevice(t->device));
    if (vind && t->n) OPB_CUDA(cudaMemcpy(vind, t->d_vind, t->n * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<KdNode> nodes((size_t)t->n_nodes);
    if (t->n_nodes) OPB_CUDA(cudaMemcpy(nodes.data(), t->d_nodes, nodes.size() * sizeof(KdNode), cudaMemcpyDeviceToHost));
    for (int i = 0; i < t->n_nodes; ++i)
    {
        if (node_ints)
        {
            node_ints[5 * i] = nodes[i].left; node_ints[5 * i + 1] = nodes[i].right; node_ints[5 * i + 2] = nodes[i].child1;
            node_ints[5 * i + 3] = nodes[i].child2; node_ints[5 * i + 4] = nodes[i].divfeat;
        }
        if (node_floats) { node_floats[2 * i] = nodes[i].divlow; node_floats[2 * i + 1] = nodes[i].divhigh; }
    }
    if (root_box)
        for (int d = 0; d < 3; ++d) { root_box[d] = t->h_ctl->root_lo[d]; root_box[3 + d] = t->h_ctl->root_hi[d]; }
    return OPB_OK;
}
// device-side radius search of the tree's own points or of `d_queries` into rows of k indices
static int kd_radius_rows(opb_kdtree *t, const float *d_queries, size_t nq, int k, float radius, int *d_index, float *d_dist, int *d_count)
{
    const int cap = (int)(size_t)(k * 2.5);
    const KdView v = kd_view(t);
    const int *order = d_queries == t->d_pts ? t->d_vind : nullptr; // the tree's own points are queried in leaf order
    // one warp per CTA with the hits of its 32 threads in shared memory (8 B x cap x 32) when that fits -- the sort walks them
    // thousands of times --, else a global scratch
    const size_t smem = (size_t)cap * 32 * 8;
    if (smem <= 96 * 1024)
    {
        OPB_CUDA(cudaFuncSetAttribute(kd_radius_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        const size_t blocks = (nq + 31) / 32, limit = (size_t)t->sm_count * 64;
        kd_radius_kernel<<<(unsigned)(blocks < limit ? blocks : limit), 32, smem, t->stream>>>(v, d_queries, order, 0, (int)nq, k, cap, radius, nullptr,
                                                                                               nullptr, 0, d_index, d_dist, d_count);
        OPB_CUDA(cudaGetLastError());
        return OPB_OK;
    }
    const size_t batch = nq < (size_t)kQueryBatch ? nq : (size_t)kQueryBatch;
    int rc = kd_reserve(&t->d_scratch, &t->scratch_bytes, batch * (size_t)cap * 8);
    if (rc) return rc;
    float *sd = (float *)t->d_scratch;
    int *si = (int *)(sd + batch * (size_t)cap);
    for (size_t q0 = 0; q0 < nq; q0 += batch)
    {
        const size_t m = nq - q0 < batch ? nq - q0 : batch;
        kd_radius_kernel<<<kd_grid(t, m, 16), kQueryThreads, 0, t->stream>>>(v, d_queries, order, (int)q0, (int)m, k, cap, radius, sd, si, (int)batch,
                                                                             d_index, d_dist, d_count);
        OPB_CUDA(cudaGetLastError());
    }
    return OPB_OK;
}
int opb_kdtree_search(opb_kdtree *t, const float *queries, size_t nq, int mode, int k, float radius, int32_t *out_index, float *out_dist,
                      int32_t *out_count)
{
    if (!t || (!queries && nq) || !out_index || !out_dist || !out_count) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    if (!t->built) { set_error("opb_kdtree_build has not succeeded"); return OPB_ERR_INVALID; }
    if (mode < 0 || mode > 2) { set_error("mode must be 0 (knn), 1 (radius) or 2 (knn within radius)"); return OPB_ERR_INVALID; }
    if (k < 1 || (mode != 1 && k > kKnnCap) || (mode == 1 && (int)(size_t)(k * 2.5) > kRadiusCapMax))
    {
        set_error("k must be 1..%d (knn) or 1..%d (radius)", kKnnCap, (int)(kRadiusCapMax / 2.5));
        return OPB_ERR_INVALID;
    }
    if (nq == 0) return OPB_OK;
    if (nq > 0x7FFFFFF0u / (size_t)k) { set_error("too many queries"); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(t->device));
    int rc;
    if ((rc = kd_reserve(&t->d_aux[0], &t->aux_bytes[0], nq * 3 * sizeof(float)))) return rc;
    if ((rc = kd_reserve(&t->d_aux[1], &t->aux_bytes[1], nq * (size_t)k * sizeof(int)))) return rc;
    if ((rc = kd_reserve(&t->d_aux[2], &t->aux_bytes[2], nq * (size_t)k * sizeof(float)))) return rc;
    if ((rc = kd_reserve(&t->d_aux[3], &t->aux_bytes[3], nq * sizeof(int)))) return rc;
    cudaStream_t s = t->stream;
    float *dq = (float *)t->d_aux[0];
    int *di = (int *)t->d_aux[1];
    float *dd = (float *)t->d_aux[2];
    int *dc = (int *)t->d_aux[3];
    OPB_CUDA(cudaMemcpyAsync(dq, queries, nq * 3 * sizeof(float), cudaMemcpyDefault, s));
    if (mode == 1)
    {
        if ((rc = kd_radius_rows(t, dq, nq, k, radius, di, dd, dc))) return rc;
    }
    else
    {
        const size_t smem = (size_t)k * kQueryThreads * 8;
        OPB_CUDA(cudaFuncSetAttribute(kd_knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kKnnCap * kQueryThreads * 8));
        kd_knn_kernel<<<kd_grid(t, nq, 4), kQueryThreads, smem, s>>>(kd_view(t), dq, (int)nq, mode, k, radius, di, dd, dc);
        OPB_CUDA(cudaGetLastError());
    }
    OPB_CUDA(cudaMemcpyAsync(out_index, di, nq * (size_t)k * sizeof(int), cudaMemcpyDefault, s));
    OPB_CUDA(cudaMemcpyAsync(out_dist, dd, nq * (size_t)k * sizeof(float), cudaMemcpyDefault, s));
    OPB_CUDA(cudaMemcpyAsync(out_count, dc, nq * sizeof(int), cudaMemcpyDefault, s));
    OPB_CUDA(cudaStreamSynchronize(s));
    return OPB_OK;
}
int opb_kdtree_estimate_normals(opb_kdtree *t, float radius, int knn, float *normals)
{
    if (!t || !normals) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    if (!t->built) { set_error("opb_kdtree_build has not succeeded"); return OPB_ERR_INVALID; }
    if (knn < 1 || knn > kKnnCap) { set_error("knn must be 1..%d", kKnnCap); return OPB_ERR_INVALID; }
    if (t->n == 0) return OPB_OK;
    OPB_CUDA(cudaSetDevice(t->device));
    int rc;
    if ((rc = kd_reserve(&t->d_aux[2], &t->aux_bytes[2], t->n * 3 * sizeof(float)))) return rc;
    float *dn = (float *)t->d_aux[2];
    const size_t smem = (size_t)knn * kQueryThreads * 8;
    OPB_CUDA(cudaFuncSetAttribute(kd_normals_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kKnnCap * kQueryThreads * 8));
    int per_sm = 0;
    OPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kd_normals_kernel, kQueryThreads, smem));
    kd_normals_kernel<<<kd_grid(t, t->n, (per_sm < 1 ? 1 : per_sm) * 2), kQueryThreads, smem, t->stream>>>(kd_view(t), knn, radius, dn);
    OPB_CUDA(cudaGetLastError());
    OPB_CUDA(cudaMemcpyAsync(normals, dn, t->n * 3 * sizeof(float), cudaMemcpyDefault, t->stream));
    OPB_CUDA(cudaStreamSynchronize(t->stream));
    return OPB_OK;
}
int opb_kdtree_fpfh(opb_kdtree *t, const float *normals, int knn, float radius, float *features)
{
    if (!t || !normals || !features) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    if (!t->built) { set_error("opb_kdtree_build has not succeeded"); return OPB_ERR_INVALID; }
    if (knn < 1 || knn > 256 || (int)(size_t)(knn * 2.5) > kRadiusCapMax) { set_error("knn must be 1..256"); return OPB_ERR_INVALID; }
    const size_t n = t->n;
    if (n == 0) return OPB_OK;
    if (n > 0x7FFFFFF0u / (size_t)(knn > 33 ? knn : 33)) { set_error("cloud too large for %d neighbours per point", knn); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(t->device));
    int rc;
    if ((rc = kd_reserve(&t->d_aux[0], &t->aux_bytes[0], n * 3 * sizeof(float)))) return rc;  // normals
    if ((rc = kd_reserve(&t->d_aux[1], &t->aux_bytes[1], n * (size_t)knn * sizeof(int)))) return rc; // neighbour rows
    if ((rc = kd_reserve(&t->d_aux[2], &t->aux_bytes[2], n * 66 * sizeof(float)))) return rc; // spfh, fpfh
    if ((rc = kd_reserve(&t->d_aux[3], &t->aux_bytes[3], n * sizeof(int)))) return rc;
    cudaStream_t s = t->stream;
    float *dn = (float *)t->d_aux[0];
    int *nbr = (int *)t->d_aux[1];
    float *spfh = (float *)t->d_aux[2], *fpfh = spfh + n * 33;
    int *cnt = (int *)t->d_aux[3];
    OPB_CUDA(cudaMemcpyAsync(dn, normals, n * 3 * sizeof(float), cudaMemcpyDefault, s));
    if ((rc = kd_radius_rows(t, t->d_pts, n, knn, radius, nbr, nullptr, cnt))) return rc;
    fpfh_spfh_kernel<<<kd_grid(t, n, 8), kQueryThreads, 0, s>>>(t->d_pts, dn, (int)n, knn, nbr, cnt, spfh);
    OPB_CUDA(cudaGetLastError());
    fpfh_combine_kernel<<<kd_grid(t, n * 32, 8), kQueryThreads, 0, s>>>(t->d_pts, (int)n, knn, nbr, cnt, spfh, fpfh);
    OPB_CUDA(cudaGetLastError());
    OPB_CUDA(cudaMemcpyAsync(features, fpfh, n * 33 * sizeof(float), cudaMemcpyDefault, s));
    OPB_CUDA(cudaStreamSynchronize(s));
    return OPB_OK;
}

int opb_kdtree_feature_matching(opb_kdtree *t, const float *src_feat33, size_t ns, const float *tgt_feat33, size_t nt, int32_t *pairs, size_t *n_pairs)
{
    if (!t || !n_pairs || (ns && (!src_feat33 || !pairs)) || (nt && !tgt_feat33)) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    *n_pairs = 0;
    if (ns == 0 || nt == 0) return OPB_OK; // an empty tree answers no query (nanoflann.hpp:1231-1232)
    if (ns > 0x7FFFFFF0u / kFeatureDim || nt > 0x7FFFFFF0u / kFeatureDim) { set_error("feature set too large"); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(t->device));
    int rc;
    if ((rc = kd_reserve(&t->d_aux[0], &t->aux_bytes[0], ns * kFeatureDim * sizeof(float)))) return rc;
    if ((rc = kd_reserve(&t->d_aux[2], &t->aux_bytes[2], nt * kFeatureDim * sizeof(float)))) return rc;
    if ((rc = kd_reserve(&t->d_aux[3], &t->aux_bytes[3], ns * sizeof(int)))) return rc;
    cudaStream_t s = t->stream;
    OPB_CUDA(cudaMemcpyAsync(t->d_aux[0], src_feat33, ns * kFeatureDim * sizeof(float), cudaMemcpyDefault, s));
    OPB_CUDA(cudaMemcpyAsync(t->d_aux[2], tgt_feat33, nt * kFeatureDim * sizeof(float), cudaMemcpyDefault, s));
    // a NaN / infinite TARGET row poisons nanoflann's boxes (the reference then prunes real neighbours); such sets take the
    // exhaustive scan, which returns the true nearest finite row.  Everything else is answered by the reference's own tree.
    int bad_rows = 0;
    OPB_CUDA(cudaMemsetAsync(t->d_ctl, 0, sizeof(int), s));
    count_nonfinite_rows_kernel<<<kd_grid(t, nt, 8), kQueryThreads, 0, s>>>((const float *)t->d_aux[2], (int)nt, kFeatureDim, (int *)t->d_ctl);
    OPB_CUDA(cudaMemcpyAsync(&bad_rows, t->d_ctl, sizeof(int), cudaMemcpyDeviceToHost, s));
    OPB_CUDA(cudaStreamSynchronize(s));
    const unsigned blocks = (unsigned)((ns + kMatchThreads - 1) / kMatchThreads);
    const char *forced = getenv("OPB_MATCH_EXHAUSTIVE"); // developer knob
    bool exhaustive = bad_rows || (forced && atoi(forced));
    if (!exhaustive)
    {
        if (!t->feature_tree)
        {
            if ((rc = opb_kdtree_create(t->device, &t->feature_tree))) return rc;
            t->feature_tree->dim = kFeatureDim;
        }
        opb_kdtree *ft = t->feature_tree;
        rc = kd_build_rows(ft, (const float *)t->d_aux[2], nt);
        if (rc == OPB_ERR_UNSUPPORTED) exhaustive = true; // a tree deeper than the search stack (never seen: 5,000 descriptors are 70-80 levels)
        else if (rc) return rc;
        else
        {
            OPB_CUDA(cudaStreamSynchronize(ft->stream));
            fpfh_match_tree_kernel<<<blocks, kMatchThreads, 0, s>>>(kd_view(ft), (const float *)t->d_aux[0], (int)ns, (int *)t->d_aux[3]);
            OPB_CUDA(cudaGetLastError());
        }
    }
    if (exhaustive)
    {
        fpfh_match_kernel<<<blocks, kMatchThreads, 0, s>>>((const float *)t->d_aux[0], (int)ns, (const float *)t->d_aux[2], (int)nt, (int *)t->d_aux[3]);
        OPB_CUDA(cudaGetLastError());
    }
    std::vector<int> nearest(ns);
    OPB_CUDA(cudaMemcpyAsync(nearest.data(), t->d_aux[3], ns * sizeof(int), cudaMemcpyDeviceToHost, s));
    OPB_CUDA(cudaStreamSynchronize(s));
    size_t m = 0;
    for (size_t i = 0; i < ns; ++i)
        if (nearest[i] >= 0) { pairs[2 * m] = (int32_t)i; pairs[2 * m + 1] = nearest[i]; ++m; }
    *n_pairs = m;
    return OPB_OK;
}

// registration::RejectMatchesRanSaPC (GlobalRegistration.cpp:75-108).  Host code on purpose: every decision consumes a
// data-dependent number of draws from ONE std::default_random_engine, so the reference's result is defined by a strictly
// sequential walk over a few thousand matches (microseconds of work).  libstdc++'s default_random_engine is minstd_rand0
// (x <- 16807 x mod 2^31 - 1, outputs in [1, 2^31 - 2]); uniform_int_distribution<int>(0, N - 1) on it scales down by integer
// division and rejects the draws past N * scaling (bits/uniform_int_dist.h, the non-power-of-two branch).
namespace
{
inline uint32_t minstd_rand0_next(uint32_t &x)
{
    x = (uint32_t)(((uint64_t)x * 16807u) % 2147483647u);
    return x;
}
inline float norm3_eigen(const float *a, const float *b)
{
    const float x = a[0] - b[0], y = a[1] - b[1], z = a[2] - b[2];
    return sqrtf(x * x + (y * y + z * z)); // Eigen's fixed-size reduction order; -ffp-contract=off keeps the products rounded
}
} // namespace
int opb_reject_matches(const float *src_xyz, size_t ns, const float *tgt_xyz, size_t nt, int32_t *pairs, size_t *n_pairs, uint32_t *engine_state,
                       int candidate_num, float difference)
{
    if (!n_pairs || !engine_state || (*n_pairs && (!src_xyz || !tgt_xyz || !pairs))) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    const size_t n = *n_pairs;
    if (n > 0x7FFFFFF0u) { set_error("too many matches"); return OPB_ERR_INVALID; }
    if (*engine_state == 0 || *engine_state >= 2147483647u) { set_error("engine state must be in [1, 2^31 - 2] (a default-constructed engine holds 1)"); return OPB_ERR_INVALID; }
    for (size_t i = 0; i < n; ++i)
        if (pairs[2 * i] < 0 || (size_t)pairs[2 * i] >= ns || pairs[2 * i + 1] < 0 || (size_t)pairs[2 * i + 1] >= nt)
        {
            set_error("match %zu names a point outside the clouds", i);
            return OPB_ERR_INVALID;
        }
    if (n == 0) return OPB_OK;
    const uint32_t urng_range = 2147483646u - 1u, ue_range = (uint32_t)n;
    const uint32_t scaling = urng_range / ue_range, past = ue_range * scaling;
    std::vector<int32_t> kept;
    kept.reserve(2 * n);
    uint32_t x = *engine_state;
    for (size_t i = 0; i < n; ++i)
    {
        const float *ref_point = src_xyz + 3 * (size_t)pairs[2 * i], *new_point = tgt_xyz + 3 * (size_t)pairs[2 * i + 1];
        bool keep = false;
        for (int j = 0; j < candidate_num; ++j)
        {
            uint32_t r;
            do r = minstd_rand0_next(x) - 1u; while (r >= past);
            const size_t c = r / scaling;
            const float d1 = norm3_eigen(src_xyz + 3 * (size_t)pairs[2 * c], ref_point);
            const float d2 = norm3_eigen(tgt_xyz + 3 * (size_t)pairs[2 * c + 1], new_point);
            if (fabs((double)(d1 - d2)) <= (double)(difference * d1)) { keep = true; break; }
        }
        if (keep) { kept.push_back(pairs[2 * i]); kept.push_back(pairs[2 * i + 1]); }
    }
    memcpy(pairs, kept.data(), kept.size() * sizeof(int32_t));
    *n_pairs = kept.size() / 2;
    *engine_state = x;
    return OPB_OK;
}

int opb_ransac_rigid_transformation(opb_kdtree *t, const float *src_xyz, const float *tgt_xyz, size_t n, int max_iteration, double threshold,
                                    uint64_t seed, const int32_t *forced_samples, float T_colmajor[16], int32_t *inlier_ids, size_t *n_inliers,
                                    int32_t *best_iteration, int32_t best_sample[8])
{
    if (!t || !T_colmajor || !n_inliers || (n && (!src_xyz || !tgt_xyz || !inlier_ids))) { set_error("NULL argument"); return OPB_ERR_INVALID; }
    *n_inliers = 0;
    if (best_iteration) *best_iteration = -1;
    for (int i = 0; i < 16; ++i) T_colmajor[i] = 0.0f;
    if (n < (size_t)kRansacSample) return OPB_OK; // "Too few canidate point pair": the zero matrix (Ransac.cpp:10-14)
    if (n == (size_t)kRansacSample) { set_error("RANSAC needs more than %d pairs (GRANSAC refuses and the reference then dereferences a null model)", kRansacSample); return OPB_ERR_INVALID; }
    if (max_iteration < 1 || max_iteration > (1 << 24)) { set_error("max_iteration must be 1..2^24"); return OPB_ERR_INVALID; }
    if (n > 0x7FFFFFF0u / 3) { set_error("too many pairs"); return OPB_ERR_INVALID; }
    if (forced_samples)
        for (size_t i = 0; i < (size_t)max_iteration * kRansacSample; ++i)
            if (forced_samples[i] < 0 || (size_t)forced_samples[i] >= n) { set_error("sample %zu names a pair outside the set", i); return OPB_ERR_INVALID; }
    OPB_CUDA(cudaSetDevice(t->device));
    int rc;
    const size_t pts_bytes = n * 3 * sizeof(float), iters = (size_t)max_iteration;
    if ((rc = kd_reserve(&t->d_aux[0], &t->aux_bytes[0], 2 * pts_bytes))) return rc;                                   // a, b
    if ((rc = kd_reserve(&t->d_aux[1], &t->aux_bytes[1], iters * (forced_samples ? kRansacSample + 1 : 1) * sizeof(int) + 64))) return rc; // scores, samples
    if ((rc = kd_reserve(&t->d_aux[3], &t->aux_bytes[3], n + 256))) return rc;                                         // flags, motion, sample, best
    cudaStream_t s = t->stream;
    float *da = (float *)t->d_aux[0], *db = da + n * 3;
    int *scores = (int *)t->d_aux[1], *forced = forced_samples ? scores + iters : nullptr;
    unsigned char *flags = (unsigned char *)t->d_aux[3];
    char *tail = (char *)t->d_aux[3] + ((n + 63) & ~(size_t)63);
    unsigned long long *best = (unsigned long long *)tail;
    float *motion = (float *)(tail + 16);
    int *sample = (int *)(tail + 16 + 12 * sizeof(float));
    OPB_CUDA(cudaMemcpyAsync(da, src_xyz, pts_bytes, cudaMemcpyDefault, s));
    OPB_CUDA(cudaMemcpyAsync(db, tgt_xyz, pts_bytes, cudaMemcpyDefault, s));
    if (forced) OPB_CUDA(cudaMemcpyAsync(forced, forced_samples, iters * kRansacSample * sizeof(int), cudaMemcpyDefault, s));
    OPB_CUDA(cudaMemsetAsync(best, 0, sizeof(unsigned long long), s));
    ransac_score_kernel<<<(unsigned)((iters + 127) / 128), 128, 0, s>>>(da, db, (int)n, max_iteration, threshold, seed, forced, scores, best);
    OPB_CUDA(cudaGetLastError());
    unsigned long long h_best = 0;
    OPB_CUDA(cudaMemcpyAsync(&h_best, best, sizeof(h_best), cudaMemcpyDeviceToHost, s));
    OPB_CUDA(cudaStreamSynchronize(s));
    const unsigned int score = (unsigned int)(h_best >> 32);
    if (score == 0) { set_error("no hypothesis has an inlier (the reference would dereference a null model)"); return OPB_ERR_INVALID; }
    const int winner = (int)(0xFFFFFFFFu - (unsigned int)(h_best & 0xFFFFFFFFu));
    ransac_winner_kernel<<<kd_grid(t, n, 8), kQueryThreads, 0, s>>>(da, db, (int)n, winner, threshold, seed, forced, flags, motion, sample);
    OPB_CUDA(cudaGetLastError());
    std::vector<unsigned char> h_flags(n);
    float h_motion[12];
    int h_sample[kRansacSample];
    OPB_CUDA(cudaMemcpyAsync(h_flags.data(), flags, n, cudaMemcpyDeviceToHost, s));
    OPB_CUDA(cudaMemcpyAsync(h_motion, motion, sizeof(h_motion), cudaMemcpyDeviceToHost, s));
    OPB_CUDA(cudaMemcpyAsync(h_sample, sample, sizeof(h_sample), cudaMemcpyDeviceToHost, s));
    OPB_CUDA(cudaStreamSynchronize(s));
    size_t m = 0;
    for (size_t i = 0; i < n; ++i)
        if (h_flags[i]) inlier_ids[m++] = (int32_t)i;
    *n_inliers = m;
    for (int r = 0; r < 3; ++r)
    {
        for (int c = 0; c < 3; ++c) T_colmajor[c * 4 + r] = h_motion[r * 3 + c];
        T_colmajor[12 + r] = h_motion[9 + r];
    }
    T_colmajor[15] = 1.0f;
    if (best_iteration) *best_iteration = winner;
    if (best_sample)
        for (int k = 0; k < kRansacSample; ++k) best_sample[k] = h_sample[k];
    return OPB_OK;
}
