// TEST INFRASTRUCTURE ONLY.  Runs the DEVICE code of onepiece_b200/csrc/opb_kdtree.cu (everything inside namespace opb, cut out
// by tests/test_kdtree_emulated.py into kdtree_device.inc) on the host emulator in cuda_emu.h, driven the way the library's host
// code drives it, so the build / search / sort / FPFH kernels can be compared with the oracle without a GPU.
#include "opb_common.cuh"
namespace opb { float knn_smem[1 << 17]; } // stands in for the kernels' dynamic shared memory
#ifndef EMU_BUILD_THREADS
#define EMU_BUILD_THREADS 64
#endif
#include "kdtree_device.inc"

using namespace opb;

struct Tree
{
    std::vector<float> pts, key, boxes;
    std::vector<float4> sorted;
    std::vector<int> vind, pos, queue[2];
    std::vector<KdNode> nodes;
    std::vector<float> sorted_rows;
    KdBuildCtl ctl;
    int n = 0, dim = 3;
    KdView view() const
    {
        KdView v;
        v.pts = pts.data(); v.sorted = sorted.data(); v.sorted_rows = sorted_rows.data(); v.vind = vind.data(); v.nodes = nodes.data(); v.n = n;
        for (int d = 0; d < kMaxDim; ++d) { v.root_lo[d] = ctl.root_lo[d]; v.root_hi[d] = ctl.root_hi[d]; }
        return v;
    }
};
static int grid_for(size_t work, int cap) { const size_t b = (work + kQueryThreads - 1) / kQueryThreads; return (int)(b < (size_t)cap ? (b ? b : 1) : cap); }

static Tree *build_rows(const float *xyz, long n, int dim)
{
    Tree *t = new Tree();
    t->n = (int)n;
    t->dim = dim;
    t->pts.assign(xyz, xyz + (size_t)dim * n);
    t->sorted_rows.resize((size_t)dim * n + 1);
    t->key.resize(n + 1); t->vind.resize(n + 1); t->pos.resize(n + 1); t->sorted.resize(n + 1);
    t->nodes.resize(2 * n + 2); t->boxes.resize((size_t)2 * dim * (2 * n + 2));
    t->queue[0].resize(2 * n + 2); t->queue[1].resize(2 * n + 2);
    memset(&t->ctl, 0, sizeof t->ctl);
    if (n == 0) return t;
    emu::launch(grid_for(n, 4), kQueryThreads, [&] { kd_iota_kernel(t->vind.data(), (int)n); });
    KdNode root;
    root.left = 0; root.right = (int)n; root.child1 = root.child2 = -1; root.divfeat = -1; root.divlow = root.divhigh = 0.0f; root.level = 0;
    t->nodes[0] = root;
    t->queue[0][0] = 0;
    t->ctl.n_nodes = 1;
    int in_count = 1, slot = 0;
    while (in_count > 0)
    {
        const int out_slot = slot ^ 1;
        t->ctl.queue_count[out_slot] = 0;
        t->ctl.queue_max[out_slot] = 0;
        emu::launch(in_count, EMU_BUILD_THREADS, [&] {
            if (dim == 3)
                kd_split_kernel<EMU_BUILD_THREADS, 3>(t->pts.data(), t->vind.data(), t->key.data(), t->pos.data(), t->nodes.data(), t->boxes.data(),
                                                      t->queue[slot].data(), t->queue[out_slot].data(), &t->ctl, out_slot);
            else
                kd_split_kernel<EMU_BUILD_THREADS, kFeatureDim>(t->pts.data(), t->vind.data(), t->key.data(), t->pos.data(), t->nodes.data(),
                                                                t->boxes.data(), t->queue[slot].data(), t->queue[out_slot].data(), &t->ctl, out_slot);
        });
        in_count = t->ctl.queue_count[out_slot];
        slot = out_slot;
    }
    if (dim == 3) emu::launch(grid_for(n, 4), kQueryThreads, [&] { kd_reorder_kernel(t->pts.data(), t->vind.data(), (int)n, t->sorted.data()); });
    else emu::launch(grid_for(n, 4), kQueryThreads, [&] { kd_reorder_rows_kernel(t->pts.data(), t->vind.data(), (int)n, dim, t->sorted_rows.data()); });
    return t;
}
extern "C"
{
void *emu_build(const float *xyz, long n) { return build_rows(xyz, n, 3); }
void *emu_build_rows(const float *rows, long n, int dim) { return build_rows(rows, n, dim); }
// what opb_kdtree_feature_matching does for finite targets: a KDTree<33> over them, one walk per source descriptor
void emu_match_tree(const float *src, long ns, const float *tgt, long nt, int32_t *nearest)
{
    int bad = 0;
    emu::launch(grid_for(nt, 2), kQueryThreads, [&] { count_nonfinite_rows_kernel(tgt, (int)nt, kFeatureDim, &bad); });
    if (bad) { for (long i = 0; i < ns; ++i) nearest[i] = -2; return; }
    Tree *t = build_rows(tgt, nt, kFeatureDim);
    const KdView v = t->view();
    emu::launch((int)((ns + kMatchThreads - 1) / kMatchThreads), kMatchThreads, [&] { fpfh_match_tree_kernel(v, src, (int)ns, nearest); });
    delete t;
}
}
extern "C"
{
void emu_destroy(void *p) { delete (Tree *)p; }
long emu_dump(void *p, int32_t *vind, int32_t *ni, float *nf, float *box)
{
    Tree *t = (Tree *)p;
    for (int i = 0; i < t->n; ++i) vind[i] = t->vind[i];
    for (int i = 0; i < t->ctl.n_nodes; ++i)
    {
        const KdNode &nd = t->nodes[i];
        ni[5 * i] = nd.left; ni[5 * i + 1] = nd.right; ni[5 * i + 2] = nd.child1; ni[5 * i + 3] = nd.child2; ni[5 * i + 4] = nd.divfeat;
        nf[2 * i] = nd.divlow; nf[2 * i + 1] = nd.divhigh;
    }
    for (int d = 0; d < t->dim; ++d) { box[d] = t->ctl.root_lo[d]; box[t->dim + d] = t->ctl.root_hi[d]; }
    return t->ctl.n_nodes * 1000L + t->ctl.max_level;
}
void emu_search(void *p, const float *queries, long nq, int mode, int k, float radius, int32_t *out_index, float *out_dist, int32_t *out_count)
{
    Tree *t = (Tree *)p;
    const KdView v = t->view();
    if (mode == 1)
    {
        const int cap = (int)(size_t)(k * 2.5);
        std::vector<float> sd((size_t)nq * cap + 1);
        std::vector<int> si((size_t)nq * cap + 1);
        if ((size_t)cap * 32 * 2 <= sizeof(knn_smem) / sizeof(float) && k != 20) // the library's one-warp CTAs with the hits in shared memory
            emu::launch((int)std::min<long>((nq + 31) / 32, 8), 32,
                        [&] { kd_radius_kernel(v, queries, nullptr, 0, (int)nq, k, cap, radius, nullptr, nullptr, 0, out_index, out_dist, out_count); });
        else
            emu::launch(grid_for(nq, 4), kQueryThreads, [&] {
                kd_radius_kernel(v, queries, nullptr, 0, (int)nq, k, cap, radius, sd.data(), si.data(), (int)nq, out_index, out_dist, out_count);
            });
    }
    else
        emu::launch(grid_for(nq, 4), kQueryThreads, [&] { kd_knn_kernel(v, queries, (int)nq, mode, k, radius, out_index, out_dist, out_count); });
}
void emu_normals(void *p, float radius, int knn, float *normals)
{
    Tree *t = (Tree *)p;
    const KdView v = t->view();
    emu::launch(grid_for(t->n, 4), kQueryThreads, [&] { kd_normals_kernel(v, knn, radius, normals); });
}
void emu_fpfh(void *p, const float *normals, int knn, float radius, float *features)
{
    Tree *t = (Tree *)p;
    const KdView v = t->view();
    const long n = t->n;
    const int cap = (int)(size_t)(knn * 2.5);
    std::vector<float> sd((size_t)n * cap + 1), spfh((size_t)n * 33 + 1);
    std::vector<int> si((size_t)n * cap + 1), nbr((size_t)n * knn + 1), cnt(n + 1);
    // the tree's own points, in leaf order, hits in shared memory: what opb_kdtree_fpfh launches
    emu::launch((int)std::min<long>((n + 31) / 32, 8), 32, [&] {
        kd_radius_kernel(v, t->pts.data(), t->vind.data(), 0, (int)n, knn, cap, radius, nullptr, nullptr, 0, nbr.data(), nullptr, cnt.data());
    });
    emu::launch(grid_for(n, 3), kQueryThreads, [&] { fpfh_spfh_kernel(t->pts.data(), normals, (int)n, knn, nbr.data(), cnt.data(), spfh.data()); });
    emu::launch(grid_for(n * 32, 5), kQueryThreads, [&] { fpfh_combine_kernel(t->pts.data(), (int)n, knn, nbr.data(), cnt.data(), spfh.data(), features); });
}
void emu_match(const float *src, long ns, const float *tgt, long nt, int32_t *nearest)
{
    emu::launch((int)((ns + kMatchThreads - 1) / kMatchThreads), kMatchThreads, [&] { fpfh_match_kernel(src, (int)ns, tgt, (int)nt, nearest); });
}
// opb_ransac_rigid_transformation's two launches; returns the winner (-1: no hypothesis with an inlier)
int emu_ransac(const float *a, const float *b, long n, int iterations, double threshold, unsigned long long seed, const int32_t *forced,
               float *motion12, unsigned char *inlier, int32_t *sample8)
{
    std::vector<int> scores(iterations + 1);
    unsigned long long best = 0;
    emu::launch((iterations + 127) / 128, 128, [&] { ransac_score_kernel(a, b, (int)n, iterations, threshold, seed, forced, scores.data(), &best); });
    if ((best >> 32) == 0) return -1;
    const int winner = (int)(0xFFFFFFFFu - (unsigned int)(best & 0xFFFFFFFFu));
    emu::launch(grid_for(n, 4), kQueryThreads, [&] { ransac_winner_kernel(a, b, (int)n, winner, threshold, seed, forced, inlier, motion12, sample8); });
    return winner;
}
}
