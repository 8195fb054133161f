// Host-side state of an opb_volume (shared by opb_volume.cu and opb_mesh.cu).
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "../../include/onepiece_b200.h"
#include "opb_volume.cuh"

namespace opb
{
struct ProfileSlot
{
    cudaEvent_t e[3]; // frame start, after cube selection, after voxel update
};
} // namespace opb

struct opb_volume
{
    opb_volume_desc desc;
    opb::VolumeDev dev = {};
    cudaStream_t stream = nullptr;      // compute
    cudaStream_t copy_stream = nullptr; // H2D staging
    bool own_stream = false;
    int sm_count = 0;
    int integrate_grid = 0;
    int integrate_pipe_grid = 0; // grid of integrate_pipelined_kernel
    int integrate_packed_grid = 0; // grid of integrate_packed_kernel
    int integrate_bulk_grid = 0;   // grid of integrate_bulk_kernel
    // double-buffered device staging for host frames
    void *stage_depth[2] = {nullptr, nullptr};
    unsigned char *stage_bgr[2] = {nullptr, nullptr};
    cudaEvent_t stage_copied[2] = {nullptr, nullptr};
    cudaEvent_t stage_consumed[2] = {nullptr, nullptr};
    unsigned long long frames_staged = 0, frames_enqueued = 0;
    // kernel timing (bench.py): CUDA events on the compute stream around K1 and K2
    bool profiling = false;
    std::vector<opb::ProfileSlot> profile_pending, profile_free;
    double prof_select_ms = 0, prof_integrate_ms = 0;
    long long prof_frames = 0;
    float last_select_ms = 0, last_integrate_ms = 0;
    // marching-cubes scratch
    void *mesh_scratch = nullptr;
    size_t mesh_scratch_bytes = 0;
    // ghost cubes received from the owner of the next slab (opb_halo.cu): slots [max_cubes - n_ghost, max_cubes)
    int n_ghost = 0;
    void *halo_scratch = nullptr;
    size_t halo_scratch_bytes = 0;
    // peer-memory exchange (opb_volume_halo_peer_*): this volume's receive box, the mapped boxes of the rank it exports to
    // (dst) and of the rank it imports from (src), the device words of an exchange, the number of exchanges started
    void *halo_box = nullptr, *halo_local = nullptr, *halo_dst = nullptr, *halo_src = nullptr;
    size_t halo_box_cap = 0, halo_dst_cap = 0;
    unsigned long long halo_epoch = 0;
    bool halo_pending = false;
    // frame ring (opb_volume_frame_ring_*): one frame uploaded in row bands by the ranks of a partitioned volume
    void *frame_ring = nullptr, *frame_ring_tickets = nullptr;
    void *frame_ring_peers[16] = {};
    int frame_ring_rank = 0, frame_ring_world = 0;
    unsigned long long frame_ring_count = 0;
    cudaEvent_t frame_ring_consumed[2] = {nullptr, nullptr};

    // pool exhaustion (the reference's unordered_map is unbounded, CubeHandler.cpp:181-191): the kernels raise sticky flags in
    // mapped host memory; the synchronous calls grow the pool and re-run the frame for the cubes that found no slot, the
    // asynchronous ones report OPB_ERR_CAPACITY at the next synchronisation
    int *h_flags = nullptr;
    long long frames_with_lost_cubes = 0;
    int grow_count = 0;
    int carry_frame_cubes = 0;                 // counters of a frame's first pass when the pool grew under it
    unsigned long long carry_updated = 0;

    int profile_acquire(opb::ProfileSlot **out);
    int profile_drain();
};

namespace opb
{
void build_frame_params(const opb_volume *v, const float *pose_cm, int depth_type, FrameParams &p);
// enqueue on the volume's stream: slots [first, first+n) back to the TSDFVoxel defaults; hash table rebuilt from slots [0, n_alloc)
int volume_reinit_slots(opb_volume *v, size_t first, size_t n);
int volume_rebuild_table(opb_volume *v, int n_alloc);
// doubles the block pool (at least min_cubes slots), keeps the cubes, rebuilds the table; OPB_ERR_CAPACITY when memory is short
int volume_grow(opb_volume *v, long long min_cubes);
int halo_drop_ghosts(opb_volume *v); // opb_halo.cu
int volume_materialize(opb_volume *v);                          // packed volumes: float mirror up to date (enqueued on the stream)
int volume_require_float(const opb_volume *v, const char *what); // OPB_ERR_UNSUPPORTED for packed volumes
// opb_meshpost.cu: TriangleMesh::ClusteringSimplify on device-resident arrays (outputs are cudaMalloc'ed), and the download
// of such a result into malloc'ed host buffers (frees the device copies)
int clustering_simplify_device(int sm_count, cudaStream_t s, const float *d_points, const float *d_colors, size_t nv, const unsigned int *d_tri,
                               size_t nt, float grid_len, float **o_points, float **o_colors, unsigned int **o_tri, size_t *o_nv, size_t *o_nt);
int mesh_result_to_host(cudaStream_t s, float *d_points, float *d_colors, unsigned int *d_tri, size_t nv, size_t nt, float **points, float **colors,
                        uint32_t **triangles);
}
