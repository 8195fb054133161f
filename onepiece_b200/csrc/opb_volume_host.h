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

    int profile_acquire(opb::ProfileSlot **out);
    int profile_drain();
};

namespace opb
{
void build_frame_params(const opb_volume *v, const float *pose_cm, int depth_type, FrameParams &p);
}
