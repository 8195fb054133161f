// Host-side state of an opb_cloud: geometry::PointCloud (reference src/Geometry/PointCloud.h:52-54) resident in HBM, together with
// the RGB-D images it was loaded from (shared by opb_cloud.cu, opb_icp.cu and opb_volume.cu).
#pragma once
#include <cuda_runtime.h>

#include "../../include/onepiece_b200.h"

struct opb_cloud
{
    int device = 0;
    cudaStream_t stream = nullptr; // uploads and the LoadFromDepth kernels run here, independent of the solver streams
    bool own_stream = false;
    cudaEvent_t ready = nullptr;   // recorded after the last enqueued operation
    // images of the frame (kept for CubeHandler::IntegrateImage on the same frame)
    void *d_depth = nullptr;
    unsigned char *d_bgr = nullptr;
    size_t cap_px = 0;
    int depth_type = 0, width = 0, height = 0;
    bool has_images = false, has_bgr = false;
    // the cloud
    float *d_xyz = nullptr, *d_nrm = nullptr;
    size_t cap_pts = 0, cap_nrm = 0;
    bool has_normals = false;
    unsigned int *d_tiles = nullptr;
    size_t cap_tiles = 0;
    unsigned int *h_count = nullptr; // mapped pinned: number of points of the last load (written by the device)
    unsigned int *d_count = nullptr; // device alias of h_count
    size_t n_host = 0;               // size when set from host arrays
    bool count_on_device = false;
};

namespace opb
{
// waits (on the host) until everything enqueued on the cloud is done; returns its size
int cloud_wait(opb_cloud *c, size_t *n);
}
